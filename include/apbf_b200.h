/*
 * apbf_b200.h -- C-ABI of libapbf_b200.so: the B200 (sm_100a) implementation of APBF's per-substep
 * particle hot path (neighbour search -> PBF incompressibility solve -> adaptive kernel width).
 *
 * Boundary.  The reference (cg-tuwien/APBF) has no FFI; its hot path is reached through
 *   (1) pbd::shader_provider::<shader>(avk::buffer...)   source/shader_provider.h:15-72  (one static function per
 *       compute shader, list lengths passed as device buffers), and
 *   (2) the operator objects' apply()                     source/neighborhood_green.h:11-14,
 *       neighborhood_binary_search.h:11-13, incompressibility.h:11-12, spread_kernel_width.h:10-11,
 *       box_collision.h:11-12, velocity_handling.h:11-13, algorithms.h:12-17.
 * Every entry point below names the reference interface it replaces.  avk::buffer becomes a plain device
 * pointer; a list length stays a 4-byte word in device memory (the host never reads it on the hot path,
 * SURVEY 3.1); *_capacity arguments are the host-known upper bounds (gpu_list::requested_length()).
 * INTEGRATION.md shows the reference-side binding.
 *
 * Conventions: all functions return 0 on success or a negative apbf_status; all work is enqueued on the
 * context's CUDA stream and is asynchronous unless stated otherwise; pointers are device pointers unless the
 * name says host.  There is no CPU fallback: without a CUDA device apbf_ctx_create fails with
 * APBF_ERR_NO_DEVICE.
 */
#ifndef APBF_B200_H
#define APBF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APBF_POS_RESOLUTION 262144.0f                 /* shaders/cpu_gpu_shared_config.h:7 */
#define APBF_KERNEL_WIDTH_RESOLUTION 262144.0f        /* :8 */
#define APBF_INCOMPRESSIBILITY_DATA_RESOLUTION 262144.0f /* :9 */
#define APBF_KERNEL_SCALE 4.0f                        /* :14 */
#define APBF_KERNEL_WIDTH_PROPAGATION_FACTOR 0.5f     /* :15 */

typedef enum apbf_status {
	APBF_OK = 0,
	APBF_ERR_NO_DEVICE = -1,
	APBF_ERR_CUDA = -2,
	APBF_ERR_INVALID = -3,
	APBF_ERR_OOM = -4,
	APBF_ERR_UNSUPPORTED = -5
} apbf_status;

/* shaders/cpu_gpu_shared_config.h:74-94 -- same fields, same order (the reference's UBO) */
typedef struct apbf_settings {
	int   mHeightKernelId;
	int   mGradientKernelId;
	int   mMerge;
	int   mSplit;
	int   mBaseKernelWidthOnTargetRadius;
	int   mBaseKernelWidthOnBoundaryDistance;
	int   mUpdateTargetRadius;
	int   mUpdateBoundariness;
	int   mNeighborListSorted;
	int   mBoundarinessCalculationMethod;
	float mBoundarinessAdaptionSpeed;
	float mKernelWidthAdaptionSpeed;
	float mBoundarinessSelfGradLengthFactor;
	float mBoundarinessUnderpressureFactor;
	float mMergeDuration;
	float mSmallestTargetRadius;
	float mTargetRadiusOffset;
	float mTargetRadiusScaleFactor;
} apbf_settings;

typedef struct apbf_ctx apbf_ctx;

/* ---- context ------------------------------------------------------------------------------------------ */
/* Replaces shader_provider::set_queue / start_recording / end_recording (source/shader_provider.cpp:10-37):
 * one ordered stream of work.  `cuda_stream` is a cudaStream_t (NULL = the legacy default stream). */
int  apbf_ctx_create(int device, void* cuda_stream, apbf_ctx** out_ctx);
void apbf_ctx_destroy(apbf_ctx* ctx);
int  apbf_ctx_set_stream(apbf_ctx* ctx, void* cuda_stream);
int  apbf_ctx_synchronize(apbf_ctx* ctx);
const char* apbf_ctx_last_error(apbf_ctx* ctx);
/* settings::update_apbf_settings_buffer (source/settings.cpp:66-90); the default is settings.cpp:5-31 */
void apbf_default_settings(apbf_settings* out);
int  apbf_ctx_set_settings(apbf_ctx* ctx, const apbf_settings* s);
/* DIMENSIONS (cpu_gpu_shared_config.h:2) as a runtime value, 2 or 3 */
int  apbf_ctx_set_dimensions(apbf_ctx* ctx, int dims);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t apbf_ctx_launch_count(apbf_ctx* ctx);
/* apbf_neighborhood_green_spread_apply never materialises the unpruned pair list of neighborhood_green; with this switch
 * on (default off) it still counts how many pairs that list would have held (apbf_sim_stats: pairs searched), at the
 * price of one more compare per distance test. */
int  apbf_ctx_set_search_stats(apbf_ctx* ctx, int enable);
/* Testing aid: cap the number of 128-entry blocks of the pair emit's hit stream (0 = sized from the list capacities).  A
 * stream that runs out of blocks makes the search fall back to its two-pass fill; results do not change. */
int  apbf_ctx_set_stream_blocks(apbf_ctx* ctx, uint32_t max_blocks);
/* Tuning / testing aid: the merge / split matching of apbf_update_transfers_split_merge_apply runs its first rounds grid-wide
 * when there are at least this many candidates (default 16384; 0 = always, 0xFFFFFFFF = never).  Results do not change. */
int  apbf_ctx_set_match_grid_min(apbf_ctx* ctx, uint32_t min_candidates);
/* Per-pass device timing with CUDA events on the context stream (replaces measurements::record_timing_interval_*,
 * source/measurements.cpp:27-62).  apbf_ctx_profile(ctx, 1) clears and starts, (ctx, 0) stops; apbf_ctx_profile_read
 * returns the accumulated milliseconds and span count of category 0 .. n-1 (APBF_ERR_INVALID past the last one) and
 * synchronises with the recorded events. */
int  apbf_ctx_profile(apbf_ctx* ctx, int enable);
int  apbf_ctx_profile_read(apbf_ctx* ctx, int category, const char** out_name, double* out_ms, uint64_t* out_calls);
/* sticky device-side status word: bit 0 = neighbour list overflow (neighbor_add.glsl:23-24 clamp hit); bit 2 = a search box was
 * at least as wide as the whole grid on some axis: the reference walks aliased cells twice there and lists those pairs twice
 * (neighborhood_green.comp:69-79), this library lists every pair once.  Synchronises. */
int  apbf_ctx_device_flags(apbf_ctx* ctx, uint32_t* out_flags);
/* which form of the passes the lists of the last search / solver iteration selected on the device (diagnostics; no counterpart in
 * the reference; synchronises): out[0] = pairs without a mirrored pair in the list, out[1] = 1 if every particle has the same
 * kernel width (bitwise), out[2] = 1 if every id had the same search / prune thresholds, out[3] = occupied grid cells. */
int  apbf_ctx_list_state(apbf_ctx* ctx, uint32_t out[4]);

/* ---- buffers: gpu_list_data::get_list best-fit pool (source/gpu_list_data.cpp:6-45) + algorithms::copy_bytes -- */
int apbf_buffer_acquire(apbf_ctx* ctx, size_t bytes, void** out_dev_ptr);
int apbf_buffer_release(apbf_ctx* ctx, void* dev_ptr);
int apbf_copy_bytes(apbf_ctx* ctx, const void* src_dev, void* dst_dev, size_t bytes);        /* algorithms.cpp:4-18 */
int apbf_copy_bytes_from_host(apbf_ctx* ctx, const void* src_host, void* dst_dev, size_t bytes); /* algorithms.cpp:20-36 */
int apbf_copy_bytes_to_host(apbf_ctx* ctx, const void* src_dev, void* dst_host, size_t bytes);   /* gpu_list::read, synchronises */

/* ---- list helper kernels (shader_provider, "gpu_lists" group) ------------------------------------------ */
/* write_sequence.comp:21-28: out[i] = start + i*step for i < *len * len_scale */
int apbf_write_sequence(apbf_ctx* ctx, uint32_t* out, const uint32_t* len, uint32_t capacity,
                        uint32_t start, uint32_t step, uint32_t len_scale);
/* write_sequence_float.comp */
int apbf_write_sequence_float(apbf_ctx* ctx, float* out, const uint32_t* len, uint32_t capacity, float start, float step);
/* copy_scattered_read.comp:21-30 (gpu_list::apply_edit, gpu_list.h:126-135): dst[i] = src[edit[i]], i < *edit_len */
int apbf_copy_scattered_read(apbf_ctx* ctx, const void* src, void* dst, const uint32_t* edit,
                             const uint32_t* edit_len, uint32_t capacity, uint32_t stride_bytes);
/* scattered_write.comp: target[index[i]] = value, i < *len */
int apbf_scattered_write(apbf_ctx* ctx, const uint32_t* index, uint32_t* target, const uint32_t* len,
                         uint32_t capacity, uint32_t value);
/* append_list.comp:21-29: target[*target_len ...] = appending[0 .. *appending_len); *new_len = min(sum, target_capacity) */
int apbf_append_list(apbf_ctx* ctx, void* target, const void* appending, const uint32_t* target_len,
                     const uint32_t* appending_len, uint32_t* new_len, uint32_t target_capacity,
                     uint32_t appending_capacity, uint32_t stride_bytes);
/* copy_with_differing_stride.comp:21-32 */
int apbf_copy_with_differing_stride(apbf_ctx* ctx, const void* src, void* dst, const uint32_t* len, uint32_t capacity,
                                    uint32_t src_stride_bytes, uint32_t dst_stride_bytes);
/* find_value_changes.comp:16-30 (indexed_list::delete_these, source/indexed_list.h:126-138): `in` is a running count;
 * out_change[k] = first i with in[i] > k; *out_change_len = in[*in_len - 1] (untouched when *in_len == 0) */
int apbf_find_value_changes(apbf_ctx* ctx, const uint32_t* in, uint32_t* out_change, const uint32_t* in_len,
                            uint32_t* out_change_len, uint32_t capacity);
/* write_increasing_sequence.comp:23-36 (shader_provider.cpp:198-219; indexed_list::increase_length, indexed_list.h:194-202):
 * L = min(target_capacity, sequence_length, value_upper_bound - *sequence_min_value); target[i] = *sequence_min_value + i
 * for i < L; *new_sequence_min_value = *sequence_min_value + L; *new_target_len = L.  The output words may alias the input. */
int apbf_write_increasing_sequence(apbf_ctx* ctx, uint32_t* target, uint32_t target_capacity, uint32_t* new_target_len,
                                   const uint32_t* sequence_min_value, uint32_t* new_sequence_min_value,
                                   uint32_t value_upper_bound, uint32_t sequence_length);
/* write_increasing_sequence_from_to.comp:19-31 (indexed_list::duplicate_these, indexed_list.h:141-152):
 * out[i] = *from + i while *from + i < *to; *out_len = *to - *from */
int apbf_write_increasing_sequence_from_to(apbf_ctx* ctx, uint32_t* out, uint32_t* out_len, const uint32_t* from, const uint32_t* to,
                                           uint32_t capacity);
/* indexed_list::apply_hidden_edit (source/indexed_list.h:289-308: atomic_swap.comp + generate_new_index_and_edit_list.comp)
 * followed by the index-list sort of indexed_list::sort (:276-286): for every new hidden slot h (ascending) and every
 * entry i of index_list with index_list[i] == edit[h], emit new_index = h and new_edit = i.  Result is ascending in
 * new_index (the order the reference reaches after its sort).  *new_len = number emitted (clamped to index_capacity). */
int apbf_apply_hidden_edit(apbf_ctx* ctx, const uint32_t* edit, const uint32_t* edit_len, uint32_t edit_capacity,
                           const uint32_t* index_list, const uint32_t* index_len, uint32_t index_capacity,
                           uint32_t hidden_capacity, uint32_t* new_index_list, uint32_t* new_edit_list, uint32_t* new_len);

/* ---- algorithms (source/algorithms.h:12-17) ------------------------------------------------------------- */
/* algorithms::sort: stable ascending sort of u32 keys with u32 payload; like the reference it may clobber the inputs
 * and only looks at the key bits below the highest set bit of upper_bound (algorithms.cpp:73).  Implemented as an
 * onesweep radix sort (8-bit digits, decoupled look-back); no helper list is needed. */
int apbf_sort(apbf_ctx* ctx, uint32_t* keys, uint32_t* values, const uint32_t* count, uint32_t max_count,
              uint32_t* out_keys, uint32_t* out_values, uint32_t upper_bound);
/* algorithms::prefix_sum: inclusive scan, in place when result == values (algorithms.cpp:93-101) */
int apbf_prefix_sum(apbf_ctx* ctx, const uint32_t* values, const uint32_t* count, uint32_t max_count, uint32_t* result);
size_t apbf_sort_calculate_needed_helper_list_length(size_t max_count);       /* algorithms.cpp:38-46 (kept for callers) */
size_t apbf_prefix_sum_calculate_needed_helper_list_length(size_t max_count); /* algorithms.cpp:48-58 */

/* ---- position keys --------------------------------------------------------------------------------------- */
/* shader_provider::calculate_position_hash (shader_provider.cpp:805; calculate_position_hash.comp:23-46) */
int apbf_calculate_position_hash(apbf_ctx* ctx, const int32_t* position4, uint32_t* out_hash, const uint32_t* len,
                                 uint32_t capacity, const float min_pos[3], const float max_pos[3], uint32_t res_log2);
/* shader_provider::calculate_position_code (shader_provider.cpp:782; calculate_position_code.comp:23-72) */
int apbf_calculate_position_code(apbf_ctx* ctx, const uint32_t* index_list, const int32_t* position4, uint32_t* out_code,
                                 const uint32_t* len, uint32_t capacity, uint32_t code_section);
/* shader_provider::find_value_ranges (shader_provider.cpp:221; find_value_ranges.comp:16-31); zero-fills both tables
 * first like neighborhood_green.cpp:61-62 */
int apbf_find_value_ranges(apbf_ctx* ctx, const uint32_t* index_list, const uint32_t* values, uint32_t* range_start,
                           uint32_t* range_end, const uint32_t* len, uint32_t capacity, uint32_t n_ranges);

/* ---- the reference's list schema as raw arrays (source/list_definitions.h:9-20) --------------------------- */
/* One array of a list that the search re-orders: `data` holds the current content; the permuted content is written
 * to `reorder_out` (a second buffer of the same capacity -- the gpu_list copy-on-write target of
 * gpu_list::apply_edit, gpu_list.h:131-133).  After a search the caller uses reorder_out as the list's buffer. */
typedef struct apbf_array {
	void* data;
	void* reorder_out;
} apbf_array;

/* pbd::particles = indexed_list<hidden_particles> */
typedef struct apbf_particles {
	apbf_array index_list;      /* u32   [capacity]          position in the index list = particle id          */
	uint32_t*  length;          /* device word: index list length                                             */
	uint32_t   capacity;
	uint32_t*  hidden_length;   /* device word: hidden list length                                            */
	uint32_t   hidden_capacity;
	apbf_array position;        /* ivec4 [hidden_capacity]   fixed point * 2^18, w unused                     */
	apbf_array velocity;        /* vec4                                                                       */
	apbf_array inverse_mass;    /* f32                                                                        */
	apbf_array radius;          /* f32                                                                        */
	apbf_array pos_backup;      /* ivec4                                                                      */
	apbf_array transferring;    /* u32                                                                        */
} apbf_particles;

/* pbd::fluid = uninterleaved_list<fluid_enum, particles, gpu_list<4> x4>; per-id arrays share particles.length */
typedef struct apbf_fluid {
	apbf_particles particle;
	apbf_array target_radius;      /* f32 [capacity] */
	apbf_array kernel_width;       /* f32 */
	apbf_array boundariness;       /* f32 */
	apbf_array boundary_distance;  /* u32 */
} apbf_fluid;

/* pbd::neighbors = gpu_list<8>: (id, idN) pairs of index-list positions */
typedef struct apbf_neighbors {
	uint32_t* pairs;      /* uvec2 [capacity] */
	uint32_t* length;     /* device word      */
	uint32_t  capacity;
} apbf_neighbors;

/* Optional read-outs of the search's intermediate results (any member may be NULL) */
typedef struct apbf_search_debug {
	uint32_t* sorted_key;     /* [hidden_capacity] sorted cell hash (Green) / unused (binary search)              */
	uint32_t* sorted_index;   /* [hidden_capacity] new hidden slot -> old hidden slot                             */
	uint32_t* cell_start;     /* [1 << (res*dims)] (Green)                                                        */
	uint32_t* cell_end;
	uint32_t* code[3];        /* [capacity] sorted 96-bit Morton code sections (binary search)                    */
	uint32_t* pair_offsets;   /* [capacity + 1] pairs of id are [pair_offsets[id], pair_offsets[id + 1]) of the list */
} apbf_search_debug;

/* ---- operators --------------------------------------------------------------------------------------------- */
/* pbd::neighborhood_green::set_data(...).set_range_scale(s).set_position_range(min,max,res).apply()
 * (source/neighborhood_green.cpp:27-77).  `range` is the per-id range list (pool.cpp:45 passes fluid.kernel_width;
 * pass the same apbf_array so that it is re-ordered with the particles).  Re-orders every array of `fluid` into its
 * reorder_out buffer, rewrites the index list, and fills `neighbors` grouped by id in the reference's discovery
 * order (valid for both values of mNeighborListSorted).  Particles must lie inside [min,max). */
int apbf_neighborhood_green_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_array* range, apbf_neighbors* neighbors,
                                  float range_scale, const float min_pos[3], const float max_pos[3], uint32_t res_log2,
                                  const apbf_search_debug* debug);
/* neighborhood_green::apply() immediately followed by spread_kernel_width::apply() on the same lists, i.e. source/pool.cpp:83-89
 * (`range` = fluid->kernel_width), in one pass: the width spread (kernel_width.comp:49-53) is gathered while the
 * candidates are tested and only the pairs that survive the prune (kernel_width.comp:57) are written.  Same kernel widths
 * and the same neighbour list, pair for pair and in the same order, as the two calls one after the other -- except that
 * `neighbors->capacity` only has to hold the pruned list (the reference clamps the unpruned one, neighbor_add.glsl:23-24).
 * out_kw_fixed (optional, device, [capacity]): the fixed-point widths after the atomicMax spread.  Single GPU only. */
int apbf_neighborhood_green_spread_apply(apbf_ctx* ctx, apbf_fluid* fluid, apbf_neighbors* neighbors, float range_scale,
                                         const float min_pos[3], const float max_pos[3], uint32_t res_log2,
                                         const apbf_search_debug* debug, uint32_t* out_kw_fixed);
/* The same fusion for the binary search (NEIGHBORHOOD_TYPE 3): neighborhood_binary_search::apply() followed by
 * spread_kernel_width::apply() on the same lists with range = kernel_width (pool.cpp:83-89).  Same lists afterwards, pair for
 * pair, as calling the two entry points in turn. */
int apbf_neighborhood_binary_search_spread_apply(apbf_ctx* ctx, apbf_fluid* fluid, apbf_neighbors* neighbors, float range_scale,
                                                 const apbf_search_debug* debug, uint32_t* out_kw_fixed);
/* pbd::neighborhood_binary_search::set_data(...).set_range_scale(s).apply() (source/neighborhood_binary_search.cpp:22-75) */
int apbf_neighborhood_binary_search_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_array* range,
                                          apbf_neighbors* neighbors, float range_scale, const apbf_search_debug* debug);
/* ---- operators over a pair list: ANY pbd::neighbors list (source/incompressibility.h:11, spread_kernel_width.h:10,
 * update_transfers.h) -- written by a search of this context, by another context, by the caller, copied, appended to, edited.
 * The sweeps run on a grouped form of the list (CSR offsets + 4 bytes per pair) that a search of this context leaves behind for
 * the buffer it filled; for every other list that form is built on the device from the (id, idN) pairs at the first use (a
 * stable sort by id; pairs with an id beyond the particle list are dropped, duplicates are kept) and remembered per pair
 * buffer -- a context holds any number of lists.  Library calls that write into a pair buffer (apbf_copy_bytes, apbf_append_list,
 * apbf_copy_scattered_read, ...) or recycle it (apbf_buffer_acquire / _release) mark what is remembered as out of date.
 * A caller that rewrites a pair buffer with its OWN kernels says so with apbf_neighbors_invalidate (pbd::gpu_list<8>::write()
 * in include/apbf_pbd.hpp does). */
/* the pair list at nb->pairs was written by somebody else since an operator last saw it */
int apbf_neighbors_invalidate(apbf_ctx* ctx, const apbf_neighbors* nb);
/* the buffer at nb->pairs is about to be freed or reused for something else */
int apbf_neighbors_release(apbf_ctx* ctx, const apbf_neighbors* nb);
/* build (or look up) the grouped form now instead of inside the next operator; out_pair_offsets (optional, device,
 * [capacity + 1]): pairs of id occupy [offsets[id], offsets[id + 1]) of the list sorted by id */
int apbf_neighbors_prepare(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, uint32_t* out_pair_offsets);
/* pbd::incompressibility::set_data(fluid, neighbors).apply() (source/incompressibility.cpp:12-45).  Works on the
 * arrays' `data` buffers in place.  Optional outputs (may be NULL): lambda[capacity], incomp_data[capacity*8]
 * ({ivec3 gradSum; uint density; uint sqGradSum; pad x3}, incompressibility_0.comp:6-13). */
int apbf_incompressibility_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* neighbors,
                                 float* out_lambda, uint32_t* out_incomp_data);
/* pbd::spread_kernel_width::set_data(fluid, neighbors).apply() (source/spread_kernel_width.cpp:12-26); the pruned pair
 * list replaces the content of `neighbors`.  Optional output kw_fixed[capacity] (the atomicMax target). */
int apbf_spread_kernel_width_apply(apbf_ctx* ctx, apbf_fluid* fluid, apbf_neighbors* neighbors, uint32_t* out_kw_fixed);
/* pbd::update_transfers::set_data(fluid, neighbors, transfers).apply() (source/update_transfers.cpp:14-54) with
 * settings::merge and settings::split off: one step of the boundary-distance flood fill and the nearest
 * neighbour (find_split_and_merge_1/2.comp), target radius, boundary-distance decay and boundariness threshold
 * (find_split_and_merge_3.comp:55-84).  The merge / split bookkeeping (:88-121) is not performed, whatever mMerge / mSplit
 * say.  Optional output nearest_neighbor[capacity] (0xFFFFFFFF for a particle without pairs; among several pairs at the
 * minimum distance the last one of the list -- the reference leaves that to a race). */
int apbf_update_transfers_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* neighbors, uint32_t* out_nearest_neighbor);
/* pbd::transfers = indexed_list<hidden_transfers> (source/list_definitions.h:16-18): rows (source, target, time_left); source and
 * target are indices into the hidden particle list (they share its data, pool.cpp:18-19); time_left > 0 is a merge, <= 0 a split
 * (particle_transfer.comp:35-37).  Like the particle arrays every list has a second buffer for the list edits. */
typedef struct apbf_transfers {
	apbf_array source;      /* u32 [capacity] */
	apbf_array target;      /* u32 */
	apbf_array time_left;   /* f32 */
	uint32_t*  length;      /* device word */
	uint32_t   capacity;    /* MAX_TRANSFERS */
} apbf_transfers;
/* pbd::update_transfers::set_data(fluid, neighbors, transfers).apply() in full (source/update_transfers.cpp:14-70): what
 * apbf_update_transfers_apply does, then the merge / split decisions of find_split_and_merge_3.comp:86-121 as mMerge / mSplit
 * say, remove_impossible_splits.comp, duplicate_these (indexed_list.h:141-152), initialize_split_particles.comp and the three
 * appends to the transfer lists.  Works in place on the `data` buffers; the list lengths grow by the number of splits.  The
 * reference resolves conflicting candidates and orders its output lists by atomics (a race); this call returns the result of
 * running the invocations in ascending id order, with copies appended in that order.  A particle without pairs never merges. */
int apbf_update_transfers_split_merge_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* neighbors, apbf_transfers* transfers,
                                            float split_duration, uint32_t* out_nearest_neighbor);
/* pbd::particle_transfer::set_data(fluid, transfers).apply(dt) (source/particle_transfer.cpp:10-28, particle_transfer.comp:30-84)
 * including deleteTransferList.delete_these() and deleteParticleList.delete_these() (indexed_list.h:126-138): finished splits
 * leave the transfer list, the sources of finished merges leave the hidden particle list, and every list that shares it follows.
 * Like a search, the call writes the edited lists of `fluid` AND `transfers` into their reorder_out buffers (surviving entries
 * keep their order); afterwards the caller uses reorder_out as the lists' buffers.  out_hidden_edit (optional, device,
 * [hidden_capacity]): new hidden slot -> old hidden slot for the first *hidden_length entries afterwards -- the edit list that
 * other lists sharing the hidden particles follow (indexed_list::apply_hidden_edit). */
int apbf_particle_transfer_apply(apbf_ctx* ctx, apbf_fluid* fluid, apbf_transfers* transfers, float dt, uint32_t* out_hidden_edit);
/* The searches permute the hidden particle list; the transfers' source / target lists share it and follow through
 * indexed_list::apply_hidden_edit (source/indexed_list.h:289-308).  sorted_index[new slot] = old slot (apbf_search_debug). */
int apbf_transfers_follow_reorder(apbf_ctx* ctx, apbf_transfers* transfers, const uint32_t* sorted_index, const uint32_t* hidden_length,
                                  uint32_t hidden_capacity);
/* pool.cpp:77-80: shader_provider::uint_to_float_with_indexed_lower_bound(boundary_distance -> kernel_width, factor
 * targetRadiusScaleFactor / POS_RESOLUTION, lower bound radius * KERNEL_SCALE, step kernelWidthAdaptionSpeed) */
int apbf_kernel_width_from_boundary_distance(apbf_ctx* ctx, apbf_fluid* fluid);
/* shader_provider::uint_to_float_with_indexed_lower_bound (source/shader_provider.h:27, uint_to_float_with_indexed_lower_bound.comp:
 * 32-44) with its own arguments: out[id] = max(move_towards(out[id], in[id] * factor, max_adaption_step),
 * lower_bound[index_list[id]] * lower_bound_factor) for id < *len */
int apbf_uint_to_float_with_indexed_lower_bound(apbf_ctx* ctx, const uint32_t* in_uint, float* out_float, const uint32_t* index_list,
                                                const float* lower_bound, const uint32_t* len, uint32_t capacity, float factor,
                                                float lower_bound_factor, float max_adaption_step);
/* pbd::box_collision::set_data(particles, boxMin, boxMax).apply() (source/box_collision.cpp:12-24); boxes are vec4
 * device arrays, n_boxes host-known (user_controlled_boxes owns them, <= 64) */
int apbf_box_collision_apply(apbf_ctx* ctx, apbf_particles* particles, const float* box_min4, const float* box_max4,
                             uint32_t n_boxes);
/* pbd::velocity_handling::apply(dt) (source/velocity_handling.cpp:15-31); last_dt is the class's mLastDeltaTime */
int apbf_velocity_handling_apply(apbf_ctx* ctx, apbf_particles* particles, float dt, float last_dt, const float accel[3]);

/* ---- a whole scene: lists resident in HBM, one substep in pool::update order (source/pool.cpp:67-106) -------- */
typedef struct apbf_sim apbf_sim;

typedef struct apbf_sim_config {
	uint32_t particle_capacity;    /* pool.cpp:7,12 (100000 there) */
	uint32_t neighbor_capacity;    /* pool.cpp:15 (10000000 there) */
	int      dims;
	int      basic_pbf;            /* settings::basicPbf */
	int      solver_iterations;    /* settings::solverIterations */
	int      use_binary_search;    /* NEIGHBORHOOD_TYPE 3 instead of 1 */
	int      integrate;            /* run velocity_handling at the start of each substep */
	float    dt;                   /* FIXED_TIME_STEP */
	float    accel[3];             /* pool.cpp:28 */
	float    min_pos[3], max_pos[3];
	uint32_t res_log2;             /* neighborhood_green::set_position_range */
	uint32_t n_boxes;
	const float* box_min4_host;    /* vec4 [n_boxes] */
	const float* box_max4_host;
	int      update_transfers;     /* pool.cpp:77-80 and :99-102: kernel width from the boundary distance before the search (default
	                                  adaptive mode) and update_transfers::apply after the solver */
	int      transfers;            /* pool.cpp:73-75 and :99-102 with settings::merge / settings::split: particle_transfer::apply after
	                                  velocity_handling and the full update_transfers::apply after the solver (as mMerge / mSplit of the
	                                  context's settings say); particle_capacity is the room for the copies */
	uint32_t transfer_capacity;    /* MAX_TRANSFERS (0: particle_capacity) */
	float    split_duration;       /* settings::splitDuration */
} apbf_sim_config;

/* host-side view of the scene's lists (the reference's list schema, plain host arrays) */
typedef struct apbf_host_state {
	uint32_t  n;                /* number of particles (hidden length == index length == fluid length) */
	uint32_t* index_list;       /* may be NULL on upload: identity */
	int32_t*  position;         /* [n*4] */
	float*    velocity;         /* [n*4] */
	float*    inverse_mass;
	float*    radius;
	int32_t*  pos_backup;       /* [n*4] */
	uint32_t* transferring;
	float*    target_radius;
	float*    kernel_width;
	float*    boundariness;
	uint32_t* boundary_distance;
} apbf_host_state;

int  apbf_sim_create(apbf_ctx* ctx, const apbf_sim_config* cfg, apbf_sim** out_sim);
void apbf_sim_destroy(apbf_sim* sim);
/* host -> device copy of all lists (members that are NULL are skipped); async on the context stream when the host
 * memory is pinned */
int  apbf_sim_upload(apbf_sim* sim, const apbf_host_state* host);
/* device -> host copy (members that are NULL are skipped); synchronises */
int  apbf_sim_download(apbf_sim* sim, apbf_host_state* host);
/* n_substeps x pool::update (pool.cpp:67-106): [velocity_handling] -> [particle_transfer, cfg.transfers] -> [kernel width from
 * the boundary distance, cfg.update_transfers] -> neighbour search -> [spread_kernel_width] -> solver_iterations x
 * (box_collision, incompressibility) -> [update_transfers, with the merge / split decisions if cfg.transfers].  Fully asynchronous.
 * The search keeps the pair list in its grouped form only; apbf_sim_neighbors() writes the (id, idN) pairs when asked. */
int  apbf_sim_substep(apbf_sim* sim, uint32_t n_substeps);
/* host buffers in -> one substep -> host buffers out, the copies overlapped with the work: positions, velocities and back-ups go up
 * first, the small lists follow on a second stream while the positions are hashed and sorted; from the search's re-order on, every
 * list the solver does not touch (56 of 80 bytes per particle) comes down while emit and solver run.  Same results as
 * apbf_sim_upload + apbf_sim_substep(1) + apbf_sim_download, which is what runs with merge / split, the binary search or slabs.
 * host_in and host_out (pinned memory) may be the same buffers; out->n is set; synchronises. */
int  apbf_sim_step_host(apbf_sim* sim, const apbf_host_state* host_in, apbf_host_state* host_out);
/* A substep is the same ~36 launches every time (lengths are device words), so after two ordinary substeps apbf_sim_substep
 * captures it as a CUDA graph -- one per buffer parity, the search swaps the two buffers of every list -- and replays that from
 * then on; a change of settings, capacities or scratch allocations falls back to ordinary launches and captures again.  On by
 * default (off while apbf_ctx_profile times the passes, with merge / split, and on slabs); enable = 0 turns it off for this scene,
 * the environment variable APBF_SIM_GRAPHS=0 for all.  apbf_sim_graph_replays: substeps that ran as a graph so far. */
int  apbf_sim_set_graphs(apbf_sim* sim, int enable);
int  apbf_sim_graph_replays(const apbf_sim* sim, uint64_t* out_count);
/* views of the device-resident lists for callers that want to run single operators on them (apbf_sim_neighbors enqueues the
 * kernel that writes the public (id, idN) list if the last substep has not) */
int  apbf_sim_fluid(apbf_sim* sim, apbf_fluid* out_fluid);
int  apbf_sim_neighbors(apbf_sim* sim, apbf_neighbors* out_neighbors);
/* pair count of the last search (device -> host read, synchronises) */
int  apbf_sim_neighbor_count(apbf_sim* sim, uint32_t* out_count);
/* the transfer list of a scene created with cfg.transfers: device view, and a synchronising read-back into host arrays of
 * transfer_capacity entries (any of them may be NULL) */
int  apbf_sim_transfers(apbf_sim* sim, apbf_transfers* out_transfers);
int  apbf_sim_download_transfers(apbf_sim* sim, uint32_t* out_n, uint32_t* source_host, uint32_t* target_host, float* time_left_host);
/* counters of the last substep (device -> host read, synchronises):
 * out[0] particles, out[1] pairs found by the search (unclamped), out[2] pairs after the spread_kernel_width prune
 * (== out[1] when it did not run), out[3] pairs without a mirrored pair */
int  apbf_sim_stats(apbf_sim* sim, uint32_t out[4]);
/* ---- one scene across the GPUs of a node: slab (brick) partition with ghost particles -------------------------------------
 * No reference counterpart (the reference is single-device); device side of the protocol described in
 * apbf_b200/csrc/mgpu.cu.  One apbf_sim per rank; the host moves the staging buffers between ranks (NCCL send/recv).
 * rank = top log2(world) bits of the particle's cell key (world in {1, 2, 4, 8}); halo_range = upper bound of
 * range_scale * kernel width over the whole scene.  The bricks are cells of the Green grid (cfg.min_pos / max_pos / res_log2)
 * whichever search the scene uses: with cfg.use_binary_search the owned particles and the ghosts are each sorted by their 96-bit
 * code (one more stable pass keeps the owned ids dense) and a chunk's 27 code ranges are looked up in both. */
int  apbf_sim_mg_enable(apbf_sim* sim, int rank, int world, float halo_range);
int  apbf_sim_mg_brick(apbf_sim* sim, int rank, uint32_t out_lo[3], uint32_t out_hi[3], uint32_t out_halo[3]);
/* list lengths = n_total (owned + ghosts), ids >= n_owned are ghosts, global id of local id 0 (box_collision hashes the id) */
int  apbf_sim_mg_set_counts(apbf_sim* sim, uint32_t n_owned, uint32_t n_total, uint32_t gid_base);
/* re-partition: group the owned particles by destination rank (stable), counts_dev[8] = group sizes */
int  apbf_sim_mg_route(apbf_sim* sim, uint32_t* counts_dev);
/* 80-byte records of every list of a particle, for the particles that change owner */
int  apbf_sim_mg_pack_state(apbf_sim* sim, uint32_t first, uint32_t count, void* out_dev);
int  apbf_sim_mg_unpack_state(apbf_sim* sim, uint32_t first, uint32_t count, const void* in_dev, int into_other);
int  apbf_sim_mg_copy_state(apbf_sim* sim, uint32_t src_first, uint32_t dst_first, uint32_t count); /* current -> other buffers */
int  apbf_sim_mg_swap(apbf_sim* sim);
/* send lists: ids_dev[world][cap_per_dest] = owned ids inside rank r's brick grown by the halo; counts_dev[8] */
int  apbf_sim_mg_halo_lists(apbf_sim* sim, uint32_t* ids_dev, uint32_t cap_per_dest, uint32_t* counts_dev);
/* what: 0 halo record (32 B), 1 kernel width (4 B), 2 packed solver position (16 B), 3 lambda (4 B) */
int  apbf_sim_mg_pack(apbf_sim* sim, int what, const uint32_t* ids_dev, uint32_t count, void* out_dev);
int  apbf_sim_mg_unpack(apbf_sim* sim, int what, const uint32_t* ids_dev, uint32_t first, uint32_t count, const void* in_dev);
/* after the search: slots before the sort -> ids after it (first != 0xFFFFFFFF: ids_dev[k] = first + k beforehand) */
int  apbf_sim_mg_remap(apbf_sim* sim, uint32_t* ids_dev, uint32_t count, uint32_t first);
/* phase: 0 integrate, 1 search, 2 spread_kernel_width, 3 solver constants, 4 iteration prologue, 5 density/lambda sweep,
 * 6 apply sweep, 7 final commit (pool.cpp:67-106 cut where the halo exchanges happen); 10 / 11 = phase 6 in the form that also
 * commits the shifts and (10) runs the next iteration's prologue where the list allows it -- phases 4 and 7 then return at once */
int  apbf_sim_mg_phase(apbf_sim* sim, int phase, int iteration);
/* The library's own NCCL communicator for the halo exchanges (libnccl.so.2 is bound at run time with dlopen; nothing links
 * against it).  Rank 0 creates the id, the caller distributes its 128 bytes, every rank of the node calls comm_init. */
int  apbf_mg_nccl_unique_id(void* out_id128);
int  apbf_sim_mg_comm_init(apbf_sim* sim, const void* id128, int rank, int world);
/* Phases 3-7 of one substep with the exchanges between them in one call, everything on the context's stream:
 * [kernel widths to the ghosts] -> solver constants -> iterations x (prologue, packed positions to the ghosts, density /
 * lambda sweep, lambdas to the ghosts, apply sweep) -> final commit.  send_ids_dev: the send lists of all destinations,
 * destination after destination; ghost_ids_dev: the ghost slots, source after source; counts: host arrays [world]. */
int  apbf_sim_mg_solve(apbf_sim* sim, const uint32_t* send_ids_dev, const uint32_t* send_counts, const uint32_t* ghost_ids_dev,
                       const uint32_t* ghost_counts, int exchange_kernel_width, int iterations);

/* ---- the slab protocol driven by the library itself --------------------------------------------------------------------------
 * apbf_sim_mg_substep = pool::update (pool.cpp:67-106) of ONE scene in bricks, with route, halo, search, solve and every exchange
 * between the ranks (grouped ncclSend / ncclRecv on the library's own communicator, apbf_sim_mg_comm_init) enqueued on the context's
 * stream by this one call.  No host read-back: every count is a device word and every message has the fixed size agreed at set-up.
 * Set-up, once per scene and rank: apbf_sim_mg_enable, apbf_sim_upload of the rank's brick, apbf_sim_mg_comm_init,
 * apbf_sim_mg_halo_counts (synchronous; the caller exchanges the numbers between the ranks and picks halo_cap[r] >= both directions
 * of the pair (rank, r), with head-room), apbf_sim_mg_loop_init. */
int  apbf_sim_mg_halo_counts(apbf_sim* sim, uint32_t n_owned, uint32_t out_counts_host[8]);
int  apbf_sim_mg_loop_init(apbf_sim* sim, uint32_t n_owned, uint32_t route_cap, const uint32_t halo_cap[8]);
int  apbf_sim_mg_loop_reset(apbf_sim* sim, uint32_t n_owned, uint32_t gid_base);
/* out[0] owned particles, [1] owned + ghosts, [2] global id of local id 0, [3] particles that left in the last substep, [4] flags
 * (1 migration buffer overflow, 2 ghost list overflow, 4 particle capacity exceeded: raise the capacities), [5] exchanges so far */
int  apbf_sim_mg_loop_stats(apbf_sim* sim, uint32_t out[8]);
int  apbf_sim_mg_substep(apbf_sim* sim, uint32_t n_substeps);
/* Peer-to-peer transport for the exchanges of apbf_sim_mg_substep (one process per GPU, all on one NVLink / NVSwitch node): after
 * apbf_sim_mg_loop_init every rank exports a description of its receive buffers (CUDA IPC handle + offsets, APBF_MG_P2P_BLOB_BYTES
 * plain bytes), the caller gathers the blobs of all ranks -- over any channel it has -- and every rank imports the table.  From then
 * on the pack kernels store each message straight into the receiver's buffer over NVLink and raise a flag there, and the first
 * kernel that reads a message waits on its own flag: no NCCL call, no staging copy, no host involvement on the data path.
 * apbf_sim_mg_p2p_import returns APBF_ERR_UNSUPPORTED when a peer's memory cannot be mapped (no peer access): the NCCL transport
 * then stays in place. */
#define APBF_MG_P2P_BLOB_BYTES 512
int  apbf_sim_mg_p2p_export(apbf_sim* sim, void* out_blob);
int  apbf_sim_mg_p2p_import(apbf_sim* sim, const void* blobs_of_all_ranks);
int  apbf_sim_mg_p2p_active(const apbf_sim* sim);

/* pinned host memory helpers for callers without a CUDA runtime of their own */
int  apbf_host_alloc_pinned(size_t bytes, void** out_host_ptr);
int  apbf_host_free_pinned(void* host_ptr);

#ifdef __cplusplus
}
#endif
#endif /* APBF_B200_H */
