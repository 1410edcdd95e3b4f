/*
 * apbf_oracle.h -- CPU restatement of the APBF per-substep particle hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under apbf_b200/ or include/ may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the
 * reported CPU baseline.
 *
 * What it restates (all paths relative to /root/reference):
 *   shaders/cpu_gpu_shared_config.h, kernels.glsl, neighbor_add.glsl,
 *   calculate_position_hash.comp, calculate_position_code.comp,
 *   find_value_ranges.comp, neighborhood_green.comp,
 *   neighborhood_binary_search.comp, neighborhood_brute_force.comp,
 *   incompressibility_{0,1,2,3}.comp, kernel_width_init.comp, kernel_width.comp,
 *   uint_to_float_but_gradual.comp, box_collision.comp, copy_scattered_read.comp,
 *   radix_sort_*.comp / prefix_sum_*.comp (semantics: stable, inclusive),
 *   infer_velocity.comp / apply_acceleration.comp / apply_velocity.comp,
 *   find_split_and_merge_1/2/3.comp, remove_impossible_splits.comp,
 *   initialize_split_particles.comp, particle_transfer.comp,
 *   uint_to_float_with_indexed_lower_bound.comp, source/update_transfers.cpp:14-70,
 *   source/particle_transfer.cpp:10-28, indexed_list.h:126-152 (delete_these /
 *   duplicate_these: stable order, copies at the end),
 *   source/algorithms.cpp, neighborhood_green.cpp, neighborhood_binary_search.cpp,
 *   incompressibility.cpp, spread_kernel_width.cpp, box_collision.cpp,
 *   velocity_handling.cpp, pool.cpp:67-106.
 *
 * Pinning status:
 *   - sort / prefix_sum / apply_edit: PINNED by the reference's own known-answer
 *     vectors (source/test.cpp:91-105,186-197,269-282,284-305,343-364,402-425)
 *     and the Z-curve vectors of the dead sortByPositions test (test.cpp:623).
 *   - neighbour search, position hash/code, incompressibility, kernel width,
 *     box collision, update_transfers, particle_transfer (merge / split): PARITY
 *     UNPINNED by the reference -- it has no enabled test for
 *     them (test.cpp:533-583 return true) and its GLSL cannot be compiled or
 *     run in this image (no glslang, no Vulkan).  Cross-checked only by the
 *     brute-force search and known answers derived by hand from the shader text
 *     (closed forms for two particles under the Gauss kernels, an exact box push-out,
 *     spread_kernel_width / velocity_handling / update_transfers / merge / split micro
 *     cases: tests/test_oracle_kats.py, tests/test_transfers.py).  Merge / split decisions are a race in the
 *     reference (atomicExchange); the restatement runs the invocations in
 *     ascending id order.
 *
 * Floating-point conventions fixed here (GLSL leaves them to the driver):
 *   IEEE binary32, round-to-nearest-even, no FMA contraction (build with
 *   -ffp-contract=off), sqrtf and '/' correctly rounded, glibc powf/expf/log2f
 *   (the merge / split passes: pow in double, rounded to float once);
 *   dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z; length(v) = sqrtf(dot(v,v));
 *   normalize(v) = v / length(v) (component-wise division); distance(a,b) =
 *   length(a-b); float->int/uint conversions truncate toward zero, uint() of a
 *   negative value saturates to 0.
 */
#ifndef APBF_ORACLE_H
#define APBF_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_POS_RESOLUTION 262144.0f        /* cpu_gpu_shared_config.h:7  */
#define ORC_KERNEL_WIDTH_RESOLUTION 262144.0f /* :8 */
#define ORC_INCOMP_RESOLUTION 262144.0f     /* :9  */
#define ORC_KERNEL_SCALE 4.0f               /* :14 */
#define ORC_KERNEL_WIDTH_PROPAGATION_FACTOR 0.5f /* :15 */

/* cpu_gpu_shared_config.h:74-94, same field order */
typedef struct orc_settings {
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
} orc_settings;

/* incompressibility_0.comp:6-13 */
typedef struct orc_incomp_data {
	int32_t  mWeightedGradSum[3];
	uint32_t mDensity;
	uint32_t mSquaredGradSum;
	uint32_t padding[3];
} orc_incomp_data;

void orc_default_settings(orc_settings* s); /* settings.cpp:5-31 */
void orc_set_threads(int n);                /* OpenMP threads for the loops that are order-independent */

/* ---- kernels.glsl ---------------------------------------------------------------- */
float orc_kernel_height(const orc_settings* s, int dims, const float r[3], float h);
void  orc_kernel_gradient(const orc_settings* s, int dims, const float r[3], float h, float out[3]);

/* ---- algorithms.cpp -------------------------------------------------------------- */
/* stable LSD radix sort, 4-bit digits, passes while (upper_bound >> off) != 0 (algorithms.cpp:59-91) */
void orc_sort(const uint32_t* keys, const uint32_t* vals, uint32_t n, uint32_t upper_bound,
              uint32_t* out_keys, uint32_t* out_vals);
/* inclusive prefix sum (algorithms.cpp:93-118) */
void orc_prefix_sum(const uint32_t* in, uint32_t n, uint32_t* out);
size_t orc_prefix_sum_helper_length(size_t max_count); /* algorithms.cpp:48-58 */
size_t orc_sort_helper_length(size_t max_count);       /* algorithms.cpp:38-46 */
/* gather: dst[i] = src[edit[i]] for elements of `stride` bytes (copy_scattered_read.comp:21-30) */
void orc_apply_edit(const void* src, void* dst, const uint32_t* edit, uint32_t n, uint32_t stride);

/* ---- position keys --------------------------------------------------------------- */
/* calculate_position_hash.comp:23-46 */
void orc_position_hash(const int32_t* pos4, uint32_t n, const float min_pos[3], const float max_pos[3],
                       uint32_t res_log2, int dims, uint32_t* out_hash);
/* calculate_position_code.comp:23-72 */
void orc_position_code(const uint32_t* index_list, const int32_t* pos4, uint32_t n, uint32_t section,
                       uint32_t* out_code);
/* find_value_ranges.comp:16-31 (tables must be pre-zeroed by the caller, neighborhood_green.cpp:61-62) */
void orc_find_value_ranges(const uint32_t* index_list, const uint32_t* values, uint32_t n,
                           uint32_t* range_start, uint32_t* range_end);

/* ---- neighbour searches: all emit (id, idN) grouped by id, discovery order -------- */
/* return the number of pairs found (clamped to cap like neighbor_add.glsl:23-24) */
uint32_t orc_neighborhood_green_pairs(const uint32_t* index_list, const int32_t* pos4, const float* range,
                                      const uint32_t* cell_start, const uint32_t* cell_end, uint32_t n,
                                      float range_scale, const float min_pos[3], const float max_pos[3],
                                      uint32_t res_log2, int dims, uint32_t* out_pairs, uint32_t cap);
uint32_t orc_neighborhood_binary_search_pairs(const uint32_t* index_list, const int32_t* pos4,
                                              const uint32_t* code0, const uint32_t* code1, const uint32_t* code2,
                                              const float* range, uint32_t n, float range_scale,
                                              uint32_t* out_pairs, uint32_t cap);
uint32_t orc_neighborhood_brute_force_pairs(const uint32_t* index_list, const int32_t* pos4, const float* range,
                                            uint32_t n, float range_scale, uint32_t* out_pairs, uint32_t cap);

/* particle state as the reference's lists (list_definitions.h:9-20); all-fluid scene */
typedef struct orc_state {
	uint32_t  n_hidden;        /* hidden list length */
	uint32_t  n;               /* index list length (== fluid length) */
	uint32_t* index_list;      /* [n] */
	int32_t*  position;        /* [n_hidden*4] */
	float*    velocity;        /* [n_hidden*4] */
	float*    inverse_mass;    /* [n_hidden] */
	float*    radius;          /* [n_hidden] */
	int32_t*  pos_backup;      /* [n_hidden*4] */
	uint32_t* transferring;    /* [n_hidden] */
	float*    target_radius;   /* [n] per id */
	float*    kernel_width;    /* [n] per id */
	float*    boundariness;    /* [n] per id */
	uint32_t* boundary_distance; /* [n] per id */
} orc_state;

/* neighborhood_green.cpp:27-77 incl. the reorder chain.  Optional outputs
 * (may be NULL): sorted_hash[n_hidden], sorted_index[n_hidden] (new slot -> old slot),
 * cell_start/cell_end[1 << (res*dims)].  Returns the pair count. */
uint32_t orc_neighborhood_green_apply(orc_state* st, const orc_settings* s, int dims, float range_scale,
                                      const float min_pos[3], const float max_pos[3], uint32_t res_log2,
                                      uint32_t* out_pairs, uint32_t cap,
                                      uint32_t* sorted_hash, uint32_t* sorted_index,
                                      uint32_t* cell_start, uint32_t* cell_end);
/* neighborhood_binary_search.cpp:22-75; optional outputs code[3][n_hidden], sorted_index */
uint32_t orc_neighborhood_binary_search_apply(orc_state* st, const orc_settings* s, float range_scale,
                                              uint32_t* out_pairs, uint32_t cap,
                                              uint32_t* code0, uint32_t* code1, uint32_t* code2,
                                              uint32_t* sorted_index);

/* ---- incompressibility (incompressibility.cpp:12-45) ------------------------------ */
void orc_incompressibility_0(const orc_state* st, const orc_settings* s, int dims, orc_incomp_data* incomp);
void orc_incompressibility_1(const orc_state* st, const orc_settings* s, int dims, const uint32_t* pairs,
                             uint32_t n_pairs, orc_incomp_data* incomp, int32_t* com4, float* grad4);
void orc_incompressibility_2(orc_state* st, const orc_settings* s, int dims, const orc_incomp_data* incomp,
                             const int32_t* com4, float* lambda);
void orc_incompressibility_3(orc_state* st, const orc_settings* s, int dims, const uint32_t* pairs,
                             uint32_t n_pairs, const float* grad4, const float* lambda,
                             const orc_incomp_data* incomp);
/* all four; optional outputs (may be NULL) incomp[n], lambda[n] */
void orc_incompressibility_apply(orc_state* st, const orc_settings* s, int dims, const uint32_t* pairs,
                                 uint32_t n_pairs, orc_incomp_data* out_incomp, float* out_lambda);

/* ---- adaptive kernel width (spread_kernel_width.cpp:12-26) ------------------------ */
/* pairs are replaced by the pruned list (order of input kept); returns new pair count */
uint32_t orc_spread_kernel_width_apply(orc_state* st, const orc_settings* s, uint32_t* pairs, uint32_t n_pairs,
                                       uint32_t* out_kw_fixed /* optional [n] */);

/* ---- box collision (box_collision.comp:36-60) ------------------------------------- */
/* update_transfers::apply with settings::merge / settings::split off (update_transfers.cpp:14-54):
 * find_split_and_merge_1/2/3.comp.  out_nearest: optional [n], 0xFFFFFFFF where a particle has no pair. */
void orc_update_transfers_apply(orc_state* st, const orc_settings* s, const uint32_t* pairs, uint32_t n_pairs,
                                uint32_t* out_nearest);
/* pool.cpp:77-80: uint_to_float_with_indexed_lower_bound.comp, boundary distance -> kernel width */
void orc_kernel_width_from_boundary_distance(orc_state* st, const orc_settings* s);

/* list_definitions.h:16-18: hidden_transfers (source and target index the hidden particle list) */
typedef struct orc_transfers {
	uint32_t  n, cap;
	uint32_t* source;    /* [cap] hidden particle index */
	uint32_t* target;    /* [cap] */
	float*    time_left; /* [cap] > 0 merge, <= 0 split (particle_transfer.comp:35-37) */
} orc_transfers;
/* update_transfers::apply (update_transfers.cpp:14-70) with merge / split as `s` says, invocations in ascending id order.
 * The arrays of `st` must have room for hidden_cap hidden entries and id_cap ids; st->n, st->n_hidden and t->n are updated. */
void orc_update_transfers_full(orc_state* st, uint32_t hidden_cap, uint32_t id_cap, const orc_settings* s, int dims,
                               const uint32_t* pairs, uint32_t n_pairs, orc_transfers* t, float split_duration,
                               uint32_t* out_nearest);
/* particle_transfer::apply(dt) (particle_transfer.cpp:10-28, particle_transfer.comp:30-84) incl. both delete_these() */
void orc_particle_transfer_apply(orc_state* st, orc_transfers* t, int dims, float dt);
/* the transfers' source / target follow a permutation of the hidden list (sorted_index[new] = old) */
void orc_transfers_follow_reorder(orc_transfers* t, const uint32_t* sorted_index, uint32_t n_hidden);

void orc_box_collision(orc_state* st, const float* box_min4, const float* box_max4, uint32_t n_boxes);

/* ---- velocity handling (velocity_handling.cpp:15-31) ------------------------------ */
void orc_velocity_handling(orc_state* st, float dt, const float accel[3]);

/* ---- one substep in pool::update order (pool.cpp:67-106), Green search ------------ */
typedef struct orc_substep_params {
	int      dims;
	int      basic_pbf;          /* settings::basicPbf */
	int      solver_iterations;
	int      use_binary_search;  /* 0: Green, 1: binary search */
	int      integrate;          /* run velocity_handling first */
	float    dt;
	float    accel[3];
	float    min_pos[3], max_pos[3];
	uint32_t res_log2;
	uint32_t n_boxes;
	const float* box_min4;
	const float* box_max4;
	int      update_transfers;   /* pool.cpp:77-80 and :99-102 */
	orc_transfers* transfers;    /* non-NULL with mMerge / mSplit: particle_transfer (pool.cpp:73-75) and the full update_transfers */
	uint32_t hidden_cap;         /* room in the arrays of the state (hidden entries and ids) */
	float    split_duration;     /* settings::splitDuration */
} orc_substep_params;
/* returns the number of pairs left in `pairs` after the substep */
uint32_t orc_substep(orc_state* st, const orc_settings* s, const orc_substep_params* p,
                     uint32_t* pairs, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif
