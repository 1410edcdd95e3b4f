// neighbors.cuh -- the grouped neighbour-list structure shared between the searches (neighbors.cu) and the sweeps
// over it (solver.cu, incompress.cu).
#pragma once
#include "common.cuh"

// The searches emit pairs grouped by id: pairs [offsets[id], offsets[id+1]) all have .x == id.  Next to the public
// (id, idN) list (the reference's pbd::neighbors layout) they write the solver's own 4-byte list
//     NB[e] = idN | (unmirrored << 31)
// where "mirrored" says that the pair (idN, id) is in the list as well; the sweeps gather through mirrored pairs and
// fall back to integer atomics only for the unmirrored ones (variable kernel widths).  ids stay below 2^31 (the pair
// counter is a u32).  The structure (offsets + NB) lives in the context's scratch memory and belongs to the
// pair buffer it was built for.
constexpr uint32_t NB_UNMIRRORED = 0x80000000u;
constexpr uint32_t NB_ID_MASK = 0x7FFFFFFFu;

// ---- whose structure is it? (nbrlist.cu) ---------------------------------------------------------------------------------
// The reference's operators take ANY pbd::neighbors list (source/incompressibility.h:11, spread_kernel_width.h:10,
// update_transfers.h): one a search of this context wrote, one of another context, one the caller built or edited by hand.
// A context therefore keeps one structure per public pair buffer (keyed by its address): the active one in SLOT_OFFSETS /
// SLOT_NB, the others parked.  A structure is VALID when it describes the buffer's current content; every library call
// that hands the buffer out for writing, writes into it, or returns it to the pool makes it invalid, and the next operator
// that needs it rebuilds it from the public list on the device (stable sort by id, then the mirrored bits).
// make nb's structure the active one (parks the one that was); after the call ctx->nbr_valid says whether it can be used
int  apbf_nbr_activate(apbf_ctx* ctx, const apbf_neighbors* nb);
// a search / prune has just written the active structure for nb
void apbf_nbr_built(apbf_ctx* ctx, const apbf_neighbors* nb, uint32_t n_cap, bool public_written);
// activate + rebuild from the public list unless valid; what every operator over a pair list calls first
int  apbf_nbr_ensure(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb);
// write the public (id, idN) list of nb from its structure if the search that built it skipped that (whole-scene path)
int  apbf_nbr_materialize(apbf_ctx* ctx, const uint32_t* len, uint32_t n_cap, const apbf_neighbors* nb);
// the buffer at `ptr` was freed or recycled: drop its structure;  a library call wrote into the buffer: structure invalid
void apbf_nbr_forget(apbf_ctx* ctx, const void* ptr);
void apbf_nbr_touch(apbf_ctx* ctx, const void* ptr);
// the particle lists the active structure was built over were compacted / re-indexed (particle_transfer): ids changed meaning
void apbf_nbr_particles_changed(apbf_ctx* ctx);
int  apbf_launch_expand_pairs(apbf_ctx* ctx, const uint32_t* offsets, const uint32_t* nbl, const uint32_t* len, uint32_t n_cap,
                              uint32_t* pairs, uint32_t pair_cap);

// the searches proper (neighbors.cu); write_public == false skips the 8-byte (id, idN) stores: the whole-scene path reads
// only the structure (apbf_sim_neighbors() materialises the list on demand)
int apbf_green_search(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_array* range, apbf_neighbors* nb, float range_scale,
                      const float min_pos[3], const float max_pos[3], uint32_t res_log2, const apbf_search_debug* dbg, bool fuse_kw,
                      uint32_t* out_kw_fixed, bool write_public);
int apbf_binary_search(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_array* range, apbf_neighbors* nb, float range_scale,
                       const apbf_search_debug* dbg, bool fuse_kw, uint32_t* out_kw_fixed, bool write_public);

#ifdef __CUDACC__
// kernel_width.comp:49-52: what a particle with original kernel width `orig` spreads onto a neighbour at distance `dist`
__device__ __forceinline__ uint32_t apbf_kw_influence(float orig, float dist)
{
	const float distanceFromKernel = dist - orig;
	const float influence = glsl_max(0.0f, 1.0f - glsl_max(0.0f, distanceFromKernel / (orig * APBF_KERNEL_WIDTH_PROPAGATION_FACTOR)));
	return f2u(orig * influence * APBF_KERNEL_WIDTH_RESOLUTION);
}
#endif

// uint_to_float_but_gradual over the fixed-point widths in SLOT_KWFX (solver.cu; tail of spread_kernel_width::apply)
int apbf_kw_finish(apbf_ctx* ctx, apbf_fluid* fluid, uint32_t* out_kw_fixed);

// one gather of up to four 16-byte and eight 4-byte arrays through the same permutation: dst[i] = src[perm[i]], i < *len
struct apbf_reorder_table {
	int          n16, n4;
	const int4*  src16[4];
	int4*        dst16[4];
	const uint32_t* src4[8];
	uint32_t*    dst4[8];
};
int apbf_launch_reorder(apbf_ctx* ctx, const apbf_reorder_table& t, const uint32_t* perm, const uint32_t* len, uint32_t cap);
