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

static inline bool apbf_nbr_struct_valid(const apbf_ctx* ctx, const apbf_neighbors* nb)
{
	return nb && nb->pairs && ctx->nbr_struct_pairs == nb->pairs;
}

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
