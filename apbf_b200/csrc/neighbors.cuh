// neighbors.cuh -- the grouped neighbour-list structure shared between the searches (neighbors.cu) and the sweeps
// over it (solver.cu).
#pragma once
#include "common.cuh"

// The searches emit pairs grouped by id: pairs [offsets[id], offsets[id+1]) all have .x == id.  One bit per pair says
// whether the mirrored pair (idN, id) is in the list as well; the sweeps gather through mirrored pairs and fall back to
// integer atomics only for the unmirrored ones (variable kernel widths).  The structure lives in the context's scratch
// memory and belongs to the pair buffer it was built for.
struct apbf_nbr_struct {
	const uint32_t* offsets; // [n + 1], unclamped running pair counts
	uint32_t*       symbits; // [ceil(capacity / 32)]
};

static inline bool apbf_nbr_struct_valid(const apbf_ctx* ctx, const apbf_neighbors* nb)
{
	return nb && nb->pairs && ctx->nbr_struct_pairs == nb->pairs;
}
