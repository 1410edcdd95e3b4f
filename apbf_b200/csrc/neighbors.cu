// neighbors.cu -- neighbour searches: uniform grid on the Z-curve hash (neighborhood_green) and 96-bit Morton code with
// per-particle power-of-two cells (neighborhood_binary_search).
//
// Replaces source/neighborhood_green.cpp:27-77 and source/neighborhood_binary_search.cpp:22-75 with their shaders
// (calculate_position_hash/_code, radix_sort_*, prefix_sum_*, copy_scattered_read, atomic_swap,
// generate_new_index_and_edit_list, find_value_ranges, neighborhood_green, neighborhood_binary_search,
// copy_with_differing_stride, linked_list_to_neighbor_list).
//
// Pipeline: key -> onesweep sort (key, slot) -> one fused gather of every hidden array -> index list / per-id arrays
// follow -> cell ranges -> count / scan / fill.  The pair list comes out grouped by id in the reference's discovery
// order, together with CSR offsets and one "mirrored" bit per pair that the solver sweeps use (solver.cu).
#include "keys.cuh"
#include "neighbors.cuh"
#include "sort.cuh"
#include <stdlib.h>
#include <algorithm>

namespace {

// ---- fused reorder of all hidden arrays -------------------------------------------------------------------------------
using reorder_table = apbf_reorder_table;

__global__ void __launch_bounds__(256) k_reorder(reorder_table t, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ len, uint32_t* clear_word)
{
	const uint32_t n = *len;
	if (clear_word && blockIdx.x == 0 && threadIdx.x == 0) *clear_word = 0u; // (MW_INDEX_NONIDENT: raised by k_mark_members, the next launch)
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t s = perm[i];
		int4 v16[4];
		uint32_t v4[8];
#pragma unroll
		for (int a = 0; a < 4; a++) if (a < t.n16) v16[a] = __ldg(t.src16[a] + s);
#pragma unroll
		for (int a = 0; a < 8; a++) if (a < t.n4) v4[a] = __ldg(t.src4[a] + s);
#pragma unroll
		for (int a = 0; a < 4; a++) if (a < t.n16) t.dst16[a][i] = v16[a];
#pragma unroll
		for (int a = 0; a < 8; a++) if (a < t.n4) t.dst4[a][i] = v4[a];
	}
}

} // namespace

int apbf_launch_reorder(apbf_ctx* ctx, const apbf_reorder_table& t, const uint32_t* perm, const uint32_t* len, uint32_t cap)
{
	k_reorder<<<apbf_grid(ctx, cap, 256), 256, 0, ctx->stream>>>(t, perm, len, nullptr);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

namespace {

// ---- index list after the hidden permutation (indexed_list.h:289-308 + :276-286 for a permutation edit) ------------------
// The usual case -- every hidden particle is a member and id i points at slot i (what a search leaves behind: the lists come out
// sorted) -- needs no marks, flags, scan or compaction: the new index list is the identity again and the ids move exactly like
// their slots.  k_mark_members finds out (MW_INDEX_NONIDENT, cleared by k_reorder before); the passes behind it return at once
// when the word is still 0, and k_compact_members writes the two trivial lists.
__global__ void k_mark_members(const uint32_t* __restrict__ index_list, const uint32_t* __restrict__ len, const uint32_t* __restrict__ hidden_len,
                               uint32_t* __restrict__ mark, uint32_t* misc)
{
	const uint32_t n = *len;
	bool other = blockIdx.x == 0 && threadIdx.x == 0 && n != *hidden_len;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t s = index_list[i];
		mark[s] = i + 1u;
		other = other || s != i;
	}
	if (other) misc[MW_INDEX_NONIDENT] = 1u;
}

__global__ void k_member_flags(const uint32_t* __restrict__ sorted_index, const uint32_t* __restrict__ mark,
                               const uint32_t* __restrict__ hidden_len, uint32_t* __restrict__ flags, const uint32_t* __restrict__ misc)
{
	if (misc[MW_INDEX_NONIDENT] == 0u) return;
	const uint32_t n = *hidden_len;
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x)
		flags[h] = mark[sorted_index[h]] != 0u ? 1u : 0u;
}

__global__ void k_compact_members(const uint32_t* __restrict__ sorted_index, const uint32_t* __restrict__ mark,
                                  const uint32_t* __restrict__ offs, const uint32_t* __restrict__ hidden_len,
                                  const uint32_t* __restrict__ index_len, uint32_t* __restrict__ new_index_list,
                                  uint32_t* __restrict__ id_perm, uint32_t index_cap, uint32_t* misc)
{
	const uint32_t n = *hidden_len;
	if (misc[MW_INDEX_NONIDENT] == 0u) { // identity in: identity out, ids follow their slots
		for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x) {
			if (h < index_cap) { new_index_list[h] = h; id_perm[h] = sorted_index[h]; }
			if (h == 0) misc[MW_IDENTITY] = 1u;
		}
		return;
	}
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x) {
		uint32_t m = mark[sorted_index[h]];
		if (m != 0u) {
			uint32_t o = offs[h];
			if (o < index_cap) {
				new_index_list[o] = h;
				id_perm[o] = m - 1u;
			}
		}
		if (h == 0) misc[MW_IDENTITY] = (offs[n] == n && *index_len == n) ? 1u : 0u;
	}
}

struct gather_table {
	int n4;
	const uint32_t* src4[8];
	uint32_t*       dst4[8];
};

__global__ void k_gather_ids(gather_table t, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ len)
{
	const uint32_t n = *len;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t s = perm[i];
#pragma unroll
		for (int a = 0; a < 8; a++) if (a < t.n4) t.dst4[a][i] = __ldg(t.src4[a] + s);
	}
}

// ---- Green pair emit (neighborhood_green.comp:50-87) ----------------------------------------------------------------
// distance(pos, posN) = sqrt(dot(d, d)), d = pos - posN, unfused, left to right (oracle convention)
__device__ __forceinline__ float dist_rn(float px, float py, float pz, float qx, float qy, float qz)
{
	float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
	float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
	return __fsqrt_rn(s);
}

// Q4[id] = {pos.x, pos.y, pos.z, T}: float position (vec3(ipos) / 2^18) and the acceptance threshold on the SQUARED
// distance.  sqrt is monotone, so "distance(pos, posN) > range" (neighborhood_green.comp:83) is the same predicate as
// "d2 > T" with T = the largest float whose correctly rounded square root is still <= range: no square root per
// candidate, bit-identical decisions.
__device__ __forceinline__ float sqrt_threshold(float r)
{
	if (r != r) return INFINITY;  // d > NaN is false: everything is accepted
	if (r < 0.0f) return -1.0f;   // d > r always: nothing is accepted
	float x = __fmul_rn(r, r);
	if (x > 3.0e38f) return 3.4028234e38f;
	for (int i = 0; i < 8 && __fsqrt_rn(x) > r; i++) x = __uint_as_float(__float_as_uint(x) - 1u);
	for (int i = 0; i < 8; i++) {
		const float y = __uint_as_float(__float_as_uint(x) + 1u);
		if (__fsqrt_rn(y) <= r) x = y; else break;
	}
	return x;
}

// per id: packed float position + threshold, and the cell key of the particle behind the id
// fused spread_kernel_width only (i4 != nullptr): integer position + original kernel width (kernel_width_init.comp:28-34)
// and the prune cutoff max(original, old kernel width) (kernel_width.comp:57)
struct build_kw_args {
	const float* radius;        // hidden
	const float* target_radius; // per id
	const float* kernel_width;  // per id, before the update
	int          base_on_target_radius;
	int4*        i4;
	float*       cutoff; // threshold on the squared integer distance equivalent to dist <= max(original, old kernel width)
	float        pmax;   // largest |coordinate| of the search grid
	uint32_t*    cell_maxw; // per cell: largest original width (fixed point) of its particles -- can this cell spread onto a query?
	float4*      qb4;    // {K, U, -, original width}: the prune decided on the float-form distance (see k_green_stream)
};

// Fused search + spread_kernel_width: the per-id record {K, U, -, original width} of the prune (kernel_width.comp:57).
// The prune compares the distance of the INTEGER difference (kernel_width.comp:36-38), the search the distance of the float
// positions (neighborhood_green.comp:83).  Both approximate the same squared distance s; with |p| the largest coordinate of
// the particle and c the cutoff, |d2_float - d2_int| <= 2^-22 (1.73 c (|p| + c) + 2 c^2) for s near c^2 (conversion of the
// coordinates to float: 2^-24 |p| each; subtraction and the two dot products: a few 2^-24 s).  hw is four times that: a pair
// with d2_float <= K = C - hw is kept for sure, one with d2_float > U = C + hw is dropped for sure, and only the band in
// between needs the integer form.  T: threshold of the range test; pmax: largest |coordinate| of the search grid (so that
// particles of equal cutoff share K and U).
__device__ __forceinline__ float4 prune_record(float T, float cut, float orig, int4 ip, float pmax)
{
	float4 qb = make_float4(T, T, 0.0f, orig); // no prune, never ambiguous (cutoff >= 1e15 or +inf)
	if (!(cut == cut) || cut < 0.0f) qb.x = qb.y = -1.0f; // keeps nothing
	else if (cut < 1.0e15f) {
		const float pm = fmaxf(fmaxf(fmaxf(fabsf((float)ip.x), fabsf((float)ip.y)), fabsf((float)ip.z)) * INV_R_POS, pmax);
		const float C = sqrt_threshold(cut);
		const float hw = fmaxf(2.3841858e-7f * cut * (7.0f * (pm + cut) + 8.0f * cut), 1.0e-20f);
		qb.x = glsl_min(T, C - hw); // kept for sure
		qb.y = glsl_min(T, C + hw); // kept at most
	}
	return qb;
}

__global__ void __launch_bounds__(256)
k_build_q4(const uint32_t* __restrict__ index_list, const int32_t* __restrict__ pos4, const uint32_t* __restrict__ hidden_key,
           const float* __restrict__ range, float range_scale, const uint32_t* __restrict__ len,
           float4* __restrict__ q4, uint32_t* __restrict__ key_id, uint32_t* __restrict__ misc, const build_kw_args K)
{
	// Device-wide scalars (occupied cells, largest initial width, smallest / largest thresholds) are reduced per thread, per warp
	// and per CTA before they touch their word -- and only if they would change it: an atomic on one address costs ~0.1 us,
	// one per warp (37 000 of them) used to be most of this kernel's time.
	__shared__ uint32_t s_red[8][10];
	const bool ident = misc[MW_IDENTITY] != 0u;
	const uint32_t n = *len;
	const unsigned lane = lane_id();
	uint32_t max_init = 0u; // fused: the largest initial width (kernel_width_init.comp:35) this thread has seen
	uint32_t occ = 0u;      // occupied cells = ids whose key differs from their predecessor's (the emit sizes its query blocks with it)
	uint32_t thr_lo[4] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, thr_hi[4] = { 0u, 0u, 0u, 0u }; // bits of {T, K, U, cutoff}
	for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) { // (warp-uniform trip count)
		const uint32_t id = base + threadIdx.x;
		const bool live = id < n;
		uint32_t key = 0xFFFFFFFFu, width_fx = 0u;
		if (live) {
			const uint32_t idx = ident ? id : index_list[id];
			const int4 ip = ldg_int4(pos4, idx);
			const float T = sqrt_threshold(range[id] * range_scale);
			q4[id] = make_float4((float)ip.x * INV_R_POS, (float)ip.y * INV_R_POS, (float)ip.z * INV_R_POS, T);
			thr_lo[0] = min(thr_lo[0], __float_as_uint(T)); thr_hi[0] = max(thr_hi[0], __float_as_uint(T));
			if (K.i4) {
				const float orig = glsl_max(K.radius[idx], K.base_on_target_radius ? K.target_radius[id] : 0.0f) * APBF_KERNEL_SCALE;
				K.i4[id] = make_int4(ip.x, ip.y, ip.z, __float_as_int(orig));
				// dist <= cutoff (kernel_width.comp:57) with dist = sqrt(d2) is d2 <= sqrt_threshold(cutoff); d2 = D2 * 2^-36 exactly,
				// D2 the same sum over the unscaled integer differences (powers of two commute with the roundings)
				const float cut = glsl_max(orig, K.kernel_width[id]);
				const float cutoff = cut == cut ? sqrt_threshold(cut) * 68719476736.0f : -1.0f; // dist <= NaN keeps nothing
				K.cutoff[id] = cutoff;
				const float4 qb = prune_record(T, cut, orig, ip, K.pmax);
				K.qb4[id] = qb;
				const uint32_t tb[3] = { __float_as_uint(qb.x), __float_as_uint(qb.y), __float_as_uint(cutoff) };
#pragma unroll
				for (int k = 0; k < 3; k++) { thr_lo[k + 1] = min(thr_lo[k + 1], tb[k]); thr_hi[k + 1] = max(thr_hi[k + 1], tb[k]); }
				width_fx = f2u(orig * APBF_KERNEL_WIDTH_RESOLUTION);
				max_init = max(max_init, width_fx);
			}
			key = hidden_key[idx];
			key_id[id] = key;
			if (id == 0u || hidden_key[ident ? id - 1u : index_list[id - 1u]] != key) occ++;
		}
		if (K.i4) {
			// largest initial width per cell (keys of ghosts index the second table): the lanes of a cell -- neighbours in the sorted
			// list -- agree on their maximum, one of them publishes it
			const uint32_t grp = __match_any_sync(0xffffffffu, key);
			const uint32_t mx = __reduce_max_sync(grp, width_fx);
			if (live && lane == (uint32_t)__ffs(grp) - 1u && mx > K.cell_maxw[key]) atomicMax(K.cell_maxw + key, mx);
		}
	}
	uint32_t red[10];
	red[0] = __reduce_add_sync(0xffffffffu, occ);
	red[1] = __reduce_max_sync(0xffffffffu, max_init);
#pragma unroll
	for (int k = 0; k < 4; k++) { red[2 + k] = __reduce_min_sync(0xffffffffu, thr_lo[k]); red[6 + k] = __reduce_max_sync(0xffffffffu, thr_hi[k]); }
	if (lane == 0u) {
#pragma unroll
		for (int k = 0; k < 10; k++) s_red[threadIdx.x >> 5][k] = red[k];
	}
	__syncthreads();
	if (threadIdx.x < 10u) {
		const uint32_t k = threadIdx.x;
		uint32_t v = s_red[0][k];
		for (uint32_t w = 1; w < blockDim.x / 32u; w++) {
			const uint32_t x = s_red[w][k];
			v = k == 0u ? v + x : (k >= 2u && k < 6u) ? min(v, x) : max(v, x);
		}
		// are the test thresholds the same for every id?  (then "the mirrored pair is kept" is the same test as "the pair is kept"
		// for every pair of the list: slabs skip their ghost queries, which exist only to find unmirrored pairs)
		volatile uint32_t* word = misc + (k == 0u ? MW_OCC_CELLS : k == 1u ? MW_MAX_INIT : k < 6u ? MW_THR_MIN + (k - 2u) : MW_THR_MAX + (k - 6u));
		if (k == 0u) { if (v) atomicAdd((uint32_t*)word, v); }
		else if (k >= 2u && k < 6u) { if (v < *word) atomicMin((uint32_t*)word, v); }
		else if (v > *word) atomicMax((uint32_t*)word, v);
	}
}

// The pair emit works on groups of neighbouring cells.  The lists were just sorted by cell key, so the particles of a
// cell -- and of an aligned 2x2x2 / 4x4x4 block of cells, which is a contiguous key range of the Z-curve -- are
// consecutive ids with nearly the same search box [gridMin, gridMax].  A warp owns 32 consecutive ids (a tile, handed
// out by a ticket) and splits them into runs of equal block key; the run's particles are the QUERIES (staged in shared
// memory).  The warp walks the union of their boxes in the reference's order (x fastest), skips the cells that are
// farther away than the largest range and flattens the occupants of 32 cells at a time into a dense CANDIDATE stream.
// Every batch of 32 candidates is handled in two phases:
//   phase 1, lane = candidate: one 16-byte load per lane, then every query is tested against the 32 candidates
//            (8 unfused flops + compare) and the ballot of the result goes to shared memory -- no branch, no counting;
//   phase 2, lane = query: each query takes its ballot, drops the invalid lanes and itself, and either counts the bits
//            or walks them in ascending order and appends the pairs behind its own offset.
// Ascending bit order inside a batch and batches in walk order keep the reference's discovery order (a candidate outside
// a query's own box cannot pass the distance test: the cell map is monotone in the position).  Work is proportional to
// the number of candidates, whatever the cell shape, occupancy or the Z-curve's jumps.  The block size adapts to the
// occupancy (k_build_q4 counts the occupied cells): sparse grids group 8 or 64 cells so that a run still has a few
// dozen queries to amortise the walk.
// FILL == false counts the accepted candidates per id, FILL == true writes them at the scanned offsets: the public
// (id, idN) pair list and the solver's internal list NB[e] = idN | (unmirrored << 31), where "mirrored" means that
// (idN, id) is in the list as well (d <= range[idN]).
// VARIANT: EMIT_PLAIN; EMIT_MG (multi-GPU slabs, ghost particles behind n_owned); EMIT_FUSED = neighborhood_green followed
// by spread_kernel_width in one go (pool.cpp:83-89): the width spread (kernel_width.comp:49-53) is GATHERED by the target
// -- a query a collects max(orig_b * influence) over every candidate b whose own range reaches a, which are exactly the
// pairs (b, a) of the unpruned list -- so there is no atomicMax, and only the pairs that survive the prune
// (kernel_width.comp:57) are ever counted and written.
constexpr int EMIT_WARPS = 8;
constexpr int EMIT_PLAIN = 0, EMIT_MG = 1, EMIT_FUSED = 2, EMIT_FUSED_MG = 3; // k_green_stream: bit 0 = slabs, bit 1 = fused prune

struct emit_args {
	const float4*   q4;
	const uint32_t* key_id;
	const float*    range;
	const uint32_t* cell_start;
	const uint32_t* cell_end;
	const uint32_t* len;
	apbf_grid_params g;
	float           range_scale;
	uint32_t*       counts;
	const uint32_t* offsets;
	uint32_t*       pairs;
	uint32_t*       nbl;
	uint32_t        cap;
	uint32_t*       misc;
	uint32_t*       ticket;
	int             cull;
	uint32_t        table_cells;
	uint32_t        layers;
	uint32_t        mg_lo[3], mg_hi[3]; // slabs: this rank's brick in cells (inclusive) -- a box inside it cannot hold a ghost
	// EMIT_FUSED
	const int4*     i4;      // {ipos.xyz, bits(original kernel width)} per id
	const float*    cutoff;  // prune threshold (kernel_width.comp:57) on the squared distance in integer units, per id
	uint32_t*       kwfx;
	// one-pass emit (k_green_stream + k_regroup)
	const float4*   qb4;
	const uint32_t* cell_maxw; // fused: per cell, the largest initial width (kernel_width_init.comp:35) of its particles
	uint32_t*       stream;
	uint32_t        stream_blocks;
	int             fallback; // two-pass fill kernel: run only if the hit stream overflowed
	// binary search (SEARCH == 1): sorted 96-bit codes per id, raw positions
	const uint32_t *c0, *c1, *c2;
	const int32_t*  pos4;
	const uint32_t* index_list;
};

template <bool FILL, int VARIANT, int DIMS>
__global__ void __launch_bounds__(EMIT_WARPS * 32)
k_green_emit(const emit_args A)
{
	constexpr bool FUSED = VARIANT == EMIT_FUSED, MG = VARIANT == EMIT_MG;
	constexpr bool NEED_M = FILL || FUSED || MG; // the reverse test d <= range[idN]
	// Multi-GPU slabs (layers >= 2): ids >= n_owned are ghosts.  Their keys carry one extra bit, so they form a second cell
	// table behind the first (offset table_cells) and every cell is walked in both.  A ghost is a query too, but only for
	// the pairs nobody else provides: unmirrored pairs onto owned particles (the scatter part of the sweeps).
	__shared__ float4 s_q[EMIT_WARPS][32];
	__shared__ uint2 s_fm[EMIT_WARPS][32]; // per query: ballots of d <= range[id] (.x) and d <= range[idN] (.y)
	__shared__ uint32_t s_cand[EMIT_WARPS][32], s_self[EMIT_WARPS][32];
	__shared__ int4 s_ci[FUSED ? EMIT_WARPS : 1][32];
	__shared__ float s_ccut[FUSED ? EMIT_WARPS : 1][32];
	if (A.fallback && A.misc[MW_STREAM_OVERFLOW] == 0u) return;
	const apbf_grid_params& g = A.g;
	const uint32_t n = *A.len;
	const uint32_t n_owned = MG ? A.misc[MW_N_OWNED] : 0xFFFFFFFFu;
	const uint32_t layers = MG ? A.layers : 1u;
	const uint32_t key_mask = A.table_cells - 1u; // table_cells is a power of two
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	const uint32_t axis_cap = 1u << g.res; // a box wider than the grid only revisits aliased cells
	// query block: 1 cell, 2^D cells or 4^D cells, from the mean occupancy of the occupied cells
	const uint32_t occupied = max(A.misc[MW_OCC_CELLS], 1u);
	const uint32_t gshift = (n > 4u * occupied) ? 0u : ((2u * n > occupied) ? (uint32_t)DIMS : 2u * (uint32_t)DIMS);
	float csz[3];
#pragma unroll
	for (int d = 0; d < 3; d++) csz[d] = g.ext[d] / g.scale;
	if (DIMS < 3) csz[2] = 0.0f; // z is not gridded in 2-D but the distance stays 3-D
	uint32_t n_asym = 0, n_searched = 0;
	for (;;) {
		uint32_t tile = 0;
		if (lane == 0) tile = atomicAdd(A.ticket, 1u);
		tile = __shfl_sync(0xffffffffu, tile, 0);
		if ((size_t)tile * 32 >= n) break;
		const uint32_t tile_first = tile * 32u;
		const uint32_t id = tile_first + lane;
		const bool in = id < n;
		float4 me = make_float4(0.f, 0.f, 0.f, -1.0f);
		uint32_t gmin[3] = { 0u, 0u, 0u }, gmax[3] = { 0u, 0u, 0u }, qc[3] = { 0u, 0u, 0u };
		uint32_t gkey = 0xFFFFFFFFu;
		float r_lane = 0.0f;
		int4 ip = make_int4(0, 0, 0, 0);
		float cutoff_a = 0.0f;
		if (in) {
			me = A.q4[id];
			gkey = A.key_id[id] >> gshift; // includes the ghost bit: owned and ghost particles never share a run
			const float r = A.range[id] * A.range_scale;
			r_lane = r == r ? fmaxf(r, 0.0f) : INFINITY; // a NaN range accepts every candidate
			qc[0] = apbf_map_axis(me.x, g, 0); qc[1] = apbf_map_axis(me.y, g, 1); qc[2] = apbf_map_axis(me.z, g, 2);
			gmin[0] = apbf_map_axis(me.x - r, g, 0); gmax[0] = apbf_map_axis(me.x + r, g, 0);
			gmin[1] = apbf_map_axis(me.y - r, g, 1); gmax[1] = apbf_map_axis(me.y + r, g, 1);
			gmin[2] = apbf_map_axis(me.z - r, g, 2); gmax[2] = apbf_map_axis(me.z + r, g, 2);
			if (DIMS < 3) { gmin[2] = 0u; gmax[2] = 0u; } // * uvec3(1, D > 1, D > 2), neighborhood_green.comp:36
			// the reference visits gridMin once even when gridMax < gridMin (its ++cell > gridMax wrap)
			gmax[0] = max(gmax[0], gmin[0]); gmax[1] = max(gmax[1], gmin[1]); gmax[2] = max(gmax[2], gmin[2]);
			if (FUSED) { ip = __ldg(A.i4 + id); cutoff_a = A.cutoff[id]; }
		}
		__syncwarp();
		s_q[w][lane] = me;
		__syncwarp();
		const uint32_t gprev = __shfl_up_sync(0xffffffffu, gkey, 1);
		uint32_t heads = __ballot_sync(0xffffffffu, in && (lane == 0u || gkey != gprev)); // runs of equal block key
		const uint32_t n_in = min(32u, n - tile_first);
		uint32_t my_off = (FILL && in) ? A.offsets[id] : 0u, my_cnt = 0u;
		// kernel_width_init.comp:35; everything a neighbour spreads is at most ITS initial value, so a candidate is only
		// looked at when that exceeds what the query already has
		uint32_t my_mx = FUSED ? f2u(__int_as_float(ip.w) * APBF_KERNEL_WIDTH_RESOLUTION) : 0u;
		while (heads) {
			const uint32_t r0 = (uint32_t)__ffs(heads) - 1u;
			heads &= heads - 1u;
			const uint32_t r1 = heads ? (uint32_t)__ffs(heads) - 1u : n_in;
			const bool valid = lane >= r0 && lane < r1;
			uint32_t umin[3], ext[3];
			int qlo[3], qhi[3];
#pragma unroll
			for (int d = 0; d < 3; d++) {
				umin[d] = __reduce_min_sync(0xffffffffu, valid ? gmin[d] : 0xFFFFFFFFu);
				ext[d] = min(__reduce_max_sync(0xffffffffu, valid ? gmax[d] : 0u) - umin[d], axis_cap - 1u) + 1u;
				qlo[d] = (int)min(__reduce_min_sync(0xffffffffu, valid ? qc[d] : 0xFFFFFFFFu), 0x7FFFFFFFu);
				qhi[d] = (int)min(__reduce_max_sync(0xffffffffu, valid ? qc[d] : 0u), 0x7FFFFFFFu);
			}
			// Cells farther from the queries' cells than the largest range cannot hold a hit: the gap between the box
			// [qlo, qhi] of the queries' cells and a cell is (distance in cells - 1) cell widths per axis; 1.01 instead of 1
			// covers the rounding of the cell map (< 2e-4 cells).
			const float r_cull = __uint_as_float(__reduce_max_sync(0xffffffffu, valid ? __float_as_uint(r_lane) : 0u));
			const float cull2 = A.cull ? r_cull * r_cull * 1.0001f : INFINITY;
			const uint32_t nxy = ext[0] * ext[1], ncell = nxy * ext[2];
			const float inv_nxy = 1.0f / (float)nxy, inv_nx = 1.0f / (float)ext[0];
			const bool ghost_run = MG && tile_first + r0 >= n_owned;
			const uint32_t n_layers = layers > 1u ? 2u : 1u; // layers == 3: two tables + ghosts keep their mirrored pairs too
			for (uint32_t cbase = 0; cbase < ncell * n_layers; cbase += 32) {
				uint32_t ci = cbase + lane;
				uint32_t c_first = 0u, c_cnt = 0u;
				if (ci < ncell * n_layers) {
					const uint32_t table_off = ci >= ncell ? A.table_cells : 0u;
					if (ci >= ncell) ci -= ncell;
					// ci = (cz * ny + cy) * nx + cx; float reciprocals + one correction step are exact for ci < 2^24
					uint32_t cz, cy;
					if (ncell <= (1u << 24)) {
						cz = (uint32_t)((float)ci * inv_nxy);
						if (cz * nxy > ci) cz--; else if ((cz + 1u) * nxy <= ci) cz++;
					} else {
						cz = ci / nxy;
					}
					const uint32_t rem = ci - cz * nxy;
					if (ncell <= (1u << 24)) {
						cy = (uint32_t)((float)rem * inv_nx);
						if (cy * ext[0] > rem) cy--; else if ((cy + 1u) * ext[0] <= rem) cy++;
					} else {
						cy = rem / ext[0];
					}
					const uint32_t cx = rem - cy * ext[0];
					const uint32_t ax = umin[0] + cx, ay = umin[1] + cy, az = umin[2] + cz;
					const int ix = (int)min(ax, 0x7FFFFFFFu), iy = (int)min(ay, 0x7FFFFFFFu), iz = (int)min(az, 0x7FFFFFFFu);
					const float gx = fmaxf((float)max(qlo[0] - ix, ix - qhi[0]) - 1.01f, 0.0f) * csz[0];
					const float gy = fmaxf((float)max(qlo[1] - iy, iy - qhi[1]) - 1.01f, 0.0f) * csz[1];
					const float gz = fmaxf((float)max(qlo[2] - iz, iz - qhi[2]) - 1.01f, 0.0f) * csz[2];
					if (!(gx * gx + gy * gy + gz * gz > cull2)) {
						const uint32_t h = (apbf_zhash<DIMS>(ax, ay, az, g.res) & key_mask) + table_off;
						c_first = __ldg(A.cell_start + h);
						c_cnt = __ldg(A.cell_end + h) - c_first;
					}
				}
				uint32_t incl = c_cnt;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
					if (lane >= (unsigned)o) incl += t;
				}
				const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
				const uint32_t c_base = c_first - (incl - c_cnt); // candidate t of this cell is id c_base + t
				for (uint32_t t0 = 0; t0 < total; t0 += 32) {
					const uint32_t t = t0 + lane;
					// the cell that holds candidate t: first lane whose inclusive count exceeds t
					uint32_t pos = 0u;
#pragma unroll
					for (int step = 16; step > 0; step >>= 1) {
						const uint32_t v = __shfl_sync(0xffffffffu, incl, (int)(pos + step - 1));
						if (v <= t) pos += step;
					}
					const uint32_t cand = __shfl_sync(0xffffffffu, c_base, (int)(pos & 31u)) + t;
					const bool cvalid = t < total;
					const uint32_t valid_mask = total - t0 >= 32u ? 0xFFFFFFFFu : (1u << (total - t0)) - 1u;
					float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
					if (cvalid) c4 = A.q4[cand];
					// ---- phase 1: lane = candidate ----------------------------------------------------------------------
					s_cand[w][lane] = cand;
					s_self[w][lane] = 0u;
					bool spread = false; // fused count pass: can a candidate of this batch raise the width of a query of this run?
					if (FUSED) {
						const int4 ci4 = cvalid ? __ldg(A.i4 + cand) : make_int4(0, 0, 0, 0);
						s_ci[w][lane] = ci4;
						if (FILL) s_ccut[w][lane] = cvalid ? A.cutoff[cand] : 0.0f;
						else {
							const uint32_t c_init = cvalid ? f2u(__int_as_float(ci4.w) * APBF_KERNEL_WIDTH_RESOLUTION) : 0u;
							spread = __reduce_max_sync(0xffffffffu, c_init) > __reduce_min_sync(0xffffffffu, valid ? my_mx : 0xFFFFFFFFu);
						}
					}
					__syncwarp();
					const uint32_t sq = cand - tile_first; // this candidate is query sq of the tile: id != idN, :83
					if (cvalid && sq < 32u) s_self[w][sq] = 1u << lane;
					if (NEED_M && (FILL || !FUSED || spread)) {
#pragma unroll 4
						for (uint32_t qi = r0; qi < r1; qi++) {
							const float4 qv = s_q[w][qi];
							const float dx = __fsub_rn(qv.x, c4.x), dy = __fsub_rn(qv.y, c4.y), dz = __fsub_rn(qv.z, c4.z);
							const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
							const uint32_t fb = __ballot_sync(0xffffffffu, !(d2 > qv.w));
							const uint32_t mb = __ballot_sync(0xffffffffu, !(d2 > c4.w));
							if (lane == 0u) s_fm[w][qi] = make_uint2(fb, mb);
						}
					} else {
#pragma unroll 4
						for (uint32_t qi = r0; qi < r1; qi++) {
							const float4 qv = s_q[w][qi];
							const float dx = __fsub_rn(qv.x, c4.x), dy = __fsub_rn(qv.y, c4.y), dz = __fsub_rn(qv.z, c4.z);
							const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
							const uint32_t fb = __ballot_sync(0xffffffffu, !(d2 > qv.w));
							if (lane == 0u) s_fm[w][qi] = make_uint2(fb, 0u);
						}
					}
					const uint32_t owned_mask = MG ? __ballot_sync(0xffffffffu, cand < n_owned) : 0xFFFFFFFFu;
					__syncwarp();
					// ---- phase 2: lane = query --------------------------------------------------------------------------
					if (valid) {
						const uint32_t live = valid_mask & ~s_self[w][lane];
						const uint2 fm = s_fm[w][lane];
						uint32_t fb = fm.x & live;
						const uint32_t mb = fm.y & live;
						if (MG && ghost_run) {
							fb &= owned_mask;
							if (layers != 3u) fb &= ~mb;
						}
						if (!FUSED) {
							if (!FILL) {
								my_cnt += __popc(fb);
							} else {
								while (fb) {
									const uint32_t j = (uint32_t)__ffs(fb) - 1u;
									fb &= fb - 1u;
									if (my_off < A.cap) {
										const uint32_t c = s_cand[w][j];
										const bool mirrored = (mb >> j) & 1u;
										*(uint2*)(A.pairs + 2 * (size_t)my_off) = make_uint2(id, c);
										A.nbl[my_off] = c | (mirrored ? 0u : NB_UNMIRRORED);
										n_asym += mirrored ? 0u : 1u;
									}
									my_off++;
								}
							}
						} else {
							if (!FILL) n_searched += __popc(fb);
							uint32_t bits = FILL ? fb : (fb | mb);
							while (bits) {
								const uint32_t j = (uint32_t)__ffs(bits) - 1u;
								bits &= bits - 1u;
								const int4 cq = s_ci[w][j];
								// kernel_width.comp:36-38: integer subtract first, then to float; D2 = dist^2 * 2^36
								const float ux = (float)(cq.x - ip.x), uy = (float)(cq.y - ip.y), uz = (float)(cq.z - ip.z);
								const float D2 = dot3(ux, uy, uz, ux, uy, uz);
								const bool keep = ((fb >> j) & 1u) && D2 <= cutoff_a; // :57
								if (!FILL) {
									my_cnt += keep ? 1u : 0u;
									// the pair (idN, id) of the unpruned list spreads idN's width onto id (:49-53)
									if (((mb >> j) & 1u) && f2u(__int_as_float(cq.w) * APBF_KERNEL_WIDTH_RESOLUTION) > my_mx) {
										const float rx = ux * INV_R_POS, ry = uy * INV_R_POS, rz = uz * INV_R_POS;
										my_mx = max(my_mx, apbf_kw_influence(__int_as_float(cq.w), sqrtf(dot3(rx, ry, rz, rx, ry, rz))));
									}
								} else if (keep) {
									if (my_off < A.cap) {
										const uint32_t c = s_cand[w][j];
										const bool mirrored = ((mb >> j) & 1u) && D2 <= s_ccut[w][j]; // (idN, id) survives the prune as well
										*(uint2*)(A.pairs + 2 * (size_t)my_off) = make_uint2(id, c);
										A.nbl[my_off] = c | (mirrored ? 0u : NB_UNMIRRORED);
										n_asym += mirrored ? 0u : 1u;
									}
									my_off++;
								}
							}
						}
					}
					__syncwarp();
				}
			}
		}
		if (!FILL && in) {
			A.counts[id] = my_cnt;
			if (FUSED) A.kwfx[id] = my_mx; // kernel_width_init.comp:35 + the atomicMax of kernel_width.comp:53, gathered
		}
	}
	if (FILL) {
		n_asym = __reduce_add_sync(0xffffffffu, n_asym);
		if (lane == 0 && n_asym) atomicAdd(A.misc + MW_N_ASYM, n_asym);
	} else if (FUSED) {
		n_searched = __reduce_add_sync(0xffffffffu, n_searched);
		if (lane == 0 && n_searched) atomicAdd(A.misc + MW_TOTAL_PAIRS, n_searched);
	}
}

// ---- 96-bit Morton code helpers of the binary search (neighborhood_binary_search.comp:166-276) ---------------------------
struct u96 { uint32_t v[3]; };
__device__ __forceinline__ u96 mk96(uint32_t a, uint32_t b, uint32_t c) { u96 r; r.v[0] = a; r.v[1] = b; r.v[2] = c; return r; }
__device__ __forceinline__ u96 plus96(u96 a, u96 b)
{
	u96 r = mk96(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]);
	bool y = r.v[0] < a.v[0];
	bool z = r.v[1] < a.v[1] || (y && r.v[1] == 0xFFFFFFFFu);
	r.v[1] += y ? 1u : 0u; r.v[2] += z ? 1u : 0u;
	return r;
}
__device__ __forceinline__ u96 minus96(u96 a, u96 b)
{
	bool y = a.v[0] < b.v[0];
	bool z = a.v[1] < b.v[1] || (y && a.v[1] == b.v[1]);
	return mk96(a.v[0] - b.v[0], a.v[1] - b.v[1] - (y ? 1u : 0u), a.v[2] - b.v[2] - (z ? 1u : 0u));
}
__device__ __forceinline__ u96 shl96_small(u96 a, uint32_t s) // s in {1, 2}
{
	return mk96(a.v[0] << s, (a.v[1] << s) | (a.v[0] >> (32u - s)), (a.v[2] << s) | (a.v[1] >> (32u - s)));
}
__device__ __forceinline__ bool greater96(u96 a, u96 b)
{
	if (a.v[2] != b.v[2]) return a.v[2] > b.v[2];
	if (a.v[1] != b.v[1]) return a.v[1] > b.v[1];
	return a.v[0] > b.v[0];
}
__device__ __forceinline__ u96 and96(u96 a, u96 b) { return mk96(a.v[0] & b.v[0], a.v[1] & b.v[1], a.v[2] & b.v[2]); }
__device__ __forceinline__ u96 or96(u96 a, u96 b) { return mk96(a.v[0] | b.v[0], a.v[1] | b.v[1], a.v[2] | b.v[2]); }
__device__ __forceinline__ u96 not96(u96 a) { return mk96(~a.v[0], ~a.v[1], ~a.v[2]); }

__device__ __forceinline__ uint32_t lower_bound96(const uint32_t* __restrict__ c0, const uint32_t* __restrict__ c1,
                                                  const uint32_t* __restrict__ c2, uint32_t n, u96 code, uint32_t from = 0u)
{ // over the ids [from, n)
	uint32_t lo = from, hi = n;
	while (lo < hi) {
		uint32_t mid = lo + ((hi - lo) >> 1);
		if (greater96(code, mk96(__ldg(c0 + mid), __ldg(c1 + mid), __ldg(c2 + mid)))) lo = mid + 1u; else hi = mid;
	}
	return lo;
}

__device__ __forceinline__ uint32_t upper_bound96(const uint32_t* __restrict__ c0, const uint32_t* __restrict__ c1,
                                                  const uint32_t* __restrict__ c2, uint32_t n, u96 code, uint32_t from = 0u)
{ // first index in [from, n) whose code is greater than `code`
	uint32_t lo = from, hi = n;
	while (lo < hi) {
		uint32_t mid = lo + ((hi - lo) >> 1);
		if (greater96(mk96(__ldg(c0 + mid), __ldg(c1 + mid), __ldg(c2 + mid)), code)) hi = mid; else lo = mid + 1u;
	}
	return lo;
}
// the cell of a particle in the binary search: level = 3 * ceil(log2(range * 2^18)) low bits masked off its Morton code (:183-190)
struct bs_cell { u96 center, mask; };
__device__ __forceinline__ bs_cell bs_cell_of(int4 ip, float r)
{
	u96 code;
	apbf_encode96(ip.x, ip.y, ip.z, code.v);
	const uint32_t digits = f2u(ceilf(log2f(r * R_POS))) * 3u;
	bs_cell c;
	c.mask.v[0] = (digits < 32u ? 1u << digits : 0u) - 1u;
	c.mask.v[1] = (digits < 64u ? 1u << (max(digits, 32u) - 32u) : 0u) - 1u;
	c.mask.v[2] = (digits < 96u ? 1u << (max(digits, 64u) - 64u) : 0u) - 1u;
	c.center = mk96(code.v[0] - (code.v[0] & c.mask.v[0]), code.v[1] - (code.v[1] & c.mask.v[1]), code.v[2] - (code.v[2] & c.mask.v[2]));
	return c;
}
// cell (cx, cy, cz) in {0, 1, 2}^3 around the centre: dilated-integer +-1 per axis (:196-220)
__device__ __forceinline__ u96 bs_neighbor_cell(const bs_cell& c, int cx, int cy, int cz)
{
	const u96 xMask3 = mk96(011111111111u, 022222222222u, 04444444444u);
	const u96 yMask3 = mk96(022222222222u, 04444444444u, 011111111111u);
	const u96 zMask3 = mk96(04444444444u, 011111111111u, 022222222222u);
	const u96 step = plus96(c.mask, mk96(1u, 0u, 0u));
	u96 x = and96(c.center, xMask3), y = and96(c.center, yMask3), z = and96(c.center, zMask3);
	if (cx == 0) x = and96(minus96(and96(c.center, xMask3), step), xMask3);
	if (cx == 2) x = and96(plus96(or96(c.center, not96(xMask3)), step), xMask3);
	if (cy == 0) y = and96(minus96(and96(c.center, yMask3), shl96_small(step, 1u)), yMask3);
	if (cy == 2) y = and96(plus96(or96(c.center, not96(yMask3)), shl96_small(step, 1u)), yMask3);
	if (cz == 0) z = and96(minus96(and96(c.center, zMask3), shl96_small(step, 2u)), zMask3);
	if (cz == 2) z = and96(plus96(or96(c.center, not96(zMask3)), shl96_small(step, 2u)), zMask3);
	return or96(or96(x, y), z);
}

// largest |coordinate| over the list (float bits; non-negative floats order like unsigned integers)
__global__ void k_max_abs_coord(const uint32_t* __restrict__ index_list, const int32_t* __restrict__ pos4, const uint32_t* __restrict__ len,
                                uint32_t* __restrict__ misc)
{
	const bool ident = misc[MW_IDENTITY] != 0u;
	const uint32_t n = *len;
	float m = 0.0f;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const int4 ip = ldg_int4(pos4, ident ? id : index_list[id]);
		m = fmaxf(m, fmaxf(fmaxf(fabsf((float)ip.x), fabsf((float)ip.y)), fabsf((float)ip.z)) * INV_R_POS);
	}
	const uint32_t am = __activemask();
	const uint32_t mx = __reduce_max_sync(am, __float_as_uint(m));
	if ((am & ((1u << lane_id()) - 1u)) == 0u) atomicMax(misc + MW_PMAX, mx);
}

// slabs: 1 for the slots behind the owned particles (the ghosts), 0 otherwise -- the key of one more stable pass after the
// three code sections, so that owned particles keep the dense ids [0, n_owned) and the ghosts follow, both in code order
__global__ void k_ghost_key(const uint32_t* __restrict__ slots, const uint32_t* __restrict__ len, const uint32_t* __restrict__ misc, uint32_t* __restrict__ key)
{
	const uint32_t n = *len, n_owned = misc[MW_N_OWNED];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) key[i] = slots[i] >= n_owned ? 1u : 0u;
}

// binary search, per id: packed float position + threshold of "d <= range" (a NaN range rejects everything here), and a flag
// where the particle's cell (level and centre) differs from its predecessor's: the stream emit takes runs of equal cells as
// its chunks of queries
__global__ void k_build_bq(const uint32_t* __restrict__ index_list, const int32_t* __restrict__ pos4, const float* __restrict__ range,
                           float range_scale, const uint32_t* __restrict__ len, float4* __restrict__ q4, uint32_t* __restrict__ head,
                           const uint32_t* __restrict__ misc, const build_kw_args K)
{
	const bool ident = misc[MW_IDENTITY] != 0u;
	const uint32_t n = *len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const uint32_t idx = ident ? id : index_list[id];
		const int4 ip = ldg_int4(pos4, idx);
		const float r = range[id] * range_scale;
		const float T = r == r ? sqrt_threshold(r) : -1.0f;
		q4[id] = make_float4((float)ip.x * INV_R_POS, (float)ip.y * INV_R_POS, (float)ip.z * INV_R_POS, T);
		if (K.i4) { // fused with spread_kernel_width (see k_build_q4)
			const float orig = glsl_max(K.radius[idx], K.base_on_target_radius ? K.target_radius[id] : 0.0f) * APBF_KERNEL_SCALE;
			K.i4[id] = make_int4(ip.x, ip.y, ip.z, __float_as_int(orig));
			const float cut = glsl_max(orig, K.kernel_width[id]);
			K.cutoff[id] = cut == cut ? sqrt_threshold(cut) * 68719476736.0f : -1.0f;
			// |p| of the band: the list's largest coordinate rounded up to a power of two (the same for everybody, and stable
			// from substep to substep, so that particles of equal cutoff share K and U)
			const float pm = exp2f(ceilf(log2f(fmaxf(__uint_as_float(misc[MW_PMAX]), 1.0f))));
			K.qb4[id] = prune_record(T, cut, orig, ip, pm);
		}
		uint32_t h = 1u;
		if (id > 0u && id != misc[MW_N_OWNED]) { // (slabs: the ghosts sit behind the owned particles, in code order of their own; no run spans both)
			const bs_cell a = bs_cell_of(ip, r);
			const bs_cell b = bs_cell_of(ldg_int4(pos4, ident ? id - 1u : index_list[id - 1u]), range[id - 1u] * range_scale);
			h = (a.center.v[0] != b.center.v[0] || a.center.v[1] != b.center.v[1] || a.center.v[2] != b.center.v[2] ||
			     a.mask.v[0] != b.mask.v[0] || a.mask.v[1] != b.mask.v[1] || a.mask.v[2] != b.mask.v[2]) ? 1u : 0u;
		}
		head[id] = h;
	}
}

// ---- one-pass pair emit: hit stream + regroup ---------------------------------------------------------------------------
// k_green_emit tests every (query, candidate) twice: once to count, once to fill.  k_green_stream tests once.
//   * QUERIES: a warp takes a window of 32 ids by ticket and handles every block of cells whose first particle lies in
//     the window -- the whole block, in chunks of up to 32 particles, also where it extends past the window.  A block is
//     an aligned group of 2^s cells (a contiguous key range of the Z-curve, an axis-aligned box), s chosen on the device
//     so that a block holds about 32 particles.  All queries of a chunk share one candidate walk.
//   * CANDIDATES: the cells of the union of the queries' boxes, x fastest as in the reference; a cell is skipped when the
//     gap between its bounds and the bounding box of the queries' POSITIONS exceeds the largest range (cell 0 of an axis
//     also holds what lies below the grid, hence its lower bound is -inf).  32 occupancies at a time become a dense
//     candidate stream (warp scan + shuffle search), one candidate per lane.
//   * TESTS, lane = candidate: one bit per query of the chunk (colK: pair kept, colM: the mirrored pair is kept too):
//     a broadcast shared-memory load, 8 unfused flops, two compares and two predicated ORs per query.
//   * the kept bits are transposed (32 x 32 bit matrix, five shuffle stages) so that lane = query again; every query
//     appends its hits {idN | unmirrored << 31, lane << 27 | rank} to the chunk's HIT STREAM: 128-entry blocks taken from
//     a global allocator, header = {first id of the chunk, entries}.  rank = how many pairs the query had before, so
//     counts[id] falls out of the same registers.
//   * after the scan of the counts, k_regroup moves every entry to offsets[id] + rank -- no order, no chains, one thread
//     per entry.  The list comes out grouped by id, within an id in discovery order, exactly like the two-pass emit,
//     for 8 + 8 bytes of extra traffic per pair instead of a second round of distance tests.
// Fused spread_kernel_width (EMIT_FUSED): the prune (kernel_width.comp:57) is decided on the float-form squared distance
// with the thresholds K = min(T, C - hw) ("kept for sure") and U = min(T, C + hw) ("kept at most") built by k_build_q4;
// a batch in which the two disagree for the query's or the candidate's prune is redone with the integer form.
// The width spread (kernel_width.comp:49-53) only matters where a candidate starts wider than a query of the chunk
// currently is; those batches take a second, exact loop.
// If the stream runs out of blocks -- which needs more pairs than the pair list can hold -- MW_STREAM_OVERFLOW is set,
// k_regroup does nothing and the fill pass of the two-pass emit writes the clamped list from the same counts.
constexpr uint32_t SB_SHIFT = 7u, SB_ENTRIES = 1u << SB_SHIFT, SB_HEADER = 8u, SB_WORDS = SB_HEADER + 2u * SB_ENTRIES; // 1056 bytes
constexpr uint32_t SB_RANK_BITS = 27u, SB_RANK_MASK = (1u << SB_RANK_BITS) - 1u;

__device__ __forceinline__ uint32_t fkey(float f) // order-preserving float -> uint
{
	const uint32_t u = __float_as_uint(f);
	return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }

// bit j of lane i  ->  bit i of lane j
__device__ __forceinline__ uint32_t transpose32(uint32_t x, unsigned lane)
{
#pragma unroll
	for (int s = 16; s > 0; s >>= 1) {
		const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
		const uint32_t t = __shfl_xor_sync(0xffffffffu, x, s);
		x = (lane & (unsigned)s) ? ((x & ~m) | ((t >> s) & m)) : ((x & m) | ((t << s) & ~m));
	}
	return x;
}

// The distance tests of one candidate (this lane) against the queries of a chunk: one bit per query.  The flops go to the
// fma pipe, compares and bit ORs to the alu pipe, each good for one warp instruction every other cycle: the loop is
// written so that a query costs 8 + (2 | 4 | 8) of them -- compare into a predicate, predicated OR (the compiler's own
// rendering of "mask |= p ? bit : 0" spends a SEL per mask on top).
//   MODE 0: colK                      (all thresholds of the batch equal: the mirrored test is the same test)
//   MODE 1: colK, colM                (plain search, per-particle ranges)
//   MODE 2: colK, colU                (fused prune, all thresholds equal)
//   MODE 3: colK, colM, colU, colMu   (fused prune, per-particle thresholds)
// "d2 > K" is false for a NaN distance, i.e. NaN is accepted, like !(distance > range) in neighborhood_green.comp:83.
// Two queries per step with packed f32x2 arithmetic (sm_100: FADD2 / FMUL2 issue at the rate of their scalar forms,
// tools/microbench/pipes.cu): the chunk's queries sit in shared memory as pairs -- sa[j] = {x0, x1, y0, y1}, sb[j] = {z0, z1, K0, K1},
// su[j] = {U0, U1} -- so that one 16-byte load yields two aligned register pairs.  Differences and squares are packed (6
// instructions for two queries instead of 12); the two sums stay SCALAR adds: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 even under -fmad=false (seen in the SASS), and a fused sum of squares is not the oracle's distance.  Same bits as the
// scalar loop: every operation is the same IEEE operation on the same operands.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void sq_diff2(unsigned long long q, float c, float& lo, float& hi) // (q - {c, c})^2 per half
{
	unsigned long long d;
	const unsigned long long cc = pack2(c, c);
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(q), "l"(cc));
	asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(d));
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d));
}

template <int MODE>
__device__ __forceinline__ void chunk_tests(const float4* __restrict__ sa, const float4* __restrict__ sb, const float2* __restrict__ su, uint32_t cnt, const float4 c4,
                                            float Kb, float Ub, uint32_t& colK, uint32_t& colM, uint32_t& colU, uint32_t& colMu)
{
	uint32_t bit = 1u;
	const uint32_t npairs = (cnt + 1u) >> 1; // (a slot behind the last query holds K = U = -1: it never passes a test)
#pragma unroll 2
	for (uint32_t j = 0; j < npairs; j++) {
		const float4 a = sa[j], b = sb[j];
		float x0, x1, y0, y1, z0, z1;
		sq_diff2(pack2(a.x, a.y), c4.x, x0, x1);
		sq_diff2(pack2(a.z, a.w), c4.y, y0, y1);
		sq_diff2(pack2(b.x, b.y), c4.z, z0, z1);
		const float d0 = __fadd_rn(__fadd_rn(x0, y0), z0), d1 = __fadd_rn(__fadd_rn(x1, y1), z1);
		const uint32_t bit1 = bit + bit;
		if (MODE == 0) {
			asm("{\n\t.reg .pred p0, p1;\n\tsetp.gt.f32 p0, %1, %2;\n\tsetp.gt.f32 p1, %3, %4;\n\t@!p0 or.b32 %0, %0, %5;\n\t@!p1 or.b32 %0, %0, %6;\n\t}"
			    : "+r"(colK) : "f"(d0), "f"(b.z), "f"(d1), "f"(b.w), "r"(bit), "r"(bit1));
		} else if (MODE == 1) {
			asm("{\n\t.reg .pred pk, pm, qk, qm;\n\tsetp.gt.f32 pk, %2, %3;\n\tsetp.gt.or.f32 pm, %2, %6, pk;\n\tsetp.gt.f32 qk, %4, %5;\n\tsetp.gt.or.f32 qm, %4, %6, qk;\n\t"
			    "@!pk or.b32 %0, %0, %7;\n\t@!pm or.b32 %1, %1, %7;\n\t@!qk or.b32 %0, %0, %8;\n\t@!qm or.b32 %1, %1, %8;\n\t}"
			    : "+r"(colK), "+r"(colM) : "f"(d0), "f"(b.z), "f"(d1), "f"(b.w), "f"(Kb), "r"(bit), "r"(bit1));
		} else if (MODE == 2) {
			const float2 u = su[j];
			asm("{\n\t.reg .pred pk, pu, qk, qu;\n\tsetp.gt.f32 pk, %2, %3;\n\tsetp.gt.f32 pu, %2, %4;\n\tsetp.gt.f32 qk, %5, %6;\n\tsetp.gt.f32 qu, %5, %7;\n\t"
			    "@!pk or.b32 %0, %0, %8;\n\t@!pu or.b32 %1, %1, %8;\n\t@!qk or.b32 %0, %0, %9;\n\t@!qu or.b32 %1, %1, %9;\n\t}"
			    : "+r"(colK), "+r"(colU) : "f"(d0), "f"(b.z), "f"(u.x), "f"(d1), "f"(b.w), "f"(u.y), "r"(bit), "r"(bit1));
		} else {
			const float2 u = su[j];
			asm("{\n\t.reg .pred pk, pm, pu, pn;\n\tsetp.gt.f32 pk, %4, %5;\n\tsetp.gt.or.f32 pm, %4, %6, pk;\n\t"
			    "setp.gt.f32 pu, %4, %7;\n\tsetp.gt.or.f32 pn, %4, %8, pu;\n\t"
			    "@!pk or.b32 %0, %0, %9;\n\t@!pm or.b32 %1, %1, %9;\n\t@!pu or.b32 %2, %2, %9;\n\t@!pn or.b32 %3, %3, %9;\n\t}"
			    : "+r"(colK), "+r"(colM), "+r"(colU), "+r"(colMu) : "f"(d0), "f"(b.z), "f"(Kb), "f"(u.x), "f"(Ub), "r"(bit));
			asm("{\n\t.reg .pred pk, pm, pu, pn;\n\tsetp.gt.f32 pk, %4, %5;\n\tsetp.gt.or.f32 pm, %4, %6, pk;\n\t"
			    "setp.gt.f32 pu, %4, %7;\n\tsetp.gt.or.f32 pn, %4, %8, pu;\n\t"
			    "@!pk or.b32 %0, %0, %9;\n\t@!pm or.b32 %1, %1, %9;\n\t@!pu or.b32 %2, %2, %9;\n\t@!pn or.b32 %3, %3, %9;\n\t}"
			    : "+r"(colK), "+r"(colM), "+r"(colU), "+r"(colMu) : "f"(d1), "f"(b.w), "f"(Kb), "f"(u.y), "f"(Ub), "r"(bit1));
		}
		bit = bit1 + bit1;
	}
	if (MODE == 0) colM = colK;
	if (MODE == 2) { colM = colK; colMu = colU; }
}

// SEARCH: 0 = uniform grid (neighborhood_green), 1 = 96-bit Morton code with per-particle power-of-two cells
// (neighborhood_binary_search): key_id holds run numbers of equal cells, a chunk's candidates are the 27 key ranges around it
template <int VARIANT, int DIMS, bool STATS, int SEARCH = 0>
__global__ void __launch_bounds__(EMIT_WARPS * 32, 4)
k_green_stream(const emit_args A)
{
	constexpr bool FUSED = (VARIANT & EMIT_FUSED) != 0, MG = (VARIANT & EMIT_MG) != 0;
	__shared__ float4 s_q[EMIT_WARPS][32];                       // {x, y, z, K}: K = T (plain) or min(T, C - hw) (fused)
	__shared__ float s_u[FUSED ? EMIT_WARPS : 1][32];            // U = min(T, C + hw)
	__shared__ float s_T[FUSED ? EMIT_WARPS : 1][32];            // threshold of the range test alone
	// the same queries as pairs for the packed test loop (chunk_tests): {x0, x1, y0, y1}, {z0, z1, K0, K1}, {U0, U1}
	__shared__ __align__(16) float4 s_qa[EMIT_WARPS][16], s_qb[EMIT_WARPS][16];
	__shared__ __align__(8) float2 s_u2[FUSED ? EMIT_WARPS : 1][16];
	const apbf_grid_params& g = A.g;
	const uint32_t n = *A.len;
	const uint32_t n_owned = MG ? A.misc[MW_N_OWNED] : 0xFFFFFFFFu;
	const uint32_t layers = MG ? A.layers : 1u;
	const uint32_t key_mask = A.table_cells - 1u;
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	const uint32_t axis_cap = 1u << g.res;
	// block = 2^gshift cells with about 32 particles, from the mean occupancy of the occupied cells
	const uint32_t occupied = max(A.misc[MW_OCC_CELLS], 1u);
	const uint32_t per32 = (uint32_t)min((32ull * occupied) / max(n, 1u), 64ull);
	const uint32_t gshift = SEARCH == 1 ? 0u : min(per32 > 1u ? 31u - (uint32_t)__clz((int)per32) : 0u, g.res * (uint32_t)DIMS);
	float csz[3], eps[3];
#pragma unroll
	for (int d = 0; d < 3; d++) {
		csz[d] = g.ext[d] / g.scale;
		eps[d] = 1.0e-3f * csz[d] + 1.0e-5f * (fabsf(g.mn[d]) + fabsf(g.ext[d])); // rounding of the cell map, generously
	}
	// Slabs: a ghost is a query only for unmirrored pairs onto owned particles.  When every id has the same thresholds (k_build_q4)
	// there is no such pair -- (idN, id) passes exactly the tests (id, idN) passes -- and the ghosts' chunks are skipped.
	bool skip_ghost_queries = false;
	if (MG && !STATS && SEARCH == 0 && (FUSED || layers != 3u)) {
		skip_ghost_queries = true;
#pragma unroll
		for (int k = 0; k < 4; k++) skip_ghost_queries = skip_ghost_queries && A.misc[MW_THR_MIN + k] >= A.misc[MW_THR_MAX + k];
	}
	uint32_t n_searched = 0;
	for (;;) {
		uint32_t win = 0;
		if (lane == 0) win = atomicAdd(A.ticket, 1u);
		win = __shfl_sync(0xffffffffu, win, 0);
		if ((size_t)win * 32 >= n) break;
		const uint32_t win_first = win * 32u;
		uint32_t wkey = 0xFFFFFFFFu;
		bool head = false;
		if (win_first + lane < n) {
			wkey = A.key_id[win_first + lane] >> gshift; // includes the ghost bit: owned and ghost particles never share a block
			head = win_first + lane == 0u || (A.key_id[win_first + lane - 1u] >> gshift) != wkey;
		}
		uint32_t heads = __ballot_sync(0xffffffffu, head);
		while (heads) {
			const uint32_t hl = (uint32_t)__ffs(heads) - 1u;
			heads &= heads - 1u;
			const uint32_t blk_first = win_first + hl;
			const uint32_t bkey = __shfl_sync(0xffffffffu, wkey, (int)hl);
			uint32_t blk_len = 1u; // particles of the block: up to the next change of the block key
			for (uint32_t b = blk_first + 1u;; b += 32u) {
				const uint32_t i = b + lane;
				const bool same = i < n && (A.key_id[i] >> gshift) == bkey;
				const uint32_t diff = __ballot_sync(0xffffffffu, !same);
				if (diff) { blk_len += (uint32_t)__ffs(diff) - 1u; break; }
				blk_len += 32u;
			}
			// chunks of equal size: 40 particles are 20 + 20, not 32 + 8 (a chunk costs a candidate walk whatever its size)
			const uint32_t n_chunks = (blk_len + 31u) >> 5, chunk_len = (blk_len + n_chunks - 1u) / n_chunks;
			for (uint32_t c0 = 0; c0 < blk_len; c0 += chunk_len) {
				// ---- one chunk: queries first .. first + cnt, lane = query ----------------------------------------------------
				const uint32_t first = blk_first + c0, cnt = min(chunk_len, blk_len - c0);
				const uint32_t id = first + lane;
				const bool valid = lane < cnt;
				if (MG && skip_ghost_queries && first >= n_owned) {
					if (valid) {
						A.counts[id] = 0u;
						if (FUSED) A.kwfx[id] = f2u(A.qb4[id].w * APBF_KERNEL_WIDTH_RESOLUTION); // (the owner's new width replaces it)
					}
					continue;
				}
				float4 me = make_float4(0.f, 0.f, 0.f, -1.0f);
				float4 qb = make_float4(-1.0f, -1.0f, 0.0f, 0.0f);
				uint32_t gmin[3] = { 0u, 0u, 0u }, gmax[3] = { 0u, 0u, 0u };
				float r_lane = 0.0f, r = 0.0f;
				if (valid) {
					me = A.q4[id];
					r = A.range[id] * A.range_scale;
					r_lane = r == r ? fmaxf(r, 0.0f) : INFINITY; // a NaN range accepts every candidate
					if (FUSED) qb = A.qb4[id];
				}
				// Fused prune, nobody in the whole list starts wider than any query of this chunk is: nothing beyond the prune
				// cutoff matters (see keep2 below), so the boxes shrink from the search range to the cutoff sqrt(U).
				if (FUSED && !STATS && A.cull) {
					const uint32_t mx0 = __reduce_min_sync(0xffffffffu, valid ? f2u(qb.w * APBF_KERNEL_WIDTH_RESOLUTION) : 0xFFFFFFFFu);
					if (A.misc[MW_MAX_INIT] <= mx0 && r == r) r = fminf(r, sqrtf(fmaxf(qb.y, 0.0f)) * 1.0001f);
				}
				if (valid && SEARCH == 0) {
					gmin[0] = apbf_map_axis(me.x - r, g, 0); gmax[0] = apbf_map_axis(me.x + r, g, 0);
					gmin[1] = apbf_map_axis(me.y - r, g, 1); gmax[1] = apbf_map_axis(me.y + r, g, 1);
					gmin[2] = apbf_map_axis(me.z - r, g, 2); gmax[2] = apbf_map_axis(me.z + r, g, 2);
					if (DIMS < 3) { gmin[2] = 0u; gmax[2] = 0u; } // * uvec3(1, D > 1, D > 2), neighborhood_green.comp:36
					// the reference visits gridMin once even when gridMax < gridMin (its ++cell > gridMax wrap)
					gmax[0] = max(gmax[0], gmin[0]); gmax[1] = max(gmax[1], gmin[1]); gmax[2] = max(gmax[2], gmin[2]);
					if (FUSED) qb = A.qb4[id];
				}
				if (SEARCH == 0) {
					// A search box at least as wide as the whole grid on some axis: the reference walks aliased cells a second time
					// there and appends those pairs twice (neighborhood_green.comp:69-79 + :40-47); every cell is
					// visited once here.  The caller is told through sticky flag bit 2 (apbf_ctx_device_flags).  (The box of the
					// reference: with the search range itself, not the cutoff the fused pass may have shrunk it to.)
					bool wide = false;
					if (valid) {
						const float pq[3] = { me.x, me.y, me.z };
#pragma unroll
						for (int d = 0; d < DIMS; d++) {
							const uint32_t lo = apbf_map_axis(pq[d] - r_lane, g, d), hi = apbf_map_axis(pq[d] + r_lane, g, d);
							wide = wide || (hi >= lo && hi - lo >= axis_cap);
						}
					}
					if (__any_sync(0xffffffffu, wide) && lane == 0u) atomicOr(A.misc + MW_FLAGS, 4u);
				}
				__syncwarp();
				s_q[w][lane] = make_float4(me.x, me.y, me.z, FUSED ? qb.x : me.w);
				if (FUSED) { s_u[w][lane] = qb.y; s_T[w][lane] = me.w; }
				{
					float* pa = (float*)&s_qa[w][lane >> 1] + (lane & 1u);
					float* pb = (float*)&s_qb[w][lane >> 1] + (lane & 1u);
					pa[0] = me.x; pa[2] = me.y; pb[0] = me.z; pb[2] = FUSED ? qb.x : me.w;
					if (FUSED) ((float*)&s_u2[w][lane >> 1])[lane & 1u] = qb.y;
				}
				__syncwarp();
				uint32_t my_mx = FUSED ? f2u(qb.w * APBF_KERNEL_WIDTH_RESOLUTION) : 0u; // kernel_width_init.comp:35
				uint32_t my_cnt = 0u;
				// do all queries of the chunk share their thresholds?  (then, for a batch of candidates that share them too,
				// "the mirrored pair is kept" is the same test as "the pair is kept")
				const float qK = __shfl_sync(0xffffffffu, FUSED ? qb.x : me.w, 0), qU = __shfl_sync(0xffffffffu, qb.y, 0);
				const bool q_uniform = __all_sync(0xffffffffu, !valid || ((FUSED ? qb.x : me.w) == qK && (!FUSED || qb.y == qU)));
				uint32_t tile_pos = 0u, n_alloc = 0u, last_blk = 0xFFFFFFFFu; // the chunk's stream: entries, blocks, newest block
				uint32_t umin[3], ext[3];
				float qlo[3], qhi[3]; // bounding box of the queries' positions
#pragma unroll
				for (int d = 0; d < 3; d++) {
					umin[d] = __reduce_min_sync(0xffffffffu, valid ? gmin[d] : 0xFFFFFFFFu);
					ext[d] = min(__reduce_max_sync(0xffffffffu, valid ? gmax[d] : 0u) - umin[d], axis_cap - 1u) + 1u;
					const float p = d == 0 ? me.x : (d == 1 ? me.y : me.z);
					qlo[d] = fkey_inv(__reduce_min_sync(0xffffffffu, valid ? fkey(p) : 0xFFFFFFFFu));
					qhi[d] = fkey_inv(__reduce_max_sync(0xffffffffu, valid ? fkey(p) : 0u));
				}
				const float r_cull = __uint_as_float(__reduce_max_sync(0xffffffffu, valid ? __float_as_uint(r_lane) : 0u));
				const float cull2 = A.cull ? r_cull * r_cull * 1.0001f : INFINITY;
				// Fused prune: a pair farther than the largest cutoff of the chunk (U, "kept at most", squared) is searched only
				// to spread the candidate's width onto the query (kernel_width.comp:49-53), which needs a candidate that starts
				// wider than the query is.  Cells beyond the cutoff whose widest particle cannot do that for any query of the chunk
				// are skipped: in a region of equal widths the walk shrinks from the search range (1.5 widths) to the cutoff.
				// (Not when the pairs of the unpruned list are being counted.)
				float keep2 = cull2;
				uint32_t min_mx0 = 0u;
				if (FUSED && !STATS && A.cull) {
					const float u_max = fkey_inv(__reduce_max_sync(0xffffffffu, valid ? fkey(qb.y) : 0u));
					keep2 = fminf(cull2, fmaxf(u_max, 0.0f) * 1.0001f);
					min_mx0 = __reduce_min_sync(0xffffffffu, valid ? my_mx : 0xFFFFFFFFu);
				}
				const uint32_t nxy = ext[0] * ext[1], ncell = SEARCH == 1 ? 27u : nxy * ext[2];
				const float inv_nxy = 1.0f / (float)nxy, inv_nx = 1.0f / (float)ext[0];
				const bool ghost_run = MG && first >= n_owned;
				// Slabs: ghosts live in the second cell table (their keys carry one extra bit).  It is walked only when it can matter:
				// a ghost's own candidates are owned particles (first table only), and a box of owned queries that lies inside this
				// rank's brick holds no cell of another rank.  Interior chunks -- nearly all of them -- walk one table like one GPU does.
				uint32_t n_layers = layers > 1u ? 2u : 1u;
				if (MG && n_layers == 2u) {
					bool inside = true;
#pragma unroll
					for (int d = 0; d < DIMS; d++) inside = inside && umin[d] >= A.mg_lo[d] && umin[d] + ext[d] - 1u <= A.mg_hi[d];
					if (SEARCH == 1) inside = false; // (no grid: the ghosts' code ranges are looked up for every chunk of owned queries)
					if (ghost_run || inside) n_layers = 1u;
				}
				uint32_t bs_first = 0u, bs_cnt = 0u, bs_first_g = 0u, bs_cnt_g = 0u;
				if (SEARCH == 1 && lane < 27u) {
					// the 27 cells around the chunk's cell, in the reference's loop order (z outer, x inner), as ranges of the
					// sorted codes: [first code >= cell, first code > cell | mask)
					const bool ident = A.misc[MW_IDENTITY] != 0u;
					const bs_cell c = bs_cell_of(ldg_int4(A.pos4, ident ? first : A.index_list[first]), A.range[first] * A.range_scale);
					const u96 cell = bs_neighbor_cell(c, (int)(lane % 3u), (int)((lane / 3u) % 3u), (int)(lane / 9u));
					const uint32_t n_own = MG ? min(n_owned, n) : n; // slabs: owned ids [0, n_own) and ghost ids [n_own, n) are each in code order
					bs_first = lower_bound96(A.c0, A.c1, A.c2, n_own, cell);
					bs_cnt = upper_bound96(A.c0, A.c1, A.c2, n_own, or96(cell, c.mask)) - bs_first;
					if (MG && n_layers == 2u) {
						bs_first_g = lower_bound96(A.c0, A.c1, A.c2, n, cell, n_own);
						bs_cnt_g = upper_bound96(A.c0, A.c1, A.c2, n, or96(cell, c.mask), n_own) - bs_first_g;
					}
				}
				for (uint32_t cbase = 0; cbase < ncell * n_layers; cbase += 32) {
					uint32_t ci = cbase + lane;
					uint32_t c_first = 0u, c_cnt = 0u;
					if (SEARCH == 1) { // first trip: the 27 ranges of owned ids; second trip (slabs): those of the ghosts
						c_first = cbase == 0u ? bs_first : bs_first_g; c_cnt = cbase == 0u ? bs_cnt : bs_cnt_g;
					} else if (ci < ncell * n_layers) {
						const uint32_t table_off = ci >= ncell ? A.table_cells : 0u;
						if (ci >= ncell) ci -= ncell;
						// ci = (cz * ny + cy) * nx + cx; float reciprocals + one correction step are exact for ci < 2^24
						uint32_t cz, cy;
						if (ncell <= (1u << 24)) {
							cz = (uint32_t)((float)ci * inv_nxy);
							if (cz * nxy > ci) cz--; else if ((cz + 1u) * nxy <= ci) cz++;
						} else {
							cz = ci / nxy;
						}
						const uint32_t rem = ci - cz * nxy;
						if (ncell <= (1u << 24)) {
							cy = (uint32_t)((float)rem * inv_nx);
							if (cy * ext[0] > rem) cy--; else if ((cy + 1u) * ext[0] <= rem) cy++;
						} else {
							cy = rem / ext[0];
						}
						const uint32_t cx = rem - cy * ext[0];
						const uint32_t a[3] = { umin[0] + cx, umin[1] + cy, umin[2] + cz };
						float gap2 = 0.0f;
#pragma unroll
						for (int d = 0; d < DIMS; d++) {
							const float lo = a[d] == 0u ? -INFINITY : g.mn[d] + (float)a[d] * csz[d];
							const float hi = g.mn[d] + ((float)a[d] + 1.0f) * csz[d];
							const float gp = fmaxf(fmaxf(lo - qhi[d], qlo[d] - hi) - eps[d], 0.0f);
							gap2 += gp * gp;
						}
						if (!(gap2 > cull2)) {
							const uint32_t h = (apbf_zhash<DIMS>(a[0], a[1], a[2], g.res) & key_mask) + table_off;
							if (!FUSED || STATS || !(gap2 > keep2) || __ldg(A.cell_maxw + h) > min_mx0) {
								c_first = __ldg(A.cell_start + h);
								c_cnt = __ldg(A.cell_end + h) - c_first;
							}
						}
					}
					uint32_t incl = c_cnt;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
						if (lane >= (unsigned)o) incl += t;
					}
					const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
					const uint32_t c_base = c_first - (incl - c_cnt); // candidate t of this cell is id c_base + t
					for (uint32_t t0 = 0; t0 < total; t0 += 32) {
						const uint32_t t = t0 + lane;
						// the cell that holds candidate t: first lane whose inclusive count exceeds t
						uint32_t pos = 0u;
#pragma unroll
						for (int step = 16; step > 0; step >>= 1) {
							const uint32_t v = __shfl_sync(0xffffffffu, incl, (int)(pos + step - 1));
							if (v <= t) pos += step;
						}
						const uint32_t cand = __shfl_sync(0xffffffffu, c_base, (int)(pos & 31u)) + t;
						const bool cvalid = t < total;
						float4 c4 = make_float4(0.f, 0.f, 0.f, -1.0f);
						float4 cb = make_float4(-1.0f, -1.0f, 0.0f, 0.0f);
						if (cvalid) {
							c4 = A.q4[cand];
							if (FUSED) cb = A.qb4[cand];
						}
						const float Kb = FUSED ? cb.x : c4.w;
						bool spread = false; // can a candidate of this batch raise the width of a query of this chunk?
						if (FUSED)
							spread = __reduce_max_sync(0xffffffffu, f2u(cb.w * APBF_KERNEL_WIDTH_RESOLUTION)) >
							         __reduce_min_sync(0xffffffffu, valid ? my_mx : 0xFFFFFFFFu);
						// ---- the tests, lane = candidate: one bit per query of the chunk ------------------------------------------
						uint32_t colK = 0u, colM = 0u, colU = 0u, colMu = 0u, colS = 0u;
						if (STATS) { // counting run: the plain loop with the range test of the unpruned list on top
							for (uint32_t qi = 0; qi < cnt; qi++) {
								const float4 qv = s_q[w][qi];
								const uint32_t bit = 1u << qi;
								const float dx = __fsub_rn(qv.x, c4.x), dy = __fsub_rn(qv.y, c4.y), dz = __fsub_rn(qv.z, c4.z);
								const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
								const bool pk = !(d2 > qv.w);
								colK |= pk ? bit : 0u;
								colM |= (pk && !(d2 > Kb)) ? bit : 0u;
								if (FUSED) {
									const bool pu = !(d2 > s_u[w][qi]);
									colU |= pu ? bit : 0u;
									colMu |= (pu && !(d2 > cb.y)) ? bit : 0u;
									colS |= !(d2 > s_T[w][qi]) ? bit : 0u;
								}
							}
						} else {
							// all thresholds of the chunk and of the batch equal: the mirrored test is the query's own test
							const bool uni = q_uniform && __all_sync(0xffffffffu, !cvalid || (Kb == qK && (!FUSED || cb.y == qU)));
							if (FUSED) {
								if (uni) chunk_tests<2>(s_qa[w], s_qb[w], s_u2[w], cnt, c4, Kb, cb.y, colK, colM, colU, colMu);
								else chunk_tests<3>(s_qa[w], s_qb[w], s_u2[w], cnt, c4, Kb, cb.y, colK, colM, colU, colMu);
							} else {
								if (uni) chunk_tests<0>(s_qa[w], s_qb[w], nullptr, cnt, c4, Kb, 0.0f, colK, colM, colU, colMu);
								else chunk_tests<1>(s_qa[w], s_qb[w], nullptr, cnt, c4, Kb, 0.0f, colK, colM, colU, colMu);
							}
						}
						const uint32_t sq = cand - first; // this candidate is query sq of the chunk: id != idN, neighborhood_green.comp:83
						const uint32_t live = cvalid ? ~(sq < 32u ? 1u << sq : 0u) : 0u;
						if (FUSED) {
							// pairs in an ambiguity band (the "for sure" and the "at most" form of a test disagree): these -- and only
							// these -- are decided again with the prune on the integer form (kernel_width.comp:36-38, :57).  On a lattice
							// whose spacing divides the cutoff those are the six axis neighbours of every particle.
							uint32_t amb = ((colU ^ colK) | (colMu ^ colM)) & live;
							if (__any_sync(0xffffffffu, amb != 0u)) {
								const int4 ci4 = cvalid ? __ldg(A.i4 + cand) : make_int4(0, 0, 0, 0);
								const float cutb = cvalid ? A.cutoff[cand] : -1.0f;
								while (amb) {
									const uint32_t qi = (uint32_t)__ffs(amb) - 1u, bit = 1u << qi;
									amb &= amb - 1u;
									const float4 qv = s_q[w][qi];
									const float dx = __fsub_rn(qv.x, c4.x), dy = __fsub_rn(qv.y, c4.y), dz = __fsub_rn(qv.z, c4.z);
									const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
									const int4 ia = __ldg(A.i4 + first + qi);
									// integer subtract first, then to float; D2 = dist^2 * 2^36
									const float ux = (float)(ci4.x - ia.x), uy = (float)(ci4.y - ia.y), uz = (float)(ci4.z - ia.z);
									const float D2 = dot3(ux, uy, uz, ux, uy, uz);
									const bool keep = !(d2 > s_T[w][qi]) && D2 <= A.cutoff[first + qi]; // :57
									const bool mir = keep && !(d2 > c4.w) && D2 <= cutb;                 // (idN, id) survives as well
									colK = keep ? (colK | bit) : (colK & ~bit);
									colM = mir ? (colM | bit) : (colM & ~bit);
								}
							}
						}
						colK &= live; colM &= live;
						if (STATS) n_searched += __popc((FUSED ? colS : colK) & live);
						if (FUSED && spread) {
							// the pair (idN, id) of the unpruned list spreads idN's width onto id (kernel_width.comp:49-53), gathered
							const int4 ci4 = cvalid ? __ldg(A.i4 + cand) : make_int4(0, 0, 0, 0);
							for (uint32_t qi = 0; qi < cnt; qi++) {
								const float4 qv = s_q[w][qi];
								const float dx = __fsub_rn(qv.x, c4.x), dy = __fsub_rn(qv.y, c4.y), dz = __fsub_rn(qv.z, c4.z);
								const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
								uint32_t val = 0u;
								if (((live >> qi) & 1u) && !(d2 > c4.w)) {
									const int4 ia = __ldg(A.i4 + first + qi);
									const float ux = (float)(ci4.x - ia.x), uy = (float)(ci4.y - ia.y), uz = (float)(ci4.z - ia.z);
									const float rx = ux * INV_R_POS, ry = uy * INV_R_POS, rz = uz * INV_R_POS;
									val = apbf_kw_influence(cb.w, sqrtf(dot3(rx, ry, rz, rx, ry, rz)));
								}
								val = __reduce_max_sync(0xffffffffu, val);
								if (lane == qi) my_mx = max(my_mx, val);
							}
						}
						if (MG && ghost_run) {
							// a ghost is a query only for the pairs nobody else provides: unmirrored pairs onto owned particles
							// (fused: colM is the mirrored bit AFTER the prune, so this is final; otherwise a following spread_kernel_width
							// can still turn mirrored pairs into unmirrored ones and layers == 3 keeps them all)
							if (cand >= n_owned) colK = 0u;
							if (FUSED || layers != 3u) colK &= ~colM;
						}
						// ---- lane = query again: append the hits to the chunk's stream ---------------------------------------------
						if (__any_sync(0xffffffffu, colK != 0u)) {
							// is the mirrored pair (idN, id) kept for every hit of the batch?  (equal widths: always)  Then no entry carries the
							// "unmirrored" flag and the mirrored bits need not be transposed.
							const bool all_mirrored = __all_sync(0xffffffffu, (colK & ~colM) == 0u);
							const uint32_t rowK = transpose32(colK, lane); // bit j: candidate j of the batch
							const uint32_t rowM = all_mirrored ? rowK : transpose32(colM, lane);
							const uint32_t c = (uint32_t)__popc(rowK);
							uint32_t inc = c;
#pragma unroll
							for (int o = 1; o < 32; o <<= 1) {
								const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
								if (lane >= (unsigned)o) inc += v;
							}
							const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
							const uint32_t n_before = n_alloc, need = ((tile_pos + tot - 1u) >> SB_SHIFT) + 1u;
							uint32_t base = 0u;
							if (need > n_before) {
								const uint32_t grp = need - n_before;
								if (lane == 0) {
									base = atomicAdd(A.misc + MW_STREAM_CURSOR, grp);
									if (base + grp > A.stream_blocks || base + grp < base) A.misc[MW_STREAM_OVERFLOW] = 1u;
								}
								base = __shfl_sync(0xffffffffu, base, 0);
								for (uint32_t i = lane; i < grp; i += 32u)
									if (base + i < A.stream_blocks) *(uint2*)(A.stream + (size_t)(base + i) * SB_WORDS) = make_uint2(first, SB_ENTRIES);
							}
							if (my_cnt + c > SB_RANK_MASK) A.misc[MW_STREAM_OVERFLOW] = 1u; // the rank field is full: two-pass fill
							// This query's hits, in candidate order, go to entries tp, tp + 1, ... of the chunk's stream as 8-byte stores
							// {idN | flag, lane | rank}.  The loop is uniform over the warp -- as many trips as the busiest query has hits, a
							// query with fewer sits out -- so the candidate's id comes by shuffle from the lane that loaded it, and a trip is
							// ~15 instructions (the per-lane `while (bits)` form with its shared-memory look-up was 40: a fifth of this
							// kernel's instructions).  A run of <= 32 entries crosses at most one block boundary: both block addresses are
							// worked out beforehand.  (After an overflow the stream is thrown away: writes without a block go to block 0.)
							const uint32_t tp = tile_pos + (inc - c), o0 = tp >> SB_SHIFT;
							const uint32_t blk0 = o0 < n_before ? last_blk : base + (o0 - n_before), blk1 = base + (o0 + 1u - n_before);
							uint2* P = (uint2*)(A.stream + (size_t)(blk0 < A.stream_blocks ? blk0 : 0u) * SB_WORDS + SB_HEADER) + (tp & (SB_ENTRIES - 1u));
							uint2* const P1 = (uint2*)(A.stream + (size_t)(blk1 < A.stream_blocks ? blk1 : 0u) * SB_WORDS + SB_HEADER);
							const uint32_t cross = SB_ENTRIES - (tp & (SB_ENTRIES - 1u)); // entries of the run that still fit its first block
							const uint32_t trips = __reduce_max_sync(0xffffffffu, c);
							uint32_t m = rowK, rk = (lane << SB_RANK_BITS) | (my_cnt & SB_RANK_MASK);
							for (uint32_t it = 0; it < trips; it++) {
								const bool act = m != 0u;
								const uint32_t j = act ? (uint32_t)__ffs(m) - 1u : 0u;
								uint32_t word = __shfl_sync(0xffffffffu, cand, (int)j);
								if (!all_mirrored && ((rowM >> j) & 1u) == 0u) word |= NB_UNMIRRORED;
								if (act) {
									if (it == cross) P = P1;
									*P = make_uint2(word, rk);
									P++; rk++;
								}
								m &= m - 1u;
							}
							my_cnt += c;
							if (need > n_before) { last_blk = base + (need - n_before) - 1u; n_alloc = need; }
							tile_pos += tot;
						}
					}
				}
				if (valid) {
					A.counts[id] = my_cnt;
					if (FUSED) A.kwfx[id] = my_mx; // kernel_width_init.comp:35 + the atomicMax of kernel_width.comp:53, gathered
				}
				if (lane == 0 && n_alloc != 0u && last_blk < A.stream_blocks)
					A.stream[(size_t)last_blk * SB_WORDS + 1u] = tile_pos - SB_ENTRIES * (n_alloc - 1u); // entries of the last block
			}
		}
	}
	if (STATS) {
		n_searched = __reduce_add_sync(0xffffffffu, n_searched);
		if (lane == 0 && n_searched) atomicAdd(A.misc + MW_TOTAL_PAIRS, n_searched);
	}
}

// second half of the one-pass emit: every entry of the hit stream -> offsets[id] + rank
__global__ void __launch_bounds__(SB_ENTRIES)
k_regroup(const uint32_t* __restrict__ stream, uint32_t stream_blocks, const uint32_t* __restrict__ offsets,
          uint32_t* __restrict__ pairs, uint32_t* __restrict__ nbl, uint32_t cap, uint32_t* misc, int no_fallback)
{
	if (misc[MW_STREAM_OVERFLOW] != 0u) {
		// (fused + slabs has no two-pass fill behind it: raise the sticky overflow flag; needs more pairs than the list holds)
		if (no_fallback && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(misc + MW_FLAGS, 1u);
		return;
	}
	const uint32_t used = min(misc[MW_STREAM_CURSOR], stream_blocks);
	uint32_t n_asym = 0u;
	constexpr int U = 4; // blocks in flight per thread: the chain header -> entry -> offset -> store is pure latency
	for (uint32_t blk0 = blockIdx.x; blk0 < used; blk0 += U * gridDim.x) {
		uint2 hdr[U];
		uint32_t word[U], rk[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t blk = blk0 + u * gridDim.x;
			hdr[u] = make_uint2(0u, 0u);
			if (blk < used) {
				const uint32_t* B = stream + (size_t)blk * SB_WORDS;
				hdr[u] = *(const uint2*)B;
				const uint2 e = ((const uint2*)(B + SB_HEADER))[threadIdx.x]; // entries are {idN | flag, lane | rank} pairs
				word[u] = e.x;
				rk[u] = e.y;
			}
		}
		uint32_t o[U], id[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			o[u] = 0xFFFFFFFFu;
			if (threadIdx.x < hdr[u].y) {
				id[u] = hdr[u].x + (rk[u] >> SB_RANK_BITS);
				o[u] = offsets[id[u]] + (rk[u] & SB_RANK_MASK);
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			if (threadIdx.x < hdr[u].y && o[u] < cap) {
				if (pairs) *(uint2*)(pairs + 2 * (size_t)o[u]) = make_uint2(id[u], word[u] & NB_ID_MASK);
				nbl[o[u]] = word[u];
				n_asym += word[u] >> 31;
			}
		}
	}
	n_asym = __reduce_add_sync(0xffffffffu, n_asym);
	if ((threadIdx.x & 31u) == 0u && n_asym) atomicAdd(misc + MW_N_ASYM, n_asym);
}

// The same pass with the stream blocks staged in shared memory by the bulk-copy engine (north_star: "TMA bulk copies where
// cells are contiguous" -- the hit-stream block is the one contiguous, aligned 1056-byte unit on the search path): one thread
// arms an mbarrier with the byte count and issues cp.async.bulk (SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64), U blocks in flight
// per CTA; every thread then waits on the barrier's phase and takes its entry from shared memory.  The default
// (APBF_REGROUP_TMA=0 selects the plain form); measured against it in profiles/r02_variants.md.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(SB_ENTRIES)
k_regroup_tma(const uint32_t* __restrict__ stream, uint32_t stream_blocks, const uint32_t* __restrict__ offsets,
              uint32_t* __restrict__ nbl, uint32_t cap, uint32_t* misc, int no_fallback)
{
	constexpr int U = 4;
	__shared__ __align__(128) uint32_t s_blk[U][SB_WORDS];
	__shared__ __align__(8) unsigned long long s_bar[U];
	if (misc[MW_STREAM_OVERFLOW] != 0u) {
		if (no_fallback && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(misc + MW_FLAGS, 1u);
		return;
	}
	const uint32_t used = min(misc[MW_STREAM_CURSOR], stream_blocks);
	if (threadIdx.x == 0) {
#pragma unroll
		for (int u = 0; u < U; u++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar[u])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	uint32_t n_asym = 0u, phase = 0u;
	for (uint32_t blk0 = blockIdx.x; blk0 < used; blk0 += U * gridDim.x) {
		if (threadIdx.x == 0) {
#pragma unroll
			for (int u = 0; u < U; u++) {
				const uint32_t blk = blk0 + u * gridDim.x;
				if (blk < used) {
					const uint32_t bar = smem_u32(&s_bar[u]);
					asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(SB_WORDS * 4u) : "memory");
					asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					             :: "r"(smem_u32(&s_blk[u][0])), "l"(stream + (size_t)blk * SB_WORDS), "r"(SB_WORDS * 4u), "r"(bar) : "memory");
				}
			}
		}
		uint32_t o[U], word[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			o[u] = 0xFFFFFFFFu;
			const uint32_t blk = blk0 + u * gridDim.x;
			if (blk >= used) continue;
			const uint32_t bar = smem_u32(&s_bar[u]);
			uint32_t done = 0u;
			while (!done)
				asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				             : "=r"(done) : "r"(bar), "r"(phase) : "memory");
			const uint2 hdr = *(const uint2*)&s_blk[u][0];
			const uint2 e = ((const uint2*)&s_blk[u][SB_HEADER])[threadIdx.x];
			word[u] = e.x;
			const uint32_t rk = e.y;
			if (threadIdx.x < hdr.y) o[u] = offsets[hdr.x + (rk >> SB_RANK_BITS)] + (rk & SB_RANK_MASK);
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			if (o[u] < cap) {
				nbl[o[u]] = word[u];
				n_asym += word[u] >> 31;
			}
		}
		phase ^= 1u;
		__syncthreads(); // everybody has read the staged blocks before the next round overwrites them
	}
	n_asym = __reduce_add_sync(0xffffffffu, n_asym);
	if ((threadIdx.x & 31u) == 0u && n_asym) atomicAdd(misc + MW_N_ASYM, n_asym);
}

// ---- binary-search pair emit, two-pass form (overflow fallback of the stream form) ----------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(128)
k_bsearch_emit(const uint32_t* __restrict__ index_list, const int32_t* __restrict__ pos4, const uint32_t* __restrict__ c0,
               const uint32_t* __restrict__ c1, const uint32_t* __restrict__ c2, const float* __restrict__ range,
               const uint32_t* __restrict__ len, float range_scale, uint32_t* __restrict__ counts,
               const uint32_t* __restrict__ offsets, uint32_t* __restrict__ pairs, uint32_t cap, uint32_t* __restrict__ nbl,
               uint32_t* misc, int fallback)
{
	if (fallback && misc[MW_STREAM_OVERFLOW] == 0u) return; // only if the hit stream of the one-pass emit overflowed
	const uint32_t n = *len;
	const bool ident = misc[MW_IDENTITY] != 0u;
	const u96 xMask3 = mk96(011111111111u, 022222222222u, 04444444444u);
	const u96 yMask3 = mk96(022222222222u, 04444444444u, 011111111111u);
	const u96 zMask3 = mk96(04444444444u, 011111111111u, 022222222222u);
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const float r = range[id] * range_scale;
		const uint32_t idx = ident ? id : index_list[id];
		const int4 ip = ldg_int4(pos4, idx);
		u96 code;
		apbf_encode96(ip.x, ip.y, ip.z, code.v);
		const uint32_t digits = f2u(ceilf(log2f(r * R_POS))) * 3u; // :183
		u96 mask;
		mask.v[0] = (digits < 32u ? 1u << digits : 0u) - 1u;
		mask.v[1] = (digits < 64u ? 1u << (max(digits, 32u) - 32u) : 0u) - 1u;
		mask.v[2] = (digits < 96u ? 1u << (max(digits, 64u) - 64u) : 0u) - 1u;
		const u96 center = mk96(code.v[0] - (code.v[0] & mask.v[0]), code.v[1] - (code.v[1] & mask.v[1]), code.v[2] - (code.v[2] & mask.v[2]));
		const u96 step = plus96(mask, mk96(1u, 0u, 0u));
		u96 xs[3], ys[3], zs[3];
		xs[0] = and96(minus96(and96(center, xMask3), step), xMask3);
		xs[2] = and96(plus96(or96(center, not96(xMask3)), step), xMask3);
		ys[0] = and96(minus96(and96(center, yMask3), shl96_small(step, 1u)), yMask3);
		ys[2] = and96(plus96(or96(center, not96(yMask3)), shl96_small(step, 1u)), yMask3);
		zs[0] = and96(minus96(and96(center, zMask3), shl96_small(step, 2u)), zMask3);
		zs[2] = and96(plus96(or96(center, not96(zMask3)), shl96_small(step, 2u)), zMask3);
		xs[1] = and96(center, xMask3);
		ys[1] = and96(center, yMask3);
		zs[1] = and96(center, zMask3);
		const float px = (float)ip.x * INV_R_POS, py = (float)ip.y * INV_R_POS, pz = (float)ip.z * INV_R_POS;
		uint32_t cnt = 0, out = 0, n_asym = 0;
		if (FILL) out = offsets[id];
		for (int cz = 0; cz < 3; cz++) for (int cy = 0; cy < 3; cy++) for (int cx = 0; cx < 3; cx++) {
			const u96 cellCode = or96(or96(xs[cx], ys[cy]), zs[cz]);
			const u96 cellLast = or96(cellCode, mask);
			for (uint32_t idN = lower_bound96(c0, c1, c2, n, cellCode); idN < n; idN++) {
				if (greater96(mk96(__ldg(c0 + idN), __ldg(c1 + idN), __ldg(c2 + idN)), cellLast)) break;
				const uint32_t idxN = ident ? idN : index_list[idN];
				const int4 iq = ldg_int4(pos4, idxN);
				const float d = dist_rn(px, py, pz, (float)iq.x * INV_R_POS, (float)iq.y * INV_R_POS, (float)iq.z * INV_R_POS);
				if (id != idN && d <= r) {
					if (FILL) {
						const uint32_t o = out + cnt;
						if (o < cap) {
							*(uint2*)(pairs + 2 * (size_t)o) = make_uint2(id, idN);
							const bool mirrored = d <= range[idN] * range_scale;
							nbl[o] = idN | (mirrored ? 0u : NB_UNMIRRORED);
							n_asym += mirrored ? 0u : 1u;
						}
					}
					cnt++;
				}
			}
		}
		if (FILL) {
			if (n_asym) atomicAdd(misc + MW_N_ASYM, n_asym);
		} else {
			counts[id] = cnt;
		}
	}
}

__global__ void k_clear_search_words(uint32_t* misc)
{
	misc[MW_N_ASYM] = 0u; misc[MW_TOTAL_PAIRS] = 0u; misc[MW_KEPT_PAIRS] = 0xFFFFFFFFu; misc[MW_OCC_CELLS] = 0u;
	misc[MW_EMIT_TICKET0] = 0u; misc[MW_EMIT_TICKET1] = 0u; misc[MW_STREAM_CURSOR] = 0u; misc[MW_STREAM_OVERFLOW] = 0u; misc[MW_MAX_INIT] = 0u; misc[MW_PMAX] = 0u;
	for (int k = 0; k < 4; k++) { misc[MW_THR_MIN + k] = 0xFFFFFFFFu; misc[MW_THR_MAX + k] = 0u; }
}

// offsets[id] for the ids between the list's length and its capacity: empty segments.  A list that grows afterwards (the copies
// of update_transfers' splits, update_transfers.cpp:64-68) then finds its new particles without pairs instead of reading
// entries the scan never wrote.
__global__ void k_offsets_tail(uint32_t* __restrict__ offsets, const uint32_t* __restrict__ len, uint32_t n_cap)
{
	const uint32_t n = min(*len, n_cap);
	const uint32_t total = offsets[n];
	for (uint32_t id = n + 1u + blockIdx.x * blockDim.x + threadIdx.x; id <= n_cap; id += gridDim.x * blockDim.x) offsets[id] = total;
}

// shared front half of both searches: gather hidden arrays by sorted_index, rebuild the index list, gather per-id arrays
int reorder_lists(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_array* range, const uint32_t* sorted_index)
{
	apbf_particles& p = fluid->particle;
	cudaStream_t st = ctx->stream;
	const uint32_t nh_cap = p.hidden_capacity, n_cap = p.capacity;
	reorder_table t;
	memset(&t, 0, sizeof t);
	const apbf_array* a16[3] = { &p.position, &p.velocity, &p.pos_backup };
	const apbf_array* a4[3] = { &p.inverse_mass, &p.radius, &p.transferring };
	for (auto a : a16) {
		APBF_REQUIRE(ctx, a->data && a->reorder_out && a->data != a->reorder_out);
		t.src16[t.n16] = (const int4*)a->data; t.dst16[t.n16] = (int4*)a->reorder_out; t.n16++;
	}
	for (auto a : a4) {
		APBF_REQUIRE(ctx, a->data && a->reorder_out && a->data != a->reorder_out);
		t.src4[t.n4] = (const uint32_t*)a->data; t.dst4[t.n4] = (uint32_t*)a->reorder_out; t.n4++;
	}
	k_reorder<<<apbf_grid(ctx, nh_cap, 256), 256, 0, st>>>(t, sorted_index, p.hidden_length, ctx->misc() + MW_INDEX_NONIDENT);
	APBF_LAUNCHED(ctx);

	// index list and the permutation of the ids
	APBF_REQUIRE(ctx, p.index_list.data && p.index_list.reorder_out && p.index_list.data != p.index_list.reorder_out);
	uint32_t* mark = (uint32_t*)ctx->scratch_get(SLOT_INV_PERM, sizeof(uint32_t) * (size_t)nh_cap);
	uint32_t* flags = (uint32_t*)ctx->scratch_get(SLOT_HIDDEN_FLAGS, sizeof(uint32_t) * (size_t)(nh_cap + 1));
	uint32_t* offs = (uint32_t*)ctx->scratch_get(SLOT_HIDDEN_OFFS, sizeof(uint32_t) * (size_t)(nh_cap + 1));
	uint32_t* id_perm = (uint32_t*)ctx->scratch_get(SLOT_TMP_VALS2, sizeof(uint32_t) * (size_t)n_cap);
	if (!mark || !flags || !offs || !id_perm) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	APBF_CUDA(ctx, cudaMemsetAsync(mark, 0, sizeof(uint32_t) * (size_t)nh_cap, st));
	k_mark_members<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>((const uint32_t*)p.index_list.data, p.length, p.hidden_length, mark, ctx->misc());
	APBF_LAUNCHED(ctx);
	k_member_flags<<<apbf_grid(ctx, nh_cap, 256), 256, 0, st>>>(sorted_index, mark, p.hidden_length, flags, ctx->misc());
	APBF_LAUNCHED(ctx);
	APBF_TRY(apbf_scan_u32(ctx, flags, offs, p.hidden_length, nh_cap, false, nullptr, 0xFFFFFFFFu, nullptr, nullptr, ctx->misc() + MW_INDEX_NONIDENT));
	k_compact_members<<<apbf_grid(ctx, nh_cap, 256), 256, 0, st>>>(sorted_index, mark, offs, p.hidden_length, p.length,
	                                                               (uint32_t*)p.index_list.reorder_out, id_perm, n_cap, ctx->misc());
	APBF_LAUNCHED(ctx);

	gather_table gt;
	memset(&gt, 0, sizeof gt);
	const apbf_array* ids[5] = { &fluid->target_radius, &fluid->kernel_width, &fluid->boundariness, &fluid->boundary_distance, range };
	for (int i = 0; i < 5; i++) {
		const apbf_array* a = ids[i];
		if (!a || !a->data) continue;
		bool dup = false;
		for (int j = 0; j < gt.n4; j++) dup = dup || gt.src4[j] == (const uint32_t*)a->data;
		if (dup) continue;
		APBF_REQUIRE(ctx, a->reorder_out && a->data != a->reorder_out);
		gt.src4[gt.n4] = (const uint32_t*)a->data; gt.dst4[gt.n4] = (uint32_t*)a->reorder_out; gt.n4++;
	}
	if (gt.n4 > 0) {
		k_gather_ids<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(gt, id_perm, p.length);
		APBF_LAUNCHED(ctx);
	}
	return APBF_OK;
}

template <bool FILL, int VARIANT>
void launch_emit(int dims, unsigned grid, cudaStream_t st, const emit_args& A)
{
	if (dims == 3) k_green_emit<FILL, VARIANT, 3><<<grid, EMIT_WARPS * 32, 0, st>>>(A);
	else k_green_emit<FILL, VARIANT, 2><<<grid, EMIT_WARPS * 32, 0, st>>>(A);
}

template <bool FILL>
void launch_emit(int variant, int dims, unsigned grid, cudaStream_t st, const emit_args& A)
{
	if (variant == EMIT_FUSED) launch_emit<FILL, EMIT_FUSED>(dims, grid, st, A);
	else if (variant == EMIT_MG) launch_emit<FILL, EMIT_MG>(dims, grid, st, A);
	else launch_emit<FILL, EMIT_PLAIN>(dims, grid, st, A);
}

template <int VARIANT, bool STATS>
void launch_stream(int dims, unsigned grid, cudaStream_t st, const emit_args& A)
{
	if (dims == 3) k_green_stream<VARIANT, 3, STATS><<<grid, EMIT_WARPS * 32, 0, st>>>(A);
	else k_green_stream<VARIANT, 2, STATS><<<grid, EMIT_WARPS * 32, 0, st>>>(A);
}
void launch_stream(int variant, bool stats, int dims, unsigned grid, cudaStream_t st, const emit_args& A)
{
	if (variant == EMIT_FUSED && stats) launch_stream<EMIT_FUSED, true>(dims, grid, st, A);
	else if (variant == EMIT_FUSED) launch_stream<EMIT_FUSED, false>(dims, grid, st, A);
	else if (variant == EMIT_FUSED_MG && stats) launch_stream<EMIT_FUSED_MG, true>(dims, grid, st, A);
	else if (variant == EMIT_FUSED_MG) launch_stream<EMIT_FUSED_MG, false>(dims, grid, st, A);
	else if (variant == EMIT_MG) launch_stream<EMIT_MG, false>(dims, grid, st, A);
	else launch_stream<EMIT_PLAIN, false>(dims, grid, st, A);
}

} // namespace

// neighborhood_green::apply (neighborhood_green.cpp:27-77); fuse_kw: followed by spread_kernel_width::apply
// (spread_kernel_width.cpp:12-26) on the same lists, with range == fluid->kernel_width as in pool.cpp:83-89
int apbf_green_search(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_array* range, apbf_neighbors* nb, float range_scale,
                      const float min_pos[3], const float max_pos[3], uint32_t res_log2, const apbf_search_debug* dbg, bool fuse_kw,
                      uint32_t* out_kw_fixed, bool write_public)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid && range && nb && min_pos && max_pos);
	APBF_REQUIRE(ctx, nb->pairs && nb->length && fluid->particle.length && fluid->particle.hidden_length);
	APBF_TRY(apbf_nbr_activate(ctx, nb)); // SLOT_OFFSETS / SLOT_NB are this list's from here on
	ctx->nbr_valid = false;
	apbf_particles& p = fluid->particle;
	cudaStream_t st = ctx->stream;
	const uint32_t nh_cap = p.hidden_capacity, n_cap = p.capacity;
	if (nh_cap == 0 || n_cap == 0) { APBF_CUDA(ctx, cudaMemsetAsync(nb->length, 0, 4, st)); return APBF_OK; }
	if (fuse_kw) APBF_REQUIRE(ctx, p.radius.data && fluid->target_radius.data && fluid->kernel_width.data);
	apbf_grid_params g;
	APBF_TRY(apbf_make_grid_params(ctx, min_pos, max_pos, res_log2, &g));
	const uint32_t max_hash = 1u << (res_log2 * (uint32_t)ctx->dims); // neighborhood_green.cpp:31
	uint32_t* misc = ctx->misc();

	uint32_t* keys = (uint32_t*)ctx->scratch_get(SLOT_SORT_KEYS_A, sizeof(uint32_t) * (size_t)nh_cap);
	uint32_t* skeys = (uint32_t*)ctx->scratch_get(SLOT_TMP_KEYS, sizeof(uint32_t) * (size_t)nh_cap);
	uint32_t* sidx = (uint32_t*)ctx->scratch_get(SLOT_TMP_VALS, sizeof(uint32_t) * (size_t)nh_cap);
	const uint32_t layers = ctx->mg_enabled ? 2u : 1u; // ghosts: second key space / cell table
	const uint32_t emit_mode = ctx->mg_enabled && ctx->mg_ghost_all_pairs && !fuse_kw ? 3u : layers;
	APBF_REQUIRE(ctx, layers == 1u || max_hash <= (1u << 30));
	uint32_t* cs = (uint32_t*)ctx->scratch_get(SLOT_CELL_START, sizeof(uint32_t) * (size_t)max_hash * layers);
	uint32_t* ce = (uint32_t*)ctx->scratch_get(SLOT_CELL_END, sizeof(uint32_t) * (size_t)max_hash * layers);
	uint32_t* counts = (uint32_t*)ctx->scratch_get(SLOT_COUNTS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	uint32_t* offsets = (uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	uint32_t* nbl = (uint32_t*)ctx->scratch_get(SLOT_NB, sizeof(uint32_t) * ((size_t)nb->capacity + 1));
	float4* q4 = (float4*)ctx->scratch_get(SLOT_Q4, sizeof(float4) * (size_t)n_cap);
	uint32_t* key_id = (uint32_t*)ctx->scratch_get(SLOT_KEY_ID, sizeof(uint32_t) * (size_t)n_cap);
	if (!keys || !skeys || !sidx || !cs || !ce || !counts || !offsets || !nbl || !misc || !q4 || !key_id)
		return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	build_kw_args K;
	memset(&K, 0, sizeof K);
	uint32_t* kwfx = nullptr;
	if (fuse_kw) {
		K.i4 = (int4*)ctx->scratch_get(SLOT_I4, sizeof(int4) * (size_t)n_cap);
		K.cutoff = (float*)ctx->scratch_get(SLOT_CUTOFF, sizeof(float) * (size_t)n_cap);
		K.qb4 = (float4*)ctx->scratch_get(SLOT_QB4, sizeof(float4) * (size_t)n_cap);
		K.cell_maxw = (uint32_t*)ctx->scratch_get(SLOT_CELL_MAXW, sizeof(uint32_t) * (size_t)max_hash * layers);
		if (!K.cell_maxw) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
		kwfx = (uint32_t*)ctx->scratch_get(SLOT_KWFX, sizeof(uint32_t) * (size_t)n_cap);
		if (!K.i4 || !K.cutoff || !K.qb4 || !kwfx) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	}
	// hit stream of the one-pass emit: every chunk of queries ends with a partly filled block, and there is at most one
	// chunk per particle, hence capacity / 128 + particles blocks would hold any list that fits the pair buffer; the usual
	// demand is pairs / 128 + particles / 20, and capacity / 128 + particles / 8 is what is reserved
	uint32_t stream_blocks = (uint32_t)std::min<size_t>((size_t)nb->capacity / SB_ENTRIES + (size_t)n_cap / 8u + 1024u, 0x7FFFFFFFu);
	if (ctx->stream_blocks_cap) stream_blocks = std::min(stream_blocks, std::max(ctx->stream_blocks_cap, 1u));
	uint32_t* stream = (uint32_t*)ctx->scratch_get(SLOT_STREAM, sizeof(uint32_t) * (size_t)stream_blocks * SB_WORDS);
	if (!stream) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);

	// hash all hidden particles (neighborhood_green.cpp:50-52), sort by hash with the slot as payload (:53)
	{
		apbf_prof_scope ps(ctx, PROF_HASH_SORT);
		APBF_TRY(apbf_launch_position_hash(ctx, (const int32_t*)p.position.data, keys, p.hidden_length, nh_cap, g, layers == 2u ? max_hash : 0u));
		APBF_TRY(apbf_radix_sort_pairs(ctx, keys, nullptr, skeys, sidx, p.hidden_length, nh_cap, apbf_reference_sort_bits(max_hash)));
	}
	// permute hidden arrays, index list and per-id arrays (:54-56)
	{
		apbf_prof_scope ps(ctx, PROF_REORDER);
		if (ctx->hook_wait_before_reorder) APBF_CUDA(ctx, cudaStreamWaitEvent(st, ctx->hook_wait_before_reorder, 0));
		APBF_TRY(reorder_lists(ctx, fluid, range, sidx));
		if (ctx->hook_record_after_reorder) APBF_CUDA(ctx, cudaEventRecord(ctx->hook_record_after_reorder, st));
	}
	const uint32_t* new_index = (const uint32_t*)p.index_list.reorder_out;
	const int32_t* new_pos = (const int32_t*)p.position.reorder_out;
	const float* new_range = (const float*)range->reorder_out;
	if (fuse_kw) {
		K.radius = (const float*)p.radius.reorder_out;
		K.target_radius = (const float*)fluid->target_radius.reorder_out;
		K.kernel_width = (const float*)fluid->kernel_width.reorder_out;
		K.base_on_target_radius = ctx->settings.mBaseKernelWidthOnTargetRadius;
		K.pmax = 0.0f;
		for (int d = 0; d < 3; d++) K.pmax = std::max(K.pmax, std::max(fabsf(min_pos[d]), fabsf(max_pos[d])));
	}
	// cell ranges (:58-63)
	{
		apbf_prof_scope ps(ctx, PROF_CELL_RANGES);
		APBF_TRY(apbf_launch_find_value_ranges(ctx, new_index, skeys, cs, ce, p.length, n_cap, max_hash * layers));
	}
	// pairs (:64-74): count, scan, fill
	k_clear_search_words<<<1, 1, 0, st>>>(misc);
	APBF_LAUNCHED(ctx);
	const unsigned egrid = apbf_grid(ctx, n_cap, EMIT_WARPS * 32, 4);
	static const int cull = getenv("APBF_NO_CULL") ? 0 : 1; // debugging aid: walk every cell of the union box
	const int variant = (fuse_kw ? EMIT_FUSED : 0) | (ctx->mg_enabled ? EMIT_MG : 0);
	emit_args A;
	memset(&A, 0, sizeof A);
	A.q4 = q4; A.key_id = key_id; A.range = new_range; A.cell_start = cs; A.cell_end = ce; A.len = p.length; A.g = g;
	A.range_scale = range_scale; A.counts = counts; A.offsets = offsets; A.pairs = nb->pairs; A.nbl = nbl; A.cap = nb->capacity;
	A.misc = misc; A.cull = cull; A.table_cells = max_hash; A.layers = emit_mode; A.i4 = K.i4; A.cutoff = K.cutoff; A.kwfx = kwfx;
	for (int d = 0; d < 3; d++) { A.mg_lo[d] = ctx->mg_lo[d]; A.mg_hi[d] = ctx->mg_hi[d]; }
	A.qb4 = K.qb4; A.cell_maxw = K.cell_maxw; A.stream = stream; A.stream_blocks = stream_blocks;
	static const int two_pass = getenv("APBF_TWO_PASS_EMIT") ? 1 : 0; // debugging aid: the count/fill emit instead of stream/regroup
	// hit-stream blocks staged by the bulk-copy engine: 25 % faster at 5 x 10^8 pairs (2.61 -> 1.97 ms), the same at 3 x 10^7 (r02c)
	static const int regroup_tma = getenv("APBF_REGROUP_TMA") ? atoi(getenv("APBF_REGROUP_TMA")) : 1;
	{
		apbf_prof_scope ps(ctx, PROF_EMIT_COUNT);
		if (fuse_kw) APBF_CUDA(ctx, cudaMemsetAsync(K.cell_maxw, 0, sizeof(uint32_t) * (size_t)max_hash * layers, st));
		k_build_q4<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(new_index, new_pos, skeys, new_range, range_scale, p.length, q4, key_id, misc, K);
		APBF_LAUNCHED(ctx);
		A.ticket = misc + MW_EMIT_TICKET0;
		if (two_pass) launch_emit<false>(variant, g.dims, egrid, st, A);
		else launch_stream(variant, ctx->search_stats, g.dims, egrid, st, A);
		APBF_LAUNCHED(ctx);
	}
	{
		apbf_prof_scope ps(ctx, PROF_EMIT_SCAN);
		// fused: the counts are those of the pruned list; the size of the unpruned one is summed up by the count pass
		APBF_TRY(apbf_scan_u32(ctx, counts, offsets, p.length, n_cap, false, nb->length, nb->capacity, misc + MW_FLAGS,
		                       misc + (fuse_kw ? MW_KEPT_PAIRS : MW_TOTAL_PAIRS)));
		k_offsets_tail<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(offsets, p.length, n_cap);
		APBF_LAUNCHED(ctx);
	}
	{
		apbf_prof_scope ps(ctx, PROF_EMIT_FILL);
		if (!two_pass) {
			if (regroup_tma) k_regroup_tma<<<ctx->num_sms * 8, SB_ENTRIES, 0, st>>>(stream, stream_blocks, offsets, nbl, nb->capacity, misc, variant == EMIT_FUSED_MG ? 1 : 0);
			else k_regroup<<<ctx->num_sms * 16, SB_ENTRIES, 0, st>>>(stream, stream_blocks, offsets, nullptr, nbl, nb->capacity, misc,
			                                                         variant == EMIT_FUSED_MG ? 1 : 0);
			APBF_LAUNCHED(ctx);
			if (write_public) APBF_TRY(apbf_launch_expand_pairs(ctx, offsets, nbl, p.length, n_cap, nb->pairs, nb->capacity));
		}
		if (variant != EMIT_FUSED_MG) { // (fused + slabs has no two-pass form: an overflow there only raises the sticky flag)
			A.ticket = misc + MW_EMIT_TICKET1;
			A.fallback = two_pass ? 0 : 1; // leaves at once unless the stream overflowed
			launch_emit<true>(variant, g.dims, egrid, st, A);
			APBF_LAUNCHED(ctx);
		}
	}
	// (the two-pass fill writes the public list itself; it only runs when the hit stream overflowed, i.e. when the list is clamped)
	apbf_nbr_built(ctx, nb, n_cap, write_public || two_pass);

	if (dbg) {
		if (dbg->sorted_key) APBF_CUDA(ctx, cudaMemcpyAsync(dbg->sorted_key, skeys, sizeof(uint32_t) * (size_t)nh_cap, cudaMemcpyDeviceToDevice, st));
		if (dbg->sorted_index) APBF_CUDA(ctx, cudaMemcpyAsync(dbg->sorted_index, sidx, sizeof(uint32_t) * (size_t)nh_cap, cudaMemcpyDeviceToDevice, st));
		if (dbg->cell_start) APBF_CUDA(ctx, cudaMemcpyAsync(dbg->cell_start, cs, sizeof(uint32_t) * (size_t)max_hash, cudaMemcpyDeviceToDevice, st));
		if (dbg->cell_end) APBF_CUDA(ctx, cudaMemcpyAsync(dbg->cell_end, ce, sizeof(uint32_t) * (size_t)max_hash, cudaMemcpyDeviceToDevice, st));
		if (dbg->pair_offsets) APBF_CUDA(ctx, cudaMemcpyAsync(dbg->pair_offsets, offsets, sizeof(uint32_t) * (size_t)(n_cap + 1), cudaMemcpyDeviceToDevice, st));
	}
	if (fuse_kw) {
		// The reordered lists are still in reorder_out (the caller swaps after the search): the width update works on them.
		apbf_fluid sorted = *fluid;
		sorted.kernel_width.data = fluid->kernel_width.reorder_out;
		apbf_prof_scope ps(ctx, PROF_KW_MISC);
		APBF_TRY(apbf_kw_finish(ctx, &sorted, out_kw_fixed));
	}
	return APBF_OK;
}

extern "C" {

int apbf_neighborhood_green_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_array* range, apbf_neighbors* nb,
                                  float range_scale, const float min_pos[3], const float max_pos[3], uint32_t res_log2,
                                  const apbf_search_debug* dbg)
{
	return apbf_green_search(ctx, fluid, range, nb, range_scale, min_pos, max_pos, res_log2, dbg, false, nullptr, true);
}

int apbf_neighborhood_green_spread_apply(apbf_ctx* ctx, apbf_fluid* fluid, apbf_neighbors* nb, float range_scale,
                                         const float min_pos[3], const float max_pos[3], uint32_t res_log2,
                                         const apbf_search_debug* dbg, uint32_t* out_kw_fixed)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid);
	return apbf_green_search(ctx, fluid, &fluid->kernel_width, nb, range_scale, min_pos, max_pos, res_log2, dbg, true, out_kw_fixed, true);
}

} // extern "C"

// neighborhood_binary_search::apply (neighborhood_binary_search.cpp:22-75); fuse_kw: followed by spread_kernel_width::apply on
// the same lists with range == fluid->kernel_width (pool.cpp:83-89 with NEIGHBORHOOD_TYPE 3)
int apbf_binary_search(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_array* range, apbf_neighbors* nb, float range_scale,
                       const apbf_search_debug* dbg, bool fuse_kw, uint32_t* out_kw_fixed, bool write_public)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid && range && nb);
	APBF_REQUIRE(ctx, nb->pairs && nb->length && fluid->particle.length && fluid->particle.hidden_length);
	// (slabs: ownership is a matter of the Green grid's key -- mgpu.cu -- whatever search runs over owned particles + ghosts)
	const bool mg = ctx->mg_enabled;
	APBF_TRY(apbf_nbr_activate(ctx, nb));
	ctx->nbr_valid = false;
	apbf_particles& p = fluid->particle;
	cudaStream_t st = ctx->stream;
	const uint32_t nh_cap = p.hidden_capacity, n_cap = p.capacity;
	if (nh_cap == 0 || n_cap == 0) { APBF_CUDA(ctx, cudaMemsetAsync(nb->length, 0, 4, st)); return APBF_OK; }
	if (fuse_kw) APBF_REQUIRE(ctx, p.radius.data && fluid->target_radius.data && fluid->kernel_width.data);
	uint32_t* misc = ctx->misc();
	uint32_t* code = (uint32_t*)ctx->scratch_get(SLOT_SORT_KEYS_A, sizeof(uint32_t) * (size_t)nh_cap);
	uint32_t* scode = (uint32_t*)ctx->scratch_get(SLOT_TMP_KEYS, sizeof(uint32_t) * (size_t)nh_cap);
	uint32_t* idx_a = (uint32_t*)ctx->scratch_get(SLOT_TMP_VALS, sizeof(uint32_t) * (size_t)nh_cap);
	uint32_t* idx_b = (uint32_t*)ctx->scratch_get(SLOT_SORT_VALS_A, sizeof(uint32_t) * (size_t)nh_cap);
	uint32_t* c[3] = { (uint32_t*)ctx->scratch_get(SLOT_CODE0, sizeof(uint32_t) * (size_t)nh_cap),
	                   (uint32_t*)ctx->scratch_get(SLOT_CODE1, sizeof(uint32_t) * (size_t)nh_cap),
	                   (uint32_t*)ctx->scratch_get(SLOT_CODE2, sizeof(uint32_t) * (size_t)nh_cap) };
	uint32_t* counts = (uint32_t*)ctx->scratch_get(SLOT_COUNTS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	uint32_t* offsets = (uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	uint32_t* nbl = (uint32_t*)ctx->scratch_get(SLOT_NB, sizeof(uint32_t) * ((size_t)nb->capacity + 1));
	if (!code || !scode || !idx_a || !idx_b || !c[0] || !c[1] || !c[2] || !counts || !offsets || !nbl || !misc)
		return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);

	// three stable 32-bit sorts, least significant section first (neighborhood_binary_search.cpp:45-51)
	const uint32_t* cur = nullptr; // nullptr == identity payload
	uint32_t* ping[2] = { idx_a, idx_b };
	{
		apbf_prof_scope ps(ctx, PROF_HASH_SORT);
		for (uint32_t sec = 0; sec < 3u; sec++) {
			APBF_TRY(apbf_launch_position_code(ctx, cur, (const int32_t*)p.position.data, code, p.hidden_length, nh_cap, sec));
			uint32_t* dst = ping[sec & 1u];
			APBF_TRY(apbf_radix_sort_pairs(ctx, code, cur, scode, dst, p.hidden_length, nh_cap, 32));
			cur = dst;
		}
		if (mg) { // owned first, ghosts behind them (idx_b = SLOT_SORT_VALS_A holds the final order)
			k_ghost_key<<<apbf_grid(ctx, nh_cap, 256), 256, 0, st>>>(cur, p.hidden_length, misc, code);
			APBF_LAUNCHED(ctx);
			uint32_t* dst = ping[1];
			APBF_TRY(apbf_radix_sort_pairs(ctx, code, cur, scode, dst, p.hidden_length, nh_cap, 1));
			cur = dst;
		}
	}
	const uint32_t* sidx = cur;
	{
		apbf_prof_scope ps(ctx, PROF_REORDER);
		APBF_TRY(reorder_lists(ctx, fluid, range, sidx)); // :53-54
	}
	const uint32_t* new_index = (const uint32_t*)p.index_list.reorder_out;
	const int32_t* new_pos = (const int32_t*)p.position.reorder_out;
	const float* new_range = (const float*)range->reorder_out;
	for (uint32_t sec = 0; sec < 3u; sec++) // :58-60
		APBF_TRY(apbf_launch_position_code(ctx, new_index, new_pos, c[sec], p.length, n_cap, sec));
	k_clear_search_words<<<1, 1, 0, st>>>(misc);
	APBF_LAUNCHED(ctx);
	const unsigned grid = apbf_grid(ctx, n_cap, 128, 16);
	static const int two_pass_env = getenv("APBF_TWO_PASS_EMIT") ? 1 : 0; // debugging aid: per-thread count/fill instead of stream/regroup
	const int two_pass = two_pass_env && !fuse_kw && !mg;
	build_kw_args K;
	memset(&K, 0, sizeof K);
	uint32_t* kwfx = nullptr;
	if (fuse_kw) {
		K.i4 = (int4*)ctx->scratch_get(SLOT_I4, sizeof(int4) * (size_t)n_cap);
		K.cutoff = (float*)ctx->scratch_get(SLOT_CUTOFF, sizeof(float) * (size_t)n_cap);
		K.qb4 = (float4*)ctx->scratch_get(SLOT_QB4, sizeof(float4) * (size_t)n_cap);
		kwfx = (uint32_t*)ctx->scratch_get(SLOT_KWFX, sizeof(uint32_t) * (size_t)n_cap);
		if (!K.i4 || !K.cutoff || !K.qb4 || !kwfx) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
		K.radius = (const float*)p.radius.reorder_out;
		K.target_radius = (const float*)fluid->target_radius.reorder_out;
		K.kernel_width = (const float*)fluid->kernel_width.reorder_out;
		K.base_on_target_radius = ctx->settings.mBaseKernelWidthOnTargetRadius;
	}
	// one-pass emit (see k_green_stream): chunks of queries = runs of equal cells, candidates = the 27 code ranges around the cell
	float4* q4 = (float4*)ctx->scratch_get(SLOT_Q4, sizeof(float4) * (size_t)n_cap);
	uint32_t* key_id = (uint32_t*)ctx->scratch_get(SLOT_KEY_ID, sizeof(uint32_t) * (size_t)n_cap);
	uint32_t* head = (uint32_t*)ctx->scratch_get(SLOT_KEEP_COUNTS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	uint32_t stream_blocks = (uint32_t)std::min<size_t>((size_t)nb->capacity / SB_ENTRIES + (size_t)n_cap / 8u + 1024u, 0x7FFFFFFFu);
	if (ctx->stream_blocks_cap) stream_blocks = std::min(stream_blocks, std::max(ctx->stream_blocks_cap, 1u));
	uint32_t* stream = (uint32_t*)ctx->scratch_get(SLOT_STREAM, sizeof(uint32_t) * (size_t)stream_blocks * SB_WORDS);
	if (!q4 || !key_id || !head || !stream) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	{
		apbf_prof_scope ps(ctx, PROF_EMIT_COUNT);
		if (two_pass) {
			k_bsearch_emit<false><<<grid, 128, 0, st>>>(new_index, new_pos, c[0], c[1], c[2], new_range, p.length, range_scale, counts,
			                                            nullptr, nullptr, 0u, nullptr, misc, 0);
			APBF_LAUNCHED(ctx);
		} else {
			if (fuse_kw) {
				k_max_abs_coord<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(new_index, new_pos, p.length, misc);
				APBF_LAUNCHED(ctx);
			}
			k_build_bq<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(new_index, new_pos, new_range, range_scale, p.length, q4, head, misc, K);
			APBF_LAUNCHED(ctx);
			APBF_TRY(apbf_scan_u32(ctx, head, key_id, p.length, n_cap, true, nullptr, 0xFFFFFFFFu, nullptr, nullptr));
			emit_args A;
			memset(&A, 0, sizeof A);
			A.q4 = q4; A.key_id = key_id; A.range = new_range; A.len = p.length; A.range_scale = range_scale; A.counts = counts;
			A.offsets = offsets; A.pairs = nb->pairs; A.nbl = nbl; A.cap = nb->capacity; A.misc = misc; A.table_cells = 1u;
			A.layers = !mg ? 1u : (ctx->mg_ghost_all_pairs && !fuse_kw ? 3u : 2u);
			A.g.ext[0] = A.g.ext[1] = A.g.ext[2] = 1.0f; A.g.scale = 1.0f; A.g.res = 1u; A.g.dims = 3; // (no grid in this search)
			A.stream = stream; A.stream_blocks = stream_blocks; A.ticket = misc + MW_EMIT_TICKET0;
			A.c0 = c[0]; A.c1 = c[1]; A.c2 = c[2]; A.pos4 = new_pos; A.index_list = new_index;
			A.i4 = K.i4; A.cutoff = K.cutoff; A.qb4 = K.qb4; A.kwfx = kwfx;
			const unsigned sgrid = apbf_grid(ctx, n_cap, EMIT_WARPS * 32, 4);
			if (mg && !fuse_kw) k_green_stream<EMIT_MG, 3, false, 1><<<sgrid, EMIT_WARPS * 32, 0, st>>>(A);
			else if (mg && ctx->search_stats) k_green_stream<EMIT_FUSED_MG, 3, true, 1><<<sgrid, EMIT_WARPS * 32, 0, st>>>(A);
			else if (mg) k_green_stream<EMIT_FUSED_MG, 3, false, 1><<<sgrid, EMIT_WARPS * 32, 0, st>>>(A);
			else if (!fuse_kw) k_green_stream<EMIT_PLAIN, 3, false, 1><<<sgrid, EMIT_WARPS * 32, 0, st>>>(A);
			else if (ctx->search_stats) k_green_stream<EMIT_FUSED, 3, true, 1><<<sgrid, EMIT_WARPS * 32, 0, st>>>(A);
			else k_green_stream<EMIT_FUSED, 3, false, 1><<<sgrid, EMIT_WARPS * 32, 0, st>>>(A);
			APBF_LAUNCHED(ctx);
		}
	}
	{
		apbf_prof_scope ps(ctx, PROF_EMIT_SCAN);
		APBF_TRY(apbf_scan_u32(ctx, counts, offsets, p.length, n_cap, false, nb->length, nb->capacity, misc + MW_FLAGS,
		                       misc + (fuse_kw ? MW_KEPT_PAIRS : MW_TOTAL_PAIRS)));
		k_offsets_tail<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(offsets, p.length, n_cap);
		APBF_LAUNCHED(ctx);
	}
	{
		apbf_prof_scope ps(ctx, PROF_EMIT_FILL);
		if (!two_pass) {
			k_regroup<<<ctx->num_sms * 16, SB_ENTRIES, 0, st>>>(stream, stream_blocks, offsets, nullptr, nbl, nb->capacity, misc, fuse_kw || mg ? 1 : 0);
			APBF_LAUNCHED(ctx);
			if (write_public) APBF_TRY(apbf_launch_expand_pairs(ctx, offsets, nbl, p.length, n_cap, nb->pairs, nb->capacity));
		}
		if (!fuse_kw && !mg) { // (the fused form and slabs have no two-pass fill behind them: a stream overflow raises the sticky flag)
			k_bsearch_emit<true><<<grid, 128, 0, st>>>(new_index, new_pos, c[0], c[1], c[2], new_range, p.length, range_scale, nullptr,
			                                           offsets, nb->pairs, nb->capacity, nbl, misc, two_pass ? 0 : 1);
			APBF_LAUNCHED(ctx);
		}
	}
	apbf_nbr_built(ctx, nb, n_cap, write_public || two_pass);
	if (dbg) {
		if (dbg->sorted_index) APBF_CUDA(ctx, cudaMemcpyAsync(dbg->sorted_index, sidx, sizeof(uint32_t) * (size_t)nh_cap, cudaMemcpyDeviceToDevice, st));
		for (int s = 0; s < 3; s++)
			if (dbg->code[s]) APBF_CUDA(ctx, cudaMemcpyAsync(dbg->code[s], c[s], sizeof(uint32_t) * (size_t)n_cap, cudaMemcpyDeviceToDevice, st));
		if (dbg->pair_offsets) APBF_CUDA(ctx, cudaMemcpyAsync(dbg->pair_offsets, offsets, sizeof(uint32_t) * (size_t)(n_cap + 1), cudaMemcpyDeviceToDevice, st));
	}
	if (fuse_kw) { // the reordered lists are still in reorder_out (the caller swaps after the search)
		apbf_fluid sorted = *fluid;
		sorted.kernel_width.data = fluid->kernel_width.reorder_out;
		apbf_prof_scope ps(ctx, PROF_KW_MISC);
		APBF_TRY(apbf_kw_finish(ctx, &sorted, out_kw_fixed));
	}
	return APBF_OK;
}

extern "C" {

int apbf_neighborhood_binary_search_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_array* range, apbf_neighbors* nb,
                                          float range_scale, const apbf_search_debug* dbg)
{
	return apbf_binary_search(ctx, fluid, range, nb, range_scale, dbg, false, nullptr, true);
}

int apbf_neighborhood_binary_search_spread_apply(apbf_ctx* ctx, apbf_fluid* fluid, apbf_neighbors* nb, float range_scale,
                                                 const apbf_search_debug* dbg, uint32_t* out_kw_fixed)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid);
	return apbf_binary_search(ctx, fluid, &fluid->kernel_width, nb, range_scale, dbg, true, out_kw_fixed, true);
}

} // extern "C"
