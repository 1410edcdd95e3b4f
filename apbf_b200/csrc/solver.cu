// solver.cu -- adaptive kernel width update, stand-alone box collision and the integrator.
//
// Replaces source/spread_kernel_width.cpp:12-26 (kernel_width_init.comp, kernel_width.comp,
// uint_to_float_but_gradual.comp), source/box_collision.cpp:12-24 (box_collision.comp) and
// source/velocity_handling.cpp:15-31 (infer_velocity / apply_acceleration / apply_velocity).
// The incompressibility sweeps live in incompress.cu.
#include "box.cuh"
#include "neighbors.cuh"
#include "sort.cuh"
#include <utility>

namespace {

constexpr int GROUP = 8;            // lanes per particle in the segment sweeps
constexpr int SWEEP_THREADS = 256;  // 32 particles per CTA
#define R_KW APBF_KERNEL_WIDTH_RESOLUTION

// Groups of one warp run different trip counts, so every warp intrinsic names only the lanes of its own group.
__device__ __forceinline__ uint32_t group_mask() { return ((1u << GROUP) - 1u) << ((threadIdx.x & 31u) & ~(GROUP - 1u)); }
__device__ __forceinline__ int group_sum(int v)
{
	const uint32_t m = group_mask();
#pragma unroll
	for (int o = GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o, GROUP);
	return v;
}
__device__ __forceinline__ uint32_t group_max(uint32_t v)
{
	const uint32_t m = group_mask();
#pragma unroll
	for (int o = GROUP / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(m, v, o, GROUP));
	return v;
}

// ---- spread_kernel_width -------------------------------------------------------------------------------------------------
struct kw_args {
	const uint32_t* index_list;
	const uint32_t* len;
	const int32_t*  pos4;
	const float*    radius;        // hidden
	const float*    target_radius; // per id
	float*          kernel_width;  // per id (old on input)
	uint32_t*       kwfx;          // per id
	const uint32_t* nbl;           // idN | (unmirrored << 31) per pair
	const uint32_t* offsets;
	uint32_t        pair_cap;
	uint32_t*       keep_counts;
	const uint32_t* keep_offsets;
	uint32_t*       out_pairs;
	uint32_t*       out_nbl;
	uint32_t*       misc;
	int             base_on_target_radius;
	float           speed;
};

__device__ __forceinline__ float kw_original(const kw_args& A, uint32_t id, uint32_t idx)
{ // kernel_width_init.comp:28-34 / kernel_width.comp:40-46
	float targetRadius = A.target_radius[id];
	if (!A.base_on_target_radius) targetRadius = 0.0f;
	return glsl_max(A.radius[idx], targetRadius) * APBF_KERNEL_SCALE;
}

__global__ void k_kw_init(kw_args A)
{
	const uint32_t n = *A.len;
	const bool ident = A.misc[MW_IDENTITY] != 0u;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const uint32_t idx = ident ? id : A.index_list[id];
		A.kwfx[id] = f2u(kw_original(A, id, idx) * R_KW);
	}
}

__device__ __forceinline__ uint32_t kw_influence(float orig, float dist) { return apbf_kw_influence(orig, dist); }

// COMPACT == false: atomicMax spread (gathered through mirrored pairs) + number of pairs to keep per particle
// COMPACT == true : write the kept pairs at the scanned offsets with their new mirrored bits
template <bool COMPACT>
__global__ void __launch_bounds__(SWEEP_THREADS) k_kw_sweep(kw_args A)
{
	const uint32_t n = *A.len;
	const uint32_t n_owned = A.misc[MW_N_OWNED]; // multi-GPU: a ghost keeps only the pairs that end up unmirrored (its scatter part)
	const bool ident = A.misc[MW_IDENTITY] != 0u;
	const unsigned sub = threadIdx.x & (GROUP - 1);
	const unsigned group_shift = (threadIdx.x & 31u) & ~(GROUP - 1u);
	const uint32_t groups_per_grid = gridDim.x * (SWEEP_THREADS / GROUP);
	for (uint32_t a = blockIdx.x * (SWEEP_THREADS / GROUP) + threadIdx.x / GROUP; a < n; a += groups_per_grid) {
		const uint32_t idx = ident ? a : A.index_list[a];
		const int4 ip = __ldg((const int4*)A.pos4 + idx);
		const float orig_a = kw_original(A, a, idx);
		const float cutoff_a = glsl_max(orig_a, A.kernel_width[a]);
		uint32_t mx = 0u, kept = 0u, n_asym = 0u;
		uint32_t out = COMPACT ? A.keep_offsets[a] : 0u;
		const uint32_t beg = min(A.offsets[a], A.pair_cap), end = min(A.offsets[a + 1], A.pair_cap);
		for (uint32_t e0 = beg; e0 < end; e0 += GROUP) {
			const uint32_t e = e0 + sub;
			bool keep = false, mirrored = false, new_mirrored = false;
			uint32_t b = 0;
			if (e < end) {
				const uint32_t nbe = A.nbl[e];
				b = nbe & NB_ID_MASK;
				const uint32_t idxN = ident ? b : A.index_list[b];
				const int4 iq = __ldg((const int4*)A.pos4 + idxN);
				const float rx = (float)(iq.x - ip.x) * INV_R_POS, ry = (float)(iq.y - ip.y) * INV_R_POS, rz = (float)(iq.z - ip.z) * INV_R_POS;
				const float dist = sqrtf(dot3(rx, ry, rz, rx, ry, rz));
				mirrored = (nbe & NB_UNMIRRORED) == 0u;
				keep = dist <= cutoff_a; // kernel_width.comp:57
				if (!COMPACT) {
					if (mirrored) mx = max(mx, kw_influence(kw_original(A, b, idxN), dist)); // the pair (b, a) spreads b's width onto a
					else atomicMax(A.kwfx + b, kw_influence(orig_a, dist));                   // (a, b) has no mirror: spread directly (:53)
				}
				if (keep) new_mirrored = mirrored && dist <= glsl_max(kw_original(A, b, idxN), A.kernel_width[b]);
				if (a >= n_owned) keep = keep && !new_mirrored;
			}
			if (!COMPACT) {
				kept += keep ? 1u : 0u;
			} else { // stable compaction inside the group: ballot of this 8-lane slice
				const uint32_t ballot = (__ballot_sync(group_mask(), keep) >> group_shift) & ((1u << GROUP) - 1u);
				const uint32_t rank = __popc(ballot & ((1u << sub) - 1u));
				if (keep) {
					const uint32_t o = out + rank;
					*(uint2*)(A.out_pairs + 2 * (size_t)o) = make_uint2(a, b);
					A.out_nbl[o] = b | (new_mirrored ? 0u : NB_UNMIRRORED);
					n_asym += new_mirrored ? 0u : 1u;
				}
				out += __popc(ballot);
			}
		}
		if (!COMPACT) {
			mx = group_max(mx);
			kept = (uint32_t)group_sum((int)kept);
			if (sub == 0) {
				if (mx) atomicMax(A.kwfx + a, mx);
				A.keep_counts[a] = kept;
			}
		} else {
			n_asym = (uint32_t)group_sum((int)n_asym);
			if (sub == 0 && n_asym) atomicAdd(A.misc + MW_N_ASYM, n_asym);
		}
	}
}

__global__ void k_kw_finish(kw_args A)
{ // uint_to_float_but_gradual.comp:21-39 with mFactor = 1/KERNEL_WIDTH_RESOLUTION, mLowerBound = -inf
	const uint32_t n = *A.len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const float newValue = (float)A.kwfx[id] * (1.0f / R_KW);
		const float oldValue = A.kernel_width[id];
		const float d = newValue - oldValue;
		const float dir = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
		const float result = oldValue * (1.0f + A.speed * dir);
		const bool reached = (oldValue < newValue) != (result < newValue);
		A.kernel_width[id] = glsl_max(reached ? newValue : result, -INFINITY);
	}
}

// ---- update_transfers (merge and split off): find_split_and_merge_1/2/3.comp as one gather over the grouped list -----------
// The reference scatters with atomicMin over the pair list (one thread per pair) and picks the nearest neighbour in a second
// pass over all pairs (every pair at the minimum distance writes; which one stays is a race).  Here GROUP lanes walk the
// particle's own segment: both minima are plain reductions, and among several pairs at the minimum distance the last one in
// list order wins (what a sequential run of the reference's second pass leaves behind).
struct ut_args {
	const uint32_t* index_list;
	const uint32_t* len;
	const int32_t*  pos4;
	const float*    radius;            // hidden
	const uint32_t* old_boundary_dist; // per id: copy taken before the pass (update_transfers.cpp:32)
	uint32_t*       boundary_dist;     // per id
	float*          target_radius;     // per id
	float*          boundariness;      // per id
	uint32_t*       nearest;           // per id, optional
	const uint32_t* nbl;
	const uint32_t* offsets;
	uint32_t        pair_cap;
	const uint32_t* misc;
	apbf_settings   s;
};

__global__ void __launch_bounds__(SWEEP_THREADS) k_update_transfers(ut_args A)
{
	const uint32_t n = min(*A.len, A.misc[MW_N_OWNED]); // slabs: a ghost's values come from its owner (its own list is incomplete)
	const bool ident = A.misc[MW_IDENTITY] != 0u;
	const unsigned sub = threadIdx.x & (GROUP - 1);
	const uint32_t gm = group_mask();
	const uint32_t groups_per_grid = gridDim.x * (SWEEP_THREADS / GROUP);
	for (uint32_t a = blockIdx.x * (SWEEP_THREADS / GROUP) + threadIdx.x / GROUP; a < n; a += groups_per_grid) {
		const uint32_t idx = ident ? a : A.index_list[a];
		const int4 ip = __ldg((const int4*)A.pos4 + idx);
		uint32_t bd = 0xFFFFFFFFu;                      // write_sequence(..., max, 0), update_transfers.cpp:39
		unsigned long long best = 0xFFFFFFFFFFFFFFFFull; // (distance, ~pair index): smallest distance, then the latest pair
		const uint32_t beg = min(A.offsets[a], A.pair_cap), end = min(A.offsets[a + 1], A.pair_cap);
		for (uint32_t e = beg + sub; e < end; e += GROUP) {
			const uint32_t b = A.nbl[e] & NB_ID_MASK;
			const uint32_t idxN = ident ? b : A.index_list[b];
			const int4 iq = __ldg((const int4*)A.pos4 + idxN);
			// uint(length(posN - pos)): integer difference, length in float (find_split_and_merge_1.comp:30-31)
			const float fx = (float)(iq.x - ip.x), fy = (float)(iq.y - ip.y), fz = (float)(iq.z - ip.z);
			const uint32_t dist = f2u(sqrtf(dot3(fx, fy, fz, fx, fy, fz)));
			bd = min(bd, A.old_boundary_dist[b] + dist); // uint arithmetic wraps like the shader's (:33)
			best = min(best, ((unsigned long long)dist << 32) | (unsigned long long)(0xFFFFFFFFu - e));
		}
#pragma unroll
		for (int o = GROUP / 2; o > 0; o >>= 1) {
			bd = min(bd, __shfl_xor_sync(gm, bd, o, GROUP));
			best = min(best, __shfl_xor_sync(gm, best, o, GROUP));
		}
		if (sub != 0) continue;
		if (A.nearest) A.nearest[a] = beg < end ? A.nbl[0xFFFFFFFFu - (uint32_t)best] & NB_ID_MASK : 0xFFFFFFFFu;
		// find_split_and_merge_3.comp:55-84
		const float radius = A.radius[idx];
		const float boundaryDistance = (float)bd / R_POS;
		float targetRadius;
		if (A.s.mUpdateTargetRadius) {
			if (A.s.mBaseKernelWidthOnBoundaryDistance) {
				targetRadius = (A.s.mTargetRadiusScaleFactor / (APBF_KERNEL_SCALE + APBF_KERNEL_SCALE * A.s.mTargetRadiusScaleFactor)) * boundaryDistance;
				targetRadius = glsl_max(targetRadius, A.s.mSmallestTargetRadius);
			} else {
				targetRadius = A.s.mSmallestTargetRadius + glsl_max(0.0f, (boundaryDistance - A.s.mTargetRadiusOffset) * A.s.mTargetRadiusScaleFactor);
			}
			A.target_radius[a] = targetRadius;
		}
		const float boundariness = A.boundariness[a] >= 1.0f ? 1.0f : 0.0f;
		// mix(boundaryDistance, radius, boundariness): boundary distance "decay" (:79)
		A.boundary_dist[a] = f2u((boundaryDistance * (1.0f - boundariness) + radius * boundariness) * R_POS);
		A.boundariness[a] = boundariness;
	}
}

// pool.cpp:77-80 -> uint_to_float_with_indexed_lower_bound.comp:32-44
__global__ void k_kw_from_boundary_distance(const uint32_t* __restrict__ index_list, const uint32_t* __restrict__ len,
                                            const uint32_t* __restrict__ boundary_dist, const float* __restrict__ radius,
                                            float* __restrict__ kernel_width, float factor, float lower_bound_factor, float speed)
{
	const uint32_t n = *len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const uint32_t idx = index_list[id]; // (runs before the search: no cached knowledge about the index list)
		const float newValue = (float)boundary_dist[id] * factor;
		const float lowerBound = radius[idx] * lower_bound_factor;
		const float oldValue = kernel_width[id];
		const float d = newValue - oldValue;
		const float dir = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
		const float result = oldValue * (1.0f + speed * dir);
		const bool reached = (oldValue < newValue) != (result < newValue);
		kernel_width[id] = glsl_max(reached ? newValue : result, lowerBound);
	}
}

__global__ void k_reset_asym(uint32_t* misc) { misc[MW_N_ASYM] = 0u; }

// dst[0 .. *len) = src[0 .. *len), 8-byte elements (the kept pairs go back into the caller's list)
__global__ void k_copy_pairs_len(const uint2* __restrict__ src, uint2* __restrict__ dst, const uint32_t* __restrict__ len)
{
	const uint32_t n = *len;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// ---- box collision (box_collision.comp:36-60) ------------------------------------------------------------------------------
__global__ void k_box_collision(const uint32_t* __restrict__ index_list, int32_t* pos4, const float* __restrict__ radius,
                                const float4* __restrict__ bmin, const float4* __restrict__ bmax, const uint32_t* __restrict__ len,
                                uint32_t n_boxes)
{
	const uint32_t n = *len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const uint32_t idx = index_list[id];
		int4 ip = *((int4*)pos4 + idx);
		float px = (float)ip.x * INV_R_POS, py = (float)ip.y * INV_R_POS, pz = (float)ip.z * INV_R_POS;
		box_push(px, py, pz, id, radius[idx], bmin, bmax, n_boxes);
		ip.x = f2i(px * R_POS); ip.y = f2i(py * R_POS); ip.z = f2i(pz * R_POS); // always written, box_collision.comp:59
		*((int4*)pos4 + idx) = ip;
	}
}

// ---- velocity handling (velocity_handling.cpp:15-31) -----------------------------------------------------------------------
__global__ void k_velocity_handling(const uint32_t* __restrict__ index_list, int32_t* pos4, float* vel4, int32_t* backup4,
                                    const uint32_t* __restrict__ len, float infer_div, float ax, float ay, float az, float step_mul)
{
	const uint32_t n = *len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const uint32_t idx = index_list[id];
		int4 p = *((int4*)pos4 + idx);
		const int4 o = *((int4*)backup4 + idx);
		float4 v = *((float4*)vel4 + idx);
		v.x = (float)(p.x - o.x) / infer_div; v.y = (float)(p.y - o.y) / infer_div; v.z = (float)(p.z - o.z) / infer_div; // infer_velocity.comp:31
		v.x += ax; v.y += ay; v.z += az;                                                                                   // apply_acceleration.comp:27
		*((int4*)backup4 + idx) = p; // posBackupList = positionList for the particles of this list
		p.x += f2i(v.x * step_mul); p.y += f2i(v.y * step_mul); p.z += f2i(v.z * step_mul);                                // apply_velocity.comp:30
		*((float4*)vel4 + idx) = v;
		*((int4*)pos4 + idx) = p;
	}
}

} // namespace

int apbf_kw_finish(apbf_ctx* ctx, apbf_fluid* fluid, uint32_t* out_kw_fixed)
{
	const uint32_t n_cap = fluid->particle.capacity;
	kw_args A;
	memset(&A, 0, sizeof A);
	A.len = fluid->particle.length;
	A.kernel_width = (float*)fluid->kernel_width.data;
	A.kwfx = (uint32_t*)ctx->scratch_get(SLOT_KWFX, sizeof(uint32_t) * (size_t)n_cap);
	A.speed = ctx->settings.mKernelWidthAdaptionSpeed;
	if (!A.kwfx) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	if (out_kw_fixed) APBF_CUDA(ctx, cudaMemcpyAsync(out_kw_fixed, A.kwfx, sizeof(uint32_t) * (size_t)n_cap, cudaMemcpyDeviceToDevice, ctx->stream));
	k_kw_finish<<<apbf_grid(ctx, n_cap, 256), 256, 0, ctx->stream>>>(A);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

extern "C" {

int apbf_spread_kernel_width_apply(apbf_ctx* ctx, apbf_fluid* fluid, apbf_neighbors* nb, uint32_t* out_kw_fixed)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid && nb);
	apbf_particles& p = fluid->particle;
	APBF_REQUIRE(ctx, p.index_list.data && p.length && p.position.data && p.radius.data && fluid->kernel_width.data && fluid->target_radius.data);
	const uint32_t n_cap = p.capacity;
	if (n_cap == 0) return APBF_OK;
	APBF_TRY(apbf_nbr_ensure(ctx, fluid, nb)); // any pair list (spread_kernel_width.h:10): foreign ones get their structure built here
	cudaStream_t st = ctx->stream;
	kw_args A;
	memset(&A, 0, sizeof A);
	A.index_list = (const uint32_t*)p.index_list.data;
	A.len = p.length;
	A.pos4 = (const int32_t*)p.position.data;
	A.radius = (const float*)p.radius.data;
	A.target_radius = (const float*)fluid->target_radius.data;
	A.kernel_width = (float*)fluid->kernel_width.data;
	A.kwfx = (uint32_t*)ctx->scratch_get(SLOT_KWFX, sizeof(uint32_t) * (size_t)n_cap);
	A.pair_cap = nb->capacity;
	uint32_t* offsets = (uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	A.offsets = offsets;
	uint32_t* nbl = (uint32_t*)ctx->scratch_get(SLOT_NB, sizeof(uint32_t) * ((size_t)nb->capacity + 1));
	A.nbl = nbl;
	A.keep_counts = (uint32_t*)ctx->scratch_get(SLOT_KEEP_COUNTS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	uint32_t* keep_offsets = (uint32_t*)ctx->scratch_get(SLOT_KEEP_OFFSETS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	A.keep_offsets = keep_offsets;
	A.out_pairs = (uint32_t*)ctx->scratch_get(SLOT_PAIRS_TMP, sizeof(uint32_t) * 2 * (size_t)nb->capacity);
	A.out_nbl = (uint32_t*)ctx->scratch_get(SLOT_NB_TMP, sizeof(uint32_t) * ((size_t)nb->capacity + 1));
	A.misc = ctx->misc();
	A.base_on_target_radius = ctx->settings.mBaseKernelWidthOnTargetRadius;
	A.speed = ctx->settings.mKernelWidthAdaptionSpeed;
	if (!A.kwfx || !offsets || !nbl || !A.keep_counts || !keep_offsets || !A.out_pairs || !A.out_nbl)
		return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	const unsigned egrid = apbf_grid(ctx, n_cap, 256);
	const unsigned sgrid = apbf_grid(ctx, (size_t)n_cap * GROUP, SWEEP_THREADS, 8);
	{
		apbf_prof_scope ps(ctx, PROF_KW_MISC);
		k_kw_init<<<egrid, 256, 0, st>>>(A);
		APBF_LAUNCHED(ctx);
	}
	{
		apbf_prof_scope ps(ctx, PROF_KW_SPREAD);
		k_kw_sweep<false><<<sgrid, SWEEP_THREADS, 0, st>>>(A);
		APBF_LAUNCHED(ctx);
	}
	apbf_prof_scope ps_compact(ctx, PROF_KW_COMPACT);
	// prune: scan the kept counts, compact into scratch, then the scratch list becomes the neighbour list
	APBF_TRY(apbf_scan_u32(ctx, A.keep_counts, keep_offsets, p.length, n_cap, false, nb->length, nb->capacity, nullptr, A.misc + MW_KEPT_PAIRS));
	k_reset_asym<<<1, 1, 0, st>>>(A.misc);
	APBF_LAUNCHED(ctx);
	k_kw_sweep<true><<<sgrid, SWEEP_THREADS, 0, st>>>(A);
	APBF_LAUNCHED(ctx);
	if (out_kw_fixed) APBF_CUDA(ctx, cudaMemcpyAsync(out_kw_fixed, A.kwfx, sizeof(uint32_t) * (size_t)n_cap, cudaMemcpyDeviceToDevice, st));
	k_kw_finish<<<egrid, 256, 0, st>>>(A);
	APBF_LAUNCHED(ctx);
	// the pruned list replaces the neighbour list: pairs go back into the caller's buffer (only the kept ones are
	// copied), the solver's own structure just swaps scratch buffers
	k_copy_pairs_len<<<apbf_grid(ctx, nb->capacity, 256, 16), 256, 0, st>>>((const uint2*)A.out_pairs, (uint2*)nb->pairs, nb->length);
	APBF_LAUNCHED(ctx);
	std::swap(ctx->scratch[SLOT_NB], ctx->scratch[SLOT_NB_TMP]);
	std::swap(ctx->scratch[SLOT_OFFSETS], ctx->scratch[SLOT_KEEP_OFFSETS]);
	apbf_nbr_built(ctx, nb, n_cap, true);
	return APBF_OK;
}

int apbf_update_transfers_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, uint32_t* out_nearest_neighbor)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid && nb);
	apbf_particles& p = fluid->particle;
	APBF_REQUIRE(ctx, p.index_list.data && p.length && p.position.data && p.radius.data && fluid->boundary_distance.data &&
	                      fluid->target_radius.data && fluid->boundariness.data);
	const uint32_t n_cap = p.capacity;
	if (n_cap == 0) return APBF_OK;
	APBF_TRY(apbf_nbr_ensure(ctx, fluid, nb)); // any pair list (update_transfers.h)
	cudaStream_t st = ctx->stream;
	apbf_prof_scope ps(ctx, PROF_UPDATE_TRANSFERS);
	ut_args A;
	memset(&A, 0, sizeof A);
	uint32_t* old_bd = (uint32_t*)ctx->scratch_get(SLOT_OLD_BOUNDARY_DIST, sizeof(uint32_t) * (size_t)n_cap);
	A.offsets = (const uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	A.nbl = (const uint32_t*)ctx->scratch_get(SLOT_NB, sizeof(uint32_t) * ((size_t)nb->capacity + 1));
	if (!old_bd || !A.offsets || !A.nbl) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	// auto oldBoundaryDistanceList = boundaryDistanceList (update_transfers.cpp:32)
	APBF_CUDA(ctx, cudaMemcpyAsync(old_bd, fluid->boundary_distance.data, sizeof(uint32_t) * (size_t)n_cap, cudaMemcpyDeviceToDevice, st));
	A.index_list = (const uint32_t*)p.index_list.data;
	A.len = p.length;
	A.pos4 = (const int32_t*)p.position.data;
	A.radius = (const float*)p.radius.data;
	A.old_boundary_dist = old_bd;
	A.boundary_dist = (uint32_t*)fluid->boundary_distance.data;
	A.target_radius = (float*)fluid->target_radius.data;
	A.boundariness = (float*)fluid->boundariness.data;
	A.nearest = out_nearest_neighbor;
	A.pair_cap = nb->capacity;
	A.misc = ctx->misc();
	A.s = ctx->settings;
	k_update_transfers<<<apbf_grid(ctx, (size_t)n_cap * GROUP, SWEEP_THREADS, 64), SWEEP_THREADS, 0, st>>>(A);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_uint_to_float_with_indexed_lower_bound(apbf_ctx* ctx, const uint32_t* in_uint, float* out_float, const uint32_t* index_list,
                                                const float* lower_bound, const uint32_t* len, uint32_t capacity, float factor,
                                                float lower_bound_factor, float max_adaption_step)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, in_uint && out_float && index_list && lower_bound && len);
	if (capacity == 0) return APBF_OK;
	apbf_prof_scope ps(ctx, PROF_KW_MISC);
	k_kw_from_boundary_distance<<<apbf_grid(ctx, capacity, 256), 256, 0, ctx->stream>>>(index_list, len, in_uint, lower_bound, out_float, factor,
	                                                                                  lower_bound_factor, max_adaption_step);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_kernel_width_from_boundary_distance(apbf_ctx* ctx, apbf_fluid* fluid)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid);
	apbf_particles& p = fluid->particle;
	APBF_REQUIRE(ctx, p.index_list.data && p.length && p.radius.data && fluid->boundary_distance.data && fluid->kernel_width.data);
	if (p.capacity == 0) return APBF_OK;
	apbf_prof_scope ps(ctx, PROF_KW_MISC);
	k_kw_from_boundary_distance<<<apbf_grid(ctx, p.capacity, 256), 256, 0, ctx->stream>>>(
	    (const uint32_t*)p.index_list.data, p.length, (const uint32_t*)fluid->boundary_distance.data, (const float*)p.radius.data,
	    (float*)fluid->kernel_width.data, ctx->settings.mTargetRadiusScaleFactor / APBF_POS_RESOLUTION, APBF_KERNEL_SCALE,
	    ctx->settings.mKernelWidthAdaptionSpeed);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_box_collision_apply(apbf_ctx* ctx, apbf_particles* particles, const float* box_min4, const float* box_max4,
                             uint32_t n_boxes)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, particles && particles->index_list.data && particles->length && particles->position.data && particles->radius.data);
	APBF_REQUIRE(ctx, n_boxes == 0 || (box_min4 && box_max4));
	if (particles->capacity == 0) return APBF_OK;
	apbf_prof_scope ps(ctx, PROF_BOX);
	k_box_collision<<<apbf_grid(ctx, particles->capacity, 256), 256, 0, ctx->stream>>>(
	    (const uint32_t*)particles->index_list.data, (int32_t*)particles->position.data, (const float*)particles->radius.data,
	    (const float4*)box_min4, (const float4*)box_max4, particles->length, n_boxes);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_velocity_handling_apply(apbf_ctx* ctx, apbf_particles* particles, float dt, float last_dt, const float accel[3])
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, particles && accel && particles->index_list.data && particles->length && particles->position.data &&
	                      particles->velocity.data && particles->pos_backup.data);
	if (particles->capacity == 0) return APBF_OK;
	if (dt == 0.0f) { // velocity_handling.cpp:21-24: posBackupList = positionList
		APBF_CUDA(ctx, cudaMemcpyAsync(particles->pos_backup.data, particles->position.data, 16 * (size_t)particles->hidden_capacity,
		                               cudaMemcpyDeviceToDevice, ctx->stream));
		return APBF_OK;
	}
	apbf_prof_scope ps(ctx, PROF_VELOCITY);
	k_velocity_handling<<<apbf_grid(ctx, particles->capacity, 256), 256, 0, ctx->stream>>>(
	    (const uint32_t*)particles->index_list.data, (int32_t*)particles->position.data, (float*)particles->velocity.data,
	    (int32_t*)particles->pos_backup.data, particles->length, last_dt * APBF_POS_RESOLUTION, accel[0] * dt, accel[1] * dt,
	    accel[2] * dt, dt * APBF_POS_RESOLUTION);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

} // extern "C"
