// transfers.cu -- split / merge proper: the merge and split bookkeeping of update_transfers::apply
// (source/update_transfers.cpp:48-70, find_split_and_merge_3.comp:86-121, remove_impossible_splits.comp,
// initialize_split_particles.comp, indexed_list::duplicate_these) and particle_transfer::apply (source/particle_transfer.cpp,
// particle_transfer.comp, indexed_list::delete_these).
//
// The reference resolves conflicting merge / split candidates with atomicExchange across threads and appends to its lists
// with atomicAdd: the winner and the list order are races.  Here the result is the one a run of the invocations in ascending
// id order produces (the oracle's serialisation), computed in parallel:
//   candidates   one pass over the ids: kind (merge / split), source and target hidden index
//   compaction   candidates in id order (scan)
//   matching     "lowest id wins": a candidate is accepted once it is the lowest live candidate on each of its particles;
//                candidates on a taken particle die; repeat until nobody is live.  This is the greedy matching in id order.
//                With many candidates the first rounds run grid-wide (three launches per round), the rest in one CTA
//                (rounds separated by __syncthreads()); the work is proportional to the number of candidates.
//   ranks        accepted merges / splits get their row in id order; a full transfer list rejects the later merges, which
//                frees their particles for later splits (find_split_and_merge_3.comp:104-109)
//   emit         rows, duplicates (appended to the hidden list, the index list and the per-id lists) in parallel
// List edits keep the order of the surviving entries (stable compaction by scan), copies go to the end.
#include "common.cuh"
#include "neighbors.cuh"
#include "sort.cuh"

namespace {

constexpr uint32_t KIND_MERGE = 0x80000000u;
constexpr uint32_t IDX_MASK = 0x7FFFFFFFu;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int MATCH_THREADS = 1024;

enum tm_word { TW_M = 0, TW_N_MERGE, TW_N_SPLIT, TW_TLEN0, TW_NEW_NH, TW_NEW_N, TW_NEW_ROWS, TW_COUNT = 16 };

struct cand_args {
	const uint32_t* index_list;
	const uint32_t* len;
	const int32_t*  pos4;
	const float*    radius;        // hidden
	const float*    target_radius; // per id (already updated by find_split_and_merge_3's first half)
	const uint32_t* nearest;       // per id, NONE without pairs
	uint32_t*       flags;         // per id: 1 if the id is a candidate
	uint32_t*       src;           // per id: source hidden index | KIND_MERGE
	uint32_t*       tgt;           // per id
	apbf_settings   s;
	float           dims;
	float           split_factor;  // float(pow(2.0, 1.0 / DIMENSIONS) * 0.99), folded in double like the shader compiler does
};

// find_split_and_merge_3.comp:57-93
__global__ void __launch_bounds__(256) k_tm_candidates(cand_args A)
{
	const uint32_t n = *A.len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const uint32_t idx = A.index_list[id];
		const float radius = A.radius[idx];
		const float targetRadius = A.target_radius[id];
		const bool split = A.s.mSplit && (targetRadius * A.split_factor <= radius);
		bool merge = false, nnLarger = false;
		uint32_t nnIdx = 0;
		const uint32_t nn = A.nearest[id];
		if (A.s.mMerge && nn != NONE) {
			nnIdx = A.index_list[nn];
			const int4 p = __ldg((const int4*)A.pos4 + idx), q = __ldg((const int4*)A.pos4 + nnIdx);
			const float fx = (float)(q.x - p.x), fy = (float)(q.y - p.y), fz = (float)(q.z - p.z);
			const float nnDist = sqrtf(dot3(fx, fy, fz, fx, fy, fz)) / R_POS;
			const float nnRadius = A.radius[nnIdx];
			nnLarger = nnRadius > radius || (nnRadius == radius && nnIdx > idx);
			merge = nnDist < radius && pow_rn(radius, A.dims) + pow_rn(nnRadius, A.dims) <= pow_rn(targetRadius, A.dims);
		}
		uint32_t f = 0u;
		if (merge) {
			A.src[id] = (nnLarger ? nnIdx : idx) | KIND_MERGE;
			A.tgt[id] = nnLarger ? idx : nnIdx;
			f = 1u;
		} else if (split) {
			A.src[id] = idx;
			A.tgt[id] = idx;
			f = 1u;
		}
		A.flags[id] = f;
	}
}

__global__ void __launch_bounds__(256) k_tm_compact(const uint32_t* __restrict__ len, const uint32_t* __restrict__ flags,
                                                    const uint32_t* __restrict__ offs, const uint32_t* __restrict__ src,
                                                    const uint32_t* __restrict__ tgt, uint32_t* __restrict__ c_id,
                                                    uint32_t* __restrict__ c_src, uint32_t* __restrict__ c_tgt)
{
	const uint32_t n = *len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		if (!flags[id]) continue;
		const uint32_t j = offs[id];
		c_id[j] = id; c_src[j] = src[id]; c_tgt[j] = tgt[id];
	}
}

// inclusive scan of one flag per thread over the CTA
__device__ __forceinline__ uint32_t block_scan_incl(uint32_t f, uint32_t* s_warp, uint32_t* total)
{
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t incl = f;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= (unsigned)o) incl += t;
	}
	if (lane == 31u) s_warp[warp] = incl;
	__syncthreads();
	uint32_t base = 0, sum = 0;
	for (unsigned w = 0; w < MATCH_THREADS / 32; w++) {
		const uint32_t v = s_warp[w];
		if (w < warp) base += v;
		sum += v;
	}
	__syncthreads();
	*total = sum;
	return base + incl;
}

struct match_args {
	const uint32_t* c_src;        // candidates in id order
	const uint32_t* c_tgt;
	uint32_t*       state;        // per candidate: 0 live, 1 accepted, 2 rejected
	uint32_t*       rank;         // per accepted candidate: position among the accepted of its kind
	uint32_t*       owner;        // per hidden particle
	uint32_t*       transferring; // hidden
	uint32_t*       words;        // TW_*
	const uint32_t* t_len;
	uint32_t        t_cap;
	const uint32_t* hidden_len;
	uint32_t        hidden_cap;
	int             split_on;     // settings::split (update_transfers.cpp:60)
};

__global__ void __launch_bounds__(256) k_tm_init(uint32_t* __restrict__ state, const uint32_t* __restrict__ words)
{
	const uint32_t M = words[TW_M];
	for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) state[j] = 0u;
}

// One round of the matching spread over the whole grid (the three phases are three launches: the kernel boundary is the
// barrier).  Only worth its launches when there are many candidates -- e.g. the substep in which the flood fill lifts the
// target radius of the whole interior at once; below `grid_min` candidates the kernels return at once and k_tm_match does all
// the rounds in one CTA.  After GRID_ROUNDS rounds the live set has shrunk geometrically and k_tm_match finishes the rest.
template <int PHASE>
__global__ void __launch_bounds__(256) k_tm_round(match_args A, uint32_t grid_min)
{
	const uint32_t M = A.words[TW_M];
	if (M < grid_min) return;
	for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
		if (A.state[j] != 0u) continue;
		const uint32_t cs = A.c_src[j], src = cs & IDX_MASK, tgt = A.c_tgt[j];
		if (PHASE == 0) {
			if (A.transferring[src] == 1u || ((cs & KIND_MERGE) && A.transferring[tgt] == 1u)) { A.state[j] = 2u; continue; }
			A.owner[src] = NONE;
			if (cs & KIND_MERGE) A.owner[tgt] = NONE;
		} else if (PHASE == 1) {
			atomicMin(&A.owner[src], j);
			if (cs & KIND_MERGE) atomicMin(&A.owner[tgt], j);
		} else {
			if (A.owner[src] == j && (!(cs & KIND_MERGE) || A.owner[tgt] == j)) {
				A.state[j] = 1u;
				A.transferring[src] = 1u;
				if (cs & KIND_MERGE) A.transferring[tgt] = 1u;
			}
		}
	}
}

__global__ void __launch_bounds__(MATCH_THREADS) k_tm_match(match_args A)
{
	__shared__ uint32_t s_warp[MATCH_THREADS / 32];
	__shared__ uint32_t s_live, s_jstar;
	const uint32_t M = A.words[TW_M];
	const uint32_t tid = threadIdx.x;
	// (the states were cleared by k_tm_init; grid-wide rounds may have decided part of the candidates already)
	// ---- greedy matching in id order: rounds of "lowest live candidate on each of its particles wins" ----
	for (;;) {
		if (tid == 0) s_live = 0u;
		__syncthreads();
		for (uint32_t j = tid; j < M; j += MATCH_THREADS) {
			if (A.state[j] != 0u) continue;
			const uint32_t cs = A.c_src[j], src = cs & IDX_MASK, tgt = A.c_tgt[j];
			if (A.transferring[src] == 1u || ((cs & KIND_MERGE) && A.transferring[tgt] == 1u)) { A.state[j] = 2u; continue; }
			A.owner[src] = NONE;
			if (cs & KIND_MERGE) A.owner[tgt] = NONE;
		}
		__syncthreads();
		for (uint32_t j = tid; j < M; j += MATCH_THREADS) {
			if (A.state[j] != 0u) continue;
			const uint32_t cs = A.c_src[j];
			atomicMin(&A.owner[cs & IDX_MASK], j);
			if (cs & KIND_MERGE) atomicMin(&A.owner[A.c_tgt[j]], j);
		}
		__syncthreads();
		for (uint32_t j = tid; j < M; j += MATCH_THREADS) {
			if (A.state[j] != 0u) continue;
			const uint32_t cs = A.c_src[j], src = cs & IDX_MASK, tgt = A.c_tgt[j];
			if (A.owner[src] == j && (!(cs & KIND_MERGE) || A.owner[tgt] == j)) {
				A.state[j] = 1u;
				A.transferring[src] = 1u;
				if (cs & KIND_MERGE) A.transferring[tgt] = 1u;
			} else {
				s_live = 1u;
			}
		}
		__syncthreads();
		if (s_live == 0u) break;
		__syncthreads();
	}
	// ---- rows of the accepted merges in id order; the transfer list may be full (find_split_and_merge_3.comp:103-109) ----
	const uint32_t tlen0 = min(*A.t_len, A.t_cap);
	const uint32_t room = A.t_cap - tlen0;
	if (tid == 0) s_jstar = NONE;
	__syncthreads();
	uint32_t running = 0;
	for (uint32_t base = 0; base < M; base += MATCH_THREADS) {
		const uint32_t j = base + tid;
		const uint32_t f = (j < M && A.state[j] == 1u && (A.c_src[j] & KIND_MERGE)) ? 1u : 0u;
		uint32_t total;
		const uint32_t incl = block_scan_incl(f, s_warp, &total);
		if (f) {
			const uint32_t r = running + incl - 1u;
			A.rank[j] = r;
			if (r == room) s_jstar = j; // the first merge that does not fit
		}
		running += total;
	}
	__syncthreads();
	const uint32_t jstar = s_jstar;
	uint32_t n_merge = running;
	if (jstar != NONE) {
		n_merge = room;
		// from jstar on no merge succeeds and none holds its particles; the splits behind are decided again
		for (uint32_t j = jstar + tid; j < M; j += MATCH_THREADS) {
			if (A.state[j] != 1u) continue;
			const uint32_t cs = A.c_src[j];
			A.state[j] = 2u;
			A.transferring[cs & IDX_MASK] = 0u;
			if (cs & KIND_MERGE) A.transferring[A.c_tgt[j]] = 0u;
		}
		__syncthreads();
		for (uint32_t j = jstar + tid; j < M; j += MATCH_THREADS) {
			const uint32_t cs = A.c_src[j];
			if ((cs & KIND_MERGE) || A.transferring[cs] == 1u) continue;
			A.state[j] = 1u;
			A.transferring[cs] = 1u;
		}
		__syncthreads();
	}
	// ---- accepted splits in id order, limited by the room in the transfer list and in the hidden particle list
	//      (remove_impossible_splits.comp:33-43) ----
	running = 0;
	for (uint32_t base = 0; base < M; base += MATCH_THREADS) {
		const uint32_t j = base + tid;
		const uint32_t f = (j < M && A.state[j] == 1u && !(A.c_src[j] & KIND_MERGE)) ? 1u : 0u;
		uint32_t total;
		const uint32_t incl = block_scan_incl(f, s_warp, &total);
		if (f) A.rank[j] = running + incl - 1u;
		running += total;
	}
	uint32_t n_split = 0;
	if (A.split_on) {
		const uint32_t nh = min(*A.hidden_len, A.hidden_cap);
		n_split = min(running, min(A.t_cap - (tlen0 + n_merge), A.hidden_cap - nh));
	}
	__syncthreads();
	if (A.split_on) {
		for (uint32_t j = tid; j < M; j += MATCH_THREADS) {
			if (A.state[j] != 1u || (A.c_src[j] & KIND_MERGE) || A.rank[j] < n_split) continue;
			A.state[j] = 2u;
			A.transferring[A.c_src[j]] = 0u; // removed splits are not transferring anymore
		}
	}
	if (tid == 0) { A.words[TW_N_MERGE] = n_merge; A.words[TW_N_SPLIT] = n_split; A.words[TW_TLEN0] = tlen0; }
}

struct emit_tm_args {
	const uint32_t* c_id;
	const uint32_t* c_src;
	const uint32_t* c_tgt;
	const uint32_t* state;
	const uint32_t* rank;
	const uint32_t* words;
	uint32_t* t_source; uint32_t* t_target; float* t_time_left;
	uint32_t* index_list; const uint32_t* len; uint32_t id_cap; const uint32_t* hidden_len;
	int4* pos4; int4* vel4; int4* backup4; float* inv_mass; float* radius; uint32_t* transferring;
	uint32_t* per_id[4];
	float merge_ttl, split_ttl;
	int split_on;
};

__global__ void __launch_bounds__(256) k_tm_emit(emit_tm_args A)
{
	const uint32_t M = A.words[TW_M], n_merge = A.words[TW_N_MERGE], tlen0 = A.words[TW_TLEN0];
	const uint32_t n = *A.len, nh = *A.hidden_len;
	for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
		if (A.state[j] != 1u) continue;
		const uint32_t cs = A.c_src[j], r = A.rank[j];
		if (cs & KIND_MERGE) {                                   // find_split_and_merge_3.comp:110-112
			A.t_source[tlen0 + r] = cs & IDX_MASK;
			A.t_target[tlen0 + r] = A.c_tgt[j];
			A.t_time_left[tlen0 + r] = A.merge_ttl;
		} else if (A.split_on) {                                 // update_transfers.cpp:60-69
			const uint32_t idx = cs, id = A.c_id[j], h = nh + r, nid = n + r;
			// duplicate_these (indexed_list.h:141-152): the copy is appended to the hidden list and to every list sharing it
			int4 p = A.pos4[idx];
			const float rad = A.radius[idx];
			p.x += f2i(rad * R_POS * 0.1f);                      // initialize_split_particles.comp:26-28
			A.pos4[h] = p;
			A.vel4[h] = A.vel4[idx];
			A.backup4[h] = A.backup4[idx];
			A.radius[h] = 0.0f;                                  // :30
			A.inv_mass[h] = 2147483648.0f;                       // :31 `1 / 0`: integer constant folded to 0x7FFFFFFF
			A.transferring[h] = 1u;
			if (nid < A.id_cap) {
				A.index_list[nid] = h;
#pragma unroll
				for (int a = 0; a < 4; a++) A.per_id[a][nid] = A.per_id[a][id];
			}
			const uint32_t row = tlen0 + n_merge + r;            // transferSourceList += splitList etc., :66-68
			A.t_source[row] = idx;
			A.t_target[row] = h;
			A.t_time_left[row] = A.split_ttl;
		}
	}
}

__global__ void k_tm_lengths(const uint32_t* words, uint32_t* t_len, uint32_t* len, uint32_t id_cap, uint32_t* hidden_len)
{
	const uint32_t ns = words[TW_N_SPLIT];
	*t_len = words[TW_TLEN0] + words[TW_N_MERGE] + ns;
	*hidden_len += ns;
	*len = min(*len + ns, id_cap);
}

// ---- particle_transfer.comp:30-84 ----------------------------------------------------------------------------------------
struct pt_args {
	const uint32_t* t_len; uint32_t t_cap;
	const uint32_t* source; const uint32_t* target; float* time_left;
	float* radius; float* inv_mass; uint32_t* transferring;
	uint32_t* del_h;    // per hidden particle: 1 = delete (source of a finished merge)
	uint32_t* del_row;  // per row: 1 = delete (finished split)
	float dt, dims, inv_dims;
};

__global__ void __launch_bounds__(256) k_particle_transfer(pt_args A)
{
	const uint32_t n = min(*A.t_len, A.t_cap);
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const float ttl_and_type = A.time_left[id];
		const float ttl = glsl_max(A.dt, fabsf(ttl_and_type));
		const bool merge = ttl_and_type > 0.0f;
		const uint32_t idS = A.source[id], idT = A.target[id];
		const float radiusS = A.radius[idS], radiusT = A.radius[idT];
		const float invMassS = A.inv_mass[idS], invMassT = A.inv_mass[idT];
		const float volS = pow_rn(radiusS, A.dims), volT = pow_rn(radiusT, A.dims);
		const float factorS = merge ? glsl_min(1.0f, A.dt / ttl) : A.dt * (volS - volT) / (2.0f * ttl * volS);
		const float transfVol = factorS * volS;
		const float normFactor = 1.0f / (invMassS + factorS * invMassT);
		A.inv_mass[idS] = invMassS / (1.0f - factorS);
		A.inv_mass[idT] = invMassS * invMassT * normFactor;
		A.radius[idS] = pow_rn(volS - transfVol, A.inv_dims);
		A.radius[idT] = pow_rn(volT + transfVol, A.inv_dims);
		A.time_left[id] = (ttl - A.dt) * (merge ? 1.0f : -1.0f);
		uint32_t dr = 0u;
		if (ttl == A.dt) {
			if (merge) A.del_h[idS] = 1u; else dr = 1u;
			A.transferring[idS] = 0u;
			A.transferring[idT] = 0u;
		}
		A.del_row[id] = dr;
	}
}

// keep flags of the three compactions (hidden particles, ids, transfer rows)
__global__ void __launch_bounds__(256) k_pt_keep(const uint32_t* __restrict__ del_h, const uint32_t* __restrict__ del_row,
                                                 const uint32_t* __restrict__ index_list, const uint32_t* __restrict__ source,
                                                 const uint32_t* hidden_len, const uint32_t* len, const uint32_t* t_len, uint32_t t_cap,
                                                 uint32_t* keep_h, uint32_t* keep_i, uint32_t* keep_r)
{
	const uint32_t nh = *hidden_len, n = *len, nr = min(*t_len, t_cap);
	const uint32_t top = max(nh, max(n, nr));
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < top; i += gridDim.x * blockDim.x) {
		if (i < nh) keep_h[i] = del_h[i] ? 0u : 1u;
		if (i < n) keep_i[i] = del_h[index_list[i]] ? 0u : 1u;
		if (i < nr) keep_r[i] = (del_row[i] || del_h[source[i]]) ? 0u : 1u;
	}
}

// perm[offs[i]] = i for the kept entries (new position -> old position)
__global__ void __launch_bounds__(256) k_pt_perm(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ offs,
                                                 const uint32_t* len, uint32_t cap, uint32_t* __restrict__ perm)
{
	const uint32_t n = min(*len, cap);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		if (keep[i]) perm[offs[i]] = i;
}

// dst[o] = map[src[perm[o]]]: an index list that follows the compaction of the list it points into
__global__ void __launch_bounds__(256) k_pt_remap(const uint32_t* __restrict__ src, const uint32_t* __restrict__ perm,
                                                  const uint32_t* __restrict__ map, const uint32_t* new_len, uint32_t* __restrict__ dst)
{
	const uint32_t n = *new_len;
	for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) dst[o] = map[src[perm[o]]];
}

__global__ void k_pt_lengths(const uint32_t* words, uint32_t* hidden_len, uint32_t* len, uint32_t* t_len)
{
	*hidden_len = words[TW_NEW_NH]; *len = words[TW_NEW_N]; *t_len = words[TW_NEW_ROWS];
}

// source / target of the transfers follow a permutation of the hidden list (sorted_index[new] = old)
__global__ void __launch_bounds__(256) k_inverse_perm(const uint32_t* __restrict__ sorted_index, const uint32_t* len, uint32_t* __restrict__ inv)
{
	const uint32_t n = *len;
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x) inv[sorted_index[h]] = h;
}
__global__ void __launch_bounds__(256) k_follow(const uint32_t* __restrict__ inv, const uint32_t* t_len, uint32_t t_cap,
                                                uint32_t* __restrict__ source, uint32_t* __restrict__ target)
{
	const uint32_t n = min(*t_len, t_cap);
	for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
		source[r] = inv[source[r]];
		target[r] = inv[target[r]];
	}
}

bool transfers_valid(const apbf_transfers* t)
{
	return t && t->length && t->capacity > 0 && t->source.data && t->target.data && t->time_left.data;
}

} // namespace

extern "C" {

int apbf_update_transfers_split_merge_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, apbf_transfers* transfers,
                                            float split_duration, uint32_t* out_nearest_neighbor)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid && nb && transfers_valid(transfers));
	if (ctx->mg_enabled) return apbf_fail(ctx, APBF_ERR_UNSUPPORTED, "merge / split is not available on slabs", __FILE__, __LINE__);
	apbf_particles& p = fluid->particle;
	const uint32_t n_cap = p.capacity, nh_cap = p.hidden_capacity;
	if (n_cap == 0) return APBF_OK;
	cudaStream_t st = ctx->stream;
	uint32_t* nearest = out_nearest_neighbor ? out_nearest_neighbor : (uint32_t*)ctx->scratch_get(SLOT_TM_NEAREST, sizeof(uint32_t) * (size_t)n_cap);
	if (!nearest) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	// find_split_and_merge_1/2 and the first half of _3 (update_transfers.cpp:36-48)
	APBF_TRY(apbf_update_transfers_apply(ctx, fluid, nb, nearest));
	const apbf_settings& s = ctx->settings;
	if (!s.mMerge && !s.mSplit) return APBF_OK;
	apbf_prof_scope ps(ctx, PROF_UPDATE_TRANSFERS);
	const size_t w = sizeof(uint32_t);
	uint32_t* flags = (uint32_t*)ctx->scratch_get(SLOT_TM_FLAGS, w * n_cap);
	uint32_t* offs = (uint32_t*)ctx->scratch_get(SLOT_TM_OFFS, w * ((size_t)n_cap + 1));
	uint32_t* src = (uint32_t*)ctx->scratch_get(SLOT_TM_SRC, w * n_cap);
	uint32_t* tgt = (uint32_t*)ctx->scratch_get(SLOT_TM_TGT, w * n_cap);
	uint32_t* c_id = (uint32_t*)ctx->scratch_get(SLOT_TM_CID, w * n_cap);
	uint32_t* c_src = (uint32_t*)ctx->scratch_get(SLOT_TM_CSRC, w * n_cap);
	uint32_t* c_tgt = (uint32_t*)ctx->scratch_get(SLOT_TM_CTGT, w * n_cap);
	uint32_t* state = (uint32_t*)ctx->scratch_get(SLOT_TM_STATE, w * n_cap);
	uint32_t* rank = (uint32_t*)ctx->scratch_get(SLOT_TM_RANK, w * n_cap);
	uint32_t* owner = (uint32_t*)ctx->scratch_get(SLOT_TM_OWNER, w * nh_cap);
	uint32_t* words = (uint32_t*)ctx->scratch_get(SLOT_TM_WORDS, w * TW_COUNT);
	if (!flags || !offs || !src || !tgt || !c_id || !c_src || !c_tgt || !state || !rank || !owner || !words)
		return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);

	cand_args C;
	memset(&C, 0, sizeof C);
	C.index_list = (const uint32_t*)p.index_list.data; C.len = p.length; C.pos4 = (const int32_t*)p.position.data;
	C.radius = (const float*)p.radius.data; C.target_radius = (const float*)fluid->target_radius.data; C.nearest = nearest;
	C.flags = flags; C.src = src; C.tgt = tgt; C.s = s; C.dims = (float)ctx->dims;
	C.split_factor = (float)(pow(2.0, 1.0 / (double)ctx->dims) * 0.99);
	k_tm_candidates<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(C);
	APBF_LAUNCHED(ctx);
	APBF_TRY(apbf_scan_u32(ctx, flags, offs, p.length, n_cap, false, words + TW_M, 0xFFFFFFFFu, nullptr, nullptr));
	k_tm_compact<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(p.length, flags, offs, src, tgt, c_id, c_src, c_tgt);
	APBF_LAUNCHED(ctx);

	match_args M;
	memset(&M, 0, sizeof M);
	M.c_src = c_src; M.c_tgt = c_tgt; M.state = state; M.rank = rank; M.owner = owner;
	M.transferring = (uint32_t*)p.transferring.data; M.words = words; M.t_len = transfers->length; M.t_cap = transfers->capacity;
	M.hidden_len = p.hidden_length; M.hidden_cap = nh_cap; M.split_on = s.mSplit;
	k_tm_init<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(state, words);
	APBF_LAUNCHED(ctx);
	constexpr int GRID_ROUNDS = 8;
	const unsigned rgrid = apbf_grid(ctx, n_cap, 256, 4);
	for (int r = 0; r < GRID_ROUNDS; r++) {
		k_tm_round<0><<<rgrid, 256, 0, st>>>(M, ctx->match_grid_min);
		APBF_LAUNCHED(ctx);
		k_tm_round<1><<<rgrid, 256, 0, st>>>(M, ctx->match_grid_min);
		APBF_LAUNCHED(ctx);
		k_tm_round<2><<<rgrid, 256, 0, st>>>(M, ctx->match_grid_min);
		APBF_LAUNCHED(ctx);
	}
	k_tm_match<<<1, MATCH_THREADS, 0, st>>>(M);
	APBF_LAUNCHED(ctx);

	emit_tm_args E;
	memset(&E, 0, sizeof E);
	E.c_id = c_id; E.c_src = c_src; E.c_tgt = c_tgt; E.state = state; E.rank = rank; E.words = words;
	E.t_source = (uint32_t*)transfers->source.data; E.t_target = (uint32_t*)transfers->target.data; E.t_time_left = (float*)transfers->time_left.data;
	E.index_list = (uint32_t*)p.index_list.data; E.len = p.length; E.id_cap = n_cap; E.hidden_len = p.hidden_length;
	E.pos4 = (int4*)p.position.data; E.vel4 = (int4*)p.velocity.data; E.backup4 = (int4*)p.pos_backup.data;
	E.inv_mass = (float*)p.inverse_mass.data; E.radius = (float*)p.radius.data; E.transferring = (uint32_t*)p.transferring.data;
	E.per_id[0] = (uint32_t*)fluid->target_radius.data; E.per_id[1] = (uint32_t*)fluid->kernel_width.data;
	E.per_id[2] = (uint32_t*)fluid->boundariness.data; E.per_id[3] = (uint32_t*)fluid->boundary_distance.data;
	APBF_REQUIRE(ctx, E.vel4 && E.backup4 && E.inv_mass && E.per_id[1]);
	E.merge_ttl = s.mMergeDuration < 0.0001f ? 0.0001f : s.mMergeDuration; // max(0.0001, mMergeDuration): 0 would read as a split
	E.split_ttl = -split_duration;                                          // write_sequence_float(-splitDuration, 0)
	E.split_on = s.mSplit;
	k_tm_emit<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(E);
	APBF_LAUNCHED(ctx);
	k_tm_lengths<<<1, 1, 0, st>>>(words, transfers->length, p.length, n_cap, p.hidden_length);
	APBF_LAUNCHED(ctx);
	// (the lists grew by the copies: they have no pairs -- the searches leave empty segments up to the capacity, k_offsets_tail --
	// and the copies' index entries keep an identity index list an identity)
	return APBF_OK;
}

int apbf_particle_transfer_apply(apbf_ctx* ctx, apbf_fluid* fluid, apbf_transfers* transfers, float dt, uint32_t* out_hidden_edit)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid && transfers_valid(transfers));
	APBF_REQUIRE(ctx, transfers->source.reorder_out && transfers->target.reorder_out && transfers->time_left.reorder_out);
	if (ctx->mg_enabled) return apbf_fail(ctx, APBF_ERR_UNSUPPORTED, "merge / split is not available on slabs", __FILE__, __LINE__);
	apbf_particles& p = fluid->particle;
	const uint32_t n_cap = p.capacity, nh_cap = p.hidden_capacity, t_cap = transfers->capacity;
	cudaStream_t st = ctx->stream;
	apbf_prof_scope ps(ctx, PROF_UPDATE_TRANSFERS);
	const size_t w = sizeof(uint32_t);
	uint32_t* del_h = (uint32_t*)ctx->scratch_get(SLOT_TM_OWNER, w * nh_cap);
	uint32_t* del_row = (uint32_t*)ctx->scratch_get(SLOT_TM_STATE, w * (size_t)(t_cap > n_cap ? t_cap : n_cap));
	uint32_t* keep_h = (uint32_t*)ctx->scratch_get(SLOT_TM_KEEP_H, w * nh_cap);
	uint32_t* offs_h = (uint32_t*)ctx->scratch_get(SLOT_TM_OFFS_H, w * ((size_t)nh_cap + 1));
	uint32_t* perm_h = (uint32_t*)ctx->scratch_get(SLOT_TM_PERM_H, w * nh_cap);
	uint32_t* keep_i = (uint32_t*)ctx->scratch_get(SLOT_TM_KEEP_I, w * n_cap);
	uint32_t* offs_i = (uint32_t*)ctx->scratch_get(SLOT_TM_OFFS_I, w * ((size_t)n_cap + 1));
	uint32_t* perm_i = (uint32_t*)ctx->scratch_get(SLOT_TM_PERM_I, w * n_cap);
	uint32_t* keep_r = (uint32_t*)ctx->scratch_get(SLOT_TM_KEEP_R, w * t_cap);
	uint32_t* offs_r = (uint32_t*)ctx->scratch_get(SLOT_TM_OFFS_R, w * ((size_t)t_cap + 1));
	uint32_t* perm_r = (uint32_t*)ctx->scratch_get(SLOT_TM_PERM_R, w * t_cap);
	uint32_t* words = (uint32_t*)ctx->scratch_get(SLOT_TM_WORDS, w * TW_COUNT);
	if (!del_h || !del_row || !keep_h || !offs_h || !perm_h || !keep_i || !offs_i || !perm_i || !keep_r || !offs_r || !perm_r || !words)
		return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	APBF_CUDA(ctx, cudaMemsetAsync(del_h, 0, w * nh_cap, st));

	pt_args A;
	memset(&A, 0, sizeof A);
	A.t_len = transfers->length; A.t_cap = t_cap;
	A.source = (const uint32_t*)transfers->source.data; A.target = (const uint32_t*)transfers->target.data; A.time_left = (float*)transfers->time_left.data;
	A.radius = (float*)p.radius.data; A.inv_mass = (float*)p.inverse_mass.data; A.transferring = (uint32_t*)p.transferring.data;
	A.del_h = del_h; A.del_row = del_row; A.dt = dt; A.dims = (float)ctx->dims; A.inv_dims = (float)(1.0 / (double)ctx->dims);
	k_particle_transfer<<<apbf_grid(ctx, t_cap, 256), 256, 0, st>>>(A);
	APBF_LAUNCHED(ctx);

	// deleteTransferList.delete_these(); deleteParticleList.delete_these() (particle_transfer.cpp:26-27): three stable compactions
	const uint32_t top = nh_cap > n_cap ? (nh_cap > t_cap ? nh_cap : t_cap) : (n_cap > t_cap ? n_cap : t_cap);
	k_pt_keep<<<apbf_grid(ctx, top, 256), 256, 0, st>>>(del_h, del_row, (const uint32_t*)p.index_list.data, A.source, p.hidden_length, p.length,
	                                                  transfers->length, t_cap, keep_h, keep_i, keep_r);
	APBF_LAUNCHED(ctx);
	APBF_TRY(apbf_scan_u32(ctx, keep_h, offs_h, p.hidden_length, nh_cap, false, words + TW_NEW_NH, 0xFFFFFFFFu, nullptr, nullptr));
	APBF_TRY(apbf_scan_u32(ctx, keep_i, offs_i, p.length, n_cap, false, words + TW_NEW_N, 0xFFFFFFFFu, nullptr, nullptr));
	APBF_TRY(apbf_scan_u32(ctx, keep_r, offs_r, transfers->length, t_cap, false, words + TW_NEW_ROWS, 0xFFFFFFFFu, nullptr, nullptr));
	k_pt_perm<<<apbf_grid(ctx, nh_cap, 256), 256, 0, st>>>(keep_h, offs_h, p.hidden_length, nh_cap, perm_h);
	APBF_LAUNCHED(ctx);
	k_pt_perm<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>(keep_i, offs_i, p.length, n_cap, perm_i);
	APBF_LAUNCHED(ctx);
	k_pt_perm<<<apbf_grid(ctx, t_cap, 256), 256, 0, st>>>(keep_r, offs_r, transfers->length, t_cap, perm_r);
	APBF_LAUNCHED(ctx);

	apbf_reorder_table t;
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
	APBF_TRY(apbf_launch_reorder(ctx, t, perm_h, words + TW_NEW_NH, nh_cap));
	memset(&t, 0, sizeof t);
	const apbf_array* ids[4] = { &fluid->target_radius, &fluid->kernel_width, &fluid->boundariness, &fluid->boundary_distance };
	for (auto a : ids) {
		APBF_REQUIRE(ctx, a->data && a->reorder_out && a->data != a->reorder_out);
		t.src4[t.n4] = (const uint32_t*)a->data; t.dst4[t.n4] = (uint32_t*)a->reorder_out; t.n4++;
	}
	APBF_TRY(apbf_launch_reorder(ctx, t, perm_i, words + TW_NEW_N, n_cap));
	APBF_REQUIRE(ctx, p.index_list.reorder_out && p.index_list.data != p.index_list.reorder_out);
	k_pt_remap<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>((const uint32_t*)p.index_list.data, perm_i, offs_h, words + TW_NEW_N, (uint32_t*)p.index_list.reorder_out);
	APBF_LAUNCHED(ctx);
	k_pt_remap<<<apbf_grid(ctx, t_cap, 256), 256, 0, st>>>(A.source, perm_r, offs_h, words + TW_NEW_ROWS, (uint32_t*)transfers->source.reorder_out);
	APBF_LAUNCHED(ctx);
	k_pt_remap<<<apbf_grid(ctx, t_cap, 256), 256, 0, st>>>(A.target, perm_r, offs_h, words + TW_NEW_ROWS, (uint32_t*)transfers->target.reorder_out);
	APBF_LAUNCHED(ctx);
	memset(&t, 0, sizeof t);
	t.src4[0] = (const uint32_t*)transfers->time_left.data; t.dst4[0] = (uint32_t*)transfers->time_left.reorder_out; t.n4 = 1;
	APBF_TRY(apbf_launch_reorder(ctx, t, perm_r, words + TW_NEW_ROWS, t_cap));
	if (out_hidden_edit) APBF_CUDA(ctx, cudaMemcpyAsync(out_hidden_edit, perm_h, w * nh_cap, cudaMemcpyDeviceToDevice, st));
	k_pt_lengths<<<1, 1, 0, st>>>(words, p.hidden_length, p.length, transfers->length);
	APBF_LAUNCHED(ctx);
	// ids and hidden slots mean something else now: no neighbour structure describes these particles any more, and the cached
	// "index list is the identity" flag of the last search is off until the next search re-derives it
	apbf_nbr_particles_changed(ctx);
	APBF_CUDA(ctx, cudaMemsetAsync(ctx->misc() + MW_IDENTITY, 0, 4, st));
	return APBF_OK;
}

int apbf_transfers_follow_reorder(apbf_ctx* ctx, apbf_transfers* transfers, const uint32_t* sorted_index, const uint32_t* hidden_length,
                                  uint32_t hidden_capacity)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, transfers_valid(transfers) && sorted_index && hidden_length);
	if (hidden_capacity == 0) return APBF_OK;
	uint32_t* inv = (uint32_t*)ctx->scratch_get(SLOT_TM_OWNER, sizeof(uint32_t) * (size_t)hidden_capacity);
	if (!inv) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	k_inverse_perm<<<apbf_grid(ctx, hidden_capacity, 256), 256, 0, ctx->stream>>>(sorted_index, hidden_length, inv);
	APBF_LAUNCHED(ctx);
	k_follow<<<apbf_grid(ctx, transfers->capacity, 256), 256, 0, ctx->stream>>>(inv, transfers->length, transfers->capacity,
	                                                                          (uint32_t*)transfers->source.data, (uint32_t*)transfers->target.data);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

} // extern "C"
