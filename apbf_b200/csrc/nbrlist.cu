// nbrlist.cu -- one neighbour-list structure per public pair buffer, and its rebuild from a foreign pair list.
//
// The reference's operators read whatever pbd::neighbors list they are given (incompressibility_1.comp:38-45,
// kernel_width.comp:27-33, find_split_and_merge_1.comp:18-35: one invocation per pair, ids straight from the buffer).  The
// sweeps here work on a grouped form of the list -- CSR offsets per id + NB[e] = idN | unmirrored << 31 (neighbors.cuh) --
// which a search of this context writes as a by-product.  For any other list (built by the caller, by another context,
// copied, appended to, edited in place) the same structure is derived from the public (id, idN) pairs:
//     key[e] = id (pairs with an id or idN beyond the particle list are dropped)  ->  stable onesweep sort by key
//     offsets[id] = first sorted position with key >= id                            (lower bound, one thread per id)
//     NB[i] = idN of the i-th sorted pair, unmirrored unless (idN, id) is in the list exactly once and (id, idN) is too
// The caller's list itself is left as it is.  All accumulators of the sweeps are integers, so the order of the pairs inside
// an id's segment (here: the order of the caller's list) does not change any result.
#include "neighbors.cuh"
#include "sort.cuh"

namespace {

constexpr size_t MAX_PARKED = 6;

// the public (id, idN) list from the grouped 4-byte list: 8 lanes per particle walk its segment, coalesced on both sides
// (writing the 8-byte pairs from the regroup scatters them over partially filled sectors and costs three times as much)
__global__ void __launch_bounds__(256)
k_expand_pairs(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ nbl, const uint32_t* __restrict__ len,
               uint32_t* __restrict__ pairs, uint32_t cap)
{
	const uint32_t n = *len;
	const unsigned sub = threadIdx.x & 7u;
	for (uint32_t a = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; a < n; a += (gridDim.x * blockDim.x) >> 3) {
		const uint32_t beg = min(offsets[a], cap), end = min(offsets[a + 1], cap);
		for (uint32_t e = beg + sub; e < end; e += 8u) *(uint2*)(pairs + 2 * (size_t)e) = make_uint2(a, nbl[e] & NB_ID_MASK);
	}
}

// ---- rebuild ---------------------------------------------------------------------------------------------------------------
__global__ void k_rebuild_begin(const uint32_t* __restrict__ pair_len, uint32_t pair_cap, const uint32_t* __restrict__ len,
                                const uint32_t* __restrict__ hidden_len, uint32_t* __restrict__ misc)
{
	misc[MW_REBUILD_LEN] = min(*pair_len, pair_cap);
	misc[MW_N_ASYM] = 0u;
	misc[MW_IDENTITY] = (hidden_len && *hidden_len == *len) ? 1u : 0u;
}

__global__ void k_rebuild_identity(const uint32_t* __restrict__ index_list, const uint32_t* __restrict__ len, uint32_t* __restrict__ misc)
{
	const uint32_t n = *len;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		if (index_list[i] != i) misc[MW_IDENTITY] = 0u;
}

__global__ void k_rebuild_keys(const uint2* __restrict__ pairs, const uint32_t* __restrict__ misc, const uint32_t* __restrict__ len,
                               uint32_t invalid_key, uint32_t* __restrict__ keys)
{
	const uint32_t m = misc[MW_REBUILD_LEN], n = *len;
	for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < m; e += gridDim.x * blockDim.x) {
		const uint2 p = pairs[e];
		keys[e] = (p.x < n && p.y < n) ? p.x : invalid_key;
	}
}

// offsets[id] = first position of the sorted keys that is >= id, for id in [0, n_cap]
__global__ void k_rebuild_offsets(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ misc, uint32_t n_cap, uint32_t* __restrict__ offsets)
{
	const uint32_t m = misc[MW_REBUILD_LEN];
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id <= n_cap; id += gridDim.x * blockDim.x) {
		uint32_t lo = 0u, hi = m;
		while (lo < hi) {
			const uint32_t mid = lo + ((hi - lo) >> 1);
			if (skeys[mid] < id) lo = mid + 1u; else hi = mid;
		}
		offsets[id] = lo;
	}
}

__global__ void k_rebuild_fill(const uint2* __restrict__ pairs, const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ perm,
                               const uint32_t* __restrict__ misc, uint32_t invalid_key, uint32_t* __restrict__ nbl)
{
	const uint32_t m = misc[MW_REBUILD_LEN];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
		if (skeys[i] != invalid_key) nbl[i] = pairs[perm[i]].y;
}

// mirrored(a, b) <=> the gather form of the sweeps may stand in for the pair (b, a): both directions exactly once
__global__ void k_rebuild_mirror(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ offsets, uint32_t* nbl,
                                 uint32_t* __restrict__ misc, uint32_t invalid_key)
{
	const uint32_t m = misc[MW_REBUILD_LEN];
	uint32_t n_asym = 0u;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
		const uint32_t a = skeys[i];
		if (a == invalid_key) continue;
		const uint32_t b = ((volatile uint32_t*)nbl)[i] & NB_ID_MASK;
		uint32_t ab = 0u, ba = 0u;
		for (uint32_t e = offsets[a], end = offsets[a + 1]; e < end; e++) ab += (((volatile uint32_t*)nbl)[e] & NB_ID_MASK) == b ? 1u : 0u;
		for (uint32_t e = offsets[b], end = offsets[b + 1]; e < end; e++) ba += (((volatile uint32_t*)nbl)[e] & NB_ID_MASK) == a ? 1u : 0u;
		if (ab != 1u || ba != 1u) { atomicOr(nbl + i, NB_UNMIRRORED); n_asym++; }
	}
	const uint32_t am = __activemask();
	n_asym = __reduce_add_sync(am, n_asym);
	if ((am & ((1u << lane_id()) - 1u)) == 0u && n_asym) atomicAdd(misc + MW_N_ASYM, n_asym);
}

int rebuild(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb)
{
	apbf_particles& p = fluid->particle;
	APBF_REQUIRE(ctx, p.length && p.index_list.data && nb->pairs && nb->length);
	const uint32_t n_cap = p.capacity, pair_cap = nb->capacity;
	cudaStream_t st = ctx->stream;
	uint32_t* misc = ctx->misc();
	uint32_t* offsets = (uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * ((size_t)n_cap + 1));
	uint32_t* nbl = (uint32_t*)ctx->scratch_get(SLOT_NB, sizeof(uint32_t) * ((size_t)pair_cap + 1));
	uint32_t* keys = (uint32_t*)ctx->scratch_get(SLOT_PAIRS_TMP, sizeof(uint32_t) * 2 * (size_t)pair_cap);
	uint32_t* skeys = (uint32_t*)ctx->scratch_get(SLOT_NB_TMP, sizeof(uint32_t) * ((size_t)pair_cap + 1));
	uint32_t* perm = (uint32_t*)ctx->scratch_get(SLOT_STREAM, sizeof(uint32_t) * ((size_t)pair_cap + 1));
	if (!misc || !offsets || !nbl || !keys || !skeys || !perm) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	const uint32_t invalid_key = n_cap; // one above the largest id: dropped pairs sort behind every segment
	int bits = 1;
	while (bits < 32 && (invalid_key >> bits) != 0u) bits++;
	k_rebuild_begin<<<1, 1, 0, st>>>(nb->length, pair_cap, p.length, p.hidden_length, misc);
	APBF_LAUNCHED(ctx);
	k_rebuild_identity<<<apbf_grid(ctx, n_cap, 256), 256, 0, st>>>((const uint32_t*)p.index_list.data, p.length, misc);
	APBF_LAUNCHED(ctx);
	if (pair_cap > 0u) {
		k_rebuild_keys<<<apbf_grid(ctx, pair_cap, 256), 256, 0, st>>>((const uint2*)nb->pairs, misc, p.length, invalid_key, keys);
		APBF_LAUNCHED(ctx);
		APBF_TRY(apbf_radix_sort_pairs(ctx, keys, nullptr, skeys, perm, misc + MW_REBUILD_LEN, pair_cap, bits));
	}
	k_rebuild_offsets<<<apbf_grid(ctx, (size_t)n_cap + 1, 256), 256, 0, st>>>(skeys, misc, n_cap, offsets);
	APBF_LAUNCHED(ctx);
	if (pair_cap > 0u) {
		k_rebuild_fill<<<apbf_grid(ctx, pair_cap, 256), 256, 0, st>>>((const uint2*)nb->pairs, skeys, perm, misc, invalid_key, nbl);
		APBF_LAUNCHED(ctx);
		k_rebuild_mirror<<<apbf_grid(ctx, pair_cap, 256, 16), 256, 0, st>>>(skeys, offsets, nbl, misc, invalid_key);
		APBF_LAUNCHED(ctx);
	}
	return APBF_OK;
}

void free_entry(apbf_nbr_entry& e)
{
	if (e.offsets.ptr) cudaFree(e.offsets.ptr);
	if (e.nbl.ptr) cudaFree(e.nbl.ptr);
	if (e.words) cudaFree(e.words);
	e = apbf_nbr_entry();
}

} // namespace

int apbf_launch_expand_pairs(apbf_ctx* ctx, const uint32_t* offsets, const uint32_t* nbl, const uint32_t* len, uint32_t n_cap,
                             uint32_t* pairs, uint32_t pair_cap)
{
	if (n_cap == 0u) return APBF_OK;
	k_expand_pairs<<<apbf_grid(ctx, (size_t)n_cap * 8, 256, 32), 256, 0, ctx->stream>>>(offsets, nbl, len, pairs, pair_cap);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_nbr_activate(apbf_ctx* ctx, const apbf_neighbors* nb)
{
	APBF_REQUIRE(ctx, nb && nb->pairs);
	if (ctx->nbr_struct_pairs == nb->pairs) return APBF_OK;
	cudaStream_t st = ctx->stream;
	uint32_t* misc = ctx->misc();
	if (!misc) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	// park the active structure (if it is worth keeping)
	if (ctx->nbr_struct_pairs && ctx->nbr_valid && ctx->scratch[SLOT_OFFSETS].ptr && ctx->scratch[SLOT_NB].ptr) {
		apbf_nbr_entry e;
		e.pairs = ctx->nbr_struct_pairs;
		e.n_cap = ctx->nbr_struct_n_cap;
		e.valid = true;
		e.public_written = ctx->nbr_public;
		e.stamp = ++ctx->nbr_clock;
		if (cudaMalloc((void**)&e.words, 256) != cudaSuccess) return apbf_fail(ctx, APBF_ERR_OOM, "cudaMalloc", __FILE__, __LINE__);
		APBF_CUDA(ctx, cudaMemcpyAsync(e.words, misc + MW_N_ASYM, 4, cudaMemcpyDeviceToDevice, st));
		e.offsets = ctx->scratch[SLOT_OFFSETS];
		e.nbl = ctx->scratch[SLOT_NB];
		ctx->scratch[SLOT_OFFSETS] = apbf_scratch();
		ctx->scratch[SLOT_NB] = apbf_scratch();
		ctx->scratch_epoch++;
		ctx->nbr_cache.push_back(e);
		if (ctx->nbr_cache.size() > MAX_PARKED) { // drop the one that has been parked longest
			size_t old = 0;
			for (size_t i = 1; i < ctx->nbr_cache.size(); i++) if (ctx->nbr_cache[i].stamp < ctx->nbr_cache[old].stamp) old = i;
			cudaStreamSynchronize(st);
			free_entry(ctx->nbr_cache[old]);
			ctx->nbr_cache.erase(ctx->nbr_cache.begin() + (long)old);
		}
	}
	ctx->nbr_struct_pairs = nb->pairs;
	ctx->nbr_struct_n_cap = 0;
	ctx->nbr_valid = false;
	ctx->nbr_public = true; // a list this context knows nothing about: its public form is all there is
	for (size_t i = 0; i < ctx->nbr_cache.size(); i++) {
		apbf_nbr_entry& e = ctx->nbr_cache[i];
		if (e.pairs != nb->pairs) continue;
		// un-park: the parked buffers become the scratch slots (whatever the slots hold now is temporary data)
		cudaStreamSynchronize(st);
		if (ctx->scratch[SLOT_OFFSETS].ptr) cudaFree(ctx->scratch[SLOT_OFFSETS].ptr);
		if (ctx->scratch[SLOT_NB].ptr) cudaFree(ctx->scratch[SLOT_NB].ptr);
		ctx->scratch[SLOT_OFFSETS] = e.offsets;
		ctx->scratch[SLOT_NB] = e.nbl;
		ctx->scratch_epoch++;
		cudaMemcpyAsync(misc + MW_N_ASYM, e.words, 4, cudaMemcpyDeviceToDevice, st);
		cudaStreamSynchronize(st);
		cudaFree(e.words);
		ctx->nbr_struct_n_cap = e.n_cap;
		ctx->nbr_valid = e.valid;
		ctx->nbr_public = e.public_written;
		ctx->nbr_cache.erase(ctx->nbr_cache.begin() + (long)i);
		break;
	}
	return APBF_OK;
}

void apbf_nbr_built(apbf_ctx* ctx, const apbf_neighbors* nb, uint32_t n_cap, bool public_written)
{
	ctx->nbr_struct_pairs = nb->pairs;
	ctx->nbr_struct_n_cap = n_cap;
	ctx->nbr_valid = true;
	ctx->nbr_public = public_written;
}

int apbf_nbr_ensure(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb)
{
	APBF_REQUIRE(ctx, fluid && nb && nb->pairs && nb->length);
	APBF_TRY(apbf_nbr_activate(ctx, nb));
	if (ctx->nbr_valid && ctx->nbr_struct_n_cap == fluid->particle.capacity) return APBF_OK;
	if (!ctx->nbr_public)
		return apbf_fail(ctx, APBF_ERR_INVALID, "neighbour list is stale and its public (id, idN) form was never written", __FILE__, __LINE__);
	APBF_TRY(rebuild(ctx, fluid, nb));
	apbf_nbr_built(ctx, nb, fluid->particle.capacity, true);
	return APBF_OK;
}

int apbf_nbr_materialize(apbf_ctx* ctx, const uint32_t* len, uint32_t n_cap, const apbf_neighbors* nb)
{
	APBF_REQUIRE(ctx, nb && nb->pairs && len);
	APBF_TRY(apbf_nbr_activate(ctx, nb));
	if (ctx->nbr_public || !ctx->nbr_valid) return APBF_OK;
	const uint32_t* offsets = (const uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * ((size_t)n_cap + 1));
	const uint32_t* nbl = (const uint32_t*)ctx->scratch_get(SLOT_NB, sizeof(uint32_t) * ((size_t)nb->capacity + 1));
	if (!offsets || !nbl) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	APBF_TRY(apbf_launch_expand_pairs(ctx, offsets, nbl, len, n_cap, nb->pairs, nb->capacity));
	ctx->nbr_public = true;
	return APBF_OK;
}

void apbf_nbr_forget(apbf_ctx* ctx, const void* ptr)
{
	if (!ptr) return;
	if (ctx->nbr_struct_pairs == ptr) { ctx->nbr_struct_pairs = nullptr; ctx->nbr_valid = false; ctx->nbr_public = false; }
	for (size_t i = 0; i < ctx->nbr_cache.size();) {
		if (ctx->nbr_cache[i].pairs == ptr) {
			cudaStreamSynchronize(ctx->stream);
			free_entry(ctx->nbr_cache[i]);
			ctx->nbr_cache.erase(ctx->nbr_cache.begin() + (long)i);
		} else i++;
	}
}

void apbf_nbr_touch(apbf_ctx* ctx, const void* ptr)
{
	if (!ptr) return;
	// (a write into the middle of a list comes with the list's base address in every library call that can do it)
	if (ctx->nbr_struct_pairs == ptr) { ctx->nbr_valid = false; ctx->nbr_public = true; }
	for (auto& e : ctx->nbr_cache) if (e.pairs == ptr) { e.valid = false; e.public_written = true; }
}

void apbf_nbr_particles_changed(apbf_ctx* ctx)
{
	// ids mean something else now: neither the structure nor the public list describes the particles any more.  The reference's
	// list is just as stale at this point (pool.cpp:73-84 searches right after particle_transfer); rebuilding from the public
	// list drops the pairs whose ids fell off the end.
	if (ctx->nbr_struct_pairs) ctx->nbr_valid = false;
	for (auto& e : ctx->nbr_cache) e.valid = false;
}

extern "C" {

int apbf_neighbors_invalidate(apbf_ctx* ctx, const apbf_neighbors* nb)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, nb);
	apbf_nbr_touch(ctx, nb->pairs);
	return APBF_OK;
}

int apbf_neighbors_release(apbf_ctx* ctx, const apbf_neighbors* nb)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, nb);
	apbf_nbr_forget(ctx, nb->pairs);
	return APBF_OK;
}

int apbf_neighbors_prepare(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, uint32_t* out_pair_offsets)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_TRY(apbf_nbr_ensure(ctx, fluid, nb));
	if (out_pair_offsets) {
		const uint32_t n_cap = fluid->particle.capacity;
		const uint32_t* offsets = (const uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * ((size_t)n_cap + 1));
		if (!offsets) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
		APBF_CUDA(ctx, cudaMemcpyAsync(out_pair_offsets, offsets, sizeof(uint32_t) * ((size_t)n_cap + 1), cudaMemcpyDeviceToDevice, ctx->stream));
	}
	return APBF_OK;
}

} // extern "C"
