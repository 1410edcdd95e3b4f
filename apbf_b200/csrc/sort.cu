// sort.cu -- onesweep-style LSD radix sort (8-bit digits, warp match ranking, shared-memory digit histograms,
// decoupled look-back across tiles) and a single-pass chained prefix sum.
//
// Replaces algorithms::sort / algorithms::prefix_sum (source/algorithms.cpp:59-118) and the four shaders behind
// them (radix_sort_apply_on_block_level.comp, radix_sort_scattered_write.comp, prefix_sum_apply_on_block_level.comp,
// prefix_sum_spread_from_block_level.comp).  Same results: the sort is stable and ascending on the key bits the
// reference's 4-bit passes cover; the scan is inclusive.
#include "sort.cuh"

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int MAX_PASSES = 4;

constexpr uint32_t FLAG_LOCAL = 1u << 30;
constexpr uint32_t FLAG_GLOBAL = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v)
{
	asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t digit_of(uint32_t key, int shift, uint32_t mask) { return (key >> shift) & mask; }

// ---- digit histograms of all passes in one read of the keys -----------------------------------------------------
__global__ void __launch_bounds__(256) k_histogram(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ count,
                                                   uint32_t* __restrict__ ghist, int passes, uint32_t last_mask)
{
	__shared__ uint32_t s_hist[MAX_PASSES][RADIX];
	for (int i = threadIdx.x; i < MAX_PASSES * RADIX; i += blockDim.x) (&s_hist[0][0])[i] = 0u;
	__syncthreads();
	const uint32_t n = *count;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		uint32_t k = keys[i];
#pragma unroll
		for (int p = 0; p < MAX_PASSES; p++) {
			if (p < passes) {
				uint32_t m = (p == passes - 1) ? last_mask : (uint32_t)(RADIX - 1);
				atomicAdd(&s_hist[p][digit_of(k, p * RADIX_BITS, m)], 1u);
			}
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) {
		uint32_t v = (&s_hist[0][0])[i];
		if (v) atomicAdd(&ghist[i], v);
	}
}

// exclusive scan of each pass's 256 bins, in place; one block of 256 threads
__global__ void __launch_bounds__(RADIX) k_scan_histograms(uint32_t* __restrict__ ghist, int passes)
{
	__shared__ uint32_t s_warp[RADIX / 32];
	for (int p = 0; p < passes; p++) {
		uint32_t v = ghist[p * RADIX + threadIdx.x];
		uint32_t incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane_id() >= (unsigned)o) incl += t;
		}
		if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
		__syncthreads();
		uint32_t base = 0;
		for (int w = 0; w < (int)(threadIdx.x >> 5); w++) base += s_warp[w];
		ghist[p * RADIX + threadIdx.x] = base + incl - v;
		__syncthreads();
	}
}

// ---- one onesweep pass ---------------------------------------------------------------------------------------------
template <int IPT, bool IOTA_VALUES>
__global__ void __launch_bounds__(SORT_THREADS)
k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ count, const uint32_t* __restrict__ gbase,
           uint32_t* status, uint32_t* ticket, int shift, uint32_t mask)
{
	constexpr int TILE = SORT_THREADS * IPT;
	__shared__ uint32_t s_warp_hist[SORT_WARPS][RADIX];
	__shared__ uint32_t s_tile_off[RADIX];
	__shared__ uint32_t s_glob[RADIX];
	__shared__ uint32_t s_stage[TILE];
	__shared__ uint32_t s_scan[RADIX / 32];
	__shared__ uint32_t s_tile;

	const uint32_t n = *count;
	const uint32_t num_tiles = (n + TILE - 1) / TILE;
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&s_warp_hist[0][0])[i] = 0u;
	__syncthreads();
	const uint32_t tile = s_tile;
	if (tile >= num_tiles) return;

	const unsigned warp = threadIdx.x >> 5, lane = lane_id();
	const uint32_t tile_base = tile * (uint32_t)TILE;
	const uint32_t n_valid = min((uint32_t)TILE, n - tile_base);

	// warp-striped load: item i of this lane sits at warp_base + i*32 + lane -> ranks follow memory order
	uint32_t key[IPT], rank[IPT];
	const uint32_t warp_base = tile_base + warp * (IPT * 32);
#pragma unroll
	for (int i = 0; i < IPT; i++) {
		uint32_t idx = warp_base + i * 32 + lane;
		key[i] = idx < n ? keys_in[idx] : 0xFFFFFFFFu;
	}
#pragma unroll
	for (int i = 0; i < IPT; i++) {
		uint32_t d = digit_of(key[i], shift, mask);
		uint32_t peers = __match_any_sync(0xffffffffu, d);
		int leader = __ffs(peers) - 1;
		uint32_t pre = 0;
		if ((int)lane == leader) {
			pre = s_warp_hist[warp][d];
			s_warp_hist[warp][d] = pre + __popc(peers);
		}
		pre = __shfl_sync(0xffffffffu, pre, leader);
		rank[i] = pre + __popc(peers & ((1u << lane) - 1u));
		__syncwarp();
	}
	__syncthreads();

	// thread d owns digit d: exclusive scan over warps, publish the tile count, look back
	{
		const uint32_t d = threadIdx.x;
		uint32_t sum = 0;
#pragma unroll
		for (int w = 0; w < SORT_WARPS; w++) {
			uint32_t c = s_warp_hist[w][d];
			s_warp_hist[w][d] = sum;
			sum += c;
		}
		const uint32_t tile_count = sum;
		uint32_t* my_status = status + (size_t)tile * RADIX + d;
		if (tile == 0) st_volatile_u32(my_status, tile_count | FLAG_GLOBAL);
		else st_volatile_u32(my_status, tile_count | FLAG_LOCAL);

		// exclusive scan of the tile counts over the 256 digits
		uint32_t incl = tile_count;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (unsigned)o) incl += t;
		}
		if (lane == 31) s_scan[warp] = incl;
		__syncthreads();
		uint32_t base = 0;
		for (unsigned w = 0; w < warp; w++) base += s_scan[w];
		const uint32_t tile_off = base + incl - tile_count;
		s_tile_off[d] = tile_off;

		uint32_t excl = 0;
		if (tile > 0) {
			int t = (int)tile - 1;
			while (true) {
				uint32_t v = ld_volatile_u32(status + (size_t)t * RADIX + d);
				uint32_t f = v & FLAG_MASK;
				if (f == 0u) continue;
				excl += v & VALUE_MASK;
				if (f == FLAG_GLOBAL) break;
				t--;
			}
			st_volatile_u32(my_status, (excl + tile_count) | FLAG_GLOBAL);
		}
		s_glob[d] = gbase[d] + excl - tile_off;
	}
	__syncthreads();

	// local scatter into shared memory in sorted order, then coalesced runs to global memory
	uint32_t lpos[IPT];
#pragma unroll
	for (int i = 0; i < IPT; i++) {
		uint32_t d = digit_of(key[i], shift, mask);
		lpos[i] = s_tile_off[d] + s_warp_hist[warp][d] + rank[i];
		s_stage[lpos[i]] = key[i];
	}
	__syncthreads();
	uint32_t gpos[IPT];
#pragma unroll
	for (int m = 0; m < IPT; m++) {
		uint32_t j = threadIdx.x + m * SORT_THREADS;
		gpos[m] = 0xFFFFFFFFu;
		if (j < n_valid) {
			uint32_t k = s_stage[j];
			gpos[m] = s_glob[digit_of(k, shift, mask)] + j;
			keys_out[gpos[m]] = k;
		}
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < IPT; i++) {
		uint32_t idx = warp_base + i * 32 + lane;
		uint32_t v = IOTA_VALUES ? idx : (idx < n ? vals_in[idx] : 0u);
		s_stage[lpos[i]] = v;
	}
	__syncthreads();
#pragma unroll
	for (int m = 0; m < IPT; m++) {
		uint32_t j = threadIdx.x + m * SORT_THREADS;
		if (j < n_valid) vals_out[gpos[m]] = s_stage[j];
	}
}

__global__ void k_copy_pairs(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                             uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                             const uint32_t* __restrict__ count)
{
	const uint32_t n = *count;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		keys_out[i] = keys_in[i];
		vals_out[i] = vals_in ? vals_in[i] : (uint32_t)i;
	}
}

// ---- chained scan -----------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;
constexpr unsigned long long SFLAG_LOCAL = 1ull << 32;
constexpr unsigned long long SFLAG_GLOBAL = 2ull << 32;

template <bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan(const uint32_t* values, uint32_t* result, const uint32_t* __restrict__ count, unsigned long long* status,
       uint32_t* ticket, uint32_t* total_out, uint32_t total_clamp, uint32_t* flags_out, uint32_t* raw_total_out, const uint32_t* __restrict__ run_if)
{
	if (run_if && *run_if == 0u) return;
	__shared__ uint32_t s_warp[SCAN_THREADS / 32];
	__shared__ uint32_t s_tile;
	__shared__ uint32_t s_excl;
	const uint32_t n = *count;
	const uint32_t num_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	if (tile >= num_tiles) {
		if (n == 0 && tile == 0 && threadIdx.x == 0) {
			if (total_out) *total_out = 0u;
			if (raw_total_out) *raw_total_out = 0u;
			if (!INCLUSIVE) result[0] = 0u;
		}
		return;
	}
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	// blocked arrangement: thread t owns items [t*IPT, t*IPT+IPT) of the tile
	const uint32_t base = tile * (uint32_t)SCAN_TILE + threadIdx.x * SCAN_IPT;
	uint32_t v[SCAN_IPT];
	if (base + SCAN_IPT <= n && (((uintptr_t)values & 15u) == 0)) {
		uint4 a = *(const uint4*)(values + base);
		uint4 b = *(const uint4*)(values + base + 4);
		v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
	} else {
#pragma unroll
		for (int i = 0; i < SCAN_IPT; i++) v[i] = base + i < n ? values[base + i] : 0u;
	}
	uint32_t thread_sum = 0;
#pragma unroll
	for (int i = 0; i < SCAN_IPT; i++) thread_sum += v[i];
	uint32_t incl = thread_sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= (unsigned)o) incl += t;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	uint32_t wbase = 0, tile_sum = 0;
#pragma unroll
	for (int w = 0; w < SCAN_THREADS / 32; w++) {
		uint32_t s = s_warp[w];
		if (w < (int)warp) wbase += s;
		tile_sum += s;
	}
	if (warp == 0) {
		// decoupled look-back by the whole warp: lane l reads the status of tile (t - l), 32 predecessors per trip, and the sums up to
		// the nearest tile whose inclusive prefix is known are added with one reduction (a single thread walking back one tile per
		// L2 round trip was 2/3 of this kernel's time at 500 tiles)
		uint32_t excl = 0;
		if (tile == 0) {
			if (lane == 0) st_volatile_u64(status + tile, SFLAG_GLOBAL | tile_sum);
		} else {
			if (lane == 0) st_volatile_u64(status + tile, SFLAG_LOCAL | tile_sum);
			int t = (int)tile - 1;
			while (true) {
				const int idx = t - (int)lane;
				const unsigned long long sv = idx >= 0 ? ld_volatile_u64(status + idx) : SFLAG_GLOBAL; // before tile 0: prefix 0, known
				const uint32_t f = (uint32_t)(sv >> 32);
				const uint32_t known = __ballot_sync(0xffffffffu, f == 2u), pending = __ballot_sync(0xffffffffu, f == 0u);
				if (known) {
					const uint32_t g = (uint32_t)__ffs(known) - 1u;         // nearest tile with a known prefix
					if (pending & ((2u << g) - 1u)) continue;               // somebody in front of it has not posted yet
					excl += __reduce_add_sync(0xffffffffu, lane <= g ? (uint32_t)sv : 0u);
					break;
				}
				if (pending) continue;
				excl += __reduce_add_sync(0xffffffffu, (uint32_t)sv);
				t -= 32;
			}
			if (lane == 0) st_volatile_u64(status + tile, SFLAG_GLOBAL | (uint32_t)(excl + tile_sum));
		}
		if (lane == 0) {
			s_excl = excl;
			if (tile == num_tiles - 1) {
				uint32_t total = excl + tile_sum;
				if (!INCLUSIVE) result[n] = total; // exclusive scans are CSR offsets: one extra entry holds the grand total
				if (raw_total_out) *raw_total_out = total;
				if (total_out) {
					if (total > total_clamp) {
						total = total_clamp;
						if (flags_out) atomicOr(flags_out, 1u);
					}
					*total_out = total;
				}
			}
		}
	}
	__syncthreads();
	uint32_t run = s_excl + wbase + incl - thread_sum;
	uint32_t o[SCAN_IPT];
#pragma unroll
	for (int i = 0; i < SCAN_IPT; i++) {
		if (INCLUSIVE) { run += v[i]; o[i] = run; }
		else { o[i] = run; run += v[i]; }
	}
	if (base + SCAN_IPT <= n && (((uintptr_t)result & 15u) == 0)) {
		*(uint4*)(result + base) = make_uint4(o[0], o[1], o[2], o[3]);
		*(uint4*)(result + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
	} else {
#pragma unroll
		for (int i = 0; i < SCAN_IPT; i++)
			if (base + i < n) result[base + i] = o[i];
	}
}

} // namespace

int apbf_reference_sort_bits(uint32_t upper_bound)
{
	int bits = 0;
	for (unsigned off = 0u; off < 32u && (upper_bound >> off) != 0u; off += 4u) bits = (int)off + 4; // algorithms.cpp:73
	return bits;
}

int apbf_radix_sort_pairs(apbf_ctx* ctx, const uint32_t* keys_in, const uint32_t* values_in, uint32_t* keys_out,
                          uint32_t* values_out, const uint32_t* count, uint32_t max_count, int key_bits)
{
	APBF_REQUIRE(ctx, keys_in && keys_out && values_out && count);
	APBF_REQUIRE(ctx, key_bits >= 0 && key_bits <= 32);
	if (max_count == 0) return APBF_OK;
	cudaStream_t st = ctx->stream;
	const int passes = (key_bits + RADIX_BITS - 1) / RADIX_BITS;
	if (passes == 0) {
		k_copy_pairs<<<apbf_grid(ctx, max_count, 256), 256, 0, st>>>(keys_in, values_in, keys_out, values_out, count);
		APBF_LAUNCHED(ctx);
		return APBF_OK;
	}
	const int last_bits = key_bits - (passes - 1) * RADIX_BITS;
	const uint32_t last_mask = (1u << last_bits) - 1u;

	// large inputs use 16 keys per thread, small ones 8 so that the grid still covers the SMs
	const bool big = max_count > (1u << 21);
	const uint32_t tile = big ? SORT_THREADS * 16 : SORT_THREADS * 8;
	const uint32_t tiles = apbf_div_up(max_count, tile);

	uint32_t* ghist = (uint32_t*)ctx->scratch_get(SLOT_SORT_HIST, sizeof(uint32_t) * MAX_PASSES * RADIX);
	uint32_t* status = (uint32_t*)ctx->scratch_get(SLOT_SORT_STATUS, sizeof(uint32_t) * (size_t)passes * tiles * RADIX);
	uint32_t* misc = ctx->misc();
	uint32_t* tmp_k = nullptr;
	uint32_t* tmp_v = nullptr;
	if (passes > 1) {
		tmp_k = (uint32_t*)ctx->scratch_get(SLOT_SORT_KEYS_B, sizeof(uint32_t) * (size_t)max_count);
		tmp_v = (uint32_t*)ctx->scratch_get(SLOT_SORT_VALS_B, sizeof(uint32_t) * (size_t)max_count);
		if (!tmp_k || !tmp_v) return apbf_fail(ctx, APBF_ERR_OOM, "sort scratch", __FILE__, __LINE__);
	}
	if (!ghist || !status || !misc) return apbf_fail(ctx, APBF_ERR_OOM, "sort scratch", __FILE__, __LINE__);

	APBF_CUDA(ctx, cudaMemsetAsync(ghist, 0, sizeof(uint32_t) * MAX_PASSES * RADIX, st));
	APBF_CUDA(ctx, cudaMemsetAsync(status, 0, sizeof(uint32_t) * (size_t)passes * tiles * RADIX, st));
	APBF_CUDA(ctx, cudaMemsetAsync(misc + MW_TICKET0, 0, sizeof(uint32_t) * 8, st));

	k_histogram<<<apbf_grid(ctx, max_count, 256, 4), 256, 0, st>>>(keys_in, count, ghist, passes, last_mask);
	APBF_LAUNCHED(ctx);
	k_scan_histograms<<<1, RADIX, 0, st>>>(ghist, passes);
	APBF_LAUNCHED(ctx);

	// buffer rotation that never writes to the inputs and ends in the outputs:
	//   odd passes : in -> out -> tmp -> out ...        even passes: in -> tmp -> out -> tmp -> out ...
	const uint32_t* src_k = keys_in;
	const uint32_t* src_v = values_in;
	for (int p = 0; p < passes; p++) {
		const bool to_out = ((passes - 1 - p) % 2) == 0;
		uint32_t* dst_k = to_out ? keys_out : tmp_k;
		uint32_t* dst_v = to_out ? values_out : tmp_v;
		const uint32_t mask = (p == passes - 1) ? last_mask : (uint32_t)(RADIX - 1);
		const bool iota = (p == 0 && values_in == nullptr);
		uint32_t* st_p = status + (size_t)p * tiles * RADIX;
#define LAUNCH_ONESWEEP(IPT, IOTA)                                                                                \
	k_onesweep<IPT, IOTA><<<tiles, SORT_THREADS, 0, st>>>(src_k, src_v, dst_k, dst_v, count, ghist + p * RADIX, st_p, \
	                                                      misc + MW_TICKET0 + p, p * RADIX_BITS, mask)
		if (big) { if (iota) LAUNCH_ONESWEEP(16, true); else LAUNCH_ONESWEEP(16, false); }
		else     { if (iota) LAUNCH_ONESWEEP(8, true);  else LAUNCH_ONESWEEP(8, false); }
#undef LAUNCH_ONESWEEP
		APBF_LAUNCHED(ctx);
		src_k = dst_k;
		src_v = dst_v;
	}
	return APBF_OK;
}

int apbf_scan_u32(apbf_ctx* ctx, const uint32_t* values, uint32_t* result, const uint32_t* count, uint32_t max_count,
                  bool inclusive, uint32_t* total_out, uint32_t total_clamp, uint32_t* flags_out, uint32_t* raw_total_out, const uint32_t* run_if)
{
	APBF_REQUIRE(ctx, values && result && count);
	cudaStream_t st = ctx->stream;
	const uint32_t tiles = max_count ? apbf_div_up(max_count, SCAN_TILE) : 1u;
	unsigned long long* status = (unsigned long long*)ctx->scratch_get(SLOT_SCAN_STATUS, sizeof(unsigned long long) * tiles);
	uint32_t* misc = ctx->misc();
	if (!status || !misc) return apbf_fail(ctx, APBF_ERR_OOM, "scan scratch", __FILE__, __LINE__);
	APBF_CUDA(ctx, cudaMemsetAsync(status, 0, sizeof(unsigned long long) * tiles, st));
	APBF_CUDA(ctx, cudaMemsetAsync(misc + MW_TICKET0 + 7, 0, sizeof(uint32_t), st));
	if (inclusive)
		k_scan<true><<<tiles, SCAN_THREADS, 0, st>>>(values, result, count, status, misc + MW_TICKET0 + 7, total_out, total_clamp, flags_out, raw_total_out, run_if);
	else
		k_scan<false><<<tiles, SCAN_THREADS, 0, st>>>(values, result, count, status, misc + MW_TICKET0 + 7, total_out, total_clamp, flags_out, raw_total_out, run_if);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

extern "C" {

int apbf_sort(apbf_ctx* ctx, uint32_t* keys, uint32_t* values, const uint32_t* count, uint32_t max_count,
              uint32_t* out_keys, uint32_t* out_values, uint32_t upper_bound)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, keys && values && out_keys && out_values && count);
	APBF_REQUIRE(ctx, keys != out_keys && values != out_values);
	return apbf_radix_sort_pairs(ctx, keys, values, out_keys, out_values, count, max_count, apbf_reference_sort_bits(upper_bound));
}

int apbf_prefix_sum(apbf_ctx* ctx, const uint32_t* values, const uint32_t* count, uint32_t max_count, uint32_t* result)
{
	if (!ctx) return APBF_ERR_INVALID;
	return apbf_scan_u32(ctx, values, result, count, max_count, true, nullptr, 0xFFFFFFFFu, nullptr, nullptr);
}

} // extern "C"
