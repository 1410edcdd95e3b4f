// keys.cu -- position keys (cell hash / 96-bit Morton code), cell ranges, and the list helper kernels.
#include "keys.cuh"
#include "neighbors.cuh"
#include "sort.cuh"

namespace {

template <int DIMS>
__global__ void k_position_hash(const int32_t* __restrict__ pos4, uint32_t* __restrict__ out, const uint32_t* __restrict__ len,
                                apbf_grid_params g, const uint32_t* __restrict__ misc, uint32_t ghost_bit)
{
	const uint32_t n = *len;
	const uint32_t n_owned = ghost_bit ? misc[MW_N_OWNED] : 0xFFFFFFFFu;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		int4 p = ldg_int4(pos4, id);
		float px = (float)p.x * INV_R_POS, py = (float)p.y * INV_R_POS, pz = (float)p.z * INV_R_POS;
		out[id] = apbf_zhash<DIMS>(apbf_map_axis(px, g, 0), apbf_map_axis(py, g, 1), apbf_map_axis(pz, g, 2), g.res) | (id >= n_owned ? ghost_bit : 0u);
	}
}

__global__ void k_position_code(const uint32_t* __restrict__ index_list, const int32_t* __restrict__ pos4,
                                uint32_t* __restrict__ out, const uint32_t* __restrict__ len, uint32_t section)
{
	const uint32_t n = *len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		uint32_t idx = index_list ? index_list[id] : id;
		int4 p = ldg_int4(pos4, idx);
		uint32_t c[3];
		apbf_encode96(p.x, p.y, p.z, c);
		out[id] = c[section];
	}
}

// find_value_ranges.comp:16-31 (the id-1 read at id 0 is not replicated)
__global__ void k_find_value_ranges(const uint32_t* __restrict__ index_list, const uint32_t* __restrict__ values,
                                    uint32_t* __restrict__ range_start, uint32_t* __restrict__ range_end,
                                    const uint32_t* __restrict__ len)
{
	const uint32_t n = *len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		uint32_t curr = values[index_list ? index_list[id] : id];
		bool differs = false;
		uint32_t prev = 0;
		if (id > 0) {
			prev = values[index_list ? index_list[id - 1] : id - 1];
			differs = curr != prev;
		}
		if (id == 0 || differs) range_start[curr] = id;
		if (differs) range_end[prev] = id;
		if (id == n - 1u) range_end[curr] = id + 1u;
	}
}

// ---- list helpers ------------------------------------------------------------------------------------------------
__global__ void k_write_sequence(uint32_t* __restrict__ out, const uint32_t* __restrict__ len, uint32_t start, uint32_t step,
                                 uint32_t len_scale)
{
	const size_t n = (size_t)*len * len_scale;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		out[i] = start + (uint32_t)i * step;
}

__global__ void k_write_sequence_float(float* __restrict__ out, const uint32_t* __restrict__ len, float start, float step)
{
	const uint32_t n = *len;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		out[i] = start + (float)i * step;
}

template <typename T>
__global__ void k_gather(const T* __restrict__ src, T* __restrict__ dst, const uint32_t* __restrict__ edit,
                         const uint32_t* __restrict__ len)
{
	const uint32_t n = *len;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[edit[i]];
}

__global__ void k_gather_words(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const uint32_t* __restrict__ edit,
                               const uint32_t* __restrict__ len, uint32_t stride_words)
{ // copy_scattered_read.comp:21-30, one thread per word
	const size_t n = (size_t)*len * stride_words;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		size_t e = i / stride_words;
		uint32_t off = (uint32_t)(i - e * stride_words);
		dst[i] = src[(size_t)edit[e] * stride_words + off];
	}
}

__global__ void k_scattered_write(const uint32_t* __restrict__ index, uint32_t* __restrict__ target,
                                  const uint32_t* __restrict__ len, uint32_t value)
{
	const uint32_t n = *len;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) target[index[i]] = value;
}

__global__ void k_append_list(uint32_t* __restrict__ target, const uint32_t* __restrict__ appending,
                              const uint32_t* __restrict__ target_len, const uint32_t* __restrict__ appending_len,
                              uint32_t* new_len, uint32_t target_capacity, uint32_t stride_words)
{ // append_list.comp:21-29; the copy is clamped to the target capacity
	const uint32_t tl = *target_len, al = *appending_len;
	const uint32_t total = min(tl + al, target_capacity);
	const size_t n = (size_t)(total > tl ? total - tl : 0u) * stride_words;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		target[(size_t)tl * stride_words + i] = appending[i];
	// new_len may alias target_len: every thread has read tl above, but other blocks may still be starting ->
	// written by a follow-up kernel instead (k_store_word)
	(void)new_len;
}

__global__ void k_append_len(const uint32_t* target_len, const uint32_t* appending_len, uint32_t* new_len, uint32_t cap)
{
	*new_len = min(*target_len + *appending_len, cap);
}

__global__ void k_copy_strided(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const uint32_t* __restrict__ len,
                               uint32_t src_words, uint32_t dst_words)
{ // copy_with_differing_stride.comp:21-32: copies min(stride) words per element
	const uint32_t w = min(src_words, dst_words);
	const size_t n = (size_t)*len * w;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		size_t e = i / w;
		uint32_t off = (uint32_t)(i - e * w);
		dst[e * dst_words + off] = src[e * src_words + off];
	}
}

// find_value_changes.comp:16-30 (the id-1 read at id 0 is not replicated): `in` is a non-decreasing running count;
// out[k] = first id whose value exceeds k, *out_len = last value
__global__ void k_find_value_changes(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ in_len,
                                     uint32_t* out_len)
{
	const uint32_t n = *in_len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const uint32_t curr = in[id], prev = id > 0 ? in[id - 1] : 0u;
		if (id == n - 1u) *out_len = curr;
		if (curr != prev) out[prev] = id;
	}
}

// write_increasing_sequence.comp:23-36: target[i] = *seq_min + i for i < L = min(capacity, seq_len, upper - *seq_min);
// *new_seq_min = *seq_min + L, *new_target_len = L.  seq_min is read by every thread before lane 0 of block 0 may
// overwrite it through an aliasing new_seq_min, hence the snapshot kernel in front.
__global__ void k_increasing_sequence_words(const uint32_t* seq_min, uint32_t* new_seq_min, uint32_t* new_target_len,
                                            uint32_t target_capacity, uint32_t upper, uint32_t seq_len, uint32_t* snapshot)
{
	const uint32_t m = *seq_min;
	const uint32_t L = min(min(target_capacity, seq_len), upper - m);
	*snapshot = m;
	*new_seq_min = m + L;
	*new_target_len = L;
}
__global__ void k_write_increasing_sequence_dev(uint32_t* __restrict__ target, uint32_t target_capacity, const uint32_t* snapshot,
                                                uint32_t upper, uint32_t seq_len)
{
	const uint32_t m = *snapshot;
	const uint32_t L = min(min(target_capacity, seq_len), upper - m);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) target[i] = m + i;
}

// write_increasing_sequence_from_to.comp:19-31
__global__ void k_write_from_to(uint32_t* __restrict__ out, uint32_t* out_len, const uint32_t* from, const uint32_t* to, uint32_t capacity)
{
	const uint32_t f = *from, t = *to;
	const uint32_t n = min(t > f ? t - f : 0u, capacity);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = f + i;
}
__global__ void k_write_from_to_len(uint32_t* out_len, const uint32_t* from, const uint32_t* to) { *out_len = *to - *from; }

// ---- general apply_hidden_edit (sort based) --------------------------------------------------------------------------
__global__ void k_hidden_edit_counts(const uint32_t* __restrict__ edit, const uint32_t* __restrict__ edit_len,
                                     const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                                     uint32_t* __restrict__ counts)
{
	const uint32_t n = *edit_len;
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x) {
		uint32_t old = edit[h];
		counts[h] = end[old] - start[old];
	}
}

__global__ void k_hidden_edit_emit(const uint32_t* __restrict__ edit, const uint32_t* __restrict__ edit_len,
                                   const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                                   const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ sorted_payload,
                                   uint32_t* __restrict__ new_index, uint32_t* __restrict__ new_edit, uint32_t cap)
{
	const uint32_t n = *edit_len;
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x) {
		uint32_t old = edit[h];
		uint32_t s = start[old], e = end[old], o = offsets[h];
		for (uint32_t j = s; j < e; j++, o++) {
			if (o < cap) {
				new_index[o] = h;
				new_edit[o] = sorted_payload[j];
			}
		}
	}
}

} // namespace

// ---- internal launchers ---------------------------------------------------------------------------------------------
int apbf_launch_position_hash(apbf_ctx* ctx, const int32_t* pos4, uint32_t* out, const uint32_t* len, uint32_t cap,
                              const apbf_grid_params& g, uint32_t ghost_bit)
{
	if (cap == 0) return APBF_OK;
	if (g.dims == 3) k_position_hash<3><<<apbf_grid(ctx, cap, 256), 256, 0, ctx->stream>>>(pos4, out, len, g, ctx->misc(), ghost_bit);
	else k_position_hash<2><<<apbf_grid(ctx, cap, 256), 256, 0, ctx->stream>>>(pos4, out, len, g, ctx->misc(), ghost_bit);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_launch_position_code(apbf_ctx* ctx, const uint32_t* index_list, const int32_t* pos4, uint32_t* out,
                              const uint32_t* len, uint32_t cap, uint32_t section)
{
	if (cap == 0) return APBF_OK;
	k_position_code<<<apbf_grid(ctx, cap, 256), 256, 0, ctx->stream>>>(index_list, pos4, out, len, section);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_launch_find_value_ranges(apbf_ctx* ctx, const uint32_t* index_list, const uint32_t* values, uint32_t* range_start,
                                  uint32_t* range_end, const uint32_t* len, uint32_t cap, uint32_t n_ranges)
{
	APBF_CUDA(ctx, cudaMemsetAsync(range_start, 0, sizeof(uint32_t) * (size_t)n_ranges, ctx->stream));
	APBF_CUDA(ctx, cudaMemsetAsync(range_end, 0, sizeof(uint32_t) * (size_t)n_ranges, ctx->stream));
	if (cap == 0) return APBF_OK;
	k_find_value_ranges<<<apbf_grid(ctx, cap, 256), 256, 0, ctx->stream>>>(index_list, values, range_start, range_end, len);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_launch_gather(apbf_ctx* ctx, const void* src, void* dst, const uint32_t* edit, const uint32_t* len, uint32_t cap,
                       uint32_t stride)
{
	if (cap == 0) return APBF_OK;
	cudaStream_t st = ctx->stream;
	const bool a16 = (((uintptr_t)src | (uintptr_t)dst) & 15u) == 0;
	const bool a8 = (((uintptr_t)src | (uintptr_t)dst) & 7u) == 0;
	if (stride == 16 && a16) k_gather<int4><<<apbf_grid(ctx, cap, 256), 256, 0, st>>>((const int4*)src, (int4*)dst, edit, len);
	else if (stride == 8 && a8) k_gather<int2><<<apbf_grid(ctx, cap, 256), 256, 0, st>>>((const int2*)src, (int2*)dst, edit, len);
	else if (stride == 4) k_gather<uint32_t><<<apbf_grid(ctx, cap, 256), 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, edit, len);
	else {
		APBF_REQUIRE(ctx, stride % 4 == 0);
		k_gather_words<<<apbf_grid(ctx, (size_t)cap * (stride / 4), 256), 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, edit, len, stride / 4);
	}
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

extern "C" {

int apbf_write_sequence(apbf_ctx* ctx, uint32_t* out, const uint32_t* len, uint32_t capacity, uint32_t start, uint32_t step,
                        uint32_t len_scale)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, out && len);
	if (capacity == 0 || len_scale == 0) return APBF_OK;
	k_write_sequence<<<apbf_grid(ctx, (size_t)capacity * len_scale, 256), 256, 0, ctx->stream>>>(out, len, start, step, len_scale);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_write_sequence_float(apbf_ctx* ctx, float* out, const uint32_t* len, uint32_t capacity, float start, float step)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, out && len);
	if (capacity == 0) return APBF_OK;
	k_write_sequence_float<<<apbf_grid(ctx, capacity, 256), 256, 0, ctx->stream>>>(out, len, start, step);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_copy_scattered_read(apbf_ctx* ctx, const void* src, void* dst, const uint32_t* edit, const uint32_t* edit_len,
                             uint32_t capacity, uint32_t stride_bytes)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, src && dst && edit && edit_len && src != dst);
	apbf_nbr_touch(ctx, dst); // (if dst is a pair list some operator knows: its content is changing)
	return apbf_launch_gather(ctx, src, dst, edit, edit_len, capacity, stride_bytes);
}

int apbf_scattered_write(apbf_ctx* ctx, const uint32_t* index, uint32_t* target, const uint32_t* len, uint32_t capacity,
                         uint32_t value)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, index && target && len);
	if (capacity == 0) return APBF_OK;
	k_scattered_write<<<apbf_grid(ctx, capacity, 256), 256, 0, ctx->stream>>>(index, target, len, value);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_append_list(apbf_ctx* ctx, void* target, const void* appending, const uint32_t* target_len,
                     const uint32_t* appending_len, uint32_t* new_len, uint32_t target_capacity,
                     uint32_t appending_capacity, uint32_t stride_bytes)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, target && appending && target_len && appending_len && new_len && stride_bytes % 4 == 0);
	apbf_nbr_touch(ctx, target);
	if (appending_capacity > 0) {
		k_append_list<<<apbf_grid(ctx, (size_t)appending_capacity * (stride_bytes / 4), 256), 256, 0, ctx->stream>>>(
		    (uint32_t*)target, (const uint32_t*)appending, target_len, appending_len, new_len, target_capacity, stride_bytes / 4);
		APBF_LAUNCHED(ctx);
	}
	k_append_len<<<1, 1, 0, ctx->stream>>>(target_len, appending_len, new_len, target_capacity);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_find_value_changes(apbf_ctx* ctx, const uint32_t* in, uint32_t* out_change, const uint32_t* in_len,
                            uint32_t* out_change_len, uint32_t capacity)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, in && out_change && in_len && out_change_len);
	if (capacity == 0) return APBF_OK;
	k_find_value_changes<<<apbf_grid(ctx, capacity, 256), 256, 0, ctx->stream>>>(in, out_change, in_len, out_change_len);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_write_increasing_sequence(apbf_ctx* ctx, uint32_t* target, uint32_t target_capacity, uint32_t* new_target_len,
                                   const uint32_t* sequence_min_value, uint32_t* new_sequence_min_value,
                                   uint32_t value_upper_bound, uint32_t sequence_length)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, target && new_target_len && sequence_min_value && new_sequence_min_value);
	uint32_t* snap = ctx->misc() + MW_SNAPSHOT;
	k_increasing_sequence_words<<<1, 1, 0, ctx->stream>>>(sequence_min_value, new_sequence_min_value, new_target_len, target_capacity,
	                                                      value_upper_bound, sequence_length, snap);
	APBF_LAUNCHED(ctx);
	const uint32_t most = target_capacity < sequence_length ? target_capacity : sequence_length;
	if (most > 0) {
		k_write_increasing_sequence_dev<<<apbf_grid(ctx, most, 256), 256, 0, ctx->stream>>>(target, target_capacity, snap, value_upper_bound,
		                                                                                  sequence_length);
		APBF_LAUNCHED(ctx);
	}
	return APBF_OK;
}

int apbf_write_increasing_sequence_from_to(apbf_ctx* ctx, uint32_t* out, uint32_t* out_len, const uint32_t* from, const uint32_t* to,
                                           uint32_t capacity)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, out && out_len && from && to && out_len != from && out_len != to);
	if (capacity > 0) {
		k_write_from_to<<<apbf_grid(ctx, capacity, 256), 256, 0, ctx->stream>>>(out, out_len, from, to, capacity);
		APBF_LAUNCHED(ctx);
	}
	k_write_from_to_len<<<1, 1, 0, ctx->stream>>>(out_len, from, to);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_copy_with_differing_stride(apbf_ctx* ctx, const void* src, void* dst, const uint32_t* len, uint32_t capacity,
                                    uint32_t src_stride_bytes, uint32_t dst_stride_bytes)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, src && dst && len && src_stride_bytes % 4 == 0 && dst_stride_bytes % 4 == 0);
	apbf_nbr_touch(ctx, dst);
	if (capacity == 0) return APBF_OK;
	uint32_t w = (src_stride_bytes < dst_stride_bytes ? src_stride_bytes : dst_stride_bytes) / 4;
	k_copy_strided<<<apbf_grid(ctx, (size_t)capacity * w, 256), 256, 0, ctx->stream>>>((const uint32_t*)src, (uint32_t*)dst, len,
	                                                                                  src_stride_bytes / 4, dst_stride_bytes / 4);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_apply_hidden_edit(apbf_ctx* ctx, const uint32_t* edit, const uint32_t* edit_len, uint32_t edit_capacity,
                           const uint32_t* index_list, const uint32_t* index_len, uint32_t index_capacity,
                           uint32_t hidden_capacity, uint32_t* new_index_list, uint32_t* new_edit_list, uint32_t* new_len)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, edit && edit_len && index_list && index_len && new_index_list && new_edit_list && new_len);
	APBF_REQUIRE(ctx, new_index_list != index_list);
	// group the index-list entries by the hidden slot they point to: stable sort (slot, entry number)
	uint32_t* sk = (uint32_t*)ctx->scratch_get(SLOT_TMP_KEYS, sizeof(uint32_t) * (size_t)(index_capacity + 1));
	uint32_t* sv = (uint32_t*)ctx->scratch_get(SLOT_TMP_VALS, sizeof(uint32_t) * (size_t)(index_capacity + 1));
	uint32_t* start = (uint32_t*)ctx->scratch_get(SLOT_HIDDEN_FLAGS, sizeof(uint32_t) * (size_t)(hidden_capacity + 1));
	uint32_t* end = (uint32_t*)ctx->scratch_get(SLOT_HIDDEN_OFFS, sizeof(uint32_t) * (size_t)(hidden_capacity + 1));
	uint32_t* counts = (uint32_t*)ctx->scratch_get(SLOT_EDIT_COUNTS, sizeof(uint32_t) * (size_t)(edit_capacity + 2));
	uint32_t* offsets = (uint32_t*)ctx->scratch_get(SLOT_EDIT_OFFSETS, sizeof(uint32_t) * (size_t)(edit_capacity + 2));
	if (!sk || !sv || !start || !end || !counts || !offsets) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	int bits = 1;
	while (bits < 32 && (hidden_capacity >> bits) != 0u) bits++;
	APBF_TRY(apbf_radix_sort_pairs(ctx, index_list, nullptr, sk, sv, index_len, index_capacity, bits));
	APBF_TRY(apbf_launch_find_value_ranges(ctx, nullptr, sk, start, end, index_len, index_capacity, hidden_capacity));
	if (edit_capacity > 0) {
		k_hidden_edit_counts<<<apbf_grid(ctx, edit_capacity, 256), 256, 0, ctx->stream>>>(edit, edit_len, start, end, counts);
		APBF_LAUNCHED(ctx);
	}
	APBF_TRY(apbf_scan_u32(ctx, counts, offsets, edit_len, edit_capacity, false, new_len, index_capacity, nullptr, nullptr));
	if (edit_capacity > 0) {
		k_hidden_edit_emit<<<apbf_grid(ctx, edit_capacity, 256), 256, 0, ctx->stream>>>(edit, edit_len, start, end, offsets, sv,
		                                                                                new_index_list, new_edit_list, index_capacity);
		APBF_LAUNCHED(ctx);
	}
	return APBF_OK;
}

int apbf_calculate_position_hash(apbf_ctx* ctx, const int32_t* position4, uint32_t* out_hash, const uint32_t* len,
                                 uint32_t capacity, const float min_pos[3], const float max_pos[3], uint32_t res_log2)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, position4 && out_hash && len && min_pos && max_pos);
	apbf_grid_params g;
	APBF_TRY(apbf_make_grid_params(ctx, min_pos, max_pos, res_log2, &g));
	return apbf_launch_position_hash(ctx, position4, out_hash, len, capacity, g);
}

int apbf_calculate_position_code(apbf_ctx* ctx, const uint32_t* index_list, const int32_t* position4, uint32_t* out_code,
                                 const uint32_t* len, uint32_t capacity, uint32_t code_section)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, index_list && position4 && out_code && len && code_section < 3);
	return apbf_launch_position_code(ctx, index_list, position4, out_code, len, capacity, code_section);
}

int apbf_find_value_ranges(apbf_ctx* ctx, const uint32_t* index_list, const uint32_t* values, uint32_t* range_start,
                           uint32_t* range_end, const uint32_t* len, uint32_t capacity, uint32_t n_ranges)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, index_list && values && range_start && range_end && len);
	return apbf_launch_find_value_ranges(ctx, index_list, values, range_start, range_end, len, capacity, n_ranges);
}

} // extern "C"

int apbf_make_grid_params(apbf_ctx* ctx, const float mn[3], const float mx[3], uint32_t res, apbf_grid_params* g)
{
	APBF_REQUIRE(ctx, res >= 1 && res * (uint32_t)ctx->dims <= 30u);
	for (int d = 0; d < 3; d++) {
		g->mn[d] = mn[d];
		g->ext[d] = mx[d] - mn[d]; // (mMaxPos - mMinPos), evaluated once in fp32 exactly as the shader does per thread
	}
	g->scale = (float)(1u << res);
	g->res = res;
	g->dims = ctx->dims;
	return APBF_OK;
}
