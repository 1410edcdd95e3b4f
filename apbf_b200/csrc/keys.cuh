// keys.cuh -- grid mapping shared by the hash kernel and the neighbour search, plus internal launchers of keys.cu
#pragma once
#include "common.cuh"

// push constants of calculate_position_hash.comp:14-18 / neighborhood_green.comp:26-31
struct apbf_grid_params {
	float    mn[3];
	float    ext[3]; // mMaxPos - mMinPos
	float    scale;  // float(1u << mResolutionLog2)
	uint32_t res;
	int      dims;
};

int apbf_make_grid_params(apbf_ctx* ctx, const float mn[3], const float mx[3], uint32_t res, apbf_grid_params* g);
// ghost_bit != 0: ids >= misc[MW_N_OWNED] get this bit added to their key (they sort behind all owned particles)
int apbf_launch_position_hash(apbf_ctx* ctx, const int32_t* pos4, uint32_t* out, const uint32_t* len, uint32_t cap,
                              const apbf_grid_params& g, uint32_t ghost_bit = 0u);
int apbf_launch_position_code(apbf_ctx* ctx, const uint32_t* index_list, const int32_t* pos4, uint32_t* out,
                              const uint32_t* len, uint32_t cap, uint32_t section);
int apbf_launch_find_value_ranges(apbf_ctx* ctx, const uint32_t* index_list, const uint32_t* values, uint32_t* range_start,
                                  uint32_t* range_end, const uint32_t* len, uint32_t cap, uint32_t n_ranges);
int apbf_launch_gather(apbf_ctx* ctx, const void* src, void* dst, const uint32_t* edit, const uint32_t* len, uint32_t cap,
                       uint32_t stride);

#ifdef __CUDACC__
// map_pos_to_grid, one axis: uint((pos - min) / (max - min) * float(1u << res))   (calculate_position_hash.comp:23-26)
__device__ __forceinline__ uint32_t apbf_map_axis(float p, const apbf_grid_params& g, int d)
{
	return f2u((p - g.mn[d]) / g.ext[d] * g.scale);
}
// ---- bit spreading ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread3_10(uint32_t v) // bit i -> bit 3i, 10 bits
{
	v &= 0x3FFu;
	v = (v | (v << 16)) & 0x030000FFu;
	v = (v | (v << 8)) & 0x0300F00Fu;
	v = (v | (v << 4)) & 0x030C30C3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}
__device__ __forceinline__ uint32_t spread2_16(uint32_t v) // bit i -> bit 2i, 16 bits
{
	v &= 0xFFFFu;
	v = (v | (v << 8)) & 0x00FF00FFu;
	v = (v | (v << 4)) & 0x0F0F0F0Fu;
	v = (v | (v << 2)) & 0x33333333u;
	v = (v | (v << 1)) & 0x55555555u;
	return v;
}
__device__ __forceinline__ unsigned long long spread3_21(unsigned long long v) // bit i -> bit 3i, 21 bits
{
	v &= 0x1FFFFFull;
	v = (v | (v << 32)) & 0x1F00000000FFFFull;
	v = (v | (v << 16)) & 0x1F0000FF0000FFull;
	v = (v | (v << 8)) & 0x100F00F00F00F00Full;
	v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
	v = (v | (v << 2)) & 0x1249249249249249ull;
	return v;
}

// z-curve hash of calculate_position_hash.comp:29-36 / neighborhood_green.comp:40-47: only the low `res` bits per axis.
// DIMS is a template parameter on purpose: with a run-time `dims` nvcc 12.9 (-O3, sm_100a) merges the tails of the
// 3-D and the 2-D bit spread into one block in which "v | (v << s)" has become "v + (v << s)" -- valid for the 3-D
// masks only -- and the 2-D hash of e.g. (0, 3) comes out as 2 instead of 10 (seen in the pair emit kernel, 2026-10).
template <int DIMS>
__device__ __forceinline__ uint32_t apbf_zhash(uint32_t gx, uint32_t gy, uint32_t gz, uint32_t res)
{
	const uint32_t m = (1u << res) - 1u;
	if (DIMS == 3) return spread3_10(gx & m) | (spread3_10(gy & m) << 1) | (spread3_10(gz & m) << 2);
	return spread2_16(gx & m) | (spread2_16(gy & m) << 1);
}

// calculate_position_code.comp:23-62: bit k of x -> bit 3k, y -> 3k+1, z -> 3k+2 of a 96-bit word (no sign bias)
__device__ __forceinline__ void apbf_encode96(int32_t x, int32_t y, int32_t z, uint32_t out[3])
{
	const uint32_t ux = (uint32_t)x, uy = (uint32_t)y, uz = (uint32_t)z;
	unsigned long long lo = spread3_21(ux) | (spread3_21(uy) << 1) | (spread3_21(uz) << 2); // bits 0..62
	unsigned long long hi = spread3_21(ux >> 21) | (spread3_21(uy >> 21) << 1) | (spread3_21(uz >> 21) << 2); // bits 63..95
	out[0] = (uint32_t)lo;
	out[1] = (uint32_t)(lo >> 32) | ((uint32_t)(hi & 1ull) << 31);
	out[2] = (uint32_t)(hi >> 1);
}
#endif
