/*
 * apbf_oracle.c -- CPU restatement of the APBF particle hot path (see apbf_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into or called from the product path.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include "apbf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define R_POS ORC_POS_RESOLUTION
#define R_INC ORC_INCOMP_RESOLUTION
#define PI 3.14159265f /* kernels.glsl:1 (a GLSL literal without suffix is a float) */

static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* ---------------------------------------------------------------------------------- */
/* GLSL helpers with the conventions stated in the header                             */
/* ---------------------------------------------------------------------------------- */
static inline float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline float length3(const float a[3]) { return sqrtf(dot3(a, a)); }
static inline void normalize3(const float a[3], float out[3])
{
	float l = length3(a);
	out[0] = a[0] / l; out[1] = a[1] / l; out[2] = a[2] / l;
}
static inline uint32_t f2u(float f) /* uint(float): truncate, negatives/NaN saturate to 0 */
{
	if (!(f > 0.0f)) return 0u;
	if (f >= 4294967296.0f) return 0xFFFFFFFFu;
	return (uint32_t)f;
}
static inline int32_t f2i(float f) /* int(float): truncate toward zero, saturating */
{
	if (f != f) return 0;
	if (f >= 2147483648.0f) return 2147483647;
	if (f <= -2147483648.0f) return (int32_t)0x80000000;
	return (int32_t)f;
}
static inline float fminf_(float a, float b) { return b < a ? b : a; } /* GLSL min(x,y) = y < x ? y : x */
static inline float fmaxf_(float a, float b) { return a < b ? b : a; } /* GLSL max(x,y) = x < y ? y : x */
static inline float fractf_(float x) { return x - floorf(x); }
static inline float signf_(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

void orc_default_settings(orc_settings* s) /* settings.cpp:5-31 */
{
	s->mHeightKernelId = 1;
	s->mGradientKernelId = 1;
	s->mMerge = 1;
	s->mSplit = 1;
	s->mBaseKernelWidthOnTargetRadius = 1;
	s->mBaseKernelWidthOnBoundaryDistance = 1;
	s->mUpdateTargetRadius = 1;
	s->mUpdateBoundariness = 1;
	s->mNeighborListSorted = 0;
	s->mBoundarinessCalculationMethod = 2;
	s->mBoundarinessAdaptionSpeed = 0.5f;
	s->mKernelWidthAdaptionSpeed = 0.01f;
	s->mBoundarinessSelfGradLengthFactor = 2.0f;
	s->mBoundarinessUnderpressureFactor = 4.0f;
	s->mMergeDuration = 2.0f;
	s->mSmallestTargetRadius = 1.0f;
	s->mTargetRadiusOffset = 10.0f;
	s->mTargetRadiusScaleFactor = 0.3f;
}

/* ---------------------------------------------------------------------------------- */
/* kernels.glsl:3-123                                                                 */
/* ---------------------------------------------------------------------------------- */
static float poly6_kernel_height(const float r[3], float h) /* kernels.glsl:3-9 */
{
	float rSquared = dot3(r, r);
	float hSquared = h * h;
	if (rSquared > hSquared) return 0.0f;
	return 315.0f * powf(hSquared - rSquared, 3.0f) / (64.0f * PI * powf(h, 9.0f));
}

static void spiky_kernel_gradient(const float r[3], float h, float o[3]) /* kernels.glsl:11-16 */
{
	float dist = length3(r);
	if (dist > h || dist < 0.0001f) { o[0] = o[1] = o[2] = 0.0f; return; }
	float f = -45.0f * powf(h - dist, 2.0f) / (PI * powf(h, 6.0f));
	o[0] = f * (r[0] / dist); o[1] = f * (r[1] / dist); o[2] = f * (r[2] / dist);
}

static float cubic_kernel_height(const float r[3], float h) /* kernels.glsl:18-31 */
{
	float dist = length3(r);
	if (dist > h) return 0.0f;
	float q = dist / h;
	float k = 8.0f / (PI * h * h * h);
	if (q <= 0.5f) {
		float q2 = q * q;
		float q3 = q * q2;
		return k * (6.0f * q3 - 6.0f * q2 + 1.0f);
	}
	return k * (2.0f * powf(1.0f - q, 3.0f));
}

static void cubic_kernel_gradient(const float r[3], float h, float o[3]) /* kernels.glsl:33-46 */
{
	float dist = length3(r);
	if (dist > h || dist < 0.0001f) { o[0] = o[1] = o[2] = 0.0f; return; }
	float q = dist / h;
	float s = 1.0f / (dist * h);
	float gradq[3] = { r[0] * s, r[1] * s, r[2] * s };
	float l = 48.0f / (PI * h * h * h);
	float f;
	if (q <= 0.5f) {
		f = l * q * (3.0f * q - 2.0f);
	} else {
		float factor = 1.0f - q;
		f = l * (-factor * factor);
	}
	o[0] = f * gradq[0]; o[1] = f * gradq[1]; o[2] = f * gradq[2];
}

static float cone_kernel_height(const float r[3], float h, int D) /* kernels.glsl:48-52 */
{
	float height = 3.0f / (PI * powf(h, (float)D));
	return fmaxf_(0.0f, (1.0f - length3(r) / h) * height);
}

static void cone_kernel_gradient(const float r[3], float h, int D, float o[3]) /* kernels.glsl:54-60 */
{
	float steepness = 3.0f / (PI * powf(h, (float)(D + 1)));
	float dist = length3(r);
	if (dist > h || dist < 0.0001f) { o[0] = o[1] = o[2] = 0.0f; return; }
	float f = steepness / dist;
	o[0] = -r[0] * f; o[1] = -r[1] * f; o[2] = -r[2] * f;
}

static float quadratic_spike_a(float h, int D) /* kernels.glsl:64-68 */
{
	if (D == 3) return 15.0f / (2.0f * PI * powf(h, 5.0f));
	return 6.0f / (PI * powf(h, 4.0f));
}

static float quadratic_spike_kernel_height(const float r[3], float h, int D) /* kernels.glsl:62-70 */
{
	float a = quadratic_spike_a(h, D);
	return a * powf(fminf_(0.0f, length3(r) - h), 2.0f);
}

static void quadratic_spike_kernel_gradient(const float r[3], float h, int D, float o[3]) /* kernels.glsl:72-82 */
{
	float dist = length3(r);
	if (dist > h || dist < 0.0001f) { o[0] = o[1] = o[2] = 0.0f; return; }
	float a = quadratic_spike_a(h, D);
	float f = -2.0f * a * fmaxf_(0.0f, h - dist) / dist;
	o[0] = f * r[0]; o[1] = f * r[1]; o[2] = f * r[2];
}

static float gauss_kernel_height(const float r[3], float height, int D) /* kernels.glsl:84-89 */
{
	float invDoubleVarianceWithoutPi = powf(height, 2.0f / (float)D);
	float invDoubleVariance = invDoubleVarianceWithoutPi * PI;
	return expf(-dot3(r, r) * invDoubleVariance) * powf(invDoubleVarianceWithoutPi, (float)D / 2.0f);
}

static void gauss_kernel_gradient(const float r[3], float height, int D, float o[3]) /* kernels.glsl:91-97 */
{
	float dist = length3(r);
	if (dist < 0.0001f) { o[0] = o[1] = o[2] = 0.0f; return; }
	float invDoubleVariance = powf(height, 2.0f / (float)D) * PI;
	float f = -gauss_kernel_height(r, height, D) * 2.0f * dist * invDoubleVariance;
	float nrm[3];
	normalize3(r, nrm);
	o[0] = f * nrm[0]; o[1] = f * nrm[1]; o[2] = f * nrm[2];
}

float orc_kernel_height(const orc_settings* s, int D, const float r[3], float w) /* kernels.glsl:99-110 */
{
	switch (s->mHeightKernelId) {
		case 0: return cubic_kernel_height(r, w);
		case 1: return gauss_kernel_height(r, 0.6f / powf(w / 2.0f, (float)D), D);
		case 2: return poly6_kernel_height(r, w);
		case 3: return cone_kernel_height(r, w, D);
		case 4: return quadratic_spike_kernel_height(r, w, D);
		default: return 0.0f;
	}
}

void orc_kernel_gradient(const orc_settings* s, int D, const float r[3], float w, float o[3]) /* kernels.glsl:112-123 */
{
	switch (s->mGradientKernelId) {
		case 0: cubic_kernel_gradient(r, w, o); return;
		case 1: gauss_kernel_gradient(r, 0.6f / powf(w / 2.0f, (float)D), D, o); return;
		case 2: spiky_kernel_gradient(r, w, o); return;
		case 3: cone_kernel_gradient(r, w, D, o); return;
		case 4: quadratic_spike_kernel_gradient(r, w, D, o); return;
		default: o[0] = o[1] = o[2] = 0.0f; return;
	}
}

/* ---------------------------------------------------------------------------------- */
/* algorithms.cpp                                                                     */
/* ---------------------------------------------------------------------------------- */
size_t orc_prefix_sum_helper_length(size_t max_count) /* algorithms.cpp:48-58 */
{
	uint32_t elementCount = (uint32_t)max_count;
	uint32_t groupsize = 512u, result = 0u;
	do {
		elementCount = (elementCount + groupsize - 1u) / groupsize;
		result += elementCount;
	} while (elementCount > 1);
	return result == 0u ? 0u : result + 10u;
}

size_t orc_sort_helper_length(size_t max_count) /* algorithms.cpp:38-46 */
{
	uint32_t bucketCount = 16u, groupsize = 512u;
	uint32_t histogramTableCount = bucketCount * (((uint32_t)max_count + groupsize - 1u) / groupsize);
	return histogramTableCount + orc_prefix_sum_helper_length(histogramTableCount);
}

/* algorithms.cpp:59-91 with radix_sort_apply_on_block_level.comp / radix_sort_scattered_write.comp:
 * per 4-bit digit a stable split; the number of passes depends only on upper_bound. */
void orc_sort(const uint32_t* keys, const uint32_t* vals, uint32_t n, uint32_t upper_bound,
              uint32_t* out_keys, uint32_t* out_vals)
{
	uint32_t* k0 = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t* v0 = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t* k1 = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t* v1 = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	memcpy(k0, keys, sizeof(uint32_t) * n);
	memcpy(v0, vals, sizeof(uint32_t) * n);
	for (uint32_t off = 0u; off < 32u && (upper_bound >> off) != 0u; off += 4u) {
		uint32_t count[16] = { 0 }, start[16];
		for (uint32_t i = 0; i < n; i++) count[(k0[i] >> off) & 15u]++;
		uint32_t sum = 0;
		for (int b = 0; b < 16; b++) { start[b] = sum; sum += count[b]; }
		for (uint32_t i = 0; i < n; i++) {
			uint32_t d = (k0[i] >> off) & 15u;
			k1[start[d]] = k0[i];
			v1[start[d]] = v0[i];
			start[d]++;
		}
		uint32_t* t;
		t = k0; k0 = k1; k1 = t;
		t = v0; v0 = v1; v1 = t;
	}
	memcpy(out_keys, k0, sizeof(uint32_t) * n);
	memcpy(out_vals, v0, sizeof(uint32_t) * n);
	free(k0); free(v0); free(k1); free(v1);
}

void orc_prefix_sum(const uint32_t* in, uint32_t n, uint32_t* out) /* algorithms.cpp:93-118: inclusive */
{
	uint32_t sum = 0u;
	for (uint32_t i = 0; i < n; i++) { sum += in[i]; out[i] = sum; }
}

void orc_apply_edit(const void* src, void* dst, const uint32_t* edit, uint32_t n, uint32_t stride)
{ /* gpu_list.h:126-135 + copy_scattered_read.comp:21-30 */
	const char* s = (const char*)src;
	char* d = (char*)dst;
	for (uint32_t i = 0; i < n; i++) memcpy(d + (size_t)i * stride, s + (size_t)edit[i] * stride, stride);
}

/* ---------------------------------------------------------------------------------- */
/* position keys                                                                      */
/* ---------------------------------------------------------------------------------- */
/* map_pos_to_grid of calculate_position_hash.comp:23-26 / neighborhood_green.comp:35-37 (without the
 * DIMENSIONS mask, which the caller applies) */
static inline void map_to_grid(const float p[3], const float mn[3], const float mx[3], uint32_t res, uint32_t g[3])
{
	float scale = (float)(1u << res);
	for (int d = 0; d < 3; d++) g[d] = f2u((p[d] - mn[d]) / (mx[d] - mn[d]) * scale);
}

static inline uint32_t zhash(const uint32_t g[3], uint32_t res, int D) /* calculate_position_hash.comp:29-36 */
{
	uint32_t result = 0u;
	for (uint32_t i = 0u; i < res; i++)
		for (int d = 0; d < D; d++) result += ((g[d] >> i) & 1u) << (i * (uint32_t)D + (uint32_t)d);
	return result;
}

void orc_position_hash(const int32_t* pos4, uint32_t n, const float mn[3], const float mx[3], uint32_t res, int D,
                       uint32_t* out_hash)
{
	for (uint32_t id = 0; id < n; id++) {
		float p[3] = { (float)pos4[4 * id] / R_POS, (float)pos4[4 * id + 1] / R_POS, (float)pos4[4 * id + 2] / R_POS };
		uint32_t g[3];
		map_to_grid(p, mn, mx, res, g);
		out_hash[id] = zhash(g, res, D);
	}
}

/* calculate_position_code.comp:23-62: bit k of x -> bit 3k, y -> 3k+1, z -> 3k+2 of a 96-bit code */
static inline void encode96(const int32_t ipos[3], uint32_t out[3])
{
	out[0] = out[1] = out[2] = 0u;
	for (uint32_t k = 0; k < 32u; k++)
		for (uint32_t d = 0; d < 3u; d++) {
			uint32_t bit = 3u * k + d;
			if (((uint32_t)ipos[d] >> k) & 1u) out[bit >> 5] += 1u << (bit & 31u);
		}
}

void orc_position_code(const uint32_t* index_list, const int32_t* pos4, uint32_t n, uint32_t section, uint32_t* out)
{
	for (uint32_t id = 0; id < n; id++) {
		uint32_t idx = index_list[id];
		uint32_t c[3];
		encode96(&pos4[4 * idx], c);
		out[id] = c[section];
	}
}

void orc_find_value_ranges(const uint32_t* index_list, const uint32_t* values, uint32_t n,
                           uint32_t* range_start, uint32_t* range_end) /* find_value_ranges.comp:16-31 */
{
	for (uint32_t id = 0; id < n; id++) {
		uint32_t curr = values[index_list[id]];
		/* id == 0 reads index_list[-1] in the reference (UB); the outcome only matters through
		 * `id == startId ||`, so treat prev as "different" for the start and as "equal" for the end write. */
		int differs = id > 0 && curr != values[index_list[id - 1]];
		if (id == 0 || differs) range_start[curr] = id;
		if (differs) range_end[values[index_list[id - 1]]] = id;
		if (id == n - 1u) range_end[curr] = id + 1u;
	}
}

/* ---------------------------------------------------------------------------------- */
/* neighbour searches                                                                 */
/* ---------------------------------------------------------------------------------- */
typedef struct { uint32_t* pairs; uint32_t cap; uint32_t len; } pair_sink;

static inline void add_pair(pair_sink* s, uint32_t id, uint32_t idN) /* neighbor_add.glsl:11-25 (unsorted layout) */
{
	uint32_t i = s->len++;
	if (i < s->cap) { s->pairs[2 * i] = id; s->pairs[2 * i + 1] = idN; }
	else s->len = s->cap;
}

static inline void posf(const int32_t* pos4, uint32_t idx, float p[3])
{
	p[0] = (float)pos4[4 * idx] / R_POS; p[1] = (float)pos4[4 * idx + 1] / R_POS; p[2] = (float)pos4[4 * idx + 2] / R_POS;
}

static inline float distance3(const float a[3], const float b[3])
{
	float d[3] = { a[0] - b[0], a[1] - b[1], a[2] - b[2] };
	return length3(d);
}

/* neighborhood_green.comp:50-87 for ONE invocation (particle id).  out == NULL counts; otherwise the pairs are written at
 * out[2 * (base + k)] for k < found while base + k < cap (neighbor_add.glsl:23-24 clamps the list at its capacity). */
static uint32_t green_one(const uint32_t* index_list, const int32_t* pos4, const float* range, const uint32_t* cell_start,
                          const uint32_t* cell_end, uint32_t id, float range_scale, const float mn[3], const float mx[3],
                          uint32_t res, int D, uint32_t* out, uint32_t base, uint32_t cap)
{
	float r = range[id] * range_scale;
	uint32_t idx = index_list[id];
	float pos[3], lo[3], hi[3];
	uint32_t found = 0u;
	posf(pos4, idx, pos);
	for (int d = 0; d < 3; d++) { lo[d] = pos[d] - r; hi[d] = pos[d] + r; }
	uint32_t gmin[3], gmax[3];
	map_to_grid(lo, mn, mx, res, gmin);
	map_to_grid(hi, mn, mx, res, gmax);
	if (D < 2) { gmin[1] = gmax[1] = 0u; }
	if (D < 3) { gmin[2] = gmax[2] = 0u; }
	/* the reference's cell walk: x fastest, then y, then z (neighborhood_green.comp:69-79) */
	for (uint32_t cz = gmin[2]; ; cz++) {
		for (uint32_t cy = gmin[1]; ; cy++) {
			for (uint32_t cx = gmin[0]; ; cx++) {
				uint32_t cell[3] = { cx, cy, cz };
				uint32_t h = zhash(cell, res, D);
				for (uint32_t idN = cell_start[h]; idN < cell_end[h]; idN++) {
					float posN[3];
					posf(pos4, index_list[idN], posN);
					if (id == idN || distance3(pos, posN) > r) continue;
					if (out && base + found < cap) { out[2 * (size_t)(base + found)] = id; out[2 * (size_t)(base + found) + 1] = idN; }
					found++;
				}
				if (cx >= gmax[0]) break;
			}
			if (cy >= gmax[1]) break;
		}
		if (cz >= gmax[2]) break;
	}
	return found;
}

/* One shader invocation per particle appends its pairs with an atomic counter (neighbor_add.glsl:11-25): the order of the list
 * across invocations is whatever the schedule gives.  The restatement returns the order of a sequential run (ascending id, each
 * id's pairs in its own discovery order).  With more than one host thread the invocations run in parallel in two passes --
 * count per id, running sum, every id writes its own segment -- which yields exactly that list. */
uint32_t orc_neighborhood_green_pairs(const uint32_t* index_list, const int32_t* pos4, const float* range,
                                      const uint32_t* cell_start, const uint32_t* cell_end, uint32_t n,
                                      float range_scale, const float mn[3], const float mx[3],
                                      uint32_t res, int D, uint32_t* out_pairs, uint32_t cap)
{ /* neighborhood_green.comp:50-87 */
	if (g_threads <= 1) {
		uint32_t len = 0u;
		for (uint32_t id = 0; id < n; id++) {
			uint32_t f = green_one(index_list, pos4, range, cell_start, cell_end, id, range_scale, mn, mx, res, D, out_pairs, len, cap);
			len = (f > cap - len) ? cap : len + f; /* saturates at the capacity like the clamp */
		}
		return len;
	}
	uint64_t* off = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)n + 1));
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 256)
	for (uint32_t id = 0; id < n; id++)
		off[id + 1] = green_one(index_list, pos4, range, cell_start, cell_end, id, range_scale, mn, mx, res, D, NULL, 0u, 0u);
	off[0] = 0u;
	for (uint32_t id = 0; id < n; id++) off[id + 1] += off[id];
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 256)
	for (uint32_t id = 0; id < n; id++)
		if (off[id] < cap) green_one(index_list, pos4, range, cell_start, cell_end, id, range_scale, mn, mx, res, D, out_pairs, (uint32_t)off[id], cap);
	uint32_t len = off[n] > cap ? cap : (uint32_t)off[n];
	free(off);
	return len;
}

uint32_t orc_neighborhood_brute_force_pairs(const uint32_t* index_list, const int32_t* pos4, const float* range,
                                            uint32_t n, float range_scale, uint32_t* out_pairs, uint32_t cap)
{ /* neighborhood_brute_force.comp:31-50 */
	pair_sink sink = { out_pairs, cap, 0u };
	for (uint32_t id = 0; id < n; id++) {
		float r = range[id] * range_scale;
		float pos[3];
		posf(pos4, index_list[id], pos);
		for (uint32_t idN = 0; idN < n; idN++) {
			float posN[3];
			posf(pos4, index_list[idN], posN);
			if (id == idN || distance3(pos, posN) > r) continue;
			add_pair(&sink, id, idN);
		}
	}
	return sink.len;
}

/* --- 96-bit helpers of neighborhood_binary_search.comp:79-113 --- */
typedef struct { uint32_t v[3]; } u96;
static inline u96 u96_make(uint32_t a, uint32_t b, uint32_t c) { u96 r = { { a, b, c } }; return r; }
static inline u96 plus96(u96 a, u96 b)
{
	u96 r = u96_make(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]);
	int y = r.v[0] < a.v[0];
	int z = r.v[1] < a.v[1] || (y && r.v[1] == 0xFFFFFFFFu);
	r.v[1] += y ? 1u : 0u; r.v[2] += z ? 1u : 0u;
	return r;
}
static inline u96 minus96(u96 a, u96 b)
{
	int y = a.v[0] < b.v[0];
	int z = a.v[1] < b.v[1] || (y && a.v[1] == b.v[1]);
	return u96_make(a.v[0] - b.v[0], a.v[1] - b.v[1] - (y ? 1u : 0u), a.v[2] - b.v[2] - (z ? 1u : 0u));
}
static inline u96 leftShift96(u96 a, uint32_t distance)
{
	u96 r = distance < 32u ? u96_make(a.v[0] << distance, a.v[1] << distance, a.v[2] << distance) : u96_make(0, 0, 0);
	if (0u < distance && distance <= 32u) { /* a >> 32 is UB in C; GLSL shift by 32 is undefined too, callers use distance<=2 */
		r.v[1] |= distance == 32u ? a.v[0] : a.v[0] >> (32u - distance);
		r.v[2] |= distance == 32u ? a.v[1] : a.v[1] >> (32u - distance);
	}
	if (32u < distance && distance < 64u) { r.v[1] |= a.v[0] << (distance - 32u); r.v[2] |= a.v[1] << (distance - 32u); }
	if (32u < distance && distance <= 64u) r.v[2] |= distance == 64u ? a.v[0] : a.v[0] >> (64u - distance);
	if (64u < distance && distance < 96u) r.v[2] |= a.v[0] << (distance - 64u);
	return r;
}
static inline int greater96(u96 a, u96 b)
{
	if (a.v[2] > b.v[2]) return 1;
	if (a.v[2] < b.v[2]) return 0;
	if (a.v[1] > b.v[1]) return 1;
	if (a.v[1] < b.v[1]) return 0;
	return a.v[0] > b.v[0];
}
static inline u96 and96(u96 a, u96 b) { return u96_make(a.v[0] & b.v[0], a.v[1] & b.v[1], a.v[2] & b.v[2]); }
static inline u96 or96(u96 a, u96 b) { return u96_make(a.v[0] | b.v[0], a.v[1] | b.v[1], a.v[2] | b.v[2]); }
static inline u96 not96(u96 a) { return u96_make(~a.v[0], ~a.v[1], ~a.v[2]); }

/* lower_bound over the sorted 96-bit codes; the reference's block-local shortcut
 * (neighborhood_binary_search.comp:119-126) only narrows the initial interval, the result is the same
 * first index with code >= query */
static uint32_t lower_bound96(const uint32_t* c0, const uint32_t* c1, const uint32_t* c2, uint32_t n, u96 code)
{
	uint32_t lo = 0u, hi = n;
	while (lo < hi) {
		uint32_t mid = lo + (hi - lo) / 2u;
		if (greater96(code, u96_make(c0[mid], c1[mid], c2[mid]))) lo = mid + 1u; else hi = mid;
	}
	return lo;
}

/* neighborhood_binary_search.comp:166-276 for ONE invocation; out == NULL counts (see green_one) */
static uint32_t bsearch_one(const uint32_t* index_list, const int32_t* pos4, const uint32_t* c0, const uint32_t* c1, const uint32_t* c2,
                            const float* range, uint32_t n, uint32_t id, float range_scale, uint32_t* out, uint32_t base, uint32_t cap)
{
	const u96 xMask3 = u96_make(011111111111u, 022222222222u, 04444444444u);
	const u96 yMask3 = u96_make(022222222222u, 04444444444u, 011111111111u);
	const u96 zMask3 = u96_make(04444444444u, 011111111111u, 022222222222u);
	uint32_t found = 0u;
	float r = range[id] * range_scale;
	uint32_t idx = index_list[id];
	const int32_t* iPos = &pos4[4 * idx];
	u96 code;
	encode96(iPos, code.v);
	uint32_t digits = f2u(ceilf(log2f(r * R_POS))) * 3u;
	u96 mask;
	mask.v[0] = (digits < 32u ? 1u << digits : 0u) - 1u;
	mask.v[1] = (digits < 64u ? 1u << ((digits > 32u ? digits : 32u) - 32u) : 0u) - 1u;
	mask.v[2] = (digits < 96u ? 1u << ((digits > 64u ? digits : 64u) - 64u) : 0u) - 1u;
	u96 center = u96_make(code.v[0] - (code.v[0] & mask.v[0]), code.v[1] - (code.v[1] & mask.v[1]), code.v[2] - (code.v[2] & mask.v[2]));
	u96 step = plus96(mask, u96_make(1u, 0u, 0u));
	u96 xs[3], ys[3], zs[3];
	xs[0] = and96(minus96(and96(center, xMask3), step), xMask3);
	xs[2] = and96(plus96(or96(center, not96(xMask3)), step), xMask3);
	ys[0] = and96(minus96(and96(center, yMask3), leftShift96(step, 1u)), yMask3);
	ys[2] = and96(plus96(or96(center, not96(yMask3)), leftShift96(step, 1u)), yMask3);
	zs[0] = and96(minus96(and96(center, zMask3), leftShift96(step, 2u)), zMask3);
	zs[2] = and96(plus96(or96(center, not96(zMask3)), leftShift96(step, 2u)), zMask3);
	xs[1] = and96(center, xMask3);
	ys[1] = and96(center, yMask3);
	zs[1] = and96(center, zMask3);
	float pos[3] = { (float)iPos[0] / R_POS, (float)iPos[1] / R_POS, (float)iPos[2] / R_POS };
	for (int cz = 0; cz < 3; cz++) for (int cy = 0; cy < 3; cy++) for (int cx = 0; cx < 3; cx++) {
		u96 cellCode = or96(or96(xs[cx], ys[cy]), zs[cz]);
		u96 cellLast = or96(cellCode, mask);
		uint32_t idN = lower_bound96(c0, c1, c2, n, cellCode);
		for (; idN < n; idN++) {
			if (greater96(u96_make(c0[idN], c1[idN], c2[idN]), cellLast)) break;
			float posN[3];
			posf(pos4, index_list[idN], posN);
			if (id != idN && distance3(pos, posN) <= r) {
				if (out && base + found < cap) { out[2 * (size_t)(base + found)] = id; out[2 * (size_t)(base + found) + 1] = idN; }
				found++;
			}
		}
	}
	return found;
}

uint32_t orc_neighborhood_binary_search_pairs(const uint32_t* index_list, const int32_t* pos4,
                                              const uint32_t* c0, const uint32_t* c1, const uint32_t* c2,
                                              const float* range, uint32_t n, float range_scale,
                                              uint32_t* out_pairs, uint32_t cap)
{ /* neighborhood_binary_search.comp:166-276 (DIMENSIONS forced to 3, :4-5); list order as in orc_neighborhood_green_pairs */
	if (g_threads <= 1) {
		uint32_t len = 0u;
		for (uint32_t id = 0; id < n; id++) {
			uint32_t f = bsearch_one(index_list, pos4, c0, c1, c2, range, n, id, range_scale, out_pairs, len, cap);
			len = (f > cap - len) ? cap : len + f;
		}
		return len;
	}
	uint64_t* off = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)n + 1));
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 256)
	for (uint32_t id = 0; id < n; id++) off[id + 1] = bsearch_one(index_list, pos4, c0, c1, c2, range, n, id, range_scale, NULL, 0u, 0u);
	off[0] = 0u;
	for (uint32_t id = 0; id < n; id++) off[id + 1] += off[id];
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 256)
	for (uint32_t id = 0; id < n; id++)
		if (off[id] < cap) bsearch_one(index_list, pos4, c0, c1, c2, range, n, id, range_scale, out_pairs, (uint32_t)off[id], cap);
	uint32_t len = off[n] > cap ? cap : (uint32_t)off[n];
	free(off);
	return len;
}

/* ---------------------------------------------------------------------------------- */
/* the reorder chain: hidden arrays gathered by sorted_index, the all-fluid index list
 * becomes the identity again after indexed_list::sort, per-id fluid arrays follow their particles.
 * General index lists (subset / permuted) are handled too: new list = sorted new slots of its
 * members, per-id arrays permuted along (indexed_list.h:289-308 + :276-286).            */
/* ---------------------------------------------------------------------------------- */
static void reorder_state(orc_state* st, const uint32_t* sorted_index)
{
	uint32_t nh = st->n_hidden, n = st->n;
	void* tmp = malloc((size_t)(nh ? nh : 1) * 16);
#define GATHER_HIDDEN(ptr, stride) do { orc_apply_edit(ptr, tmp, sorted_index, nh, stride); memcpy(ptr, tmp, (size_t)nh * (stride)); } while (0)
	GATHER_HIDDEN(st->position, 16);
	GATHER_HIDDEN(st->velocity, 16);
	GATHER_HIDDEN(st->inverse_mass, 4);
	GATHER_HIDDEN(st->radius, 4);
	GATHER_HIDDEN(st->pos_backup, 16);
	GATHER_HIDDEN(st->transferring, 4);
#undef GATHER_HIDDEN
	free(tmp);
	/* inverse permutation: old hidden slot -> new hidden slot */
	uint32_t* inv = (uint32_t*)malloc(sizeof(uint32_t) * (nh ? nh : 1));
	for (uint32_t k = 0; k < nh; k++) inv[sorted_index[k]] = k;
	/* new index values with their old list position as payload, then sort ascending (stable) */
	uint32_t* nk = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t* nv = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t* sk = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t* sv = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	for (uint32_t i = 0; i < n; i++) { nk[i] = inv[st->index_list[i]]; nv[i] = i; }
	orc_sort(nk, nv, n, 0xFFFFFFFFu, sk, sv);
	memcpy(st->index_list, sk, sizeof(uint32_t) * n);
	float* tf = (float*)malloc(sizeof(float) * (n ? n : 1));
#define GATHER_ID(ptr) do { orc_apply_edit(ptr, tf, sv, n, 4); memcpy(ptr, tf, (size_t)n * 4); } while (0)
	GATHER_ID(st->target_radius);
	GATHER_ID(st->kernel_width);
	GATHER_ID(st->boundariness);
	GATHER_ID(st->boundary_distance);
#undef GATHER_ID
	free(tf); free(inv); free(nk); free(nv); free(sk); free(sv);
}

uint32_t orc_neighborhood_green_apply(orc_state* st, const orc_settings* s, int D, float range_scale,
                                      const float mn[3], const float mx[3], uint32_t res,
                                      uint32_t* out_pairs, uint32_t cap,
                                      uint32_t* sorted_hash, uint32_t* sorted_index,
                                      uint32_t* cell_start, uint32_t* cell_end)
{ /* neighborhood_green.cpp:27-77 */
	(void)s;
	uint32_t nh = st->n_hidden;
	uint32_t maxHash = 1u << (res * (uint32_t)D);
	uint32_t* hash = (uint32_t*)malloc(sizeof(uint32_t) * (nh ? nh : 1));
	uint32_t* iota = (uint32_t*)malloc(sizeof(uint32_t) * (nh ? nh : 1));
	uint32_t* shash = sorted_hash ? sorted_hash : (uint32_t*)malloc(sizeof(uint32_t) * (nh ? nh : 1));
	uint32_t* sidx = sorted_index ? sorted_index : (uint32_t*)malloc(sizeof(uint32_t) * (nh ? nh : 1));
	uint32_t* cs = cell_start ? cell_start : (uint32_t*)malloc(sizeof(uint32_t) * maxHash);
	uint32_t* ce = cell_end ? cell_end : (uint32_t*)malloc(sizeof(uint32_t) * maxHash);
	for (uint32_t i = 0; i < nh; i++) iota[i] = i;                         /* :51 */
	orc_position_hash(st->position, nh, mn, mx, res, D, hash);              /* :52 */
	orc_sort(hash, iota, nh, maxHash, shash, sidx);                         /* :53 */
	reorder_state(st, sidx);                                                /* :54-56 */
	memset(cs, 0, sizeof(uint32_t) * maxHash);                              /* :61-62 */
	memset(ce, 0, sizeof(uint32_t) * maxHash);
	orc_find_value_ranges(st->index_list, shash, st->n, cs, ce);            /* :63 */
	uint32_t np = orc_neighborhood_green_pairs(st->index_list, st->position, st->kernel_width, cs, ce, st->n,
	                                           range_scale, mn, mx, res, D, out_pairs, cap); /* :64 */
	free(hash); free(iota);
	if (!sorted_hash) free(shash);
	if (!sorted_index) free(sidx);
	if (!cell_start) free(cs);
	if (!cell_end) free(ce);
	return np;
}

uint32_t orc_neighborhood_binary_search_apply(orc_state* st, const orc_settings* s, float range_scale,
                                              uint32_t* out_pairs, uint32_t cap,
                                              uint32_t* code0, uint32_t* code1, uint32_t* code2, uint32_t* sorted_index)
{ /* neighborhood_binary_search.cpp:22-75 */
	(void)s;
	uint32_t nh = st->n_hidden;
	size_t bytes = sizeof(uint32_t) * (nh ? nh : 1);
	uint32_t* unsortedIdx = (uint32_t*)malloc(bytes);
	uint32_t* sortedIdx = sorted_index ? sorted_index : (uint32_t*)malloc(bytes);
	uint32_t* codeU = (uint32_t*)malloc(bytes);
	uint32_t* codeS = (uint32_t*)malloc(bytes);
	uint32_t* c[3] = { code0 ? code0 : (uint32_t*)malloc(bytes), code1 ? code1 : (uint32_t*)malloc(bytes), code2 ? code2 : (uint32_t*)malloc(bytes) };
	for (uint32_t i = 0; i < nh; i++) unsortedIdx[i] = i;                    /* :45 */
	for (uint32_t sec = 0; sec < 3u; sec++) {                                /* :47-51: three stable 32-bit sorts, LSD */
		orc_position_code(unsortedIdx, st->position, nh, sec, codeU);
		orc_sort(codeU, unsortedIdx, nh, 0xFFFFFFFFu, codeS, sortedIdx);
		memcpy(unsortedIdx, sortedIdx, bytes);
	}
	reorder_state(st, sortedIdx);                                            /* :53-54 */
	for (uint32_t sec = 0; sec < 3u; sec++) orc_position_code(st->index_list, st->position, st->n, sec, c[sec]); /* :58-60 */
	uint32_t np = orc_neighborhood_binary_search_pairs(st->index_list, st->position, c[0], c[1], c[2], st->kernel_width,
	                                                   st->n, range_scale, out_pairs, cap);  /* :62 */
	free(unsortedIdx); free(codeU); free(codeS);
	if (!sorted_index) free(sortedIdx);
	if (!code0) free(c[0]);
	if (!code1) free(c[1]);
	if (!code2) free(c[2]);
	return np;
}

/* ---------------------------------------------------------------------------------- */
/* incompressibility                                                                  */
/* ---------------------------------------------------------------------------------- */
void orc_incompressibility_0(const orc_state* st, const orc_settings* s, int D, orc_incomp_data* incomp)
{ /* incompressibility_0.comp:33-45 */
	const float zero[3] = { 0.0f, 0.0f, 0.0f };
	for (uint32_t id = 0; id < st->n; id++) {
		uint32_t idx = st->index_list[id];
		float invMass = st->inverse_mass[idx];
		float kw = st->kernel_width[id];
		incomp[id].mWeightedGradSum[0] = incomp[id].mWeightedGradSum[1] = incomp[id].mWeightedGradSum[2] = 0;
		incomp[id].mDensity = f2u(orc_kernel_height(s, D, zero, kw) / invMass * R_INC);
		incomp[id].mSquaredGradSum = 0u;
		incomp[id].padding[0] = incomp[id].padding[1] = incomp[id].padding[2] = 0u;
	}
}

void orc_incompressibility_1(const orc_state* st, const orc_settings* s, int D, const uint32_t* pairs, uint32_t n_pairs,
                             orc_incomp_data* incomp, int32_t* com4, float* grad4)
{ /* incompressibility_1.comp:38-69.  Integer accumulation makes the result independent of pair order, so
   * the OpenMP build uses atomics like the shader does. */
#pragma omp parallel for num_threads(g_threads) schedule(static)
	for (uint32_t e = 0; e < n_pairs; e++) {
		uint32_t n0 = pairs[2 * e], n1 = pairs[2 * e + 1];
		uint32_t idx = st->index_list[n0], idxN = st->index_list[n1];
		const int32_t* pos = &st->position[4 * idx];
		const int32_t* posN = &st->position[4 * idxN];
		float invMassN = st->inverse_mass[idxN];
		float kw = st->kernel_width[n0];
		float diff[3] = { (float)(posN[0] - pos[0]) / R_POS, (float)(posN[1] - pos[1]) / R_POS, (float)(posN[2] - pos[2]) / R_POS };
		float g[3];
		orc_kernel_gradient(s, D, diff, kw, g);
		float wg[3] = { g[0] / invMassN, g[1] / invMassN, g[2] / invMassN };
		uint32_t dDens = f2u(orc_kernel_height(s, D, diff, kw) / invMassN * R_INC);
		int32_t dG[3] = { f2i(wg[0] * R_INC), f2i(wg[1] * R_INC), f2i(wg[2] * R_INC) };
		uint32_t dSq = f2u(dot3(g, g) / invMassN * R_INC);
#pragma omp atomic
		incomp[n0].mDensity += dDens;
#pragma omp atomic
		incomp[n0].mWeightedGradSum[0] += dG[0];
#pragma omp atomic
		incomp[n0].mWeightedGradSum[1] += dG[1];
#pragma omp atomic
		incomp[n0].mWeightedGradSum[2] += dG[2];
#pragma omp atomic
		incomp[n0].mSquaredGradSum += dSq;
		grad4[4 * (size_t)e] = g[0]; grad4[4 * (size_t)e + 1] = g[1]; grad4[4 * (size_t)e + 2] = g[2]; grad4[4 * (size_t)e + 3] = 0.0f;
		if (s->mBoundarinessCalculationMethod == 1) { /* :62-68 */
			float div = invMassN * kw;
			int32_t wd[3] = { f2i((float)(posN[0] - pos[0]) / div), f2i((float)(posN[1] - pos[1]) / div), f2i((float)(posN[2] - pos[2]) / div) };
			int32_t wm = f2i(R_INC / invMassN);
#pragma omp atomic
			com4[4 * n0] += wd[0];
#pragma omp atomic
			com4[4 * n0 + 1] += wd[1];
#pragma omp atomic
			com4[4 * n0 + 2] += wd[2];
#pragma omp atomic
			com4[4 * n0 + 3] += wm;
		}
	}
}

static inline float move_towards_abs(float oldValue, float newValue, float maxStep) /* incompressibility_2.comp:37-41 */
{
	float step = newValue - oldValue;
	return oldValue + fminf_(maxStep, fmaxf_(-maxStep, step));
}

void orc_incompressibility_2(orc_state* st, const orc_settings* s, int D, const orc_incomp_data* incomp,
                             const int32_t* com4, float* lambda)
{ /* incompressibility_2.comp:72-110 */
#pragma omp parallel for num_threads(g_threads) schedule(static)
	for (uint32_t id = 0; id < st->n; id++) {
		uint32_t idx = st->index_list[id];
		float radius = st->radius[idx];
		float invMass = st->inverse_mass[idx];
		float kernelWidth = st->kernel_width[id];
		float invRestDensity = powf(2.0f * radius, (float)D) * invMass;
		float density = (float)incomp[id].mDensity / R_INC;
		float wgs[3] = { (float)incomp[id].mWeightedGradSum[0] / R_INC, (float)incomp[id].mWeightedGradSum[1] / R_INC, (float)incomp[id].mWeightedGradSum[2] / R_INC };
		float squaredGradSum = (float)incomp[id].mSquaredGradSum / R_INC;
		float selfGrad[3] = { -wgs[0] * invRestDensity, -wgs[1] * invRestDensity, -wgs[2] * invRestDensity };
		float wgs2 = dot3(wgs, wgs);
		float selfGradLength = sqrtf(wgs2) * invRestDensity;
		squaredGradSum += wgs2 * invMass;
		float underpressure = 1.0f - density * invRestDensity;
		if (s->mUpdateBoundariness) { /* compute_boundariness, :43-69 */
			float sgl = selfGradLength * kernelWidth;
			float up = underpressure;
			float b = 0.0f;
			switch (s->mBoundarinessCalculationMethod) {
				case 0:
				case 2:
					sgl *= s->mBoundarinessSelfGradLengthFactor;
					up = fmaxf_(0.0f, up) * s->mBoundarinessUnderpressureFactor;
					b = sgl + up;
					break;
				case 1: {
					float totalMass = (float)com4[4 * id + 3] / R_INC + 1.0f / invMass;
					float c[3] = { (float)com4[4 * id], (float)com4[4 * id + 1], (float)com4[4 * id + 2] };
					float dev = length3(c) / (R_POS * totalMass);
					dev *= s->mBoundarinessSelfGradLengthFactor;
					b = dev;
					break;
				}
			}
			b = b >= 1.0f ? 1.0f : 0.0f;
			b = move_towards_abs(st->boundariness[id], b, s->mBoundarinessAdaptionSpeed);
			st->boundariness[id] = fminf_(1.0f, b);
		}
		float lam = underpressure / (invRestDensity * invRestDensity * (squaredGradSum + 0.01f));
		lam /= powf(2.0f * s->mSmallestTargetRadius, (float)D) / invRestDensity * invMass;
		lambda[id] = lam;
		if (lam >= 0.0f) continue;
		float f = lam * invMass * R_POS;
		st->position[4 * idx] += f2i(selfGrad[0] * f);
		st->position[4 * idx + 1] += f2i(selfGrad[1] * f);
		st->position[4 * idx + 2] += f2i(selfGrad[2] * f);
	}
}

void orc_incompressibility_3(orc_state* st, const orc_settings* s, int D, const uint32_t* pairs, uint32_t n_pairs,
                             const float* grad4, const float* lambda, const orc_incomp_data* incomp)
{ /* incompressibility_3.comp:49-67 */
	(void)D;
#pragma omp parallel for num_threads(g_threads) schedule(static)
	for (uint32_t e = 0; e < n_pairs; e++) {
		uint32_t n0 = pairs[2 * e], n1 = pairs[2 * e + 1];
		uint32_t idxN = st->index_list[n1];
		float lam = lambda[n0];
		float sg[3] = { grad4[4 * (size_t)e], grad4[4 * (size_t)e + 1], grad4[4 * (size_t)e + 2] };
		if (s->mBoundarinessCalculationMethod == 2) { /* filter_boundariness, :41-46 */
			float gs[3] = { (float)incomp[n0].mWeightedGradSum[0], (float)incomp[n0].mWeightedGradSum[1], (float)incomp[n0].mWeightedGradSum[2] };
			float e0[3], nd[3];
			normalize3(gs, e0);
			e0[0] = -e0[0]; e0[1] = -e0[1]; e0[2] = -e0[2];
			normalize3(sg, nd);
			if (dot3(e0, nd) > 0.6f) {
#pragma omp atomic write
				st->boundariness[n0] = 0.0f;
			}
		}
		if (lam >= 0.0f) continue;
		float f = lam * R_POS;
		int32_t d0 = f2i(sg[0] * f), d1 = f2i(sg[1] * f), d2 = f2i(sg[2] * f);
#pragma omp atomic
		st->position[4 * idxN] += d0;
#pragma omp atomic
		st->position[4 * idxN + 1] += d1;
#pragma omp atomic
		st->position[4 * idxN + 2] += d2;
	}
}

void orc_incompressibility_apply(orc_state* st, const orc_settings* s, int D, const uint32_t* pairs, uint32_t n_pairs,
                                 orc_incomp_data* out_incomp, float* out_lambda)
{ /* incompressibility.cpp:12-45 */
	uint32_t n = st->n;
	orc_incomp_data* incomp = out_incomp ? out_incomp : (orc_incomp_data*)malloc(sizeof(orc_incomp_data) * (n ? n : 1));
	float* lambda = out_lambda ? out_lambda : (float*)malloc(sizeof(float) * (n ? n : 1));
	float* grad4 = (float*)malloc(sizeof(float) * 4 * (size_t)(n_pairs ? n_pairs : 1));
	int32_t* com4 = (int32_t*)calloc((size_t)(n ? n : 1) * 4, sizeof(int32_t)); /* :28-32 */
	orc_incompressibility_0(st, s, D, incomp);
	orc_incompressibility_1(st, s, D, pairs, n_pairs, incomp, com4, grad4);
	orc_incompressibility_2(st, s, D, incomp, com4, lambda);
	orc_incompressibility_3(st, s, D, pairs, n_pairs, grad4, lambda, incomp);
	free(grad4); free(com4);
	if (!out_incomp) free(incomp);
	if (!out_lambda) free(lambda);
}

/* ---------------------------------------------------------------------------------- */
/* adaptive kernel width                                                              */
/* ---------------------------------------------------------------------------------- */
static inline float move_towards_rel(float oldValue, float newValue, float maxStep) /* uint_to_float_but_gradual.comp:21-27 */
{
	float dir = signf_(newValue - oldValue);
	float result = oldValue * (1.0f + maxStep * dir);
	int reached = (oldValue < newValue) != (result < newValue);
	return reached ? newValue : result;
}

uint32_t orc_spread_kernel_width_apply(orc_state* st, const orc_settings* s, uint32_t* pairs, uint32_t n_pairs,
                                       uint32_t* out_kw_fixed)
{ /* spread_kernel_width.cpp:12-26 */
	uint32_t n = st->n;
	uint32_t* kwfx = out_kw_fixed ? out_kw_fixed : (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	for (uint32_t id = 0; id < n; id++) { /* kernel_width_init.comp:22-36 */
		uint32_t idx = st->index_list[id];
		float radius = st->radius[idx];
		float targetRadius = st->target_radius[id];
		if (!s->mBaseKernelWidthOnTargetRadius) targetRadius = 0.0f;
		float orig = fmaxf_(radius, targetRadius) * ORC_KERNEL_SCALE;
		kwfx[id] = f2u(orig * ORC_KERNEL_WIDTH_RESOLUTION);
	}
	uint32_t kept = 0u;
	/* kernel_width.comp:27-61, one invocation per pair.  The atomicMax result does not depend on the order; the kept pairs
	 * are returned in the order of the input list (the reference appends them in schedule order).  Parallel form: keep flags,
	 * running sum, scatter -- the same list as the sequential loop. */
	uint8_t* keep = (uint8_t*)malloc(n_pairs ? n_pairs : 1);
#pragma omp parallel for num_threads(g_threads) schedule(static)
	for (uint32_t e = 0; e < n_pairs; e++) {
		uint32_t n0 = pairs[2 * (size_t)e], n1 = pairs[2 * (size_t)e + 1];
		uint32_t idx = st->index_list[n0], idxN = st->index_list[n1];
		const int32_t* pos = &st->position[4 * idx];
		const int32_t* posN = &st->position[4 * idxN];
		float radius = st->radius[idx];
		float targetRadius = st->target_radius[n0];
		float oldKw = st->kernel_width[n0];
		float diff[3] = { (float)(posN[0] - pos[0]) / R_POS, (float)(posN[1] - pos[1]) / R_POS, (float)(posN[2] - pos[2]) / R_POS };
		float dist = length3(diff);
		if (!s->mBaseKernelWidthOnTargetRadius) targetRadius = 0.0f;
		float orig = fmaxf_(radius, targetRadius) * ORC_KERNEL_SCALE;
		float cutoff = fmaxf_(orig, oldKw);
		float distanceFromKernel = dist - orig;
		float influence = fmaxf_(0.0f, 1.0f - fmaxf_(0.0f, distanceFromKernel / (orig * ORC_KERNEL_WIDTH_PROPAGATION_FACTOR)));
		uint32_t v = f2u(orig * influence * ORC_KERNEL_WIDTH_RESOLUTION);
		uint32_t cur = __atomic_load_n(&kwfx[n1], __ATOMIC_RELAXED);           /* atomicMax, kernel_width.comp:53 */
		while (v > cur && !__atomic_compare_exchange_n(&kwfx[n1], &cur, v, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
		keep[e] = dist <= cutoff;                                               /* :57 */
	}
	for (uint32_t e = 0; e < n_pairs; e++)
		if (keep[e]) { pairs[2 * (size_t)kept] = pairs[2 * (size_t)e]; pairs[2 * (size_t)kept + 1] = pairs[2 * (size_t)e + 1]; kept++; }
	free(keep);
	for (uint32_t id = 0; id < n; id++) { /* uint_to_float_but_gradual.comp:29-39; mLowerBound = -inf (shader_provider.h:26) */
		float value = (float)kwfx[id] * (1.0f / ORC_KERNEL_WIDTH_RESOLUTION);
		value = move_towards_rel(st->kernel_width[id], value, s->mKernelWidthAdaptionSpeed);
		st->kernel_width[id] = fmaxf_(value, -INFINITY);
	}
	if (!out_kw_fixed) free(kwfx);
	return kept;
}

/* ---------------------------------------------------------------------------------- */
/* update_transfers::apply with settings::merge and settings::split off               */
/* (update_transfers.cpp:14-54) and the kernel width of the default adaptive mode     */
/* (pool.cpp:77-80)                                                                   */
/* ---------------------------------------------------------------------------------- */
static inline uint32_t pair_dist_units(const orc_state* st, uint32_t n0, uint32_t n1)
{ /* find_split_and_merge_1.comp:25-31 / _2.comp:27-33: uint(length(posN - pos)), integer difference, length in float */
	uint32_t idx = st->index_list[n0], idxN = st->index_list[n1];
	const int32_t* pos = &st->position[4 * idx];
	const int32_t* posN = &st->position[4 * idxN];
	float diff[3] = { (float)(posN[0] - pos[0]), (float)(posN[1] - pos[1]), (float)(posN[2] - pos[2]) };
	return f2u(length3(diff));
}

void orc_update_transfers_apply(orc_state* st, const orc_settings* s, const uint32_t* pairs, uint32_t n_pairs,
                                uint32_t* out_nearest)
{
	uint32_t n = st->n;
	uint32_t* old_bd = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t* min_nd = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	uint32_t* nearest = out_nearest ? out_nearest : (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
	memcpy(old_bd, st->boundary_distance, sizeof(uint32_t) * n);      /* oldBoundaryDistanceList, update_transfers.cpp:32 */
	for (uint32_t id = 0; id < n; id++) {                             /* write_sequence(..., max, 0), :39 and :46 */
		st->boundary_distance[id] = 0xFFFFFFFFu;
		min_nd[id] = 0xFFFFFFFFu;
		nearest[id] = 0xFFFFFFFFu;                                    /* not initialised by the reference */
	}
#pragma omp parallel for num_threads(g_threads) schedule(static)
	for (uint32_t e = 0; e < n_pairs; e++) {                          /* find_split_and_merge_1.comp:21-34 (atomicMin: any order) */
		uint32_t n0 = pairs[2 * (size_t)e], n1 = pairs[2 * (size_t)e + 1];
		uint32_t dist = pair_dist_units(st, n0, n1);
		uint32_t v = old_bd[n1] + dist;                               /* uint arithmetic: wraps like the shader */
		uint32_t cur = __atomic_load_n(&st->boundary_distance[n0], __ATOMIC_RELAXED);
		while (v < cur && !__atomic_compare_exchange_n(&st->boundary_distance[n0], &cur, v, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
		cur = __atomic_load_n(&min_nd[n0], __ATOMIC_RELAXED);
		while (dist < cur && !__atomic_compare_exchange_n(&min_nd[n0], &cur, dist, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
	}
	{ /* find_split_and_merge_2.comp:21-37: every pair at the minimum distance writes; in list order the last one stays.
	   * Parallel form: the largest pair index at the minimum distance per id, then its idN. */
		uint32_t* last = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
		for (uint32_t id = 0; id < n; id++) last[id] = 0xFFFFFFFFu;
#pragma omp parallel for num_threads(g_threads) schedule(static)
		for (uint32_t e = 0; e < n_pairs; e++) {
			uint32_t n0 = pairs[2 * (size_t)e], n1 = pairs[2 * (size_t)e + 1];
			if (pair_dist_units(st, n0, n1) != min_nd[n0]) continue;
			uint32_t cur = __atomic_load_n(&last[n0], __ATOMIC_RELAXED);
			while ((cur == 0xFFFFFFFFu || e > cur) && !__atomic_compare_exchange_n(&last[n0], &cur, e, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
		}
		for (uint32_t id = 0; id < n; id++) if (last[id] != 0xFFFFFFFFu) nearest[id] = pairs[2 * (size_t)last[id] + 1];
		free(last);
	}
	for (uint32_t id = 0; id < n; id++) {                             /* find_split_and_merge_3.comp:55-84 */
		uint32_t idx = st->index_list[id];
		float radius = st->radius[idx];
		float boundaryDistance = (float)st->boundary_distance[id] / R_POS;
		float boundariness = st->boundariness[id];
		float targetRadius;
		if (s->mUpdateTargetRadius) {
			if (s->mBaseKernelWidthOnBoundaryDistance) {
				targetRadius = (s->mTargetRadiusScaleFactor / (ORC_KERNEL_SCALE + ORC_KERNEL_SCALE * s->mTargetRadiusScaleFactor)) * boundaryDistance;
				targetRadius = fmaxf_(targetRadius, s->mSmallestTargetRadius);
			} else {
				targetRadius = s->mSmallestTargetRadius + fmaxf_(0.0f, (boundaryDistance - s->mTargetRadiusOffset) * s->mTargetRadiusScaleFactor);
			}
			st->target_radius[id] = targetRadius;
		}
		boundariness = boundariness >= 1.0f ? 1.0f : 0.0f;
		/* mix(x, y, a) = x * (1 - a) + y * a */
		st->boundary_distance[id] = f2u((boundaryDistance * (1.0f - boundariness) + radius * boundariness) * R_POS);
		st->boundariness[id] = boundariness;
		/* :88-121 merge / split bookkeeping: both switches are off in this scope */
	}
	free(old_bd); free(min_nd);
	if (!out_nearest) free(nearest);
}

/* ---------------------------------------------------------------------------------- */
/* update_transfers::apply with merge and split as the settings say                   */
/* (update_transfers.cpp:14-70) and particle_transfer::apply (particle_transfer.cpp)   */
/*                                                                                    */
/* Serialisation.  The reference decides merges and splits with atomicExchange /      */
/* atomicAdd across threads (find_split_and_merge_3.comp:95-120), so which of two     */
/* conflicting candidates wins and the order of the output lists are races.  The      */
/* restatement runs the invocations in ascending id order -- one of the orders the    */
/* reference may take; list edits (delete_these / duplicate_these, indexed_list.h:    */
/* 126-152) keep the surviving entries in their order and append copies at the end.   */
/* A particle without any pair has no nearest neighbour (the reference reads an       */
/* uninitialised list entry there): it never merges.                                  */
/* `1 / 0` in initialize_split_particles.comp:30 is an INTEGER constant expression;   */
/* glslang folds an integer division by zero to 0x7FFFFFFF, so the inverse mass of a  */
/* fresh duplicate is float(0x7FFFFFFF) = 2^31 (assumption, named in DESIGN.md).      */
/* pow(2.0, 1.0 / DIMENSIONS) * 0.99 (find_split_and_merge_3.comp:88) is folded by    */
/* the compiler in double precision and rounded to float once.                        */
/* ---------------------------------------------------------------------------------- */
/* pow() of the merge / split passes: evaluated in double and rounded to float once (what the CUDA path's pow_rn does).
 * GLSL leaves pow to the driver; glibc's powf is within 0.52 ulp but not always correctly rounded, so "powf on both sides"
 * is not a reproducible convention for x^2, x^(1/2) -- this one is (double rounding is harmless at 53 -> 24 bits here). */
static inline float pow_rn(float x, float y) { return (float)pow((double)x, (double)y); }

static void find_split_and_merge_3_decide(orc_state* st, const orc_settings* s, int D, const uint32_t* nearest,
                                          orc_transfers* t, uint32_t* split_ids, uint32_t* n_split, uint32_t max_split)
{ /* find_split_and_merge_3.comp:86-121, invocations in ascending id order */
	const float splitFactor = (float)(pow(2.0, 1.0 / (double)D) * 0.99);
	uint32_t n0 = st->n;
	*n_split = 0;
	for (uint32_t id = 0; id < n0; id++) {
		uint32_t idx = st->index_list[id];
		float radius = st->radius[idx];
		float targetRadius = st->target_radius[id];
		int split = s->mSplit && (targetRadius * splitFactor <= radius);
		int merge = 0, nnLarger = 0;
		uint32_t nnIdx = 0;
		if (s->mMerge && nearest[id] != 0xFFFFFFFFu) {
			nnIdx = st->index_list[nearest[id]];
			const int32_t* pos = &st->position[4 * idx];
			const int32_t* posN = &st->position[4 * nnIdx];
			float diff[3] = { (float)(posN[0] - pos[0]), (float)(posN[1] - pos[1]), (float)(posN[2] - pos[2]) };
			float nnDist = length3(diff) / R_POS;
			float nnRadius = st->radius[nnIdx];
			nnLarger = nnRadius > radius || (nnRadius == radius && nnIdx > idx);
			merge = nnDist < radius && pow_rn(radius, (float)D) + pow_rn(nnRadius, (float)D) <= pow_rn(targetRadius, (float)D);
		}
		if (!merge && !split) continue;
		uint32_t sourceIdx = (merge && nnLarger) ? nnIdx : idx;
		uint32_t targetIdx = (merge && nnLarger) ? idx : nnIdx;
		if (st->transferring[sourceIdx] == 1u) continue;          /* atomicExchange(..., 1) == 1 */
		st->transferring[sourceIdx] = 1u;
		if (merge) {
			if (st->transferring[targetIdx] == 1u) { st->transferring[sourceIdx] = 0u; continue; }
			st->transferring[targetIdx] = 1u;
			if (t->n >= t->cap) {                                  /* :104-109 */
				t->n = t->cap;
				st->transferring[sourceIdx] = 0u;
				st->transferring[targetIdx] = 0u;
				continue;
			}
			t->source[t->n] = sourceIdx;
			t->target[t->n] = targetIdx;
			t->time_left[t->n] = fmaxf_(0.0001f, s->mMergeDuration);
			t->n++;
		} else {
			if (*n_split >= max_split) { *n_split = max_split; st->transferring[sourceIdx] = 0u; continue; }
			split_ids[(*n_split)++] = id;                          /* outSplit[sIdx] = idx: kept as the id, the idx follows */
		}
	}
}

void orc_update_transfers_full(orc_state* st, uint32_t hidden_cap, uint32_t id_cap, const orc_settings* s, int D,
                               const uint32_t* pairs, uint32_t n_pairs, orc_transfers* t, float split_duration,
                               uint32_t* out_nearest)
{
	uint32_t n0 = st->n;
	uint32_t* nearest = out_nearest ? out_nearest : (uint32_t*)malloc(sizeof(uint32_t) * (n0 ? n0 : 1));
	uint32_t* split_ids = (uint32_t*)malloc(sizeof(uint32_t) * (n0 ? n0 : 1));
	uint32_t n_split = 0;
	orc_update_transfers_apply(st, s, pairs, n_pairs, nearest);              /* update_transfers.cpp:36-48 */
	if (s->mMerge || s->mSplit)
		find_split_and_merge_3_decide(st, s, D, nearest, t, split_ids, &n_split, id_cap); /* :48, mMaxSplitLength = requested_length() */
	if (s->mSplit) {                                                        /* update_transfers.cpp:60-70 */
		/* remove_impossible_splits.comp:33-43 */
		uint32_t newLen = n_split;
		if (t->cap - t->n < newLen) newLen = t->cap - t->n;
		if (hidden_cap - st->n_hidden < newLen) newLen = hidden_cap - st->n_hidden;
		for (uint32_t k = newLen; k < n_split; k++) st->transferring[st->index_list[split_ids[k]]] = 0u;
		n_split = newLen;
		for (uint32_t k = 0; k < n_split; k++) {
			uint32_t id = split_ids[k], idx = st->index_list[id];
			uint32_t h = st->n_hidden + k, nid = st->n + k;
			/* duplicate_these (indexed_list.h:141-152): the copy goes to the end of the hidden list, and every list that
			 * shares the hidden data gains an entry for it -- the fluid's per-id values are copied as well */
			memcpy(&st->position[4 * h], &st->position[4 * idx], 16);
			memcpy(&st->velocity[4 * h], &st->velocity[4 * idx], 16);
			memcpy(&st->pos_backup[4 * h], &st->pos_backup[4 * idx], 16);
			st->inverse_mass[h] = st->inverse_mass[idx];
			st->radius[h] = st->radius[idx];
			st->transferring[h] = st->transferring[idx];
			if (nid < id_cap) {
				st->index_list[nid] = h;
				st->target_radius[nid] = st->target_radius[id];
				st->kernel_width[nid] = st->kernel_width[id];
				st->boundariness[nid] = st->boundariness[id];
				st->boundary_distance[nid] = st->boundary_distance[id];
			}
			/* initialize_split_particles.comp:24-31 */
			st->position[4 * h] += f2i(st->radius[h] * R_POS * 0.1f);
			st->radius[h] = 0.0f;
			st->inverse_mass[h] = 2147483648.0f;                            /* float(1 / 0), see the header of this section */
			/* transferSourceList += splitList etc., update_transfers.cpp:66-68 */
			t->source[t->n + k] = idx;
			t->target[t->n + k] = h;
			t->time_left[t->n + k] = -split_duration;                       /* write_sequence_float(-splitDuration, 0) */
		}
		t->n += n_split;
		st->n_hidden += n_split;
		st->n = st->n + n_split < id_cap ? st->n + n_split : id_cap;
	}
	free(split_ids);
	if (!out_nearest) free(nearest);
}

void orc_particle_transfer_apply(orc_state* st, orc_transfers* t, int D, float dt)
{ /* particle_transfer.cpp:10-28 + particle_transfer.comp:30-84 */
	uint32_t nh = st->n_hidden;
	uint8_t* del_h = (uint8_t*)calloc(nh ? nh : 1, 1);
	uint8_t* del_row = (uint8_t*)calloc(t->n ? t->n : 1, 1);
	const float invD = (float)(1.0 / (double)D);
	for (uint32_t id = 0; id < t->n; id++) {
		float ttl_and_type = t->time_left[id];
		float ttl = fmaxf_(dt, fabsf(ttl_and_type));
		int merge = ttl_and_type > 0.0f;
		uint32_t idS = t->source[id], idT = t->target[id];
		float radiusS = st->radius[idS], radiusT = st->radius[idT];
		float invMassS = st->inverse_mass[idS], invMassT = st->inverse_mass[idT];
		float volS = pow_rn(radiusS, (float)D), volT = pow_rn(radiusT, (float)D);
		float factorS = merge ? fminf_(1.0f, dt / ttl) : dt * (volS - volT) / (2.0f * ttl * volS);
		float transfVol = factorS * volS;
		float normFactor = 1.0f / (invMassS + factorS * invMassT);
		st->inverse_mass[idS] = invMassS / (1.0f - factorS);
		st->inverse_mass[idT] = invMassS * invMassT * normFactor;
		st->radius[idS] = pow_rn(volS - transfVol, invD);
		st->radius[idT] = pow_rn(volT + transfVol, invD);
		t->time_left[id] = (ttl - dt) * (merge ? 1.0f : -1.0f);
		if (ttl == dt) {
			if (merge) del_h[idS] = 1; else del_row[id] = 1;
			st->transferring[idS] = 0u;
			st->transferring[idT] = 0u;
		}
	}
	/* deleteTransferList.delete_these(), then deleteParticleList.delete_these() (particle_transfer.cpp:26-27): the hidden
	 * particle list is compacted; every list sharing it (the fluid's index list with its per-id values, the transfers' source
	 * and target lists with the rows they belong to) loses the entries that pointed at a deleted particle */
	uint32_t* map = (uint32_t*)malloc(sizeof(uint32_t) * (nh ? nh : 1));
	uint32_t w = 0;
	for (uint32_t h = 0; h < nh; h++) {
		map[h] = w;
		if (del_h[h]) continue;
		if (w != h) {
			memcpy(&st->position[4 * w], &st->position[4 * h], 16);
			memcpy(&st->velocity[4 * w], &st->velocity[4 * h], 16);
			memcpy(&st->pos_backup[4 * w], &st->pos_backup[4 * h], 16);
			st->inverse_mass[w] = st->inverse_mass[h];
			st->radius[w] = st->radius[h];
			st->transferring[w] = st->transferring[h];
		}
		w++;
	}
	st->n_hidden = w;
	w = 0;
	for (uint32_t id = 0; id < st->n; id++) {
		uint32_t idx = st->index_list[id];
		if (del_h[idx]) continue;
		st->index_list[w] = map[idx];
		st->target_radius[w] = st->target_radius[id];
		st->kernel_width[w] = st->kernel_width[id];
		st->boundariness[w] = st->boundariness[id];
		st->boundary_distance[w] = st->boundary_distance[id];
		w++;
	}
	st->n = w;
	w = 0;
	for (uint32_t id = 0; id < t->n; id++) {
		if (del_row[id] || del_h[t->source[id]]) continue;
		t->source[w] = map[t->source[id]];
		t->target[w] = map[t->target[id]];
		t->time_left[w] = t->time_left[id];
		w++;
	}
	t->n = w;
	free(del_h); free(del_row); free(map);
}

void orc_transfers_follow_reorder(orc_transfers* t, const uint32_t* sorted_index, uint32_t n_hidden)
{ /* the search permutes the hidden list (hidden'[h] = hidden[sorted_index[h]]); the transfers' source and target lists
     share that hidden data and are re-pointed by indexed_list::apply_hidden_edit (indexed_list.h:289-308) */
	uint32_t* inv = (uint32_t*)malloc(sizeof(uint32_t) * (n_hidden ? n_hidden : 1));
	for (uint32_t h = 0; h < n_hidden; h++) inv[sorted_index[h]] = h;
	for (uint32_t r = 0; r < t->n; r++) { t->source[r] = inv[t->source[r]]; t->target[r] = inv[t->target[r]]; }
	free(inv);
}

void orc_kernel_width_from_boundary_distance(orc_state* st, const orc_settings* s)
{ /* pool.cpp:77-80 -> uint_to_float_with_indexed_lower_bound.comp:32-44 */
	const float factor = s->mTargetRadiusScaleFactor / R_POS, lowerBoundFactor = ORC_KERNEL_SCALE;
	for (uint32_t id = 0; id < st->n; id++) {
		uint32_t idx = st->index_list[id];
		float value = (float)st->boundary_distance[id] * factor;
		float lowerBound = st->radius[idx] * lowerBoundFactor;
		value = move_towards_rel(st->kernel_width[id], value, s->mKernelWidthAdaptionSpeed);
		st->kernel_width[id] = fmaxf_(value, lowerBound);
	}
}

/* ---------------------------------------------------------------------------------- */
/* box collision                                                                      */
/* ---------------------------------------------------------------------------------- */
static inline void hash31(float p, float o[3]) /* box_collision.comp:20-25 */
{
	float p3[3] = { fractf_(p * .1031f), fractf_(p * .1030f), fractf_(p * .0973f) };
	float q[3] = { p3[1] + 33.33f, p3[2] + 33.33f, p3[0] + 33.33f }; /* p3.yzx + 33.33 */
	float d = dot3(p3, q);
	p3[0] += d; p3[1] += d; p3[2] += d;
	/* (p3.xxy + p3.yzz) * p3.zyx */
	o[0] = fractf_((p3[0] + p3[1]) * p3[2]);
	o[1] = fractf_((p3[0] + p3[2]) * p3[1]);
	o[2] = fractf_((p3[1] + p3[2]) * p3[0]);
}

static inline void smallest_component(const float v[3], float m[3]) /* box_collision.comp:27-33 */
{
	int b1 = v[0] <= v[1], b2 = v[0] <= v[2], b3 = v[1] <= v[2];
	m[0] = (b1 && b2) ? 1.0f : 0.0f;
	m[1] = (!b1 && b3) ? 1.0f : 0.0f;
	m[2] = (!b2 && !b3) ? 1.0f : 0.0f;
}

void orc_box_collision(orc_state* st, const float* box_min4, const float* box_max4, uint32_t n_boxes)
{ /* box_collision.comp:36-60; box_collision.cpp:14 skips empty particle lists */
#pragma omp parallel for num_threads(g_threads) schedule(static)
	for (uint32_t id = 0; id < st->n; id++) {
		uint32_t idx = st->index_list[id];
		float pos[3];
		posf(st->position, idx, pos);
		float radius = st->radius[idx];
		for (uint32_t i = 0; i < n_boxes; i++) {
			float h0[3], h1[3], bMin[3], bMax[3], toMin[3], toMax[3], neg[3], m0[3], m1[3];
			hash31((float)id * pos[0] - pos[1] - pos[2], h0);
			hash31((float)id * pos[1] + pos[0] + pos[2], h1);
			for (int d = 0; d < 3; d++) {
				bMin[d] = box_min4[4 * i + d] - radius - h0[d] * 0.05f;
				bMax[d] = box_max4[4 * i + d] + radius + h1[d] * 0.05f;
				toMin[d] = bMin[d] - pos[d];
				toMax[d] = bMax[d] - pos[d];
				neg[d] = -toMin[d];
			}
			smallest_component(neg, m0);
			smallest_component(toMax, m1);
			for (int d = 0; d < 3; d++) { toMin[d] *= m0[d]; toMax[d] *= m1[d]; }
			float distToMin = -((toMin[0] * 1.0f + toMin[1] * 1.0f) + toMin[2] * 1.0f);
			float distToMax = (toMax[0] * 1.0f + toMax[1] * 1.0f) + toMax[2] * 1.0f;
			if (distToMin <= 0.0f || distToMax <= 0.0f) continue;
			const float* sh = distToMin < distToMax ? toMin : toMax;
			pos[0] += sh[0]; pos[1] += sh[1]; pos[2] += sh[2];
		}
		st->position[4 * idx] = f2i(pos[0] * R_POS);
		st->position[4 * idx + 1] = f2i(pos[1] * R_POS);
		st->position[4 * idx + 2] = f2i(pos[2] * R_POS);
	}
}

/* ---------------------------------------------------------------------------------- */
/* velocity handling (velocity_handling.cpp:15-31); mLastDeltaTime == dt for a fixed time step
 * except on the very first call where it is 1.0f -- callers pass the dt to infer with.     */
/* ---------------------------------------------------------------------------------- */
void orc_velocity_handling(orc_state* st, float dt, const float accel[3])
{
	if (dt == 0.0f) { memcpy(st->pos_backup, st->position, sizeof(int32_t) * 4 * (size_t)st->n_hidden); return; }
	float f = dt * R_POS;
	float a[3] = { accel[0] * dt, accel[1] * dt, accel[2] * dt }; /* mAcceleration * aDeltaTime, :28 */
	for (uint32_t id = 0; id < st->n; id++) {
		uint32_t idx = st->index_list[id];
		for (int d = 0; d < 3; d++) { /* infer_velocity.comp:24-32 */
			st->velocity[4 * idx + d] = (float)(st->position[4 * idx + d] - st->pos_backup[4 * idx + d]) / f;
		}
	}
	memcpy(st->pos_backup, st->position, sizeof(int32_t) * 4 * (size_t)st->n_hidden); /* posBackupList = positionList, :27 */
	for (uint32_t id = 0; id < st->n; id++) {
		uint32_t idx = st->index_list[id];
		for (int d = 0; d < 3; d++) {
			st->velocity[4 * idx + d] += a[d];                                   /* apply_acceleration.comp:20-28 */
			st->position[4 * idx + d] += f2i(st->velocity[4 * idx + d] * f);     /* apply_velocity.comp:23-31 */
		}
	}
}

/* ---------------------------------------------------------------------------------- */
/* one substep, pool.cpp:67-106                                                       */
/* ---------------------------------------------------------------------------------- */
uint32_t orc_substep(orc_state* st, const orc_settings* s, const orc_substep_params* p, uint32_t* pairs, uint32_t cap)
{
	const int transfers = p->transfers != NULL && !p->basic_pbf && (s->mMerge || s->mSplit);
	if (p->integrate) orc_velocity_handling(st, p->dt, p->accel);                        /* pool.cpp:71 */
	if (transfers) orc_particle_transfer_apply(st, p->transfers, p->dims, p->dt);        /* pool.cpp:73-75 */
	if (p->update_transfers && !p->basic_pbf && s->mBaseKernelWidthOnBoundaryDistance)
		orc_kernel_width_from_boundary_distance(st, s);                                  /* pool.cpp:77-80 */
	int adaptive = !p->basic_pbf && !s->mBaseKernelWidthOnBoundaryDistance;
	float scale = (p->basic_pbf || s->mBaseKernelWidthOnBoundaryDistance) ? 1.0f : 1.5f; /* pool.cpp:83 */
	uint32_t np;
	uint32_t* sidx = transfers ? (uint32_t*)malloc(sizeof(uint32_t) * (st->n_hidden ? st->n_hidden : 1)) : NULL;
	if (p->use_binary_search) np = orc_neighborhood_binary_search_apply(st, s, scale, pairs, cap, NULL, NULL, NULL, sidx);
	else np = orc_neighborhood_green_apply(st, s, p->dims, scale, p->min_pos, p->max_pos, p->res_log2, pairs, cap, NULL, sidx, NULL, NULL);
	if (transfers) { orc_transfers_follow_reorder(p->transfers, sidx, st->n_hidden); free(sidx); }
	if (adaptive) np = orc_spread_kernel_width_apply(st, s, pairs, np, NULL);            /* pool.cpp:87-89 */
	for (int i = 0; i < p->solver_iterations; i++) {                                     /* pool.cpp:92-95 */
		if (st->n > 0) orc_box_collision(st, p->box_min4, p->box_max4, p->n_boxes);
		orc_incompressibility_apply(st, s, p->dims, pairs, np, NULL, NULL);
	}
	if (transfers)                                                                       /* pool.cpp:99-102 */
		orc_update_transfers_full(st, p->hidden_cap, p->hidden_cap, s, p->dims, pairs, np, p->transfers, p->split_duration, NULL);
	else if (p->update_transfers && !p->basic_pbf) orc_update_transfers_apply(st, s, pairs, np, NULL);
	return np;
}
