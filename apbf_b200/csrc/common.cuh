// common.cuh -- context, scratch memory and small device helpers shared by all translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/apbf_b200.h"

#define APBF_NUM_SMS_DEFAULT 148

// named scratch slots (grown on demand, never shrunk: the steady state allocates nothing)
enum apbf_scratch_slot {
	SLOT_SORT_KEYS_A = 0, SLOT_SORT_KEYS_B, SLOT_SORT_VALS_A, SLOT_SORT_VALS_B, SLOT_SORT_HIST, SLOT_SORT_STATUS,
	SLOT_SCAN_STATUS, SLOT_MISC_WORDS, SLOT_CELL_START, SLOT_CELL_END, SLOT_COUNTS, SLOT_OFFSETS, SLOT_NB,
	SLOT_INV_PERM, SLOT_TMP_KEYS, SLOT_TMP_VALS, SLOT_TMP_VALS2, SLOT_CODE0, SLOT_CODE1, SLOT_CODE2,
	SLOT_P4, SLOT_L4, SLOT_G4, SLOT_DELTA, SLOT_PUSH, SLOT_RADIUS_ID, SLOT_KWFX, SLOT_KEEP_COUNTS, SLOT_KEEP_OFFSETS,
	SLOT_PAIRS_TMP, SLOT_NB_TMP, SLOT_BOXES, SLOT_HIDDEN_FLAGS, SLOT_HIDDEN_OFFS, SLOT_IDENTITY, SLOT_COM4, SLOT_KG, SLOT_KH, SLOT_Q4, SLOT_KEY_ID, SLOT_EDIT_COUNTS, SLOT_EDIT_OFFSETS, SLOT_MG_INV, SLOT_I4, SLOT_CUTOFF, SLOT_QB4, SLOT_STREAM, SLOT_TILE_FIRST, SLOT_TILE_TOTAL, SLOT_OLD_BOUNDARY_DIST, SLOT_CELL_MAXW, SLOT_MG_SEND, SLOT_MG_RECV,
	SLOT_TM_NEAREST, SLOT_TM_FLAGS, SLOT_TM_OFFS, SLOT_TM_SRC, SLOT_TM_TGT, SLOT_TM_CID, SLOT_TM_CSRC, SLOT_TM_CTGT, SLOT_TM_STATE, SLOT_TM_RANK,
	SLOT_TM_OWNER, SLOT_TM_WORDS, SLOT_TM_KEEP_H, SLOT_TM_OFFS_H, SLOT_TM_PERM_H, SLOT_TM_KEEP_I, SLOT_TM_OFFS_I, SLOT_TM_PERM_I, SLOT_TM_KEEP_R,
	SLOT_TM_OFFS_R, SLOT_TM_PERM_R, SLOT_KP, SLOT_PL, SLOT_MG_TILES,
	SLOT_COUNT
};

struct apbf_scratch {
	void*  ptr = nullptr;
	size_t bytes = 0;
};

struct apbf_pool_block {
	void*  ptr;
	size_t bytes;
	bool   in_use;
};

// words in the SLOT_MISC_WORDS buffer (device-side scalars)
enum apbf_misc_word {
	MW_FLAGS = 0,        // sticky status bits (bit 0: neighbour list overflow)
	MW_N_ASYM = 1,       // number of pairs (a,b) without a mirrored pair (b,a) in the current neighbour list
	MW_TOTAL_PAIRS = 2,  // unclamped pair count of the last search
	MW_IDENTITY = 3,     // 1 if the index list is the identity over all hidden particles
	MW_KEPT_PAIRS = 4,   // pair count after the last spread_kernel_width prune
	MW_OCC_CELLS = 5,    // number of occupied grid cells seen by the last Green search
	MW_SNAPSHOT = 6,     // scratch word of the list helpers
	MW_N_OWNED = 7,      // multi-GPU: ids >= this are ghost particles (0xFFFFFFFF: none)
	MW_TICKET0 = 8,      // tile tickets of the chained-scan kernels (8 words)
	MW_SCAN_TOTAL = 16,
	MW_GID_BASE = 17,    // multi-GPU: global id of local id 0 (box_collision hashes the id)
	MW_EMIT_TICKET0 = 18, // tile tickets of the pair emit (count pass, fill pass)
	MW_EMIT_TICKET1 = 19,
	MW_STREAM_CURSOR = 20,   // block allocator of the pair emit's hit stream
	MW_PMAX = 23,            // binary search fused with the spread: largest |coordinate| of the list (float bits)
	MW_MAX_INIT = 22,        // fused search + spread: largest initial kernel width (fixed point) over all particles
	MW_STREAM_OVERFLOW = 21, // the hit stream ran out of blocks (only when the pair list overflows): the two-pass fill takes over
	MW_REBUILD_LEN = 24,     // rebuild of a structure from a foreign pair list: min(list length, capacity)
	MW_H_NONUNIFORM = 25,    // solver constants: 1 if two particles of the list differ in kernel width (bitwise), else 0
	MW_THR_MIN = 26,         // last Green search: smallest / largest bit pattern over all ids of {T, K, U, cutoff} (4 + 4 words): every
	MW_THR_MAX = 30,         //   test threshold is the same for everybody -- hence every kept pair is mirrored -- iff min >= max in all four
	MW_INDEX_NONIDENT = 34,  // re-order of the lists: 1 if the index list going in is NOT the identity over all hidden particles
	MW_WORDS = 64
};

// per-kernel-category device timing (CUDA events on the context stream), off by default
enum apbf_prof_cat {
	PROF_HASH_SORT = 0, PROF_REORDER, PROF_CELL_RANGES, PROF_EMIT_COUNT, PROF_EMIT_SCAN, PROF_EMIT_FILL, PROF_KW_SPREAD,
	PROF_KW_COMPACT, PROF_KW_MISC, PROF_BOX, PROF_DENSITY_LAMBDA, PROF_APPLY_DELTA, PROF_COMMIT, PROF_VELOCITY, PROF_SOLVER_PREPARE, PROF_UPDATE_TRANSFERS,
	PROF_MG_ROUTE, PROF_MG_HALO, PROF_MG_EXCHANGE, PROF_COUNT
};
struct apbf_prof_span { int cat; cudaEvent_t beg, end; };

// a neighbour-list structure (CSR offsets + 4-byte list, neighbors.cuh) that is parked while another pair buffer's
// structure is the active one
struct apbf_nbr_entry {
	const uint32_t* pairs = nullptr; // the public pair buffer the structure describes
	apbf_scratch    offsets, nbl;
	uint32_t*       words = nullptr; // parked device words: [0] = MW_N_ASYM
	uint32_t        n_cap = 0;
	bool            valid = false, public_written = false;
	uint64_t        stamp = 0;
};

struct apbf_ctx {
	bool          prof_on = false;
	std::vector<apbf_prof_span> prof_spans;
	std::vector<cudaEvent_t>    prof_free;
	double        prof_ms[PROF_COUNT] = {};
	uint64_t      prof_calls[PROF_COUNT] = {};
	int           prof_open = -1;
	int           device = 0;
	cudaStream_t  stream = nullptr;
	int           num_sms = APBF_NUM_SMS_DEFAULT;
	int           dims = 3;
	apbf_settings settings;
	uint64_t      launches = 0;
	std::string   last_error;
	apbf_scratch  scratch[SLOT_COUNT];
	std::vector<apbf_pool_block> pool;
	// The ACTIVE neighbour-list structure (SLOT_OFFSETS, SLOT_NB, misc[MW_N_ASYM]) belongs to this public pair buffer; the
	// structures of other pair buffers are parked in nbr_cache (neighbors.cuh: apbf_nbr_activate / apbf_nbr_ensure).
	const uint32_t* nbr_struct_pairs = nullptr;
	uint32_t        nbr_struct_n_cap = 0;
	bool            nbr_valid = false;   // offsets + NB describe the list in nbr_struct_pairs
	bool            nbr_public = false;  // ... and the public (id, idN) list has been written as well
	uint64_t        nbr_clock = 0;
	std::vector<apbf_nbr_entry> nbr_cache;
	uint64_t        scratch_epoch = 0;   // bumped whenever a scratch slot is (re)allocated: captured graphs go stale
	// multi-GPU slabs: ghost particles sort into a second key space and are searched through a second cell table
	bool            mg_enabled = false;
	uint32_t        mg_lo[3] = { 0u, 0u, 0u }, mg_hi[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }; // this rank's brick in cells, inclusive
	// a following spread_kernel_width prunes pairs, which can turn a mirrored pair into an unmirrored one: ghosts then
	// keep ALL their pairs onto owned particles until the prune has decided
	bool            mg_ghost_all_pairs = false;
	uint32_t        stream_blocks_cap = 0; // testing aid: upper bound on the hit stream's blocks (0 = automatic)
	uint32_t        match_grid_min = 16384; // merge / split matching: grid-wide rounds from this many candidates on
	bool            search_stats = false; // fused search + spread: count the pairs of the unpruned list as well
	// apbf_sim_step_host: the Green search waits for this event before it re-orders the lists (the small lists are still being
	// uploaded on a second stream while positions are hashed and sorted) and records that one right after (from there on every list
	// but positions, widths and boundariness is final: their download overlaps the solver)
	cudaEvent_t     hook_wait_before_reorder = nullptr, hook_record_after_reorder = nullptr;

	void* scratch_get(int slot, size_t bytes);
	uint32_t* misc() { return (uint32_t*)scratch_get(SLOT_MISC_WORDS, MW_WORDS * sizeof(uint32_t)); }
};

int apbf_fail(apbf_ctx* ctx, int code, const char* what, const char* file, int line);
int  apbf_prof_begin(apbf_ctx* ctx, int cat);   // returns the span's number
void apbf_prof_end(apbf_ctx* ctx, int span);
struct apbf_prof_scope { // RAII: times everything enqueued while it is alive; scopes may nest (the slab sections hold the search's)
	apbf_ctx* c;
	int       span = -1;
	apbf_prof_scope(apbf_ctx* ctx, int cat) : c(ctx) { if (c->prof_on) span = apbf_prof_begin(c, cat); }
	~apbf_prof_scope() { if (span >= 0) apbf_prof_end(c, span); }
};

#define APBF_CUDA(ctx, expr)                                                                   \
	do {                                                                                       \
		cudaError_t e__ = (expr);                                                              \
		if (e__ != cudaSuccess) return apbf_fail(ctx, APBF_ERR_CUDA, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)
#define APBF_REQUIRE(ctx, cond)                                                                \
	do {                                                                                       \
		if (!(cond)) return apbf_fail(ctx, APBF_ERR_INVALID, #cond, __FILE__, __LINE__);       \
	} while (0)
#define APBF_TRY(expr)                                                                         \
	do {                                                                                       \
		int r__ = (expr);                                                                      \
		if (r__ != 0) return r__;                                                              \
	} while (0)
// after a kernel launch
#define APBF_LAUNCHED(ctx)                                                                     \
	do {                                                                                       \
		(ctx)->launches++;                                                                     \
		cudaError_t e__ = cudaGetLastError();                                                  \
		if (e__ != cudaSuccess) return apbf_fail(ctx, APBF_ERR_CUDA, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

static inline unsigned apbf_div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
// grid for a grid-stride kernel over up to `n` items: enough CTAs for n, capped at a multiple of the SM count
static inline unsigned apbf_grid(const apbf_ctx* ctx, size_t n, unsigned block, unsigned ctas_per_sm = 8)
{
	size_t need = (n + block - 1) / block;
	size_t cap = (size_t)ctx->num_sms * ctas_per_sm;
	if (need < 1) need = 1;
	return (unsigned)(need < cap ? need : cap);
}

// ---- device helpers ---------------------------------------------------------------------------------------
// The library is compiled with -fmad=false: a*b+c is never contracted, like the oracle's -ffp-contract=off.
#ifdef __CUDACC__
#define R_POS APBF_POS_RESOLUTION
#define INV_R_POS (1.0f / 262144.0f) // exact power of two: x / 2^18 == x * 2^-18 bit for bit

__device__ __forceinline__ float glsl_min(float x, float y) { return y < x ? y : x; }
__device__ __forceinline__ float glsl_max(float x, float y) { return x < y ? y : x; }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
	return (ax * bx + ay * by) + az * bz;
}
// pow(x, y) rounded once from double precision: equals a correctly rounded powf (what glibc's powf returns, error
// < 0.52 ulp) except on near-ties, whereas CUDA's powf is only good to a few ulp.  Every value derived from it feeds a
// float -> fixed-point truncation, where a last-bit difference becomes a whole unit.
__device__ __forceinline__ float pow_rn(float x, float y) { return (float)pow((double)x, (double)y); }
// The exponents the per-particle constants meet are DIMENSIONS, 2 / DIMENSIONS and DIMENSIONS / 2 with DIMENSIONS 2 or 3.  Where the
// exponent is exact in float (2, 3, 1, 1.5) a product of two or three factors resp. x * sqrt(x) in double precision is within an ulp
// or two of the true power, like pow() itself, so after the ONE rounding to float they agree with it except where the true value
// lies within ~2^-50 relative of a float rounding boundary (probability ~1e-8 per call, the same as CUDA's pow against glibc's).
// 2.0f / 3.0f is NOT 2/3 -- the shader's pow(x, 2.0 / DIMENSIONS) has the rounded exponent, and x^(that) differs from cbrt(x)^2 in
// the last float bit every other time -- so that one stays a pow().  pow() costs ~200 FP64 instructions per call, the products a
// handful (solver_prepare: 62 -> 33 us per 10^6 particles).
__device__ __forceinline__ float pow_int_rn(float x, float d) // x^d for d = DIMENSIONS
{
	const double t = (double)x;
	if (d == 3.0f) return (float)(t * t * t);
	if (d == 2.0f) return (float)(t * t);
	return pow_rn(x, d);
}
__device__ __forceinline__ float pow_2_over_d_rn(float x, float d) // x^(2/d)
{
	if (d == 2.0f) return x;
	return pow_rn(x, 2.0f / d);
}
__device__ __forceinline__ float pow_d_over_2_rn(float x, float d) // x^(d/2)
{
	if (d == 3.0f && x >= 0.0f) { const double t = (double)x; return (float)(t * sqrt(t)); }
	if (d == 2.0f) return x;
	return pow_rn(x, d / 2.0f);
}
__device__ __forceinline__ uint32_t f2u(float f) { return (uint32_t)f; } // cvt.rzi.u32.f32: negative/NaN -> 0, saturating
__device__ __forceinline__ int32_t f2i(float f) { return (int32_t)f; }   // cvt.rzi.s32.f32: NaN -> 0, saturating

__device__ __forceinline__ int4 ldg_int4(const int32_t* p4, uint32_t i) { return __ldg((const int4*)p4 + i); }
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
#endif
