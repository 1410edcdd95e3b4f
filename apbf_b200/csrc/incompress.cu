// incompress.cu -- the PBF incompressibility constraint (source/incompressibility.cpp:12-45, incompressibility_0..3.comp)
// as two sweeps over the grouped neighbour list.
//
// The reference scatters every pair's contribution with integer atomics (5 per pair in pass 1, 3 per pair in pass 3) and
// spills a 16-byte gradient per pair between the passes.  All accumulators are integers, so any summation order gives
// the same bits.  Here a warp owns 32 consecutive particles; 4 lanes walk one particle's contiguous pair segment at a
// time and accumulate in registers, then the per-particle sums are transposed so that all 32 lanes finish one particle
// each:
//   prologue  (k_begin_iteration): [pending deltas] + [box_collision] + pack {position, mass} per id into P4
//   sweep T1  (k_density_lambda) = passes 0+1+2: density, gradient sums, lambda, boundariness, the particle's own shift
//   sweep T2  (k_apply_delta)    = pass 3 as a GATHER: a particle collects the shifts its neighbours push onto it through
//             the mirrored pairs ((b,a) exists iff the bit of (a,b) is set), recomputing the gradient from the packed
//             positions; only unmirrored pairs (variable kernel widths) still use integer atomics
//   epilogue  (k_commit_delta): position += delta.  Positions stay untouched between T1 and T2, as they are between the
//             reference's passes 1 and 3.
#include "box.cuh"
#include "neighbors.cuh"
#include "solver.cuh"

namespace {

constexpr int SWEEP_THREADS = 256; // 8 warps = 256 particles per CTA tile
constexpr int SWEEP_ILP = 4;       // pairs one lane has in flight
// Lanes per particle.  Four lanes walk one particle's segment, eight particles per round, four rounds per warp tile: a segment of
// 25 pairs (the dam break's mean) costs 7 slots per lane = 28 pair slots, where eight lanes per particle cost 4 x 8 = 32; the
// shuffle reductions shrink from three stages to two and the transposes from eight rounds to four (r02: -12 % on both sweeps).
constexpr int LPP = 4, PPR = 32 / LPP, ROUNDS = 32 / PPR;
#define R_INC APBF_INCOMPRESSIBILITY_DATA_RESOLUTION
#define R2_MIN_UNSCALED (1.0e-8f * 68719476736.0f) // "dist < 0.0001" (kernels.glsl:93) on the squared distance of the unscaled integer difference
#define FULL 0xffffffffu

struct sweep_args {
	const uint32_t* index_list;
	const uint32_t* len;
	int32_t*        pos4;          // hidden
	const float*    inv_mass;      // hidden
	const float*    radius;        // hidden
	const float*    kernel_width;  // per id
	float*          boundariness;  // per id
	const uint32_t* nbl;           // idN | (unmirrored << 31) per pair
	const uint32_t* offsets;
	uint32_t        pair_cap;
	int4*           P4;   // {x, y, z, bits(1 / inverse mass)} per id (written by the prologue only)
	float4*         KG;   // {h, gradient c0, gradient c1, invRestDensity}
	float4*         KH;   // {height c0, height c1, lambda divisor, inverse mass}
	float4*         L4;   // {lambda, h, gradient c0, gradient c1}
	float4*         E4;   // {emptyDirection.xyz, 0}  (boundariness method 2)
	int4*           delta;
	int4*           push;
	int4*           PL;   // {x, y, z, bits(lambda < 0 ? lambda * 2^18 : 0)}: all the apply sweep needs of a neighbour when every width is equal
	float*          kp;   // Gauss gradient kernel only: per pair (a, b) the scalar k with grad W_a(r) = k * r, written by T1, reused by T2
	const float4*   bmin;
	const float4*   bmax;
	uint32_t        n_boxes;
	uint32_t*       misc;
	float*          out_lambda;
	uint32_t*       out_incomp;
	apbf_settings   s;
	float           D;
	const uint32_t* tile_flags;     // slabs: 1 per boundary tile (see ITER_TILES_*)
	int             tile_filter;    // 0 = every tile, 1 = interior tiles only, 2 = boundary tiles only
	int             t2_tail;        // apply sweep, equal-width form: 0 = store the shift, 1 = commit it + next prologue (box, pack), 2 = commit it
	int             skip_if_t2_did; // prologue / commit launch: the apply sweep before was asked to do this work (t2_tail != 0)
};

// does the apply sweep run in its equal-width form (apply_delta_body<GK, false, true>)?  The same words k_apply_delta dispatches on.
__device__ __forceinline__ bool t2_uniform_form(const sweep_args& A)
{
	return A.s.mGradientKernelId == 1 && A.misc[MW_N_ASYM] == 0u && A.misc[MW_H_NONUNIFORM] == 0u;
}

__device__ __forceinline__ float move_towards_abs(float oldValue, float newValue, float maxStep) // incompressibility_2.comp:37-41
{
	float step = newValue - oldValue;
	return oldValue + glsl_min(maxStep, glsl_max(-maxStep, step));
}

__device__ __forceinline__ int sum_group(int v) // the LPP lanes of a group hold the same total afterwards
{
#pragma unroll
	for (int o = 1; o < LPP; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
	return v;
}

// ---- per-particle constants (exact: double-precision pow rounded once) ------------------------------------------------------
template <int HK, int GK>
__global__ void k_prepare_consts(sweep_args A)
{
	const uint32_t n = *A.len;
	const bool ident = A.misc[MW_IDENTITY] != 0u;
	const float w0 = n ? A.kernel_width[0] : 0.0f;
	// The constants are functions of (kernel width, inverse mass, radius) alone, and ~600 instructions of double-precision
	// arithmetic per particle.  A thread keeps its last inputs and results: while the whole warp sees the inputs it saw in its
	// previous trip (one particle class: every trip but the first), nothing is recomputed -- same bits by construction.
	float pw = 0.0f, pim = 0.0f, pr = 0.0f;
	float4 kg = make_float4(0.f, 0.f, 0.f, 0.f), kh = kg;
	bool have = false;
	for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) { // (warp-uniform trip count)
		const uint32_t a = base + threadIdx.x;
		const bool live = a < n;
		float w = pw, invMass = pim, radius = pr;
		if (live) {
			const uint32_t idx = ident ? a : A.index_list[a];
			w = A.kernel_width[a]; invMass = A.inv_mass[idx]; radius = A.radius[idx];
			if (__float_as_uint(w) != __float_as_uint(w0)) A.misc[MW_H_NONUNIFORM] = 1u; // (cleared by the host before the launch)
		}
		const bool same = have && __float_as_uint(w) == __float_as_uint(pw) && __float_as_uint(invMass) == __float_as_uint(pim) &&
		                  __float_as_uint(radius) == __float_as_uint(pr);
		if (!__all_sync(FULL, same)) {
			const kpar hp = height_params<HK>(w, A.D);
			const kpar gp = grad_params<GK>(w, A.D);
			const float invRestDensity = pow_int_rn(2.0f * radius, A.D) * invMass;                               // incompressibility_2.comp:81
			const float lamDiv = pow_int_rn(2.0f * A.s.mSmallestTargetRadius, A.D) / invRestDensity * invMass;  // :98
			kg = make_float4(w, gp.c0, gp.c1, invRestDensity);
			kh = make_float4(hp.c0, hp.c1, lamDiv, invMass);
			pw = w; pim = invMass; pr = radius; have = true;
		}
		if (live) { A.KG[a] = kg; A.KH[a] = kh; }
	}
}

// ---- prologue: [commit pending deltas] [box collision] pack -------------------------------------------------------------------
template <bool COMMIT, bool BOX>
__global__ void k_begin_iteration(sweep_args A)
{
	if (A.skip_if_t2_did && t2_uniform_form(A)) return; // the apply sweep before has committed, collided and packed already
	const uint32_t n = *A.len;
	const bool ident = A.misc[MW_IDENTITY] != 0u;
	const bool has_asym = A.misc[MW_N_ASYM] != 0u;
	const uint32_t n_own = min(n, A.misc[MW_N_OWNED]); // ghosts: their packed records arrive by halo exchange
	const uint32_t gid_base = A.misc[MW_GID_BASE];
	for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n_own; a += gridDim.x * blockDim.x) {
		const uint32_t idx = ident ? a : A.index_list[a];
		int4 p = *((const int4*)A.pos4 + idx);
		if (COMMIT) {
			int4 d = A.delta[a];
			if (has_asym) { const int4 q = A.push[a]; d.x += q.x; d.y += q.y; d.z += q.z; }
			p.x += d.x; p.y += d.y; p.z += d.z;
		}
		if (BOX) { // box_collision.comp:36-60
			float px = (float)p.x * INV_R_POS, py = (float)p.y * INV_R_POS, pz = (float)p.z * INV_R_POS;
			box_push(px, py, pz, a + gid_base, A.radius[idx], A.bmin, A.bmax, A.n_boxes);
			p.x = f2i(px * R_POS); p.y = f2i(py * R_POS); p.z = f2i(pz * R_POS);
		}
		if (COMMIT || BOX) *((int4*)A.pos4 + idx) = p;
		A.P4[a] = make_int4(p.x, p.y, p.z, __float_as_int(1.0f / A.KH[a].w));
	}
}

// ---- T1 -------------------------------------------------------------------------------------------------------------------------
template <int HK, int GK, bool COM>
__global__ void __launch_bounds__(SWEEP_THREADS, (HK == 1 && GK == 1 && !COM) ? 5 : 1) k_density_lambda(sweep_args A)
{
	const uint32_t n = min(*A.len, A.misc[MW_N_OWNED]); // owned particles only; a ghost's lambda arrives by halo exchange
	const bool has_asym = A.misc[MW_N_ASYM] != 0u;
	const unsigned lane = threadIdx.x & 31u, sub = lane & (LPP - 1u), grp = lane / LPP;
	const uint32_t warps_per_grid = gridDim.x * (SWEEP_THREADS / 32);
	for (uint32_t tile = blockIdx.x * (SWEEP_THREADS / 32) + (threadIdx.x >> 5); (size_t)tile * 32 < n; tile += warps_per_grid) {
		const uint32_t base = tile * 32u;
		if (A.tile_filter != 0 && (A.tile_flags[tile] != 0u || base + 31u >= n) != (A.tile_filter == 2)) continue;
		int my_dens = 0, my_sq = 0, my_gx = 0, my_gy = 0, my_gz = 0, my_cx = 0, my_cy = 0, my_cz = 0, my_cw = 0;
#pragma unroll 1
		for (int r = 0; r < ROUNDS; r++) { // round r: group g sweeps particle base + PPR r + g
			const uint32_t a = base + r * PPR + grp;
			int dens = 0, sq = 0, gx = 0, gy = 0, gz = 0, cx = 0, cy = 0, cz = 0, cw = 0;
			if (a < n) {
				// (measured, r02c: the branch-free K-slot form of the apply sweep -- all gathers of a trip issued up front, t2_pairs --
				// needs 57 registers here instead of 40 and runs 7 % SLOWER: this sweep is bound by instruction issue, the occupancy
				// of the 40-register form hides the gather latency already.  The apply sweep gains 5 % from it and keeps it.)
				const int4 ip = A.P4[a];
				const float4 kg = A.KG[a], kh = A.KH[a];
				kpar hp, gp;
				hp.w = kg.x; hp.c0 = kh.x; hp.c1 = kh.y;
				gp.w = kg.x; gp.c0 = kg.y; gp.c1 = kg.z;
				const float c0s = gp.c0 * (INV_R_POS * INV_R_POS); // Gauss: c0 * 2^-36, see below
				const uint32_t beg = min(A.offsets[a], A.pair_cap), end = min(A.offsets[a + 1], A.pair_cap);
				for (uint32_t e0 = beg + sub; e0 < end; e0 += LPP * SWEEP_ILP) { // SWEEP_ILP pairs per lane in flight
					uint32_t bb[SWEEP_ILP];
					int4 qq[SWEEP_ILP];
#pragma unroll
					for (int u = 0; u < SWEEP_ILP; u++) bb[u] = (e0 + LPP * u < end) ? A.nbl[e0 + LPP * u] & NB_ID_MASK : a;
#pragma unroll
					for (int u = 0; u < SWEEP_ILP; u++) qq[u] = A.P4[bb[u]];
#pragma unroll
					for (int u = 0; u < SWEEP_ILP; u++) {
						if (e0 + LPP * u >= end) break;
						const int4 iq = qq[u];
						const float mN = __int_as_float(iq.w); // neighbour's mass = 1 / inverse mass
						const int dxi = iq.x - ip.x, dyi = iq.y - ip.y, dzi = iq.z - ip.z; // int subtract first, incompressibility_1.comp:51
						float W, gk; vec3f g;
						if (HK == 1 && GK == 1) {
							// Gauss / Gauss, the reference's default: the 2^-18 of the fixed-point positions is a power of two and commutes with
							// every rounding on the way, so it is applied to the per-particle constant and to the gradient scalar instead of to
							// the three differences (identical bits: F2 = r2 * 2^36, c0s = c0 * 2^-36, gks = gk * 2^-18, all exact scalings).
							const float fx = (float)dxi, fy = (float)dyi, fz = (float)dzi;
							const float F2 = dot3(fx, fy, fz, fx, fy, fz);
							W = expf(-F2 * c0s) * gp.c1;
							gk = F2 >= R2_MIN_UNSCALED ? (-2.0f * W * gp.c0) * INV_R_POS : 0.0f;
							g.x = gk * fx; g.y = gk * fy; g.z = gk * fz;
						} else {
							const float rx = (float)dxi * INV_R_POS, ry = (float)dyi * INV_R_POS, rz = (float)dzi * INV_R_POS;
							const float r2 = dot3(rx, ry, rz, rx, ry, rz);
							pair_eval<HK, GK>(hp, gp, rx, ry, rz, r2, W, g, gk);
							gk *= INV_R_POS;
						}
						if (GK == 1) A.kp[e0 + LPP * u] = gk; // (scaled by 2^-18: times the INTEGER difference it is the apply sweep's gradient of this pair and, between equal widths, of its mirror)
						// x / invMassN * 2^18 (:55-59): 2^18 is a power of two, so (x * mN) * 2^18 == x * (mN * 2^18) bit for bit.
						// (x * (1 / invMassN) instead of x / invMassN: identical whenever the mass is a power of two -- every scene seeded
						// by initialize.cpp:16-27 with r a power of two; otherwise the product can differ from the quotient in the last
						// bit, i.e. by one unit of 2^-18 in a truncated term now and then: inside the accumulators' parity bar, and
						// measured by tests/test_gpu_parity.py::test_incompressibility_odd_masses.)
						const float mNR = mN * R_INC;
						dens += (int)f2u(W * mNR);                                               // :55
						gx += f2i(g.x * mNR); gy += f2i(g.y * mNR); gz += f2i(g.z * mNR);        // :56-58
						sq += (int)f2u(dot3(g.x, g.y, g.z, g.x, g.y, g.z) * mNR);                // :59
						if (COM) { // boundariness method 1, :62-68
							const float div = kg.x / mN; // invMassN * kernelWidth
							cx += f2i((float)dxi / div); cy += f2i((float)dyi / div); cz += f2i((float)dzi / div);
							cw += f2i(R_INC * mN);
						}
					}
				}
			}
			dens = sum_group(dens); sq = sum_group(sq); gx = sum_group(gx); gy = sum_group(gy); gz = sum_group(gz);
			if (COM) { cx = sum_group(cx); cy = sum_group(cy); cz = sum_group(cz); cw = sum_group(cw); }
			// transpose: lane j finishes particle base + j; its sums sit in group j % PPR during round j / PPR
			const int src = (int)(lane % PPR) * LPP;
			const bool mine = (lane / PPR) == (unsigned)r;
			int t;
			t = __shfl_sync(FULL, dens, src); if (mine) my_dens = t;
			t = __shfl_sync(FULL, sq, src);   if (mine) my_sq = t;
			t = __shfl_sync(FULL, gx, src);   if (mine) my_gx = t;
			t = __shfl_sync(FULL, gy, src);   if (mine) my_gy = t;
			t = __shfl_sync(FULL, gz, src);   if (mine) my_gz = t;
			if (COM) {
				t = __shfl_sync(FULL, cx, src); if (mine) my_cx = t;
				t = __shfl_sync(FULL, cy, src); if (mine) my_cy = t;
				t = __shfl_sync(FULL, cz, src); if (mine) my_cz = t;
				t = __shfl_sync(FULL, cw, src); if (mine) my_cw = t;
			}
		}
		const uint32_t a = base + lane;
		if (a >= n) continue;

		const float4 kg = A.KG[a], kh = A.KH[a];
		const float kw = kg.x, invRestDensity = kg.w, invMass = kh.w;
		kpar hp; hp.w = kw; hp.c0 = kh.x; hp.c1 = kh.y;
		// incompressibility_0.comp:33-45: the particle's own contribution
		uint32_t dens = (uint32_t)my_dens + f2u(kheight<HK>(hp, 0.0f, 0.0f) / invMass * R_INC);
		const uint32_t sq = (uint32_t)my_sq;
		if (A.out_incomp) {
			uint4* o = (uint4*)A.out_incomp + 2 * (size_t)a;
			o[0] = make_uint4((uint32_t)my_gx, (uint32_t)my_gy, (uint32_t)my_gz, dens);
			o[1] = make_uint4(sq, 0u, 0u, 0u);
		}
		// incompressibility_2.comp:72-110
		const float density = (float)dens / R_INC;
		const float wx = (float)my_gx / R_INC, wy = (float)my_gy / R_INC, wz = (float)my_gz / R_INC;
		float squaredGradSum = (float)sq / R_INC;
		const float wgs2 = dot3(wx, wy, wz, wx, wy, wz);
		const float selfGradLength = sqrtf(wgs2) * invRestDensity;
		squaredGradSum += wgs2 * invMass;
		const float underpressure = 1.0f - density * invRestDensity;
		if (A.s.mUpdateBoundariness) { // compute_boundariness :43-69
			float sgl = selfGradLength * kw, up = underpressure, bn = 0.0f;
			if (COM) {
				float totalMass = (float)my_cw / R_INC + 1.0f / invMass;
				float fx = (float)my_cx, fy = (float)my_cy, fz = (float)my_cz;
				float dev = sqrtf(dot3(fx, fy, fz, fx, fy, fz)) / (R_POS * totalMass);
				bn = dev * A.s.mBoundarinessSelfGradLengthFactor;
			} else {
				sgl *= A.s.mBoundarinessSelfGradLengthFactor;
				up = glsl_max(0.0f, up) * A.s.mBoundarinessUnderpressureFactor;
				bn = sgl + up;
			}
			bn = bn >= 1.0f ? 1.0f : 0.0f;
			bn = move_towards_abs(A.boundariness[a], bn, A.s.mBoundarinessAdaptionSpeed);
			A.boundariness[a] = glsl_min(1.0f, bn);
		}
		float lam = underpressure / (invRestDensity * invRestDensity * (squaredGradSum + 0.01f)); // :96
		lam /= kh.z;                                                                              // :97-98
		if (A.out_lambda) A.out_lambda[a] = lam;
		A.L4[a] = make_float4(lam, kw, kg.y, kg.z);
		if (GK == 1) { // the one-gather record of the apply sweep (equal widths)
			const int4 pa = A.P4[a];
			A.PL[a] = make_int4(pa.x, pa.y, pa.z, __float_as_int(lam < 0.0f ? lam * R_POS : 0.0f));
		}
		// the particle's own shift (:100-109); the neighbours' shifts are added by T2
		int sx = 0, sy = 0, sz = 0;
		if (lam < 0.0f) {
			const float f = lam * invMass * R_POS;
			sx = f2i(-wx * invRestDensity * f); sy = f2i(-wy * invRestDensity * f); sz = f2i(-wz * invRestDensity * f);
		}
		A.delta[a] = make_int4(sx, sy, sz, 0);
		if (A.s.mBoundarinessCalculationMethod == 2) { // emptyDirection = -normalize(vec3(gradSum)), incompressibility_3.comp:43
			const float ex = (float)my_gx, ey = (float)my_gy, ez = (float)my_gz;
			const float l = sqrtf(dot3(ex, ey, ez, ex, ey, ez));
			A.E4[a] = make_float4(-(ex / l), -(ey / l), -(ez / l), 0.0f);
		}
		if (has_asym) A.push[a] = make_int4(0, 0, 0, 0);
	}
}

// K x LPP pairs of one particle's segment per call: lane `sub` of the group takes the pairs e0, e0 + LPP, ...  All 2 K gathers are
// issued before the first pair is evaluated, and there is no branch in here for the compiler to sink a load into (with a `break`
// per pair it did: LDG.128, ~60 instructions, LDG.128, ... -- the L2 latencies of a lane in a row).  A slot behind the segment's
// end gathers the particle itself: r = 0, every gradient kernel is zero there, the slot adds nothing (and its "mirrored" bit is
// clear, so it never takes the push path).
template <int K, int GK, bool ASYM>
__device__ __forceinline__ void t2_pairs(const sweep_args& A, uint32_t e0, uint32_t end, uint32_t a, const int4 ip, const float4 la, const float4 e0v,
                                         bool filter, bool push_a, bool kp_ok, int& sx, int& sy, int& sz, int& hit)
{
	uint32_t nn[K];
	int4 qq[K];
	float4 ll[K];
	float kk[K];
#pragma unroll
	for (int u = 0; u < K; u++) nn[u] = (e0 + LPP * u < end) ? __ldg(A.nbl + e0 + LPP * u) : a;
	if (GK == 1) {
#pragma unroll
		for (int u = 0; u < K; u++) kk[u] = (e0 + LPP * u < end) ? __ldg(A.kp + e0 + LPP * u) : 0.0f;
	}
#pragma unroll
	for (int u = 0; u < K; u++) { qq[u] = __ldg(A.P4 + (nn[u] & NB_ID_MASK)); ll[u] = __ldg(A.L4 + (nn[u] & NB_ID_MASK)); }
#pragma unroll
	for (int u = 0; u < K; u++) {
		const uint32_t b = nn[u] & NB_ID_MASK;
		const int4 iq = qq[u];
		const float4 lb = ll[u];
		const bool mirrored = !ASYM || (nn[u] & NB_UNMIRRORED) == 0u; // (no unmirrored pair in the list: nothing to test)
		const int dxi = iq.x - ip.x, dyi = iq.y - ip.y, dzi = iq.z - ip.z;
		const float rx = (float)dxi * INV_R_POS, ry = (float)dyi * INV_R_POS, rz = (float)dzi * INV_R_POS;
		const float r2 = dot3(rx, ry, rz, rx, ry, rz);
		if (filter) {
			// filter_boundariness (incompressibility_3.comp:41-46): dot(emptyDirection, normalize(gradient)) > 0.6.
			// Every kernel's gradient is a non-positive multiple of r, so normalize(gradient) = -r / |r| wherever
			// the gradient is not zero: the test is dot(e0, r) < -0.6 |r|, evaluated without the square root.
			bool nz = r2 >= 1.0e-8f;
			// compact support: the gradient is zero outside h, and at |r| == h for all but the cone kernel
			if (GK != 1) nz = nz && (GK == 3 ? !(sqrtf(r2) > la.y) : sqrtf(r2) < la.y);
			const float d = dot3(e0v.x, e0v.y, e0v.z, rx, ry, rz);
			if (nz && d < 0.0f && d * d > 0.36f * r2) hit = 1;
		}
		if (mirrored) { // the pair (b, a): b shifts a with b's lambda and b's kernel width, diff = pos_a - pos_b
			kpar gp_b; gp_b.w = lb.y; gp_b.c0 = lb.z; gp_b.c1 = lb.w;
			vec3f g;
			if (GK == 1) {
				// Gauss: grad W_b(-r) = k_b * (-r), and k_b depends on b's width and |r|^2 alone.  Between particles of equal width
				// it is the very number T1 computed for (a, b) -- same constants, same r2 bits -- so it is read back (4 bytes,
				// coalesced) instead of evaluated again (dot product, expf, three products: 19 of this loop's 50 instructions).
				float k = kk[u]; // (T1 stored it times 2^-18: it multiplies the unscaled difference)
				if (lb.y != la.y) k = gauss_k(gp_b, r2) * INV_R_POS;
				g.x = k * -(float)dxi; g.y = k * -(float)dyi; g.z = k * -(float)dzi;
			} else {
				g = kgrad_fast<GK>(gp_b, -rx, -ry, -rz, r2);
			}
			const float f = lb.x < 0.0f ? lb.x * R_POS : 0.0f; // (lambda >= 0 pushes nothing: x * 0 truncates to 0)
			sx += f2i(g.x * f); sy += f2i(g.y * f); sz += f2i(g.z * f);
		} else if (push_a) { // nobody gathers (a, b): push it like the reference does (:63-66)
			kpar gp_a; gp_a.w = la.y; gp_a.c0 = la.z; gp_a.c1 = la.w;
			vec3f g;
			if (GK == 1) { // a's own kernel: what T1 stored -- unless a is a ghost (slabs), whose segment T1 never walked
				const float k = kp_ok ? kk[u] : gauss_k(gp_a, r2) * INV_R_POS;
				g.x = k * (float)dxi; g.y = k * (float)dyi; g.z = k * (float)dzi;
			} else g = kgrad_fast<GK>(gp_a, rx, ry, rz, r2);
			const float f = la.x * R_POS;
			atomicAdd(&A.push[b].x, f2i(g.x * f));
			atomicAdd(&A.push[b].y, f2i(g.y * f));
			atomicAdd(&A.push[b].z, f2i(g.z * f));
		}
	}
}

// ---- list segments staged in shared memory by the bulk-copy engine ------------------------------------------------------------------
// A round of the sweep walks the segments of 8 consecutive particles: ONE contiguous piece of the pair list (~200 entries).  Read
// directly, the 8 groups of a load instruction touch 8 different 128-byte lines -- a line of 32 entries costs 8 L1 wavefronts instead
// of one, and the list loads are two fifths of all wavefronts of this sweep, which is what bounds it (ncu: l1tex data-pipe wavefronts
// 65 % of peak, issue slots 50 %).  So one lane arms an mbarrier and issues cp.async.bulk for the round's piece of NB and of the
// gradient scalars (16-byte aligned start: the piece begins up to three entries early); the lanes then read their entries from
// shared memory.  Rounds longer than STAGE_CAP entries read the lists directly as before.
constexpr uint32_t STAGE_CAP = 256u;
struct sweep_stage {
	uint32_t           nb[SWEEP_THREADS / 32][STAGE_CAP + 4u];
	float              kp[SWEEP_THREADS / 32][STAGE_CAP + 4u];
	unsigned long long bar[SWEEP_THREADS / 32];
};
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_init(sweep_stage& S)
{
	if ((threadIdx.x & 31u) == 0u) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&S.bar[threadIdx.x >> 5])));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();
}
// all lanes of the warp call; E0 .. E1: the round's entries.  Returns true if they are in S (entry e at index e - A0, A0 = E0 & ~3).
__device__ __forceinline__ bool stage_round(sweep_stage& S, const uint32_t* nbl, const float* kp, uint32_t E0, uint32_t E1, uint32_t& phase)
{
	const uint32_t A0 = E0 & ~3u;
	if (E1 <= E0 || E1 - A0 > STAGE_CAP) return false;
	const unsigned w = threadIdx.x >> 5;
	__syncwarp(); // everybody is done with the previous round's entries
	if ((threadIdx.x & 31u) == 0u) {
		const uint32_t bytes = ((E1 - A0) * 4u + 15u) & ~15u, bar = smem_addr(&S.bar[w]);
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy reads before, async-proxy writes after
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(2u * bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             :: "r"(smem_addr(&S.nb[w][0])), "l"(nbl + A0), "r"(bytes), "r"(bar) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             :: "r"(smem_addr(&S.kp[w][0])), "l"(kp + A0), "r"(bytes), "r"(bar) : "memory");
	}
	const uint32_t bar = smem_addr(&S.bar[w]);
	uint32_t done = 0u;
	while (!done)
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done) : "r"(bar), "r"(phase) : "memory");
	phase ^= 1u;
	return true;
}

// The same for a list in which every particle has the same kernel width (and no unmirrored pair), Gauss gradient: the scalar of
// every pair is what T1 stored, so a neighbour contributes its position and its lambda -- ONE 16-byte gather per pair instead of two.
// The sweeps are bound by the L1's gather wavefronts (ncu: 77 % of peak in the two-gather form), not by instruction issue.
// (nbp / kpp: the lists -- in global memory, or the round's staged piece shifted so that the same entry numbers index it)
template <int K>
__device__ __forceinline__ void t2_pairs_uniform(const sweep_args& A, const uint32_t* nbp, const float* kpp, uint32_t e0, uint32_t end, uint32_t a,
                                                 const int4 ip, const float4 e0v, bool filter, int& sx, int& sy, int& sz, int& hit)
{
	uint32_t nn[K];
	int4 qq[K];
	float kk[K];
#pragma unroll
	for (int u = 0; u < K; u++) nn[u] = (e0 + LPP * u < end) ? nbp[e0 + LPP * u] & NB_ID_MASK : a;
#pragma unroll
	for (int u = 0; u < K; u++) kk[u] = (e0 + LPP * u < end) ? kpp[e0 + LPP * u] : 0.0f;
#pragma unroll
	for (int u = 0; u < K; u++) qq[u] = __ldg(A.PL + nn[u]);
#pragma unroll
	for (int u = 0; u < K; u++) {
		const int4 iq = qq[u];
		const int dxi = iq.x - ip.x, dyi = iq.y - ip.y, dzi = iq.z - ip.z;
		const float fx = (float)dxi, fy = (float)dyi, fz = (float)dzi; // unscaled: the 2^-18 sits in k (exact, see T1)
		if (filter) { // (see t2_pairs; both sides of every comparison carry the same power of two)
			const float F2 = dot3(fx, fy, fz, fx, fy, fz);
			const float d = dot3(e0v.x, e0v.y, e0v.z, fx, fy, fz);
			if (F2 >= R2_MIN_UNSCALED && d < 0.0f && d * d > 0.36f * F2) hit = 1;
		}
		const float k = kk[u], f = __int_as_float(iq.w);
		sx += f2i((k * -fx) * f); sy += f2i((k * -fy) * f); sz += f2i((k * -fz) * f);
	}
}

// ---- T2 -------------------------------------------------------------------------------------------------------------------------
// ASYM: the list holds unmirrored pairs (variable kernel widths), which push with integer atomics like the reference.  The host
// does not know the list's state without a read-back, so the kernel holds both forms and takes the one that matches (one launch:
// the second, empty launch of 3907 CTAs used to cost 5.5 us per iteration; both forms fit the same 64 registers).
template <int GK, bool ASYM, bool UNIFORM>
__device__ __forceinline__ void apply_delta_body(const sweep_args& A, sweep_stage& S)
{
	uint32_t phase = 0u;
	const uint32_t n = *A.len;
	const uint32_t n_own = min(n, A.misc[MW_N_OWNED]); // a ghost's segment holds only unmirrored pairs onto owned particles: push part only
	const bool filter = A.s.mBoundarinessCalculationMethod == 2;
	constexpr bool has_asym = ASYM;
	const unsigned lane = threadIdx.x & 31u, sub = lane & (LPP - 1u), grp = lane / LPP;
	const uint32_t warps_per_grid = gridDim.x * (SWEEP_THREADS / 32);
	// (without unmirrored pairs the ghosts' segments are empty: the tiles behind the owned particles have nothing to do)
	const uint32_t n_walk = ASYM ? n : n_own;
	for (uint32_t tile = blockIdx.x * (SWEEP_THREADS / 32) + (threadIdx.x >> 5); (size_t)tile * 32 < n_walk; tile += warps_per_grid) {
		const uint32_t base = tile * 32u;
		if (A.tile_filter != 0 && (A.tile_flags[tile] != 0u || base + 31u >= n_own) != (A.tile_filter == 2)) continue;
		int my_sx = 0, my_sy = 0, my_sz = 0, my_hit = 0;
#pragma unroll 1
		for (int r = 0; r < ROUNDS; r++) {
			const uint32_t a = base + r * PPR + grp;
			int sx = 0, sy = 0, sz = 0, hit = 0;
			{
				const bool live = a < n;
				const uint32_t ac = live ? a : base;
				const int4 ip = A.P4[ac];
				const float4 la = A.L4[ac];
				float4 e0v = make_float4(0.f, 0.f, 0.f, 0.f);
				if (filter) e0v = A.E4[ac];
				const bool push_a = has_asym && la.x < 0.0f;
				const bool kp_ok = ac < n_own; // T1 walked this particle's segment: its Gauss scalars are in A.kp
				const uint32_t beg = live ? min(A.offsets[a], A.pair_cap) : 0u, end = live ? min(A.offsets[a + 1], A.pair_cap) : 0u;
				// the PPR groups of the warp walk their segments in lock step, LPP x SWEEP_ILP pairs per trip, as many trips as the
				// longest of them needs; the number of slots of the last trip is the same for the whole warp (no divergence)
				const uint32_t longest = __reduce_max_sync(FULL, end - beg);
				const uint32_t* nbp = A.nbl;
				const float* kpp = A.kp;
				if (UNIFORM) { // the round's piece of the lists: staged by the copy engine (see stage_round)
					const uint32_t E0 = __reduce_min_sync(FULL, live ? beg : 0xFFFFFFFFu), E1 = __reduce_max_sync(FULL, end);
					if (stage_round(S, A.nbl, A.kp, E0, E1, phase)) {
						const unsigned w = threadIdx.x >> 5;
						nbp = S.nb[w] - (E0 & ~3u);
						kpp = S.kp[w] - (E0 & ~3u);
					}
				}
				for (uint32_t done = 0; done < longest; done += LPP * SWEEP_ILP) {
					const uint32_t e0 = beg + done + sub, left = longest - done;
					if (UNIFORM) {
						if (left > 3u * LPP) t2_pairs_uniform<4>(A, nbp, kpp, e0, end, ac, ip, e0v, filter, sx, sy, sz, hit);
						else if (left > 2u * LPP) t2_pairs_uniform<3>(A, nbp, kpp, e0, end, ac, ip, e0v, filter, sx, sy, sz, hit);
						else if (left > 1u * LPP) t2_pairs_uniform<2>(A, nbp, kpp, e0, end, ac, ip, e0v, filter, sx, sy, sz, hit);
						else t2_pairs_uniform<1>(A, nbp, kpp, e0, end, ac, ip, e0v, filter, sx, sy, sz, hit);
					}
					else if (left > 3u * LPP) t2_pairs<4, GK, ASYM>(A, e0, end, ac, ip, la, e0v, filter, push_a, kp_ok, sx, sy, sz, hit);
					else if (left > 2u * LPP) t2_pairs<3, GK, ASYM>(A, e0, end, ac, ip, la, e0v, filter, push_a, kp_ok, sx, sy, sz, hit);
					else if (left > 1u * LPP) t2_pairs<2, GK, ASYM>(A, e0, end, ac, ip, la, e0v, filter, push_a, kp_ok, sx, sy, sz, hit);
					else t2_pairs<1, GK, ASYM>(A, e0, end, ac, ip, la, e0v, filter, push_a, kp_ok, sx, sy, sz, hit);
				}
			}
			sx = sum_group(sx); sy = sum_group(sy); sz = sum_group(sz); hit = sum_group(hit);
			const int src = (int)(lane % PPR) * LPP;
			const bool mine = (lane / PPR) == (unsigned)r;
			int t;
			t = __shfl_sync(FULL, sx, src);  if (mine) my_sx = t;
			t = __shfl_sync(FULL, sy, src);  if (mine) my_sy = t;
			t = __shfl_sync(FULL, sz, src);  if (mine) my_sz = t;
			t = __shfl_sync(FULL, hit, src); if (mine) my_hit = t;
		}
		const uint32_t a = base + lane;
		if (a >= n_own) continue;
		int4 d = A.delta[a]; // the particle's own shift from T1
		d.x += my_sx; d.y += my_sy; d.z += my_sz;
		if (UNIFORM && A.t2_tail != 0) {
			// the shift is complete (no pushes in this form): commit it, and run the next iteration's prologue for this particle
			// (k_begin_iteration<true, true>, which returns at once when it sees that this form ran).  Nobody else reads P4[a]
			// or the position in this sweep: neighbours are gathered from PL.
			const uint32_t idx = A.misc[MW_IDENTITY] != 0u ? a : A.index_list[a];
			int4 p = *((const int4*)A.pos4 + idx);
			p.x += d.x; p.y += d.y; p.z += d.z;
			if (A.t2_tail == 1) {
				float px = (float)p.x * INV_R_POS, py = (float)p.y * INV_R_POS, pz = (float)p.z * INV_R_POS; // box_collision.comp:36-60
				box_push(px, py, pz, a + A.misc[MW_GID_BASE], A.radius[idx], A.bmin, A.bmax, A.n_boxes);
				p.x = f2i(px * R_POS); p.y = f2i(py * R_POS); p.z = f2i(pz * R_POS);
				A.P4[a] = make_int4(p.x, p.y, p.z, A.P4[a].w);
			}
			*((int4*)A.pos4 + idx) = p;
		} else A.delta[a] = d;
		if (filter && my_hit) A.boundariness[a] = 0.0f;
	}
}

template <int GK>
__global__ void __launch_bounds__(SWEEP_THREADS, 4) k_apply_delta(sweep_args A)
{
	__shared__ __align__(16) sweep_stage S;
	stage_init(S);
	if (A.misc[MW_N_ASYM] != 0u) apply_delta_body<GK, true, false>(A, S);
	else if (GK == 1 && A.misc[MW_H_NONUNIFORM] == 0u) apply_delta_body<GK, false, true>(A, S);
	else apply_delta_body<GK, false, false>(A, S);
}

// position += delta (+ pushes); xyz only, w is the caller's
__global__ void k_commit_delta(sweep_args A)
{
	if (A.skip_if_t2_did && t2_uniform_form(A)) return; // the apply sweep has committed already
	const uint32_t n = min(*A.len, A.misc[MW_N_OWNED]);
	const bool ident = A.misc[MW_IDENTITY] != 0u;
	const bool has_asym = A.misc[MW_N_ASYM] != 0u;
	for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
		const uint32_t idx = ident ? a : A.index_list[a];
		int4 d = A.delta[a];
		if (has_asym) { const int4 q = A.push[a]; d.x += q.x; d.y += q.y; d.z += q.z; }
		int4 p = *((int4*)A.pos4 + idx);
		p.x += d.x; p.y += d.y; p.z += d.z;
		*((int4*)A.pos4 + idx) = p;
	}
}

int fill_args(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, sweep_args& A)
{
	apbf_particles& p = fluid->particle;
	APBF_REQUIRE(ctx, p.index_list.data && p.length && p.position.data && p.inverse_mass.data && p.radius.data);
	APBF_REQUIRE(ctx, fluid->kernel_width.data && fluid->boundariness.data);
	const uint32_t n_cap = p.capacity;
	memset(&A, 0, sizeof A);
	A.index_list = (const uint32_t*)p.index_list.data;
	A.len = p.length;
	A.pos4 = (int32_t*)p.position.data;
	A.inv_mass = (const float*)p.inverse_mass.data;
	A.radius = (const float*)p.radius.data;
	A.kernel_width = (const float*)fluid->kernel_width.data;
	A.boundariness = (float*)fluid->boundariness.data;
	if (nb) A.pair_cap = nb->capacity;
	A.nbl = (const uint32_t*)ctx->scratch_get(SLOT_NB, 4);
	A.offsets = (const uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	A.P4 = (int4*)ctx->scratch_get(SLOT_P4, sizeof(int4) * (size_t)n_cap);
	A.KG = (float4*)ctx->scratch_get(SLOT_KG, sizeof(float4) * (size_t)n_cap);
	A.KH = (float4*)ctx->scratch_get(SLOT_KH, sizeof(float4) * (size_t)n_cap);
	A.L4 = (float4*)ctx->scratch_get(SLOT_L4, sizeof(float4) * (size_t)n_cap);
	A.E4 = (float4*)ctx->scratch_get(SLOT_G4, sizeof(float4) * (size_t)n_cap);
	A.delta = (int4*)ctx->scratch_get(SLOT_DELTA, sizeof(int4) * (size_t)n_cap);
	A.push = (int4*)ctx->scratch_get(SLOT_PUSH, sizeof(int4) * (size_t)n_cap);
	if (nb && ctx->settings.mGradientKernelId == 1) {
		A.kp = (float*)ctx->scratch_get(SLOT_KP, sizeof(float) * (size_t)(nb->capacity ? nb->capacity : 1u));
		A.PL = (int4*)ctx->scratch_get(SLOT_PL, sizeof(int4) * (size_t)n_cap);
		if (!A.kp || !A.PL) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	}
	A.misc = ctx->misc();
	A.s = ctx->settings;
	A.D = (float)ctx->dims;
	if (!A.offsets || !A.nbl || !A.P4 || !A.KG || !A.KH || !A.L4 || !A.E4 || !A.delta || !A.push || !A.misc)
		return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	return APBF_OK;
}

template <int HK>
int launch_prepare(apbf_ctx* ctx, const sweep_args& A, unsigned grid)
{
	switch (A.s.mGradientKernelId) {
		case 0: k_prepare_consts<HK, 0><<<grid, 256, 0, ctx->stream>>>(A); break;
		case 1: k_prepare_consts<HK, 1><<<grid, 256, 0, ctx->stream>>>(A); break;
		case 2: k_prepare_consts<HK, 2><<<grid, 256, 0, ctx->stream>>>(A); break;
		case 3: k_prepare_consts<HK, 3><<<grid, 256, 0, ctx->stream>>>(A); break;
		case 4: k_prepare_consts<HK, 4><<<grid, 256, 0, ctx->stream>>>(A); break;
		default: return apbf_fail(ctx, APBF_ERR_INVALID, "gradient kernel id", __FILE__, __LINE__);
	}
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

template <int HK, int GK>
void launch_t1(apbf_ctx* ctx, const sweep_args& A, unsigned grid)
{
	if (A.s.mBoundarinessCalculationMethod == 1) k_density_lambda<HK, GK, true><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A);
	else k_density_lambda<HK, GK, false><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A);
}

template <int HK>
int launch_density_lambda(apbf_ctx* ctx, const sweep_args& A, unsigned grid)
{
	switch (A.s.mGradientKernelId) {
		case 0: launch_t1<HK, 0>(ctx, A, grid); break;
		case 1: launch_t1<HK, 1>(ctx, A, grid); break;
		case 2: launch_t1<HK, 2>(ctx, A, grid); break;
		case 3: launch_t1<HK, 3>(ctx, A, grid); break;
		case 4: launch_t1<HK, 4>(ctx, A, grid); break;
		default: return apbf_fail(ctx, APBF_ERR_INVALID, "gradient kernel id", __FILE__, __LINE__);
	}
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

} // namespace

int apbf_solver_prepare(apbf_ctx* ctx, apbf_fluid* fluid)
{
	sweep_args A;
	APBF_TRY(fill_args(ctx, fluid, nullptr, A));
	const uint32_t n_cap = fluid->particle.capacity;
	if (n_cap == 0) return APBF_OK;
	apbf_prof_scope ps(ctx, PROF_SOLVER_PREPARE);
	const unsigned grid = apbf_grid(ctx, n_cap, 256, 4); // (few, long-lived threads: k_prepare_consts re-uses its last results)
	APBF_CUDA(ctx, cudaMemsetAsync(A.misc + MW_H_NONUNIFORM, 0, sizeof(uint32_t), ctx->stream));
	switch (A.s.mHeightKernelId) {
		case 0: return launch_prepare<0>(ctx, A, grid);
		case 1: return launch_prepare<1>(ctx, A, grid);
		case 2: return launch_prepare<2>(ctx, A, grid);
		case 3: return launch_prepare<3>(ctx, A, grid);
		case 4: return launch_prepare<4>(ctx, A, grid);
	}
	return apbf_fail(ctx, APBF_ERR_INVALID, "height kernel id", __FILE__, __LINE__);
}

int apbf_solver_iteration(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, int flags, const float* box_min4,
                          const float* box_max4, uint32_t n_boxes, float* out_lambda, uint32_t* out_incomp)
{
	APBF_TRY(apbf_nbr_ensure(ctx, fluid, nb)); // any pair list (incompressibility.h:11): foreign ones get their structure built here
	sweep_args A;
	APBF_TRY(fill_args(ctx, fluid, nb, A));
	const uint32_t n_cap = fluid->particle.capacity;
	if (n_cap == 0) return APBF_OK;
	A.bmin = (const float4*)box_min4;
	A.bmax = (const float4*)box_max4;
	A.n_boxes = n_boxes;
	A.out_lambda = out_lambda;
	A.out_incomp = out_incomp;
	cudaStream_t st = ctx->stream;
	const unsigned egrid = apbf_grid(ctx, n_cap, 256);
	const bool run_all = (flags & (ITER_RUN_BEGIN | ITER_RUN_T1 | ITER_RUN_T2)) == 0;
	if (flags & (ITER_TILES_INTERIOR | ITER_TILES_BOUNDARY)) {
		A.tile_filter = (flags & ITER_TILES_BOUNDARY) ? 2 : 1;
		A.tile_flags = (const uint32_t*)ctx->scratch_get(SLOT_MG_TILES, sizeof(uint32_t) * ((size_t)n_cap / 32u + 2u));
		if (!A.tile_flags) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	}
	A.skip_if_t2_did = (flags & ITER_SKIP_IF_T2_DID) ? 1 : 0;
	A.t2_tail = (flags & ITER_T2_COMMIT) ? ((flags & ITER_T2_NEXT_BOX) ? 1 : 2) : 0;
	if (run_all || (flags & ITER_RUN_BEGIN)) {
		apbf_prof_scope ps(ctx, (flags & ITER_BEGIN_BOX) ? PROF_BOX : PROF_COMMIT);
		const bool c = flags & ITER_BEGIN_COMMIT, b = flags & ITER_BEGIN_BOX;
		if (c && b) k_begin_iteration<true, true><<<egrid, 256, 0, st>>>(A);
		else if (c) k_begin_iteration<true, false><<<egrid, 256, 0, st>>>(A);
		else if (b) k_begin_iteration<false, true><<<egrid, 256, 0, st>>>(A);
		else k_begin_iteration<false, false><<<egrid, 256, 0, st>>>(A);
		APBF_LAUNCHED(ctx);
	}
	// one 256-particle tile per CTA up to 64 CTAs per SM: the hardware scheduler evens out the segments' lengths (a grid of
	// exactly-resident CTAs with a static tile stride left a third of the warp slots idle at the tail)
	const unsigned grid = apbf_grid(ctx, n_cap, SWEEP_THREADS, 64);
	if (run_all || (flags & ITER_RUN_T1)) {
		apbf_prof_scope ps(ctx, PROF_DENSITY_LAMBDA);
		switch (A.s.mHeightKernelId) {
			case 0: APBF_TRY(launch_density_lambda<0>(ctx, A, grid)); break;
			case 1: APBF_TRY(launch_density_lambda<1>(ctx, A, grid)); break;
			case 2: APBF_TRY(launch_density_lambda<2>(ctx, A, grid)); break;
			case 3: APBF_TRY(launch_density_lambda<3>(ctx, A, grid)); break;
			case 4: APBF_TRY(launch_density_lambda<4>(ctx, A, grid)); break;
			default: return apbf_fail(ctx, APBF_ERR_INVALID, "height kernel id", __FILE__, __LINE__);
		}
	}
	if (run_all || (flags & ITER_RUN_T2)) {
		apbf_prof_scope ps(ctx, PROF_APPLY_DELTA);
		switch (A.s.mGradientKernelId) {
			case 0: k_apply_delta<0><<<grid, SWEEP_THREADS, 0, st>>>(A); break;
			case 1: k_apply_delta<1><<<grid, SWEEP_THREADS, 0, st>>>(A); break;
			case 2: k_apply_delta<2><<<grid, SWEEP_THREADS, 0, st>>>(A); break;
			case 3: k_apply_delta<3><<<grid, SWEEP_THREADS, 0, st>>>(A); break;
			case 4: k_apply_delta<4><<<grid, SWEEP_THREADS, 0, st>>>(A); break;
		}
		APBF_LAUNCHED(ctx);
	}
	if (flags & ITER_END_COMMIT) {
		apbf_prof_scope ps(ctx, PROF_COMMIT);
		A.skip_if_t2_did = (flags & ITER_T2_COMMIT) ? 1 : 0; // (this call's own apply sweep)
		k_commit_delta<<<egrid, 256, 0, st>>>(A);
		APBF_LAUNCHED(ctx);
	}
	return APBF_OK;
}

extern "C" int apbf_incompressibility_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, float* out_lambda,
                                            uint32_t* out_incomp_data)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid && nb);
	// a stand-alone call knows nothing about what changed since the last one: recompute the constants every time
	APBF_TRY(apbf_solver_prepare(ctx, fluid));
	return apbf_solver_iteration(ctx, fluid, nb, ITER_END_COMMIT | ITER_T2_COMMIT, nullptr, nullptr, 0u, out_lambda, out_incomp_data);
}
