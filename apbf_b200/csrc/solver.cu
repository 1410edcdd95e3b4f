// solver.cu -- the PBF incompressibility constraint, the adaptive kernel width update, box collision and the integrator
// as sweeps over the grouped neighbour list.
//
// Replaces source/incompressibility.cpp:12-45 (incompressibility_0..3.comp), source/spread_kernel_width.cpp:12-26
// (kernel_width_init.comp, kernel_width.comp, uint_to_float_but_gradual.comp), source/box_collision.cpp:12-24
// (box_collision.comp) and source/velocity_handling.cpp:15-31 (infer_velocity / apply_acceleration / apply_velocity).
//
// Formulation.  The reference scatters every pair's contribution with integer atomics (5 per pair in pass 1, 3 per pair
// in pass 3).  All accumulators are integers, so any summation order gives the same bits; here a group of lanes owns
// one particle, walks its contiguous pair segment and accumulates in registers:
//   sweep T1 (= passes 0+1+2): density, gradient sums, lambda, boundariness  -> L4 = {lambda, h, c0, c1}, G4 = {gradSum, 1/rho0}
//   sweep T2 (= pass 3 + the self shift of pass 2): a particle GATHERS the shifts its neighbours push onto it through
//     the mirrored pairs ((b,a) exists iff bit set on (a,b)), recomputing the gradient instead of reading the 16 B/pair
//     spill of the reference; only unmirrored pairs (variable kernel widths) still use integer atomics.
//   commit: position += delta (positions stay untouched between T1 and T2, as they are between the reference's passes).
#include "kernels.cuh"
#include "neighbors.cuh"
#include "sort.cuh"

namespace {

constexpr int GROUP = 8;            // lanes per particle in the segment sweeps
constexpr int SWEEP_THREADS = 256;  // 32 particles per CTA
#define R_INC APBF_INCOMPRESSIBILITY_DATA_RESOLUTION
#define R_KW APBF_KERNEL_WIDTH_RESOLUTION

struct sweep_args {
	const uint32_t* index_list;
	const uint32_t* len;
	int32_t*        pos4;
	const float*    inv_mass;   // hidden
	const float*    radius;     // hidden
	float*          kernel_width;   // per id
	float*          boundariness;   // per id
	const float*    target_radius;  // per id
	const uint32_t* pairs;
	const uint32_t* offsets;
	uint32_t*       symbits;
	uint32_t        pair_cap;
	float4*         L4;         // {lambda, h, c0, c1}
	int4*           G4;         // {gradSum.xyz, bits(invRestDensity)}
	int4*           delta;      // position shift of the iteration
	int4*           push;       // shifts pushed through unmirrored pairs
	int4*           com4;       // centre-of-mass sums (boundariness method 1)
	uint32_t*       misc;
	float*          out_lambda;
	uint32_t*       out_incomp;
	apbf_settings   s;
	float           D;
};

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

__device__ __forceinline__ float move_towards_abs(float oldValue, float newValue, float maxStep) // incompressibility_2.comp:37-41
{
	float step = newValue - oldValue;
	return oldValue + glsl_min(maxStep, glsl_max(-maxStep, step));
}

// ---- T1: incompressibility_0 + _1 + _2 -------------------------------------------------------------------------------
template <int HK, int GK>
__global__ void __launch_bounds__(SWEEP_THREADS) k_density_lambda(sweep_args A)
{
	const uint32_t n = *A.len;
	const bool ident = A.misc[MW_IDENTITY] != 0u;
	const bool has_asym = A.misc[MW_N_ASYM] != 0u;
	const unsigned sub = threadIdx.x & (GROUP - 1);
	const uint32_t groups_per_grid = gridDim.x * (SWEEP_THREADS / GROUP);
	for (uint32_t a = blockIdx.x * (SWEEP_THREADS / GROUP) + threadIdx.x / GROUP; a < n; a += groups_per_grid) {
		const uint32_t idx = ident ? a : A.index_list[a];
		const int4 ip = *((const int4*)A.pos4 + idx);
		const float kw = A.kernel_width[a];
		const kpar hp = height_params<HK>(kw, A.D);
		const kpar gp = grad_params<GK>(kw, A.D);
		uint32_t dens = 0u, sq = 0u;
		int gx = 0, gy = 0, gz = 0;
		int cx = 0, cy = 0, cz = 0, cw = 0;
		const uint32_t beg = min(A.offsets[a], A.pair_cap), end = min(A.offsets[a + 1], A.pair_cap);
		for (uint32_t e = beg + sub; e < end; e += GROUP) {
			const uint32_t b = A.pairs[2 * (size_t)e + 1];
			const uint32_t idxN = ident ? b : A.index_list[b];
			const int4 iq = *((const int4*)A.pos4 + idxN);
			const float mN = A.inv_mass[idxN];
			const int dxi = iq.x - ip.x, dyi = iq.y - ip.y, dzi = iq.z - ip.z; // int subtract first, incompressibility_1.comp:51
			const float rx = (float)dxi * INV_R_POS, ry = (float)dyi * INV_R_POS, rz = (float)dzi * INV_R_POS;
			const float r2 = dot3(rx, ry, rz, rx, ry, rz);
			const float dist = sqrtf(r2);
			const vec3f g = kgrad<GK>(gp, rx, ry, rz, r2, dist);
			const float W = kheight<HK>(hp, r2, dist);
			dens += f2u(W / mN * R_INC);
			gx += f2i(g.x / mN * R_INC);
			gy += f2i(g.y / mN * R_INC);
			gz += f2i(g.z / mN * R_INC);
			sq += f2u(dot3(g.x, g.y, g.z, g.x, g.y, g.z) / mN * R_INC);
			if (A.s.mBoundarinessCalculationMethod == 1) { // incompressibility_1.comp:62-68
				const float div = mN * kw;
				cx += f2i((float)dxi / div); cy += f2i((float)dyi / div); cz += f2i((float)dzi / div);
				cw += f2i(R_INC / mN);
			}
		}
		dens = (uint32_t)group_sum((int)dens); sq = (uint32_t)group_sum((int)sq);
		gx = group_sum(gx); gy = group_sum(gy); gz = group_sum(gz);
		if (A.s.mBoundarinessCalculationMethod == 1) { cx = group_sum(cx); cy = group_sum(cy); cz = group_sum(cz); cw = group_sum(cw); }
		if (sub != 0) continue;

		// incompressibility_0.comp:33-45: the particle's own contribution
		const float invMass = A.inv_mass[idx];
		const float radius = A.radius[idx];
		dens += f2u(kheight<HK>(hp, 0.0f, 0.0f) / invMass * R_INC);
		if (A.out_incomp) {
			uint4* o = (uint4*)A.out_incomp + 2 * (size_t)a;
			o[0] = make_uint4((uint32_t)gx, (uint32_t)gy, (uint32_t)gz, dens);
			o[1] = make_uint4(sq, 0u, 0u, 0u);
		}
		// incompressibility_2.comp:72-110
		const float invRestDensity = pow_rn(2.0f * radius, A.D) * invMass;
		const float density = (float)dens / R_INC;
		const float wx = (float)gx / R_INC, wy = (float)gy / R_INC, wz = (float)gz / R_INC;
		float squaredGradSum = (float)sq / R_INC;
		const float wgs2 = dot3(wx, wy, wz, wx, wy, wz);
		const float selfGradLength = sqrtf(wgs2) * invRestDensity;
		squaredGradSum += wgs2 * invMass;
		const float underpressure = 1.0f - density * invRestDensity;
		if (A.s.mUpdateBoundariness) { // compute_boundariness :43-69
			float sgl = selfGradLength * kw, up = underpressure, bn = 0.0f;
			if (A.s.mBoundarinessCalculationMethod == 1) {
				float totalMass = (float)cw / R_INC + 1.0f / invMass;
				float fx = (float)cx, fy = (float)cy, fz = (float)cz;
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
		float lam = underpressure / (invRestDensity * invRestDensity * (squaredGradSum + 0.01f));
		lam /= pow_rn(2.0f * A.s.mSmallestTargetRadius, A.D) / invRestDensity * invMass;
		if (A.out_lambda) A.out_lambda[a] = lam;
		A.L4[a] = make_float4(lam, gp.w, gp.c0, gp.c1);
		A.G4[a] = make_int4(gx, gy, gz, __float_as_int(invRestDensity));
		if (has_asym) A.push[a] = make_int4(0, 0, 0, 0);
	}
}

// ---- T2: incompressibility_3 as a gather (+ the self shift of incompressibility_2) -----------------------------------------
template <int GK>
__global__ void __launch_bounds__(SWEEP_THREADS) k_apply_delta(sweep_args A)
{
	const uint32_t n = *A.len;
	const bool ident = A.misc[MW_IDENTITY] != 0u;
	const bool filter = A.s.mBoundarinessCalculationMethod == 2;
	const unsigned sub = threadIdx.x & (GROUP - 1);
	const uint32_t groups_per_grid = gridDim.x * (SWEEP_THREADS / GROUP);
	for (uint32_t a = blockIdx.x * (SWEEP_THREADS / GROUP) + threadIdx.x / GROUP; a < n; a += groups_per_grid) {
		const uint32_t idx = ident ? a : A.index_list[a];
		const int4 ip = *((const int4*)A.pos4 + idx);
		const float4 la = A.L4[a];
		const int4 ga = A.G4[a];
		kpar gp_a; gp_a.w = la.y; gp_a.c0 = la.z; gp_a.c1 = la.w;
		const float lam_a = la.x;
		// emptyDirection = -normalize(vec3(gradSum)), incompressibility_3.comp:43
		float ex = (float)ga.x, ey = (float)ga.y, ez = (float)ga.z;
		{
			const float l = sqrtf(dot3(ex, ey, ez, ex, ey, ez));
			ex = -(ex / l); ey = -(ey / l); ez = -(ez / l);
		}
		int sx = 0, sy = 0, sz = 0, hit = 0;
		const uint32_t beg = min(A.offsets[a], A.pair_cap), end = min(A.offsets[a + 1], A.pair_cap);
		for (uint32_t e = beg + sub; e < end; e += GROUP) {
			const uint32_t b = A.pairs[2 * (size_t)e + 1];
			const uint32_t idxN = ident ? b : A.index_list[b];
			const int4 iq = *((const int4*)A.pos4 + idxN);
			const bool mirrored = (A.symbits[e >> 5] >> (e & 31u)) & 1u;
			const int dxi = iq.x - ip.x, dyi = iq.y - ip.y, dzi = iq.z - ip.z;
			if (filter || (!mirrored && lam_a < 0.0f)) { // the pair (a, b) itself: gradient with a's kernel width
				const float rx = (float)dxi * INV_R_POS, ry = (float)dyi * INV_R_POS, rz = (float)dzi * INV_R_POS;
				const float r2 = dot3(rx, ry, rz, rx, ry, rz);
				const float dist = sqrtf(r2);
				const vec3f g = kgrad<GK>(gp_a, rx, ry, rz, r2, dist);
				if (filter) { // filter_boundariness :41-46
					const float gl = sqrtf(dot3(g.x, g.y, g.z, g.x, g.y, g.z));
					const float nx = g.x / gl, ny = g.y / gl, nz = g.z / gl;
					if (dot3(ex, ey, ez, nx, ny, nz) > 0.6f) hit = 1;
				}
				if (!mirrored && lam_a < 0.0f) { // nobody gathers this pair: push it like the reference does (:63-66)
					const float f = lam_a * R_POS;
					atomicAdd(&A.push[b].x, f2i(g.x * f));
					atomicAdd(&A.push[b].y, f2i(g.y * f));
					atomicAdd(&A.push[b].z, f2i(g.z * f));
				}
			}
			if (mirrored) { // the pair (b, a): b shifts a with b's lambda and b's kernel width
				const float4 lb = A.L4[b];
				if (lb.x < 0.0f) {
					kpar gp_b; gp_b.w = lb.y; gp_b.c0 = lb.z; gp_b.c1 = lb.w;
					const float rx = (float)(-dxi) * INV_R_POS, ry = (float)(-dyi) * INV_R_POS, rz = (float)(-dzi) * INV_R_POS;
					const float r2 = dot3(rx, ry, rz, rx, ry, rz);
					const float dist = sqrtf(r2);
					const vec3f g = kgrad<GK>(gp_b, rx, ry, rz, r2, dist);
					const float f = lb.x * R_POS;
					sx += f2i(g.x * f); sy += f2i(g.y * f); sz += f2i(g.z * f);
				}
			}
		}
		sx = group_sum(sx); sy = group_sum(sy); sz = group_sum(sz); hit = group_sum(hit);
		if (sub != 0) continue;
		if (lam_a < 0.0f) { // incompressibility_2.comp:100-109
			const float invMass = A.inv_mass[idx];
			const float invRestDensity = __int_as_float(ga.w);
			const float wx = (float)ga.x / R_INC, wy = (float)ga.y / R_INC, wz = (float)ga.z / R_INC;
			const float f = lam_a * invMass * R_POS;
			sx += f2i(-wx * invRestDensity * f);
			sy += f2i(-wy * invRestDensity * f);
			sz += f2i(-wz * invRestDensity * f);
		}
		A.delta[a] = make_int4(sx, sy, sz, 0);
		if (filter && hit) A.boundariness[a] = 0.0f;
	}
}

// position += delta (+ pushes); 12 of the 16 bytes are rewritten, w is the caller's
__global__ void k_commit_delta(sweep_args A)
{
	const uint32_t n = *A.len;
	const bool ident = A.misc[MW_IDENTITY] != 0u;
	const bool has_asym = A.misc[MW_N_ASYM] != 0u;
	for (uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
		const uint32_t idx = ident ? a : A.index_list[a];
		int4 d = A.delta[a];
		if (has_asym) { const int4 p = A.push[a]; d.x += p.x; d.y += p.y; d.z += p.z; }
		int4 p = *((int4*)A.pos4 + idx);
		p.x += d.x; p.y += d.y; p.z += d.z;
		*((int4*)A.pos4 + idx) = p;
	}
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
	const uint32_t* pairs;
	const uint32_t* offsets;
	uint32_t*       symbits;
	uint32_t        pair_cap;
	uint32_t*       keep_counts;
	const uint32_t* keep_offsets;
	uint32_t*       out_pairs;
	uint32_t*       out_symbits;
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

__device__ __forceinline__ uint32_t kw_influence(float orig, float dist) // kernel_width.comp:49-52
{
	const float distanceFromKernel = dist - orig;
	const float influence = glsl_max(0.0f, 1.0f - glsl_max(0.0f, distanceFromKernel / (orig * APBF_KERNEL_WIDTH_PROPAGATION_FACTOR)));
	return f2u(orig * influence * R_KW);
}

// COMPACT == false: atomicMax spread (gathered through mirrored pairs) + number of pairs to keep per particle
// COMPACT == true : write the kept pairs at the scanned offsets with their new mirrored bits
template <bool COMPACT>
__global__ void __launch_bounds__(SWEEP_THREADS) k_kw_sweep(kw_args A)
{
	const uint32_t n = *A.len;
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
				b = A.pairs[2 * (size_t)e + 1];
				const uint32_t idxN = ident ? b : A.index_list[b];
				const int4 iq = __ldg((const int4*)A.pos4 + idxN);
				const float rx = (float)(iq.x - ip.x) * INV_R_POS, ry = (float)(iq.y - ip.y) * INV_R_POS, rz = (float)(iq.z - ip.z) * INV_R_POS;
				const float dist = sqrtf(dot3(rx, ry, rz, rx, ry, rz));
				mirrored = (A.symbits[e >> 5] >> (e & 31u)) & 1u;
				keep = dist <= cutoff_a; // kernel_width.comp:57
				if (!COMPACT) {
					if (mirrored) mx = max(mx, kw_influence(kw_original(A, b, idxN), dist)); // the pair (b, a) spreads b's width onto a
					else atomicMax(A.kwfx + b, kw_influence(orig_a, dist));                   // (a, b) has no mirror: spread directly (:53)
				} else if (keep) {
					new_mirrored = mirrored && dist <= glsl_max(kw_original(A, b, idxN), A.kernel_width[b]);
				}
			}
			if (!COMPACT) {
				kept += keep ? 1u : 0u;
			} else { // stable compaction inside the group: ballot of this 8-lane slice
				const uint32_t ballot = (__ballot_sync(group_mask(), keep) >> group_shift) & ((1u << GROUP) - 1u);
				const uint32_t rank = __popc(ballot & ((1u << sub) - 1u));
				if (keep) {
					const uint32_t o = out + rank;
					*(uint2*)(A.out_pairs + 2 * (size_t)o) = make_uint2(a, b);
					if (new_mirrored) atomicOr(A.out_symbits + (o >> 5), 1u << (o & 31u));
					else n_asym++;
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

__global__ void k_reset_asym(uint32_t* misc) { misc[MW_N_ASYM] = 0u; }

// ---- box collision (box_collision.comp:36-60) ------------------------------------------------------------------------------
__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }
__device__ __forceinline__ vec3f hash31(float p) // box_collision.comp:20-25
{
	float ax = fractf(p * .1031f), ay = fractf(p * .1030f), az = fractf(p * .0973f);
	const float d = dot3(ax, ay, az, ay + 33.33f, az + 33.33f, ax + 33.33f);
	ax += d; ay += d; az += d;
	vec3f o;
	o.x = fractf((ax + ay) * az);
	o.y = fractf((ax + az) * ay);
	o.z = fractf((ay + az) * ax);
	return o;
}

__device__ __forceinline__ void box_push(float& px, float& py, float& pz, uint32_t id, float radius, const float4* bmin,
                                         const float4* bmax, uint32_t n_boxes)
{
	for (uint32_t i = 0; i < n_boxes; i++) {
		const vec3f h0 = hash31((float)id * px - py - pz);
		const vec3f h1 = hash31((float)id * py + px + pz);
		const float4 lo = bmin[i], hi = bmax[i];
		float tminx = (lo.x - radius - h0.x * 0.05f) - px, tminy = (lo.y - radius - h0.y * 0.05f) - py, tminz = (lo.z - radius - h0.z * 0.05f) - pz;
		float tmaxx = (hi.x + radius + h1.x * 0.05f) - px, tmaxy = (hi.y + radius + h1.y * 0.05f) - py, tmaxz = (hi.z + radius + h1.z * 0.05f) - pz;
		{ // toMin *= vec3(smallestComponent(-toMin))
			const float vx = -tminx, vy = -tminy, vz = -tminz;
			const bool b1 = vx <= vy, b2 = vx <= vz, b3 = vy <= vz;
			tminx *= (b1 && b2) ? 1.0f : 0.0f; tminy *= (!b1 && b3) ? 1.0f : 0.0f; tminz *= (!b2 && !b3) ? 1.0f : 0.0f;
		}
		{
			const bool b1 = tmaxx <= tmaxy, b2 = tmaxx <= tmaxz, b3 = tmaxy <= tmaxz;
			tmaxx *= (b1 && b2) ? 1.0f : 0.0f; tmaxy *= (!b1 && b3) ? 1.0f : 0.0f; tmaxz *= (!b2 && !b3) ? 1.0f : 0.0f;
		}
		const float distToMin = -((tminx * 1.0f + tminy * 1.0f) + tminz * 1.0f);
		const float distToMax = (tmaxx * 1.0f + tmaxy * 1.0f) + tmaxz * 1.0f;
		if (distToMin <= 0.0f || distToMax <= 0.0f) continue;
		if (distToMin < distToMax) { px += tminx; py += tminy; pz += tminz; }
		else { px += tmaxx; py += tmaxy; pz += tmaxz; }
	}
}

__global__ void k_box_collision(const uint32_t* __restrict__ index_list, int32_t* pos4, const float* __restrict__ radius,
                                const float4* __restrict__ bmin, const float4* __restrict__ bmax, const uint32_t* __restrict__ len,
                                uint32_t n_boxes, const uint32_t* __restrict__ misc, int use_ident)
{
	const uint32_t n = *len;
	const bool ident = use_ident && misc[MW_IDENTITY] != 0u;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const uint32_t idx = ident ? id : index_list[id];
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

template <int HK>
int launch_density_lambda(apbf_ctx* ctx, const sweep_args& A, unsigned grid)
{
	switch (A.s.mGradientKernelId) {
		case 0: k_density_lambda<HK, 0><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
		case 1: k_density_lambda<HK, 1><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
		case 2: k_density_lambda<HK, 2><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
		case 3: k_density_lambda<HK, 3><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
		case 4: k_density_lambda<HK, 4><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
		default: return apbf_fail(ctx, APBF_ERR_INVALID, "gradient kernel id", __FILE__, __LINE__);
	}
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

} // namespace

// one incompressibility::apply() on the public arrays
int apbf_incompressibility_run(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, float* out_lambda,
                               uint32_t* out_incomp)
{
	apbf_particles& p = fluid->particle;
	APBF_REQUIRE(ctx, p.index_list.data && p.length && p.position.data && p.inverse_mass.data && p.radius.data);
	APBF_REQUIRE(ctx, fluid->kernel_width.data && fluid->boundariness.data);
	if (!apbf_nbr_struct_valid(ctx, nb))
		return apbf_fail(ctx, APBF_ERR_UNSUPPORTED, "neighbour list was not produced by this context's search", __FILE__, __LINE__);
	const uint32_t n_cap = p.capacity;
	if (n_cap == 0) return APBF_OK;
	sweep_args A;
	memset(&A, 0, sizeof A);
	A.index_list = (const uint32_t*)p.index_list.data;
	A.len = p.length;
	A.pos4 = (int32_t*)p.position.data;
	A.inv_mass = (const float*)p.inverse_mass.data;
	A.radius = (const float*)p.radius.data;
	A.kernel_width = (float*)fluid->kernel_width.data;
	A.boundariness = (float*)fluid->boundariness.data;
	A.target_radius = (const float*)fluid->target_radius.data;
	A.pairs = nb->pairs;
	A.pair_cap = nb->capacity;
	A.offsets = (const uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	A.symbits = (uint32_t*)ctx->scratch_get(SLOT_SYMBITS, 4);
	A.L4 = (float4*)ctx->scratch_get(SLOT_L4, sizeof(float4) * (size_t)n_cap);
	A.G4 = (int4*)ctx->scratch_get(SLOT_G4, sizeof(int4) * (size_t)n_cap);
	A.delta = (int4*)ctx->scratch_get(SLOT_DELTA, sizeof(int4) * (size_t)n_cap);
	A.push = (int4*)ctx->scratch_get(SLOT_PUSH, sizeof(int4) * (size_t)n_cap);
	A.com4 = nullptr;
	A.misc = ctx->misc();
	A.out_lambda = out_lambda;
	A.out_incomp = out_incomp;
	A.s = ctx->settings;
	A.D = (float)ctx->dims;
	if (!A.offsets || !A.symbits || !A.L4 || !A.G4 || !A.delta || !A.push) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	const unsigned grid = apbf_grid(ctx, (size_t)n_cap * GROUP, SWEEP_THREADS, 8);
	{
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
	{
		apbf_prof_scope ps(ctx, PROF_APPLY_DELTA);
		switch (A.s.mGradientKernelId) {
			case 0: k_apply_delta<0><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
			case 1: k_apply_delta<1><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
			case 2: k_apply_delta<2><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
			case 3: k_apply_delta<3><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
			case 4: k_apply_delta<4><<<grid, SWEEP_THREADS, 0, ctx->stream>>>(A); break;
		}
		APBF_LAUNCHED(ctx);
	}
	{
		apbf_prof_scope ps(ctx, PROF_COMMIT);
		k_commit_delta<<<apbf_grid(ctx, n_cap, 256), 256, 0, ctx->stream>>>(A);
		APBF_LAUNCHED(ctx);
	}
	return APBF_OK;
}

extern "C" {

int apbf_incompressibility_apply(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, float* out_lambda,
                                 uint32_t* out_incomp_data)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid && nb);
	return apbf_incompressibility_run(ctx, fluid, nb, out_lambda, out_incomp_data);
}

int apbf_spread_kernel_width_apply(apbf_ctx* ctx, apbf_fluid* fluid, apbf_neighbors* nb, uint32_t* out_kw_fixed)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, fluid && nb);
	apbf_particles& p = fluid->particle;
	APBF_REQUIRE(ctx, p.index_list.data && p.length && p.position.data && p.radius.data && fluid->kernel_width.data && fluid->target_radius.data);
	if (!apbf_nbr_struct_valid(ctx, nb))
		return apbf_fail(ctx, APBF_ERR_UNSUPPORTED, "neighbour list was not produced by this context's search", __FILE__, __LINE__);
	const uint32_t n_cap = p.capacity;
	if (n_cap == 0) return APBF_OK;
	cudaStream_t st = ctx->stream;
	const size_t sym_words = ((size_t)nb->capacity + 31) / 32 + 1;
	kw_args A;
	memset(&A, 0, sizeof A);
	A.index_list = (const uint32_t*)p.index_list.data;
	A.len = p.length;
	A.pos4 = (const int32_t*)p.position.data;
	A.radius = (const float*)p.radius.data;
	A.target_radius = (const float*)fluid->target_radius.data;
	A.kernel_width = (float*)fluid->kernel_width.data;
	A.kwfx = (uint32_t*)ctx->scratch_get(SLOT_KWFX, sizeof(uint32_t) * (size_t)n_cap);
	A.pairs = nb->pairs;
	A.pair_cap = nb->capacity;
	uint32_t* offsets = (uint32_t*)ctx->scratch_get(SLOT_OFFSETS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	A.offsets = offsets;
	A.symbits = (uint32_t*)ctx->scratch_get(SLOT_SYMBITS, sizeof(uint32_t) * sym_words);
	A.keep_counts = (uint32_t*)ctx->scratch_get(SLOT_KEEP_COUNTS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	uint32_t* keep_offsets = (uint32_t*)ctx->scratch_get(SLOT_KEEP_OFFSETS, sizeof(uint32_t) * (size_t)(n_cap + 1));
	A.keep_offsets = keep_offsets;
	A.out_pairs = (uint32_t*)ctx->scratch_get(SLOT_PAIRS_TMP, sizeof(uint32_t) * 2 * (size_t)nb->capacity);
	A.out_symbits = (uint32_t*)ctx->scratch_get(SLOT_SYMBITS_TMP, sizeof(uint32_t) * sym_words);
	A.misc = ctx->misc();
	A.base_on_target_radius = ctx->settings.mBaseKernelWidthOnTargetRadius;
	A.speed = ctx->settings.mKernelWidthAdaptionSpeed;
	if (!A.kwfx || !offsets || !A.symbits || !A.keep_counts || !keep_offsets || !A.out_pairs || !A.out_symbits)
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
	APBF_CUDA(ctx, cudaMemsetAsync(A.out_symbits, 0, sizeof(uint32_t) * sym_words, st));
	k_reset_asym<<<1, 1, 0, st>>>(A.misc);
	APBF_LAUNCHED(ctx);
	k_kw_sweep<true><<<sgrid, SWEEP_THREADS, 0, st>>>(A);
	APBF_LAUNCHED(ctx);
	if (out_kw_fixed) APBF_CUDA(ctx, cudaMemcpyAsync(out_kw_fixed, A.kwfx, sizeof(uint32_t) * (size_t)n_cap, cudaMemcpyDeviceToDevice, st));
	k_kw_finish<<<egrid, 256, 0, st>>>(A);
	APBF_LAUNCHED(ctx);
	APBF_CUDA(ctx, cudaMemcpyAsync(nb->pairs, A.out_pairs, sizeof(uint32_t) * 2 * (size_t)nb->capacity, cudaMemcpyDeviceToDevice, st));
	APBF_CUDA(ctx, cudaMemcpyAsync(A.symbits, A.out_symbits, sizeof(uint32_t) * sym_words, cudaMemcpyDeviceToDevice, st));
	APBF_CUDA(ctx, cudaMemcpyAsync(offsets, keep_offsets, sizeof(uint32_t) * (size_t)(n_cap + 1), cudaMemcpyDeviceToDevice, st));
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
	    (const float4*)box_min4, (const float4*)box_max4, particles->length, n_boxes, ctx->misc(), 0);
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
