// mgpu.cu -- slab partition of one scene across the GPUs of a node (no reference counterpart: the reference is
// single-device).  Device side of the protocol; the exchanges themselves (NCCL send/recv between the ranks' staging
// buffers) are driven by the host (apbf_b200/multi_gpu.py).
//
// Partition: rank = top log2(world) bits of the particle's Z-curve cell key, i.e. axis-aligned bricks that are contiguous
// key ranges.  A rank's particles in global sorted order are then exactly its owned particles in local sorted order, so
// global id = gid_base + local id (box_collision hashes the id) and N-GPU results equal 1-GPU results bit for bit: every
// accumulator on the path is an integer.
//
// Per substep (host order): integrate owned -> ROUTE (particles that left the brick move, full state) -> HALO (owned
// particles inside another rank's brick grown by the halo width are sent as ghosts: position, mass, radius, widths) ->
// search over owned + ghosts (ghosts sort into a second key space, neighbors.cu) -> spread_kernel_width, then the owners'
// new widths overwrite the ghosts' -> per iteration: owners' packed positions refresh the ghosts before the density sweep,
// owners' lambdas refresh them before the apply sweep.  A ghost's own pair list only holds the unmirrored pairs onto owned
// particles (the scatter part of the sweeps); everything else about a ghost comes from its owner.
#include "keys.cuh"
#include "neighbors.cuh"
#include "sim.cuh"
#include "solver.cuh"
#include "sort.cuh"
#include <dlfcn.h>
#include <algorithm>

namespace {

struct mg_boxes {
	uint32_t lo[8][3], hi[8][3]; // bricks grown by the halo width, in cells, inclusive
	int      world, rank;
};

__device__ __forceinline__ void cell_of(const int32_t* pos4, uint32_t id, const apbf_grid_params& g, uint32_t c[3])
{
	const int4 p = ldg_int4(pos4, id);
	const uint32_t m = (1u << g.res) - 1u;
	c[0] = apbf_map_axis((float)p.x * INV_R_POS, g, 0) & m;
	c[1] = apbf_map_axis((float)p.y * INV_R_POS, g, 1) & m;
	c[2] = g.dims == 3 ? apbf_map_axis((float)p.z * INV_R_POS, g, 2) & m : 0u;
}

// destination rank of every owned particle + particles per destination
template <int DIMS>
__global__ void k_route_keys(const int32_t* __restrict__ pos4, const uint32_t* __restrict__ len, apbf_grid_params g, uint32_t key_bits,
                             uint32_t log2_world, uint32_t* __restrict__ dest, uint32_t* __restrict__ counts)
{
	const uint32_t n = *len;
	for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		uint32_t c[3];
		cell_of(pos4, id, g, c);
		const uint32_t key = apbf_zhash<DIMS>(c[0], c[1], c[2], g.res);
		const uint32_t d = log2_world ? key >> (key_bits - log2_world) : 0u;
		dest[id] = d;
		const uint32_t peers = __match_any_sync(__activemask(), d);
		if ((peers & ((1u << lane_id()) - 1u)) == 0u) atomicAdd(counts + d, (uint32_t)__popc(peers));
	}
}

// 80-byte record of one particle: every list of the scene (list_definitions.h:9-20)
struct state_lists {
	int4 *pos, *vel, *backup;
	uint32_t *inv_mass, *radius, *transferring, *target_radius, *kernel_width, *boundariness, *boundary_distance, *index_list;
};

__global__ void k_pack_state(state_lists L, uint32_t first, uint32_t count, int4* __restrict__ out)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const uint32_t id = first + k;
		int4* o = out + 5 * (size_t)k;
		o[0] = L.pos[id]; o[1] = L.vel[id]; o[2] = L.backup[id];
		o[3] = make_int4((int)L.inv_mass[id], (int)L.radius[id], (int)L.transferring[id], (int)L.target_radius[id]);
		o[4] = make_int4((int)L.kernel_width[id], (int)L.boundariness[id], (int)L.boundary_distance[id], 0);
	}
}

__global__ void k_unpack_state(state_lists L, uint32_t first, uint32_t count, const int4* __restrict__ in)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const uint32_t id = first + k;
		const int4* o = in + 5 * (size_t)k;
		L.pos[id] = o[0]; L.vel[id] = o[1]; L.backup[id] = o[2];
		const int4 a = o[3], b = o[4];
		L.inv_mass[id] = (uint32_t)a.x; L.radius[id] = (uint32_t)a.y; L.transferring[id] = (uint32_t)a.z; L.target_radius[id] = (uint32_t)a.w;
		L.kernel_width[id] = (uint32_t)b.x; L.boundariness[id] = (uint32_t)b.y; L.boundary_distance[id] = (uint32_t)b.z;
		L.index_list[id] = id;
	}
}

__global__ void k_copy_state(state_lists S, state_lists D, uint32_t src_first, uint32_t dst_first, uint32_t count)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const uint32_t s = src_first + k, d = dst_first + k;
		D.pos[d] = S.pos[s]; D.vel[d] = S.vel[s]; D.backup[d] = S.backup[s];
		D.inv_mass[d] = S.inv_mass[s]; D.radius[d] = S.radius[s]; D.transferring[d] = S.transferring[s];
		D.target_radius[d] = S.target_radius[s]; D.kernel_width[d] = S.kernel_width[s]; D.boundariness[d] = S.boundariness[s];
		D.boundary_distance[d] = S.boundary_distance[s];
		D.index_list[d] = d;
	}
}

// send lists: owned ids whose cell lies inside another rank's grown brick
__global__ void k_halo_lists(const int32_t* __restrict__ pos4, uint32_t n_owned, apbf_grid_params g, mg_boxes B, uint32_t* __restrict__ ids,
                             uint32_t cap, uint32_t* __restrict__ counts, uint32_t* flags)
{
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t base = blockIdx.x * blockDim.x; base < n_owned; base += stride) { // whole warps stay in the loop for the ballots
		const uint32_t id = base + threadIdx.x;
		uint32_t c[3] = { 0u, 0u, 0u };
		if (id < n_owned) cell_of(pos4, id, g, c);
		for (int r = 0; r < B.world; r++) {
			if (r == B.rank) continue;
			const bool inside = id < n_owned && c[0] >= B.lo[r][0] && c[0] <= B.hi[r][0] && c[1] >= B.lo[r][1] && c[1] <= B.hi[r][1] &&
			                    c[2] >= B.lo[r][2] && c[2] <= B.hi[r][2];
			const uint32_t m = __ballot_sync(0xffffffffu, inside);
			if (m == 0u) continue;
			uint32_t slot = 0u;
			if (lane_id() == 0u) slot = atomicAdd(counts + r, (uint32_t)__popc(m));
			slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(m & ((1u << lane_id()) - 1u));
			if (inside) {
				if (slot < cap) ids[(size_t)r * cap + slot] = id;
				else atomicOr(flags, 2u); // halo list overflow
			}
		}
	}
}

struct halo_lists {
	const int4* pos;
	const uint32_t *inv_mass, *radius, *kernel_width, *target_radius, *boundary_distance;
};
struct halo_lists_out {
	int4* pos;
	uint32_t *inv_mass, *radius, *kernel_width, *target_radius, *boundary_distance, *index_list;
};

__global__ void k_pack_halo(halo_lists L, const uint32_t* __restrict__ ids, uint32_t count, int4* __restrict__ out)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const uint32_t id = ids[k];
		out[3 * (size_t)k] = L.pos[id];
		out[3 * (size_t)k + 1] = make_int4((int)L.inv_mass[id], (int)L.radius[id], (int)L.kernel_width[id], (int)L.target_radius[id]);
		out[3 * (size_t)k + 2] = make_int4((int)L.boundary_distance[id], 0, 0, 0); // (update_transfers floods it across the bricks)
	}
}
__global__ void k_unpack_halo(halo_lists_out L, uint32_t first, uint32_t count, const int4* __restrict__ in)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const uint32_t id = first + k;
		L.pos[id] = in[3 * (size_t)k];
		const int4 a = in[3 * (size_t)k + 1];
		L.inv_mass[id] = (uint32_t)a.x; L.radius[id] = (uint32_t)a.y; L.kernel_width[id] = (uint32_t)a.z; L.target_radius[id] = (uint32_t)a.w;
		L.boundary_distance[id] = (uint32_t)in[3 * (size_t)k + 2].x;
		L.index_list[id] = id;
	}
}

__global__ void k_pack_u32(const uint32_t* __restrict__ src, uint32_t src_stride_words, const uint32_t* __restrict__ ids, uint32_t count,
                           uint32_t* __restrict__ out)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) out[k] = src[(size_t)ids[k] * src_stride_words];
}
__global__ void k_unpack_u32(uint32_t* __restrict__ dst, const uint32_t* __restrict__ ids, uint32_t count, const uint32_t* __restrict__ in)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) dst[ids[k]] = in[k];
}
__global__ void k_pack_16(const int4* __restrict__ src, const uint32_t* __restrict__ ids, uint32_t count, int4* __restrict__ out)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) out[k] = src[ids[k]];
}
__global__ void k_unpack_16(int4* __restrict__ dst, const uint32_t* __restrict__ ids, uint32_t count, const int4* __restrict__ in)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) dst[ids[k]] = in[k];
}
// a ghost's record for the apply sweep: {lambda from its owner, h, gradient c0, gradient c1 from the local constants}
// (PL: the one-gather record {position, lambda * 2^18 or 0} of the apply sweep's equal-width form, incompress.cu; the ghost's packed
// position arrived with the previous exchange)
__global__ void k_unpack_lambda(float4* __restrict__ L4, const float4* __restrict__ KG, const uint32_t* __restrict__ ids, uint32_t count,
                                const float* __restrict__ in, const int4* __restrict__ P4, int4* __restrict__ PL)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const uint32_t id = ids[k];
		const float4 kg = KG[id];
		const float lam = in[k];
		L4[id] = make_float4(lam, kg.x, kg.y, kg.z);
		const int4 p = P4[id];
		PL[id] = make_int4(p.x, p.y, p.z, __float_as_int(lam < 0.0f ? lam * R_POS : 0.0f));
	}
}

__global__ void k_inverse_perm(const uint32_t* __restrict__ sorted_index, const uint32_t* __restrict__ len, uint32_t* __restrict__ inv)
{
	const uint32_t n = *len;
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x) inv[sorted_index[h]] = h;
}
__global__ void k_remap(uint32_t* __restrict__ ids, uint32_t count, const uint32_t* __restrict__ inv)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) ids[k] = inv[ids[k]];
}
__global__ void k_iota_from(uint32_t* __restrict__ ids, uint32_t count, uint32_t first)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) ids[k] = first + k;
}

__global__ void k_set_counts(uint32_t* index_len, uint32_t* hidden_len, uint32_t* misc, uint32_t n_owned, uint32_t n_total, uint32_t gid_base)
{
	*index_len = n_total; *hidden_len = n_total;
	misc[MW_N_OWNED] = n_owned; misc[MW_GID_BASE] = gid_base;
}

state_lists lists_of(apbf_sim* sim, bool other)
{
	apbf_fluid& f = sim->fluid;
	auto pick = [&](apbf_array& a) { return other ? a.reorder_out : a.data; };
	state_lists L;
	L.pos = (int4*)pick(f.particle.position); L.vel = (int4*)pick(f.particle.velocity); L.backup = (int4*)pick(f.particle.pos_backup);
	L.inv_mass = (uint32_t*)pick(f.particle.inverse_mass); L.radius = (uint32_t*)pick(f.particle.radius);
	L.transferring = (uint32_t*)pick(f.particle.transferring); L.target_radius = (uint32_t*)pick(f.target_radius);
	L.kernel_width = (uint32_t*)pick(f.kernel_width); L.boundariness = (uint32_t*)pick(f.boundariness);
	L.boundary_distance = (uint32_t*)pick(f.boundary_distance); L.index_list = (uint32_t*)pick(f.particle.index_list);
	return L;
}

int grid_of(apbf_sim* sim, apbf_grid_params* g)
{
	APBF_TRY(apbf_ctx_set_dimensions(sim->ctx, sim->cfg.dims));
	return apbf_make_grid_params(sim->ctx, sim->cfg.min_pos, sim->cfg.max_pos, sim->cfg.res_log2, g);
}

} // namespace

extern "C" {

int apbf_sim_mg_enable(apbf_sim* sim, int rank, int world, float halo_range)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, world == 1 || world == 2 || world == 4 || world == 8);
	APBF_REQUIRE(ctx, rank >= 0 && rank < world && halo_range >= 0.0f);
	const int dims = sim->cfg.dims;
	const uint32_t res = sim->cfg.res_log2, cells = 1u << res;
	int lw = 0;
	while ((1 << lw) < world) lw++;
	APBF_REQUIRE(ctx, (uint32_t)lw <= res * (uint32_t)dims);
	apbf_mg_state& m = sim->mg;
	m.enabled = true; m.rank = rank; m.world = world;
	// the top key bit is bit res-1 of the last axis, then of the axis before it, ... (bit i of axis d -> bit i*D+d)
	for (int r = 0; r < world; r++) {
		uint32_t lo[3] = { 0u, 0u, 0u }, size[3] = { cells, cells, dims == 3 ? cells : 1u };
		for (int b = 0; b < lw; b++) {
			const int level = b / dims, axis = dims - 1 - (b % dims);
			(void)level;
			size[axis] >>= 1;
			if ((r >> (lw - 1 - b)) & 1) lo[axis] += size[axis];
		}
		for (int d = 0; d < 3; d++) { m.lo[r][d] = lo[d]; m.hi[r][d] = lo[d] + size[d] - 1u; }
	}
	for (int d = 0; d < 3; d++) {
		const float cell = (sim->cfg.max_pos[d] - sim->cfg.min_pos[d]) / (float)cells;
		m.halo[d] = (d < dims && cell > 0.0f) ? (uint32_t)ceilf(halo_range / cell) + 1u : 0u; // +1: rounding of the cell map
	}
	ctx->mg_enabled = world > 1;
	for (int d = 0; d < 3; d++) { ctx->mg_lo[d] = m.lo[rank][d]; ctx->mg_hi[d] = m.hi[rank][d]; }
	return APBF_OK;
}

int apbf_sim_mg_brick(apbf_sim* sim, int rank, uint32_t out_lo[3], uint32_t out_hi[3], uint32_t out_halo[3])
{
	if (!sim || !sim->mg.enabled || rank < 0 || rank >= sim->mg.world) return APBF_ERR_INVALID;
	for (int d = 0; d < 3; d++) { out_lo[d] = sim->mg.lo[rank][d]; out_hi[d] = sim->mg.hi[rank][d]; out_halo[d] = sim->mg.halo[d]; }
	return APBF_OK;
}

int apbf_sim_mg_set_counts(apbf_sim* sim, uint32_t n_owned, uint32_t n_total, uint32_t gid_base)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, n_owned <= n_total && n_total <= sim->cfg.particle_capacity);
	sim->mg.n_owned = n_owned; sim->mg.n_total = n_total;
	k_set_counts<<<1, 1, 0, ctx->stream>>>(sim->fluid.particle.length, sim->fluid.particle.hidden_length, ctx->misc(), n_owned, n_total, gid_base);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

// Groups the owned particles by destination rank (stable: local order is kept inside a group) into the lists' other buffers
// and swaps; counts_dev[world] receives the group sizes.  Call with lengths == n_owned.
// destination rank of every owned particle, particles per destination, and the stable order by destination (perm: SLOT_TMP_VALS).
// move_lists: also bring every list into that order (the Python-driven protocol packs contiguous segments); the library's own loop
// packs and compacts THROUGH perm instead, which saves one copy of the whole state per substep.
static int route_plan_and_move(apbf_sim* sim, uint32_t* counts_dev, bool move_lists)
{
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mg.enabled);
	cudaStream_t st = ctx->stream;
	apbf_grid_params g;
	APBF_TRY(grid_of(sim, &g));
	const uint32_t cap = sim->cfg.particle_capacity;
	int lw = 0;
	while ((1 << lw) < sim->mg.world) lw++;
	uint32_t* dest = (uint32_t*)ctx->scratch_get(SLOT_SORT_KEYS_A, sizeof(uint32_t) * (size_t)cap);
	uint32_t* sdest = (uint32_t*)ctx->scratch_get(SLOT_TMP_KEYS, sizeof(uint32_t) * (size_t)cap);
	uint32_t* perm = (uint32_t*)ctx->scratch_get(SLOT_TMP_VALS, sizeof(uint32_t) * (size_t)cap);
	if (!dest || !sdest || !perm) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	APBF_CUDA(ctx, cudaMemsetAsync(counts_dev, 0, sizeof(uint32_t) * 8, st));
	const uint32_t key_bits = sim->cfg.res_log2 * (uint32_t)sim->cfg.dims;
	const uint32_t* len = sim->fluid.particle.length;
	if (g.dims == 3) k_route_keys<3><<<apbf_grid(ctx, cap, 256), 256, 0, st>>>((const int32_t*)sim->fluid.particle.position.data, len, g, key_bits, (uint32_t)lw, dest, counts_dev);
	else k_route_keys<2><<<apbf_grid(ctx, cap, 256), 256, 0, st>>>((const int32_t*)sim->fluid.particle.position.data, len, g, key_bits, (uint32_t)lw, dest, counts_dev);
	APBF_LAUNCHED(ctx);
	APBF_TRY(apbf_radix_sort_pairs(ctx, dest, nullptr, sdest, perm, len, cap, lw > 0 ? lw : 1));
	if (!move_lists) return APBF_OK;
	apbf_fluid& f = sim->fluid;
	apbf_reorder_table t; // every list of the scene in one pass (16-byte loads), like the search's own reorder
	memset(&t, 0, sizeof t);
	apbf_array* a16[3] = { &f.particle.position, &f.particle.velocity, &f.particle.pos_backup };
	apbf_array* a4[7] = { &f.particle.inverse_mass, &f.particle.radius, &f.particle.transferring, &f.target_radius, &f.kernel_width,
	                      &f.boundariness, &f.boundary_distance };
	for (auto a : a16) { t.src16[t.n16] = (const int4*)a->data; t.dst16[t.n16] = (int4*)a->reorder_out; t.n16++; }
	for (auto a : a4) { t.src4[t.n4] = (const uint32_t*)a->data; t.dst4[t.n4] = (uint32_t*)a->reorder_out; t.n4++; }
	APBF_TRY(apbf_launch_reorder(ctx, t, perm, len, cap));
	APBF_TRY(apbf_write_sequence(ctx, (uint32_t*)f.particle.index_list.reorder_out, len, cap, 0u, 1u, 1u));
	apbf_sim_swap_buffers(sim);
	return APBF_OK;
}

int apbf_sim_mg_route(apbf_sim* sim, uint32_t* counts_dev)
{
	if (!sim || !counts_dev) return APBF_ERR_INVALID;
	return route_plan_and_move(sim, counts_dev, true);
}

int apbf_sim_mg_pack_state(apbf_sim* sim, uint32_t first, uint32_t count, void* out)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	if (count == 0) return APBF_OK;
	APBF_REQUIRE(ctx, out && (size_t)first + count <= sim->cfg.particle_capacity);
	k_pack_state<<<apbf_grid(ctx, count, 256), 256, 0, ctx->stream>>>(lists_of(sim, false), first, count, (int4*)out);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

// into_other != 0: write into the lists' other buffers (the re-partitioned scene is assembled there, then apbf_sim_mg_swap)
int apbf_sim_mg_unpack_state(apbf_sim* sim, uint32_t first, uint32_t count, const void* in, int into_other)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	if (count == 0) return APBF_OK;
	APBF_REQUIRE(ctx, in && (size_t)first + count <= sim->cfg.particle_capacity);
	k_unpack_state<<<apbf_grid(ctx, count, 256), 256, 0, ctx->stream>>>(lists_of(sim, into_other != 0), first, count, (const int4*)in);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_sim_mg_copy_state(apbf_sim* sim, uint32_t src_first, uint32_t dst_first, uint32_t count)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	if (count == 0) return APBF_OK;
	APBF_REQUIRE(ctx, (size_t)src_first + count <= sim->cfg.particle_capacity && (size_t)dst_first + count <= sim->cfg.particle_capacity);
	k_copy_state<<<apbf_grid(ctx, count, 256), 256, 0, ctx->stream>>>(lists_of(sim, false), lists_of(sim, true), src_first, dst_first, count);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

int apbf_sim_mg_swap(apbf_sim* sim)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_sim_swap_buffers(sim);
	return APBF_OK;
}

// ids_dev[world][cap_per_dest]: owned ids every other rank needs as ghosts; counts_dev[world] (8 words are cleared)
int apbf_sim_mg_halo_lists(apbf_sim* sim, uint32_t* ids_dev, uint32_t cap_per_dest, uint32_t* counts_dev)
{
	if (!sim || !ids_dev || !counts_dev) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mg.enabled);
	apbf_grid_params g;
	APBF_TRY(grid_of(sim, &g));
	const apbf_mg_state& m = sim->mg;
	const uint32_t cells = 1u << sim->cfg.res_log2;
	mg_boxes B;
	memset(&B, 0, sizeof B);
	B.world = m.world; B.rank = m.rank;
	for (int r = 0; r < m.world; r++)
		for (int d = 0; d < 3; d++) {
			B.lo[r][d] = m.lo[r][d] > m.halo[d] ? m.lo[r][d] - m.halo[d] : 0u;
			B.hi[r][d] = (d < sim->cfg.dims) ? (m.hi[r][d] + m.halo[d] < cells - 1u ? m.hi[r][d] + m.halo[d] : cells - 1u) : 0u;
		}
	APBF_CUDA(ctx, cudaMemsetAsync(counts_dev, 0, sizeof(uint32_t) * 8, ctx->stream));
	if (m.n_owned > 0) {
		k_halo_lists<<<apbf_grid(ctx, m.n_owned, 256), 256, 0, ctx->stream>>>((const int32_t*)sim->fluid.particle.position.data, m.n_owned, g, B, ids_dev,
		                                                                     cap_per_dest, counts_dev, ctx->misc() + MW_FLAGS);
		APBF_LAUNCHED(ctx);
	}
	return APBF_OK;
}

// what: 0 halo record (48 B: position, inverse mass, radius, kernel width, target radius, boundary distance), 1 kernel width
//       (4 B), 2 packed position of the solver (16 B), 3 lambda (4 B), 4 committed position (16 B)
int apbf_sim_mg_pack(apbf_sim* sim, int what, const uint32_t* ids_dev, uint32_t count, void* out)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	if (count == 0) return APBF_OK;
	APBF_REQUIRE(ctx, ids_dev && out);
	apbf_fluid& f = sim->fluid;
	const unsigned grid = apbf_grid(ctx, count, 256);
	const uint32_t cap = sim->cfg.particle_capacity;
	cudaStream_t st = ctx->stream;
	if (what == 0) {
		halo_lists L{ (const int4*)f.particle.position.data, (const uint32_t*)f.particle.inverse_mass.data, (const uint32_t*)f.particle.radius.data,
		              (const uint32_t*)f.kernel_width.data, (const uint32_t*)f.target_radius.data, (const uint32_t*)f.boundary_distance.data };
		k_pack_halo<<<grid, 256, 0, st>>>(L, ids_dev, count, (int4*)out);
	} else if (what == 1) {
		k_pack_u32<<<grid, 256, 0, st>>>((const uint32_t*)f.kernel_width.data, 1u, ids_dev, count, (uint32_t*)out);
	} else if (what == 2) {
		k_pack_16<<<grid, 256, 0, st>>>((const int4*)ctx->scratch_get(SLOT_P4, sizeof(int4) * (size_t)cap), ids_dev, count, (int4*)out);
	} else if (what == 3) {
		k_pack_u32<<<grid, 256, 0, st>>>((const uint32_t*)ctx->scratch_get(SLOT_L4, sizeof(float4) * (size_t)cap), 4u, ids_dev, count, (uint32_t*)out);
	} else if (what == 4) { // committed positions (update_transfers measures distances after the solver)
		k_pack_16<<<grid, 256, 0, st>>>((const int4*)f.particle.position.data, ids_dev, count, (int4*)out);
	} else {
		return apbf_fail(ctx, APBF_ERR_INVALID, "what", __FILE__, __LINE__);
	}
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

// what == 0: ghosts are appended at [first, first + count) (ids_dev unused); otherwise ids_dev[k] is the ghost's id
int apbf_sim_mg_unpack(apbf_sim* sim, int what, const uint32_t* ids_dev, uint32_t first, uint32_t count, const void* in)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	if (count == 0) return APBF_OK;
	APBF_REQUIRE(ctx, in && (what == 0 || ids_dev));
	apbf_fluid& f = sim->fluid;
	const unsigned grid = apbf_grid(ctx, count, 256);
	const uint32_t cap = sim->cfg.particle_capacity;
	cudaStream_t st = ctx->stream;
	if (what == 0) {
		APBF_REQUIRE(ctx, (size_t)first + count <= cap);
		halo_lists_out L{ (int4*)f.particle.position.data, (uint32_t*)f.particle.inverse_mass.data, (uint32_t*)f.particle.radius.data,
		                  (uint32_t*)f.kernel_width.data, (uint32_t*)f.target_radius.data, (uint32_t*)f.boundary_distance.data,
		                  (uint32_t*)f.particle.index_list.data };
		k_unpack_halo<<<grid, 256, 0, st>>>(L, first, count, (const int4*)in);
	} else if (what == 1) {
		k_unpack_u32<<<grid, 256, 0, st>>>((uint32_t*)f.kernel_width.data, ids_dev, count, (const uint32_t*)in);
	} else if (what == 2) {
		k_unpack_16<<<grid, 256, 0, st>>>((int4*)ctx->scratch_get(SLOT_P4, sizeof(int4) * (size_t)cap), ids_dev, count, (const int4*)in);
	} else if (what == 3) {
		k_unpack_lambda<<<grid, 256, 0, st>>>((float4*)ctx->scratch_get(SLOT_L4, sizeof(float4) * (size_t)cap),
		                                      (const float4*)ctx->scratch_get(SLOT_KG, sizeof(float4) * (size_t)cap), ids_dev, count, (const float*)in,
		                                      (const int4*)ctx->scratch_get(SLOT_P4, sizeof(int4) * (size_t)cap), (int4*)ctx->scratch_get(SLOT_PL, sizeof(int4) * (size_t)cap));
	} else if (what == 4) {
		k_unpack_16<<<grid, 256, 0, st>>>((int4*)f.particle.position.data, ids_dev, count, (const int4*)in);
	} else {
		return apbf_fail(ctx, APBF_ERR_INVALID, "what", __FILE__, __LINE__);
	}
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

// after the search: ids_dev[k] (slots before the sort) -> ids after the sort; first != 0xFFFFFFFF: ids_dev[k] = first + k first
int apbf_sim_mg_remap(apbf_sim* sim, uint32_t* ids_dev, uint32_t count, uint32_t first)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	if (count == 0) return APBF_OK;
	APBF_REQUIRE(ctx, ids_dev);
	const uint32_t* inv = (const uint32_t*)ctx->scratch_get(SLOT_MG_INV, sizeof(uint32_t) * (size_t)sim->cfg.particle_capacity);
	if (!inv) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	if (first != 0xFFFFFFFFu) {
		k_iota_from<<<apbf_grid(ctx, count, 256), 256, 0, ctx->stream>>>(ids_dev, count, first);
		APBF_LAUNCHED(ctx);
	}
	k_remap<<<apbf_grid(ctx, count, 256), 256, 0, ctx->stream>>>(ids_dev, count, inv);
	APBF_LAUNCHED(ctx);
	return APBF_OK;
}

// phase: 0 integrate, 1 search, 2 spread_kernel_width, 3 solver constants, 4 iteration prologue (commit of the previous
// iteration, box collision, pack), 5 density/lambda sweep, 6 apply sweep, 7 final commit, 8 kernel width from the boundary
// distance, 9 update_transfers
int apbf_sim_mg_phase(apbf_sim* sim, int phase, int iteration)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	const apbf_sim_config& c = sim->cfg;
	APBF_TRY(apbf_ctx_set_dimensions(ctx, c.dims));
	const apbf_settings& s = ctx->settings;
	const bool unit_scale = c.basic_pbf || s.mBaseKernelWidthOnBoundaryDistance;
	const float* bmin = sim->boxes;
	const float* bmax = sim->boxes ? sim->boxes + 4 * (size_t)c.n_boxes : nullptr;
	switch (phase) {
		case 0:
			APBF_TRY(apbf_velocity_handling_apply(ctx, &sim->fluid.particle, c.dt, sim->last_dt, c.accel));
			sim->last_dt = c.dt;
			return APBF_OK;
		case 1: {
			const bool adaptive = !c.basic_pbf && !s.mBaseKernelWidthOnBoundaryDistance; // spread_kernel_width will prune
			sim->mg_fused = adaptive && !sim->no_fuse;
			ctx->mg_ghost_all_pairs = adaptive && !sim->mg_fused;
			// (the sweeps and a separate spread work on NB + offsets: no public pair list, write_public = false)
			if (c.use_binary_search) // NEIGHBORHOOD_TYPE 3 over owned particles + ghosts (the bricks themselves are cells of the Green grid)
				APBF_TRY(apbf_binary_search(ctx, &sim->fluid, &sim->fluid.kernel_width, &sim->nb, sim->mg_fused ? 1.5f : (unit_scale ? 1.0f : 1.5f), nullptr, sim->mg_fused, nullptr, false));
			else if (sim->mg_fused) // search + spread_kernel_width in one pass (pool.cpp:83-89), ghosts included
				APBF_TRY(apbf_green_search(ctx, &sim->fluid, &sim->fluid.kernel_width, &sim->nb, 1.5f, c.min_pos, c.max_pos, c.res_log2, nullptr, true, nullptr, false));
			else
				APBF_TRY(apbf_green_search(ctx, &sim->fluid, &sim->fluid.kernel_width, &sim->nb, unit_scale ? 1.0f : 1.5f, c.min_pos, c.max_pos, c.res_log2, nullptr, false, nullptr, false));
			apbf_sim_swap_buffers(sim);
			// old slot -> new id, for the send lists and the ghost slots (the search's sorted_index is still in scratch)
			const uint32_t cap = c.particle_capacity;
			uint32_t* inv = (uint32_t*)ctx->scratch_get(SLOT_MG_INV, sizeof(uint32_t) * (size_t)cap);
			const uint32_t* sidx = (const uint32_t*)ctx->scratch_get(c.use_binary_search ? SLOT_SORT_VALS_A : SLOT_TMP_VALS, sizeof(uint32_t) * (size_t)cap);
			if (!inv || !sidx) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
			k_inverse_perm<<<apbf_grid(ctx, cap, 256), 256, 0, ctx->stream>>>(sidx, sim->fluid.particle.hidden_length, inv);
			APBF_LAUNCHED(ctx);
			return APBF_OK;
		}
		case 2: return sim->mg_fused ? APBF_OK : apbf_spread_kernel_width_apply(ctx, &sim->fluid, &sim->nb, nullptr);
		case 3: return apbf_solver_prepare(ctx, &sim->fluid);
		// (phases 10 / 11: the apply sweep commits and runs the next prologue itself where it can -- solver.cuh, ITER_T2_COMMIT; the
		// prologue / commit launch behind it is told to return at once in that case)
		case 4: {
			const int skip = sim->mg_t2_tail_pending && iteration > 0 ? ITER_SKIP_IF_T2_DID : 0;
			sim->mg_t2_tail_pending = false;
			return apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_BEGIN | ITER_BEGIN_BOX | (iteration > 0 ? ITER_BEGIN_COMMIT : 0) | skip, bmin, bmax, c.n_boxes, nullptr, nullptr);
		}
		case 5: return apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_T1, bmin, bmax, c.n_boxes, nullptr, nullptr);
		case 6: return apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_T2, bmin, bmax, c.n_boxes, nullptr, nullptr);
		case 10: // apply sweep, another iteration follows     case 11: apply sweep of the last iteration
		case 11:
			sim->mg_t2_tail_pending = true;
			return apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_T2 | ITER_T2_COMMIT | (phase == 10 ? ITER_T2_NEXT_BOX : 0), bmin, bmax, c.n_boxes, nullptr, nullptr);
		case 7: {
			const int skip = sim->mg_t2_tail_pending ? ITER_SKIP_IF_T2_DID : 0;
			sim->mg_t2_tail_pending = false;
			return apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_BEGIN | ITER_BEGIN_COMMIT | skip, nullptr, nullptr, 0u, nullptr, nullptr);
		}
		case 8: return apbf_kernel_width_from_boundary_distance(ctx, &sim->fluid); // pool.cpp:77-80 (owned particles; call before ROUTE)
		case 9: return apbf_update_transfers_apply(ctx, &sim->fluid, &sim->nb, nullptr); // pool.cpp:99-102, after the ghosts' positions came in
	}
	return apbf_fail(ctx, APBF_ERR_INVALID, "phase", __FILE__, __LINE__);
}

} // extern "C"

// ---- the library's own NCCL communicator ---------------------------------------------------------------------------------------
// Driving every exchange from the host language costs a pack call, a torch collective (with two stream hops) and an unpack
// call per exchange, ten times per substep.  With its own communicator the library runs pack -> ncclSend/ncclRecv -> unpack
// on the context's stream, and the solver loop of a substep is ONE call (apbf_sim_mg_solve).  NCCL is bound at run time
// (dlopen of the libnccl.so.2 the process already uses): the shared library itself does not link against it.
namespace {
struct nccl_uid { char internal[128]; };
typedef void* nccl_comm;
struct nccl_api {
	void* so = nullptr;
	int (*GetUniqueId)(nccl_uid*) = nullptr;
	int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
	int (*CommDestroy)(nccl_comm) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	int (*Send)(const void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};
constexpr int NCCL_INT32 = 2; // ncclInt32 (nccl.h: ncclInt8 0, ncclUint8 1, ncclInt32 2)

nccl_api* nccl()
{
	static nccl_api api;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (so) {
			api.GetUniqueId = (int (*)(nccl_uid*))dlsym(so, "ncclGetUniqueId");
			api.CommInitRank = (int (*)(nccl_comm*, int, nccl_uid, int))dlsym(so, "ncclCommInitRank");
			api.CommDestroy = (int (*)(nccl_comm))dlsym(so, "ncclCommDestroy");
			api.GroupStart = (int (*)())dlsym(so, "ncclGroupStart");
			api.GroupEnd = (int (*)())dlsym(so, "ncclGroupEnd");
			api.Send = (int (*)(const void*, size_t, int, int, nccl_comm, cudaStream_t))dlsym(so, "ncclSend");
			api.Recv = (int (*)(void*, size_t, int, int, nccl_comm, cudaStream_t))dlsym(so, "ncclRecv");
			api.GetErrorString = (const char* (*)(int))dlsym(so, "ncclGetErrorString");
			if (api.GetUniqueId && api.CommInitRank && api.GroupStart && api.GroupEnd && api.Send && api.Recv) api.so = so;
		}
	}
	return api.so ? &api : nullptr;
}

#define APBF_NCCL(ctx, expr)                                                                                            \
	do {                                                                                                                \
		int r__ = (expr);                                                                                               \
		if (r__ != 0) return apbf_fail(ctx, APBF_ERR_CUDA, nccl()->GetErrorString ? nccl()->GetErrorString(r__) : "nccl", __FILE__, __LINE__); \
	} while (0)

struct mg_lists {
	const uint32_t* send_ids;    // concatenated send lists, destination after destination
	const uint32_t* ghost_ids;   // ghost slots, source after source
	uint32_t send_counts[8], ghost_counts[8];
	uint32_t n_send, n_ghost;
};

// one halo exchange of `what` (1 kernel width, 2 packed position, 3 lambda) on the context's stream
int mg_exchange(apbf_sim* sim, const mg_lists& L, int what)
{
	apbf_ctx* ctx = sim->ctx;
	nccl_api* N = nccl();
	if (!N || !sim->nccl_comm) return apbf_fail(ctx, APBF_ERR_UNSUPPORTED, "apbf_sim_mg_comm_init has not been called", __FILE__, __LINE__);
	const uint32_t words = what == 2 ? 4u : 1u;
	uint32_t* stage_s = (uint32_t*)ctx->scratch_get(SLOT_MG_SEND, sizeof(uint32_t) * 4 * (size_t)(L.n_send ? L.n_send : 1));
	uint32_t* stage_r = (uint32_t*)ctx->scratch_get(SLOT_MG_RECV, sizeof(uint32_t) * 4 * (size_t)(L.n_ghost ? L.n_ghost : 1));
	if (!stage_s || !stage_r) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	APBF_TRY(apbf_sim_mg_pack(sim, what, L.send_ids, L.n_send, stage_s));
	APBF_NCCL(ctx, N->GroupStart());
	size_t so = 0, ro = 0;
	int first_err = 0; // the group is closed whatever a call inside it returned
	for (int r = 0; r < sim->mg.world; r++) {
		if (r == sim->mg.rank) continue;
		if (L.send_counts[r]) { const int e = N->Send(stage_s + so * words, (size_t)L.send_counts[r] * words, NCCL_INT32, r, sim->nccl_comm, ctx->stream); if (e && !first_err) first_err = e; }
		if (L.ghost_counts[r]) { const int e = N->Recv(stage_r + ro * words, (size_t)L.ghost_counts[r] * words, NCCL_INT32, r, sim->nccl_comm, ctx->stream); if (e && !first_err) first_err = e; }
		so += L.send_counts[r]; ro += L.ghost_counts[r];
	}
	{ const int e = N->GroupEnd(); if (e && !first_err) first_err = e; }
	APBF_NCCL(ctx, first_err);
	APBF_TRY(apbf_sim_mg_unpack(sim, what, L.ghost_ids, 0u, L.n_ghost, stage_r));
	return APBF_OK;
}
} // namespace

extern "C" {

int apbf_mg_nccl_unique_id(void* out_id128)
{
	nccl_api* N = nccl();
	if (!N || !out_id128) return APBF_ERR_UNSUPPORTED;
	nccl_uid id;
	if (N->GetUniqueId(&id) != 0) return APBF_ERR_CUDA;
	memcpy(out_id128, &id, sizeof id);
	return APBF_OK;
}

void apbf_sim_mg_comm_destroy(apbf_sim* sim)
{
	if (sim) { // the other ranks' arenas (peer-to-peer transport)
		for (int r = 0; r < 8; r++)
			if (sim->mgl.peer_base[r]) { cudaIpcCloseMemHandle(sim->mgl.peer_base[r]); sim->mgl.peer_base[r] = nullptr; }
		sim->mgl.p2p = false;
	}
	nccl_api* N = nccl();
	if (sim && sim->nccl_comm && N && N->CommDestroy) N->CommDestroy(sim->nccl_comm);
	if (sim) sim->nccl_comm = nullptr;
}

int apbf_sim_mg_comm_init(apbf_sim* sim, const void* id128, int rank, int world)
{
	if (!sim || !id128) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	nccl_api* N = nccl();
	if (!N) return apbf_fail(ctx, APBF_ERR_UNSUPPORTED, "libnccl.so.2 not found", __FILE__, __LINE__);
	APBF_REQUIRE(ctx, sim->mg.enabled && rank == sim->mg.rank && world == sim->mg.world && !sim->nccl_comm);
	APBF_CUDA(ctx, cudaSetDevice(ctx->device));
	nccl_uid id;
	memcpy(&id, id128, sizeof id);
	APBF_NCCL(ctx, N->CommInitRank(&sim->nccl_comm, world, id, rank));
	return APBF_OK;
}

// spread_kernel_width's new widths to the ghosts (if adaptive), solver constants, then `iterations` x (prologue, packed positions
// to the ghosts, density/lambda sweep, lambdas to the ghosts, apply sweep) and the final commit: phases 3-7 of apbf_sim_mg_phase
// with the exchanges in between, in one call.  send_ids: the send lists of all destinations, concatenated; counts per rank.
int apbf_sim_mg_solve(apbf_sim* sim, const uint32_t* send_ids_dev, const uint32_t* send_counts, const uint32_t* ghost_ids_dev,
                      const uint32_t* ghost_counts, int exchange_kernel_width, int iterations)
{
	if (!sim || !send_counts || !ghost_counts) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mg.enabled && sim->mg.world <= 8);
	mg_lists L;
	memset(&L, 0, sizeof L);
	L.send_ids = send_ids_dev; L.ghost_ids = ghost_ids_dev;
	for (int r = 0; r < sim->mg.world; r++) {
		L.send_counts[r] = r == sim->mg.rank ? 0u : send_counts[r];
		L.ghost_counts[r] = r == sim->mg.rank ? 0u : ghost_counts[r];
		L.n_send += L.send_counts[r]; L.n_ghost += L.ghost_counts[r];
	}
	APBF_REQUIRE(ctx, (L.n_send == 0 || send_ids_dev) && (L.n_ghost == 0 || ghost_ids_dev));
	if (exchange_kernel_width) APBF_TRY(mg_exchange(sim, L, 1));
	APBF_TRY(apbf_sim_mg_phase(sim, 3, 0));
	for (int it = 0; it < iterations; it++) {
		APBF_TRY(apbf_sim_mg_phase(sim, 4, it));
		APBF_TRY(mg_exchange(sim, L, 2));
		APBF_TRY(apbf_sim_mg_phase(sim, 5, it));
		APBF_TRY(mg_exchange(sim, L, 3));
		APBF_TRY(apbf_sim_mg_phase(sim, it + 1 < iterations ? 10 : 11, it));
	}
	return apbf_sim_mg_phase(sim, 7, 0);
}

} // extern "C"


// =====================================================================================================================================
// The whole substep of one scene in bricks inside the library (apbf_sim_mg_substep): pool::update (source/pool.cpp:67-106) cut where
// data of other ranks is needed, everything -- kernels AND exchanges -- enqueued on the context's stream, no host read-back anywhere.
//
//   integrate owned -> ROUTE -> HALO -> search over owned + ghosts -> [widths to the ghosts] -> constants ->
//   iterations x (prologue, packed positions to the ghosts, density / lambda sweep, lambdas to the ghosts, apply sweep) -> commit
//
// The host never learns how many particles migrate or how many ghosts there are: every count is a device word, every kernel is sized
// by a capacity and reads its count on the device, and every message has a FIXED size (the capacity agreed at set-up, count in a
// 16-byte header).  NVLink moves the padding for free (a few MB per exchange); what a substep pays for is the latency of eleven
// grouped ncclSend / ncclRecv rounds.  Same order of the lists as the host-driven protocol (apbf_b200/multi_gpu.py), hence the same
// results: N ranks == 1 rank, bit for bit.
// =====================================================================================================================================
namespace {

enum mgl_word {
	MGL_N_OWNED = 0, MGL_N_TOTAL = 1, MGL_GID_BASE = 2, MGL_MIGRATED = 3, MGL_FLAGS = 4, MGL_STAY_SRC = 5, MGL_STAY_DST = 6, MGL_STAY = 7,
	MGL_ROUTE = 8,        // [8] particles of this rank per destination after the integrator
	MGL_ARRIVE = 16,      // [8] arrivals per source
	MGL_ARRIVE_DST = 24,  // [8] where the arrivals of source r go in the new order
	MGL_HALO_SEND = 32,   // [8] ghosts this rank provides to rank r
	MGL_HALO_RECV = 40,   // [8] ghosts this rank holds from rank r
	MGL_GHOST_FIRST = 48, // [8] id of the first ghost from rank r (before the search's sort)
	MGL_WORDS = 64
};
constexpr uint32_t MGL_FLAG_ROUTE_OVERFLOW = 1u, MGL_FLAG_HALO_OVERFLOW = 2u, MGL_FLAG_CAPACITY = 4u;
constexpr uint32_t STATE_INT4 = 5u, HALO_INT4 = 3u; // 16-byte units per record

struct mgl_bufs { int4* p[8]; };
struct mgl_caps { uint32_t cap[8], off[8]; int world, rank; };

// ---- peer-to-peer transport: store into the receiver's buffer, raise a flag there, wait on one's own flags -------------------------
// One exchange = every rank sends one message to every other rank.  Exchange number `seq` (1, 2, ... -- the same on all ranks,
// the protocol is symmetric) uses buffer seq & 1 of each (source, destination) pair and sets the destination's flag [source][seq & 1]
// to seq once the message is complete.  Why two buffers are enough: rank A starts writing message seq only after its own consumer
// of seq - 1 ran (stream order), which waited for B's message seq - 1, which B packed after ITS consumer of seq - 2 -- the last
// reader of the buffer A is about to overwrite.  Nobody waits inside a pack kernel, so the waits cannot form a cycle.
struct mgl_sig {
	uint32_t*       remote_flag[8]; // in rank r's arena: the two flags for messages from this rank
	const uint32_t* local_flag;     // this rank's flags[8][2]
	uint32_t*       done;           // CTAs of the pack kernel that have finished
	uint32_t        seq;
	int             world, rank, on;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// reads of a received message: with the pull transport the message sits in the SENDER's memory and is read over NVLink; ld.cv
// never takes a line an earlier kernel left in this SM's L1 (the buffer is rewritten every other exchange)
__device__ __forceinline__ int4 msg_ld(const int4* p) { return __ldcv(p); }
__device__ __forceinline__ uint32_t msg_ld(const uint32_t* p) { return __ldcv(p); }

// tail of a pack kernel (every thread of every CTA gets here): the CTA's stores into the peers' buffers are ordered before thread
// 0 by the barrier and made visible at system scope by ITS fence (cumulativity: one MEMBAR.SYS per CTA); the CTA that finishes last
// raises the flags.  Measured (profiles/r02_variants.md, 2 GPUs, 49 000 ghosts each way): a pack kernel takes 19 us, 5 of them the
// fence, ~0 the counter, and the same whether a message is 0.2 or 0.8 MB; the unpack kernel with its wait 5.6 us.
__device__ __forceinline__ void mgl_signal(const mgl_sig& S)
{
	if (!S.on) return;
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();
		if (atomicAdd(S.done, 1u) == gridDim.x - 1u) {
			atomicExch(S.done, 0u);
			__threadfence_system();
			for (int r = 0; r < S.world; r++)
				if (r != S.rank) st_release_sys(S.remote_flag[r] + (S.seq & 1u), S.seq);
		}
	}
}

// head of the first kernel that reads an exchange's messages: until every peer's message `seq` is complete
__device__ __forceinline__ void mgl_wait(const mgl_sig& S)
{
	if (!S.on) return;
	for (int r = (int)threadIdx.x; r < S.world; r += (int)blockDim.x) {
		if (r == S.rank) continue;
		const uint32_t* f = S.local_flag + 2 * r + (S.seq & 1u);
		while ((int32_t)(ld_acquire_sys(f) - S.seq) < 0) { }
	}
	__syncthreads();
}

__global__ void k_mgl_begin(uint32_t* __restrict__ words, uint32_t* len, uint32_t* hidden_len, uint32_t* misc)
{
	const uint32_t n = words[MGL_N_OWNED];
	*len = n; *hidden_len = n;
	misc[MW_N_OWNED] = n; misc[MW_GID_BASE] = words[MGL_GID_BASE];
}

// segment r of the lists (grouped by destination) -> send buffer r; thread (r, k)
// (perm: the lists' stable order by destination; segment r of it = the particles that go to rank r)
__global__ void k_mgl_pack_route(state_lists L, const uint32_t* __restrict__ perm, uint32_t* __restrict__ words, mgl_bufs B, int world, int rank, uint32_t route_cap, mgl_sig S)
{
	const uint32_t total = (uint32_t)world * route_cap;
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
		const int r = (int)(t / route_cap);
		const uint32_t k = t - (uint32_t)r * route_cap;
		if (r == rank) continue;
		uint32_t start = 0u;
		for (int q = 0; q < r; q++) start += words[MGL_ROUTE + q];
		const uint32_t cnt = words[MGL_ROUTE + r];
		if (k == 0u) {
			B.p[r][0] = make_int4((int)min(cnt, route_cap), 0, 0, 0);
			if (cnt > route_cap) atomicOr(words + MGL_FLAGS, MGL_FLAG_ROUTE_OVERFLOW);
		}
		if (k >= cnt) continue;
		const uint32_t id = perm[start + k];
		int4* o = B.p[r] + 1 + STATE_INT4 * (size_t)k;
		o[0] = L.pos[id]; o[1] = L.vel[id]; o[2] = L.backup[id];
		o[3] = make_int4((int)L.inv_mass[id], (int)L.radius[id], (int)L.transferring[id], (int)L.target_radius[id]);
		o[4] = make_int4((int)L.kernel_width[id], (int)L.boundariness[id], (int)L.boundary_distance[id], 0);
	}
	mgl_signal(S);
}

// new order of this rank's lists: [arrivals from lower ranks, rank after rank | stayers | arrivals from higher ranks] -- the order
// the single-GPU sort sees them in (its sort is stable and the previous order was sorted by key, i.e. by rank first)
__global__ void k_mgl_route_plan(uint32_t* __restrict__ words, mgl_bufs R, int world, int rank, uint32_t capacity, mgl_sig S)
{
	mgl_wait(S);
	if (threadIdx.x != 0u) return;
	uint32_t low = 0u, high = 0u, start = 0u, total_out = 0u;
	for (int r = 0; r < world; r++) {
		const uint32_t a = r == rank ? 0u : (uint32_t)msg_ld(R.p[r]).x;
		words[MGL_ARRIVE + r] = a;
		if (r < rank) low += a; else if (r > rank) high += a;
		if (r < rank) start += words[MGL_ROUTE + r];
		if (r != rank) total_out += words[MGL_ROUTE + r];
	}
	const uint32_t stay = words[MGL_ROUTE + rank];
	uint32_t off_low = 0u, off_high = low + stay;
	for (int r = 0; r < world; r++) {
		if (r < rank) { words[MGL_ARRIVE_DST + r] = off_low; off_low += words[MGL_ARRIVE + r]; }
		else if (r > rank) { words[MGL_ARRIVE_DST + r] = off_high; off_high += words[MGL_ARRIVE + r]; }
	}
	words[MGL_STAY_SRC] = start; words[MGL_STAY_DST] = low; words[MGL_STAY] = stay;
	uint32_t n = low + stay + high;
	if (n > capacity) { atomicOr(words + MGL_FLAGS, MGL_FLAG_CAPACITY); n = capacity; }
	words[MGL_N_OWNED] = n;
	words[MGL_MIGRATED] = total_out;
}

__global__ void k_mgl_copy_stayers(state_lists S, state_lists D, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ words, uint32_t capacity)
{
	const uint32_t src0 = words[MGL_STAY_SRC], dst0 = words[MGL_STAY_DST], count = words[MGL_STAY];
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const uint32_t s = perm[src0 + k], d = dst0 + k; // (the stayers keep their order: perm is increasing over them)
		if (d >= capacity) continue;
		D.pos[d] = S.pos[s]; D.vel[d] = S.vel[s]; D.backup[d] = S.backup[s];
		D.inv_mass[d] = S.inv_mass[s]; D.radius[d] = S.radius[s]; D.transferring[d] = S.transferring[s];
		D.target_radius[d] = S.target_radius[s]; D.kernel_width[d] = S.kernel_width[s]; D.boundariness[d] = S.boundariness[s];
		D.boundary_distance[d] = S.boundary_distance[s];
		D.index_list[d] = d;
	}
}

__global__ void k_mgl_unpack_arrivals(state_lists D, const uint32_t* __restrict__ words, mgl_bufs R, int world, int rank, uint32_t route_cap, uint32_t capacity)
{
	const uint32_t total = (uint32_t)world * route_cap;
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
		const int r = (int)(t / route_cap);
		const uint32_t k = t - (uint32_t)r * route_cap;
		if (r == rank || k >= words[MGL_ARRIVE + r]) continue;
		const uint32_t id = words[MGL_ARRIVE_DST + r] + k;
		if (id >= capacity) continue;
		const int4* o = R.p[r] + 1 + STATE_INT4 * (size_t)k;
		D.pos[id] = msg_ld(o); D.vel[id] = msg_ld(o + 1); D.backup[id] = msg_ld(o + 2);
		const int4 a = msg_ld(o + 3), b = msg_ld(o + 4);
		D.inv_mass[id] = (uint32_t)a.x; D.radius[id] = (uint32_t)a.y; D.transferring[id] = (uint32_t)a.z; D.target_radius[id] = (uint32_t)a.w;
		D.kernel_width[id] = (uint32_t)b.x; D.boundariness[id] = (uint32_t)b.y; D.boundary_distance[id] = (uint32_t)b.z;
		D.index_list[id] = id;
	}
}

// send lists: owned ids whose cell lies inside another rank's grown brick (block r of `ids` at C.off[r], at most C.cap[r] entries)
__global__ void k_mgl_halo_lists(const int32_t* __restrict__ pos4, uint32_t* __restrict__ words, apbf_grid_params g, mg_boxes B, mgl_caps C,
                                 uint32_t* __restrict__ ids)
{
	const uint32_t n_owned = words[MGL_N_OWNED];
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t base = blockIdx.x * blockDim.x; base < n_owned; base += stride) { // whole warps stay in the loop for the ballots
		const uint32_t id = base + threadIdx.x;
		uint32_t c[3] = { 0u, 0u, 0u };
		if (id < n_owned) cell_of(pos4, id, g, c);
		for (int r = 0; r < B.world; r++) {
			if (r == B.rank) continue;
			const bool inside = id < n_owned && c[0] >= B.lo[r][0] && c[0] <= B.hi[r][0] && c[1] >= B.lo[r][1] && c[1] <= B.hi[r][1] &&
			                    c[2] >= B.lo[r][2] && c[2] <= B.hi[r][2];
			const uint32_t m = __ballot_sync(0xffffffffu, inside);
			if (m == 0u) continue;
			uint32_t slot = 0u;
			if (lane_id() == 0u) slot = atomicAdd(words + MGL_HALO_SEND + r, (uint32_t)__popc(m));
			slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(m & ((1u << lane_id()) - 1u));
			if (inside) {
				if (slot < C.cap[r]) ids[C.off[r] + slot] = id;
				else atomicOr(words + MGL_FLAGS, MGL_FLAG_HALO_OVERFLOW);
			}
		}
	}
}

__global__ void k_mgl_pack_halo(halo_lists L, uint32_t* __restrict__ words, const uint32_t* __restrict__ ids, mgl_caps C, mgl_bufs B, uint32_t total, mgl_sig S)
{
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
		int r = 0;
		while (r + 1 < C.world && t >= C.off[r + 1]) r++;
		if (r == C.rank) continue;
		const uint32_t k = t - C.off[r];
		if (k >= C.cap[r]) continue;
		const uint32_t cnt = min(words[MGL_HALO_SEND + r], C.cap[r]);
		if (k == 0u) {
			B.p[r][0] = make_int4((int)cnt, (int)words[MGL_N_OWNED], 0, 0);
			words[MGL_HALO_SEND + r] = cnt; // (clamped: the exchanges of the solver loop use it)
		}
		if (k >= cnt) continue;
		const uint32_t id = ids[C.off[r] + k];
		int4* o = B.p[r] + 1 + HALO_INT4 * (size_t)k;
		o[0] = L.pos[id];
		o[1] = make_int4((int)L.inv_mass[id], (int)L.radius[id], (int)L.kernel_width[id], (int)L.target_radius[id]);
		o[2] = make_int4((int)L.boundary_distance[id], 0, 0, 0); // (update_transfers floods it across the bricks)
	}
	mgl_signal(S);
}

__global__ void k_mgl_halo_plan(uint32_t* __restrict__ words, mgl_bufs R, mgl_caps C, uint32_t capacity, uint32_t* len, uint32_t* hidden_len, uint32_t* misc, mgl_sig S)
{
	mgl_wait(S);
	if (threadIdx.x != 0u) return;
	const uint32_t n_owned = words[MGL_N_OWNED];
	uint32_t first = n_owned, gid = 0u;
	for (int r = 0; r < C.world; r++) {
		uint32_t cnt = 0u;
		if (r != C.rank) {
			const int4 hdr = msg_ld(R.p[r]);
			cnt = min((uint32_t)hdr.x, C.cap[r]);
			if (first + cnt > capacity) { atomicOr(words + MGL_FLAGS, MGL_FLAG_CAPACITY); cnt = capacity - first; }
			if (r < C.rank) gid += (uint32_t)hdr.y;
		}
		words[MGL_HALO_RECV + r] = cnt;
		words[MGL_GHOST_FIRST + r] = first;
		first += cnt;
	}
	words[MGL_N_TOTAL] = first;
	words[MGL_GID_BASE] = gid;
	*len = first; *hidden_len = first;
	misc[MW_N_OWNED] = n_owned; misc[MW_GID_BASE] = gid;
}

__global__ void k_mgl_unpack_halo(halo_lists_out L, const uint32_t* __restrict__ words, mgl_bufs R, mgl_caps C, uint32_t* __restrict__ ghost_ids, uint32_t total)
{
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
		int r = 0;
		while (r + 1 < C.world && t >= C.off[r + 1]) r++;
		if (r == C.rank) continue;
		const uint32_t k = t - C.off[r];
		if (k >= words[MGL_HALO_RECV + r]) continue;
		const uint32_t id = words[MGL_GHOST_FIRST + r] + k;
		const int4* in = R.p[r] + 1 + HALO_INT4 * (size_t)k;
		L.pos[id] = msg_ld(in);
		const int4 a = msg_ld(in + 1);
		L.inv_mass[id] = (uint32_t)a.x; L.radius[id] = (uint32_t)a.y; L.kernel_width[id] = (uint32_t)a.z; L.target_radius[id] = (uint32_t)a.w;
		L.boundary_distance[id] = (uint32_t)msg_ld(in + 2).x;
		L.index_list[id] = id;
		ghost_ids[C.off[r] + k] = id;
	}
}

// slots before the search's sort -> ids after it, for the send lists (which = MGL_HALO_SEND) or the ghost slots (MGL_HALO_RECV)
// tiles: (send lists only) flags the tile of 32 consecutive ids each sent particle is in -- an owned particle can only have a ghost
// neighbour if it is some other rank's ghost itself (the halo is as wide on both sides of a face), so these are the BOUNDARY tiles
// of the sweeps and every other tile of owned particles works without the halo exchange that is in flight
__global__ void k_mgl_remap(uint32_t* __restrict__ ids, const uint32_t* __restrict__ inv, const uint32_t* __restrict__ words, int which, mgl_caps C, uint32_t total,
                            uint32_t* __restrict__ tiles)
{
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
		int r = 0;
		while (r + 1 < C.world && t >= C.off[r + 1]) r++;
		if (r == C.rank || t - C.off[r] >= words[which + r]) continue;
		if (tiles) tiles[inv[ids[t]] >> 5] = 1u;
		ids[t] = inv[ids[t]];
	}
}

// one quantity of the owners for their ghosts elsewhere: what = 1 kernel width, 2 packed solver position, 3 lambda, 4 position
__global__ void k_mgl_pack(int what, const uint32_t* __restrict__ src4, const int4* __restrict__ src16, uint32_t stride4, const uint32_t* __restrict__ ids,
                           const uint32_t* __restrict__ words, mgl_caps C, mgl_bufs B, uint32_t total, mgl_sig S)
{
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
		int r = 0;
		while (r + 1 < C.world && t >= C.off[r + 1]) r++;
		const uint32_t k = t - C.off[r];
		if (r == C.rank || k >= words[MGL_HALO_SEND + r]) continue;
		const uint32_t id = ids[t];
		if (what == 2 || what == 4) B.p[r][k] = src16[id];
		else ((uint32_t*)B.p[r])[k] = src4[(size_t)id * stride4];
	}
	mgl_signal(S);
}

__global__ void k_mgl_unpack(int what, uint32_t* __restrict__ dst4, int4* __restrict__ dst16, float4* __restrict__ L4, const float4* __restrict__ KG,
                             const int4* __restrict__ P4, int4* __restrict__ PL, const uint32_t* __restrict__ ghost_ids, const uint32_t* __restrict__ words, mgl_caps C, mgl_bufs R, uint32_t total, mgl_sig S)
{
	mgl_wait(S);
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
		int r = 0;
		while (r + 1 < C.world && t >= C.off[r + 1]) r++;
		const uint32_t k = t - C.off[r];
		if (r == C.rank || k >= words[MGL_HALO_RECV + r]) continue;
		const uint32_t id = ghost_ids[t];
		if (what == 2 || what == 4) dst16[id] = msg_ld(R.p[r] + k);
		else if (what == 1) dst4[id] = msg_ld((const uint32_t*)R.p[r] + k);
		else { // a ghost's record for the apply sweep: {lambda from its owner, h, gradient c0, gradient c1 from the local constants}
			const float4 kg = KG[id];
			const float lam = __uint_as_float(msg_ld((const uint32_t*)R.p[r] + k));
			L4[id] = make_float4(lam, kg.x, kg.y, kg.z);
			const int4 p = P4[id]; // (refreshed by the exchange before the density sweep)
			PL[id] = make_int4(p.x, p.y, p.z, __float_as_int(lam < 0.0f ? lam * R_POS : 0.0f));
		}
	}
}

mgl_bufs bufs_of(void* const p[8]) { mgl_bufs b; for (int r = 0; r < 8; r++) b.p[r] = (int4*)p[r]; return b; }

// A new exchange: its number, and (peer-to-peer transport) where the flags are.  With the NCCL transport `on` is 0 and the kernels
// neither signal nor wait.
mgl_sig mgl_begin_exchange(apbf_sim* sim)
{
	apbf_mg_loop& M = sim->mgl;
	mgl_sig S;
	memset(&S, 0, sizeof S);
	S.seq = ++M.seq;
	S.world = sim->mg.world; S.rank = sim->mg.rank;
	S.on = M.p2p ? 1 : 0;
	if (M.p2p) {
		for (int r = 0; r < 8; r++) S.remote_flag[r] = M.remote_flag[r];
		S.local_flag = (const uint32_t*)((const char*)M.arena + M.flag_off);
		S.done = M.done_counter;
	}
	return S;
}
// where the pack kernels of exchange S write / where its messages are read
// Peer to peer, two ways round.  PUSH: a message is stored into the RECEIVER's arena (slot [source][parity]) and read there.  PULL
// (APBF_MG_P2P_PULL=1): it is written into the SENDER's own arena (slot [destination][parity]; the pairs size their messages
// alike, so the same offsets serve) and the receiver's unpack kernel reads it over NVLink once its flag is up -- the pack kernel
// then has no remote store to drain but the flag.  The two-buffer argument at mgl_sig holds either way: a slot is rewritten two
// exchanges later, after this rank has waited for the reader's next flag, which the reader raised after it had read.
bool mgl_pull()
{
	static const bool pull = getenv("APBF_MG_P2P_PULL") != nullptr && atoi(getenv("APBF_MG_P2P_PULL")) != 0;
	return pull;
}
int4* mgl_local_slot(const apbf_sim* sim, int r, uint32_t parity)
{
	return (r != sim->mg.rank && r < sim->mg.world) ? (int4*)((char*)sim->mgl.arena + sim->mgl.recv_off[r][parity]) : nullptr;
}
mgl_bufs send_bufs(const apbf_sim* sim, const mgl_sig& S)
{
	mgl_bufs b;
	for (int r = 0; r < 8; r++)
		b.p[r] = !sim->mgl.p2p ? (int4*)sim->mgl.send_buf[r] : mgl_pull() ? mgl_local_slot(sim, r, S.seq & 1u) : (int4*)sim->mgl.remote_recv[r][S.seq & 1u];
	return b;
}
mgl_bufs recv_bufs(const apbf_sim* sim, const mgl_sig& S)
{
	mgl_bufs b;
	for (int r = 0; r < 8; r++)
		b.p[r] = !sim->mgl.p2p ? (int4*)sim->mgl.recv_buf[r] : mgl_pull() ? (int4*)sim->mgl.remote_recv[r][S.seq & 1u] : mgl_local_slot(sim, r, S.seq & 1u);
	return b;
}
mgl_caps caps_of(const apbf_sim* sim)
{
	mgl_caps c;
	for (int r = 0; r < 8; r++) { c.cap[r] = sim->mgl.halo_cap[r]; c.off[r] = sim->mgl.halo_off[r]; }
	c.world = sim->mg.world; c.rank = sim->mg.rank;
	return c;
}

// every peer gets `bytes_of(r)` bytes from send_buf[r] and delivers as many into recv_buf[r]: one grouped round on the context's stream.
// The group is always closed, whatever an individual call returned.
template <class F>
int mgl_exchange(apbf_sim* sim, F bytes_of)
{
	apbf_ctx* ctx = sim->ctx;
	if (sim->mgl.p2p) { // the pack kernel has stored the messages where they belong and raised the flags
		sim->mgl.exchanges++;
		return APBF_OK;
	}
	nccl_api* N = nccl();
	if (!N || !sim->nccl_comm) return apbf_fail(ctx, APBF_ERR_UNSUPPORTED, "apbf_sim_mg_comm_init has not been called", __FILE__, __LINE__);
	int first_err = N->GroupStart();
	if (first_err != 0) return apbf_fail(ctx, APBF_ERR_CUDA, N->GetErrorString ? N->GetErrorString(first_err) : "ncclGroupStart", __FILE__, __LINE__);
	for (int r = 0; r < sim->mg.world; r++) {
		if (r == sim->mg.rank) continue;
		const size_t words = (bytes_of(r) + 3) / 4;
		int e = N->Send(sim->mgl.send_buf[r], words, NCCL_INT32, r, sim->nccl_comm, ctx->stream);
		if (e != 0 && first_err == 0) first_err = e;
		e = N->Recv(sim->mgl.recv_buf[r], words, NCCL_INT32, r, sim->nccl_comm, ctx->stream);
		if (e != 0 && first_err == 0) first_err = e;
	}
	const int e = N->GroupEnd();
	if (e != 0 && first_err == 0) first_err = e;
	if (first_err != 0) return apbf_fail(ctx, APBF_ERR_CUDA, N->GetErrorString ? N->GetErrorString(first_err) : "nccl", __FILE__, __LINE__);
	sim->mgl.exchanges++;
	return APBF_OK;
}

// one quantity of the owners to their ghosts elsewhere, in two halves: send = pack kernel (stores into the peers' buffers and raises
// their flags; NCCL transport: into the send buffers), recv = [NCCL group] + unpack kernel (waits for the peers' flags).  Whatever is
// enqueued between the two runs while the messages travel.
int mgl_refresh_part(apbf_sim* sim, int what, bool send, bool recv, mgl_sig& S)
{
	apbf_ctx* ctx = sim->ctx;
	const mgl_caps C = caps_of(sim);
	const uint32_t total = sim->mgl.halo_total, cap = sim->cfg.particle_capacity;
	if (total == 0u) return APBF_OK;
	apbf_fluid& f = sim->fluid;
	const uint32_t* src4 = nullptr; const int4* src16 = nullptr; uint32_t stride4 = 1u;
	uint32_t* dst4 = nullptr; int4* dst16 = nullptr; float4* L4 = nullptr; const float4* KG = nullptr;
	const int4* P4 = nullptr; int4* PL = nullptr;
	if (what == 1) { src4 = (const uint32_t*)f.kernel_width.data; dst4 = (uint32_t*)f.kernel_width.data; }
	else if (what == 2) { src16 = dst16 = (int4*)ctx->scratch_get(SLOT_P4, sizeof(int4) * (size_t)cap); }
	else if (what == 3) {
		L4 = (float4*)ctx->scratch_get(SLOT_L4, sizeof(float4) * (size_t)cap);
		KG = (const float4*)ctx->scratch_get(SLOT_KG, sizeof(float4) * (size_t)cap);
		P4 = (const int4*)ctx->scratch_get(SLOT_P4, sizeof(int4) * (size_t)cap);
		PL = (int4*)ctx->scratch_get(SLOT_PL, sizeof(int4) * (size_t)cap);
		if (!P4 || !PL) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
		src4 = (const uint32_t*)L4; stride4 = 4u;
	} else if (what == 4) { src16 = dst16 = (int4*)f.particle.position.data; }
	else return apbf_fail(ctx, APBF_ERR_INVALID, "what", __FILE__, __LINE__);
	// (at most two CTAs per SM: every CTA of a pack kernel ends with an atomic on one counter, and a message is a few hundred KB)
	const unsigned grid = apbf_grid(ctx, total, 256, 2);
	apbf_prof_scope ps(ctx, PROF_MG_EXCHANGE);
	if (send) {
		S = mgl_begin_exchange(sim);
		k_mgl_pack<<<grid, 256, 0, ctx->stream>>>(what, src4, src16, stride4, sim->mgl.send_ids, sim->mgl.words, C, send_bufs(sim, S), total, S);
		APBF_LAUNCHED(ctx);
	}
	if (recv) {
		const size_t elem = (what == 2 || what == 4) ? 16u : 4u;
		APBF_TRY(mgl_exchange(sim, [&](int r) { return elem * (size_t)sim->mgl.halo_cap[r]; }));
		k_mgl_unpack<<<grid, 256, 0, ctx->stream>>>(what, dst4, dst16, L4, KG, P4, PL, sim->mgl.ghost_ids, sim->mgl.words, C, recv_bufs(sim, S), total, S);
		APBF_LAUNCHED(ctx);
	}
	return APBF_OK;
}

int mgl_refresh(apbf_sim* sim, int what)
{
	mgl_sig S;
	return mgl_refresh_part(sim, what, true, true, S);
}

mg_boxes grown_boxes(const apbf_sim* sim)
{
	const apbf_mg_state& m = sim->mg;
	const uint32_t cells = 1u << sim->cfg.res_log2;
	mg_boxes B;
	memset(&B, 0, sizeof B);
	B.world = m.world; B.rank = m.rank;
	for (int r = 0; r < m.world; r++)
		for (int d = 0; d < 3; d++) {
			B.lo[r][d] = m.lo[r][d] > m.halo[d] ? m.lo[r][d] - m.halo[d] : 0u;
			B.hi[r][d] = (d < sim->cfg.dims) ? (m.hi[r][d] + m.halo[d] < cells - 1u ? m.hi[r][d] + m.halo[d] : cells - 1u) : 0u;
		}
	return B;
}

} // namespace

extern "C" {

// Set-up helper (synchronises): how many of this rank's first n_owned particles every other rank needs as ghosts right now.  The
// caller exchanges these numbers between the ranks and derives the per-peer message capacities from them.
int apbf_sim_mg_halo_counts(apbf_sim* sim, uint32_t n_owned, uint32_t out_counts_host[8])
{
	if (!sim || !out_counts_host) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mg.enabled && n_owned <= sim->cfg.particle_capacity);
	apbf_grid_params g;
	APBF_TRY(grid_of(sim, &g));
	uint32_t* words = (uint32_t*)ctx->scratch_get(SLOT_MG_SEND, sizeof(uint32_t) * MGL_WORDS);
	uint32_t* ids = (uint32_t*)ctx->scratch_get(SLOT_MG_RECV, 256);
	if (!words || !ids) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
	APBF_CUDA(ctx, cudaMemsetAsync(words, 0, sizeof(uint32_t) * MGL_WORDS, ctx->stream));
	APBF_CUDA(ctx, cudaMemcpyAsync(words + MGL_N_OWNED, &n_owned, 4, cudaMemcpyHostToDevice, ctx->stream));
	mgl_caps C;
	memset(&C, 0, sizeof C); // capacity 0 everywhere: count only
	C.world = sim->mg.world; C.rank = sim->mg.rank;
	if (n_owned > 0) {
		k_mgl_halo_lists<<<apbf_grid(ctx, n_owned, 256), 256, 0, ctx->stream>>>((const int32_t*)sim->fluid.particle.position.data, words, g, grown_boxes(sim), C, ids);
		APBF_LAUNCHED(ctx);
	}
	APBF_CUDA(ctx, cudaMemcpyAsync(out_counts_host, words + MGL_HALO_SEND, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return APBF_OK;
}

// Allocates the staging buffers of the library loop.  route_cap: particles per peer and substep that may change owner; halo_cap[r]:
// ghost records exchanged with rank r -- the SAME number on both sides of a pair (it is the size of their messages).
int apbf_sim_mg_loop_init(apbf_sim* sim, uint32_t n_owned, uint32_t route_cap, const uint32_t halo_cap[8])
{
	if (!sim || !halo_cap) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mg.enabled && !sim->mgl.ready && n_owned <= sim->cfg.particle_capacity && route_cap > 0u);
	apbf_mg_loop& M = sim->mgl;
	M.route_cap = route_cap;
	uint32_t off = 0u;
	for (int r = 0; r < 8; r++) {
		M.halo_cap[r] = (r < sim->mg.world && r != sim->mg.rank) ? halo_cap[r] : 0u;
		M.halo_off[r] = off;
		off += M.halo_cap[r];
	}
	M.halo_total = off;
	auto dev_alloc = [&](void** p, size_t bytes) {
		if (cudaMalloc(p, bytes ? bytes : 16) != cudaSuccess) return false;
		sim->owned.push_back(*p);
		return cudaMemsetAsync(*p, 0, bytes ? bytes : 16, ctx->stream) == cudaSuccess;
	};
	bool ok = dev_alloc((void**)&M.words, sizeof(uint32_t) * MGL_WORDS) && dev_alloc((void**)&M.send_ids, sizeof(uint32_t) * (size_t)(off + 1)) &&
	          dev_alloc((void**)&M.ghost_ids, sizeof(uint32_t) * (size_t)(off + 1));
	for (int r = 0; r < sim->mg.world && ok; r++) {
		if (r == sim->mg.rank) continue;
		const size_t bytes = 16 + std::max((size_t)route_cap * STATE_INT4 * 16, (size_t)M.halo_cap[r] * HALO_INT4 * 16);
		M.buf_bytes[r] = bytes;
		ok = dev_alloc(&M.send_buf[r], bytes) && dev_alloc(&M.recv_buf[r], bytes);
	}
	if (!ok) return apbf_fail(ctx, APBF_ERR_OOM, "cudaMalloc", __FILE__, __LINE__);
	APBF_CUDA(ctx, cudaMemcpyAsync(M.words + MGL_N_OWNED, &n_owned, 4, cudaMemcpyHostToDevice, ctx->stream));
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	M.ready = true;
	return APBF_OK;
}

// ---- peer-to-peer transport set-up -----------------------------------------------------------------------------------------------
// apbf_sim_mg_p2p_export: allocates this rank's receive arena (two buffers per source rank + the flags) and describes it in a
// 512-byte blob {CUDA IPC handle, offsets}.  The caller gathers the blobs of all ranks (any transport: they are plain bytes) and
// hands the table to apbf_sim_mg_p2p_import, which opens the other ranks' arenas.  From then on apbf_sim_mg_substep exchanges
// without NCCL: pack kernels store into the receiver's buffer over NVLink and raise its flag, unpack kernels wait on their own.
struct mgl_blob {
	cudaIpcMemHandle_t handle;      // 64 bytes
	uint64_t recv_off[8][2];
	uint64_t flag_off;
	uint64_t buf_bytes[8];
	uint64_t arena_bytes;
	uint32_t rank, world;
};
static_assert(sizeof(mgl_blob) <= APBF_MG_P2P_BLOB_BYTES, "blob size is part of the C-ABI");

int apbf_sim_mg_p2p_export(apbf_sim* sim, void* out_blob)
{
	if (!sim || !out_blob) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mg.enabled && sim->mgl.ready && !sim->mgl.arena);
	apbf_mg_loop& M = sim->mgl;
	size_t off = 0;
	for (int r = 0; r < sim->mg.world; r++) {
		if (r == sim->mg.rank) continue;
		for (int p = 0; p < 2; p++) { M.recv_off[r][p] = off; off += (M.buf_bytes[r] + 255) & ~(size_t)255; }
	}
	M.flag_off = off;
	off += 256; // uint32 flags[8][2]
	M.arena_bytes = off;
	APBF_CUDA(ctx, cudaMalloc(&M.arena, M.arena_bytes));
	sim->owned.push_back(M.arena);
	APBF_CUDA(ctx, cudaMemsetAsync(M.arena, 0, M.arena_bytes, ctx->stream));
	APBF_CUDA(ctx, cudaMalloc((void**)&M.done_counter, 256));
	sim->owned.push_back(M.done_counter);
	APBF_CUDA(ctx, cudaMemsetAsync(M.done_counter, 0, 256, ctx->stream));
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	mgl_blob b;
	memset(&b, 0, sizeof b);
	APBF_CUDA(ctx, cudaIpcGetMemHandle(&b.handle, M.arena));
	for (int r = 0; r < 8; r++) { b.recv_off[r][0] = M.recv_off[r][0]; b.recv_off[r][1] = M.recv_off[r][1]; b.buf_bytes[r] = M.buf_bytes[r]; }
	b.flag_off = M.flag_off; b.arena_bytes = M.arena_bytes;
	b.rank = (uint32_t)sim->mg.rank; b.world = (uint32_t)sim->mg.world;
	memset(out_blob, 0, APBF_MG_P2P_BLOB_BYTES);
	memcpy(out_blob, &b, sizeof b);
	return APBF_OK;
}

// blobs: world x APBF_MG_P2P_BLOB_BYTES bytes, entry r = what rank r exported.  Call on every rank, after a barrier that follows the exports.
int apbf_sim_mg_p2p_import(apbf_sim* sim, const void* blobs)
{
	if (!sim || !blobs) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mg.enabled && sim->mgl.ready && sim->mgl.arena && !sim->mgl.p2p);
	apbf_mg_loop& M = sim->mgl;
	const int me = sim->mg.rank;
	for (int r = 0; r < sim->mg.world; r++) {
		if (r == me) continue;
		mgl_blob b;
		memcpy(&b, (const char*)blobs + (size_t)r * APBF_MG_P2P_BLOB_BYTES, sizeof b);
		APBF_REQUIRE(ctx, (int)b.rank == r && (int)b.world == sim->mg.world && b.buf_bytes[me] == M.buf_bytes[r]); // both ends size a pair's messages alike
		void* base = nullptr;
		const cudaError_t e = cudaIpcOpenMemHandle(&base, b.handle, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) {
			cudaGetLastError();
			for (int q = 0; q < r; q++) if (M.peer_base[q]) { cudaIpcCloseMemHandle(M.peer_base[q]); M.peer_base[q] = nullptr; }
			return apbf_fail(ctx, APBF_ERR_UNSUPPORTED, cudaGetErrorString(e), __FILE__, __LINE__); // (the NCCL transport stays in place)
		}
		M.peer_base[r] = base;
		M.remote_recv[r][0] = (char*)base + b.recv_off[me][0];
		M.remote_recv[r][1] = (char*)base + b.recv_off[me][1];
		M.remote_flag[r] = (uint32_t*)((char*)base + b.flag_off) + 2 * me;
	}
	M.seq = 0;
	M.p2p = true;
	return APBF_OK;
}

// 1 if the exchanges of apbf_sim_mg_substep go peer to peer, 0 if through NCCL
int apbf_sim_mg_p2p_active(const apbf_sim* sim) { return sim && sim->mgl.p2p ? 1 : 0; }

// the lists were uploaded afresh: this rank owns their first n_owned entries, global ids start at gid_base
int apbf_sim_mg_loop_reset(apbf_sim* sim, uint32_t n_owned, uint32_t gid_base)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mgl.ready && n_owned <= sim->cfg.particle_capacity);
	const uint32_t w[3] = { n_owned, n_owned, gid_base };
	APBF_CUDA(ctx, cudaMemcpyAsync(sim->mgl.words + MGL_N_OWNED, w, sizeof w, cudaMemcpyHostToDevice, ctx->stream));
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // (w lives on this stack frame)
	return APBF_OK;
}

// out[0] owned particles, [1] owned + ghosts, [2] global id of local id 0, [3] particles that left in the last substep,
// [4] flags (1 migration buffer overflow, 2 ghost list overflow, 4 particle capacity exceeded), [5] exchanges so far; synchronises
int apbf_sim_mg_loop_stats(apbf_sim* sim, uint32_t out[8])
{
	if (!sim || !out) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mgl.ready);
	uint32_t w[8];
	APBF_CUDA(ctx, cudaMemcpyAsync(w, sim->mgl.words, sizeof w, cudaMemcpyDeviceToHost, ctx->stream));
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	out[0] = w[MGL_N_OWNED]; out[1] = w[MGL_N_TOTAL]; out[2] = w[MGL_GID_BASE]; out[3] = w[MGL_MIGRATED]; out[4] = w[MGL_FLAGS];
	out[5] = (uint32_t)sim->mgl.exchanges; out[6] = 0u; out[7] = 0u;
	return APBF_OK;
}

int apbf_sim_mg_substep(apbf_sim* sim, uint32_t n_substeps)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->mg.enabled && sim->mgl.ready);
	apbf_mg_loop& M = sim->mgl;
	const apbf_sim_config& c = sim->cfg;
	const apbf_settings& s = ctx->settings;
	const int world = sim->mg.world, rank = sim->mg.rank;
	const uint32_t cap = c.particle_capacity;
	cudaStream_t st = ctx->stream;
	apbf_grid_params g;
	APBF_TRY(grid_of(sim, &g));
	const mgl_caps C = caps_of(sim);
	const bool adaptive = !c.basic_pbf && !s.mBaseKernelWidthOnBoundaryDistance;
	const bool default_mode = c.update_transfers && !c.basic_pbf && s.mBaseKernelWidthOnBoundaryDistance;
	uint32_t* misc = ctx->misc();
	for (uint32_t step = 0; step < n_substeps; step++) {
		apbf_fluid& f = sim->fluid;
		// ghosts of the last substep are dropped: the lists are the owned particles again
		k_mgl_begin<<<1, 1, 0, st>>>(M.words, f.particle.length, f.particle.hidden_length, misc);
		APBF_LAUNCHED(ctx);
		if (c.integrate) APBF_TRY(apbf_sim_mg_phase(sim, 0, 0));
		if (default_mode) APBF_TRY(apbf_sim_mg_phase(sim, 8, 0));
		if (world > 1) {
			// ---- ROUTE: particles that left the brick change owner with their full state -------------------------------------------
			{
			apbf_prof_scope ps_route(ctx, PROF_MG_ROUTE);
			APBF_TRY(route_plan_and_move(sim, M.words + MGL_ROUTE, false)); // perm = stable order by destination; counts on the device
			const uint32_t* perm = (const uint32_t*)ctx->scratch_get(SLOT_TMP_VALS, sizeof(uint32_t) * (size_t)cap);
			const mgl_sig SR = mgl_begin_exchange(sim);
			k_mgl_pack_route<<<apbf_grid(ctx, (size_t)world * M.route_cap, 256, 2), 256, 0, st>>>(lists_of(sim, false), perm, M.words, send_bufs(sim, SR), world, rank, M.route_cap, SR);
			APBF_LAUNCHED(ctx);
			APBF_TRY(mgl_exchange(sim, [&](int) { return 16 + (size_t)M.route_cap * STATE_INT4 * 16; }));
			k_mgl_route_plan<<<1, 32, 0, st>>>(M.words, recv_bufs(sim, SR), world, rank, cap, SR);
			APBF_LAUNCHED(ctx);
			k_mgl_copy_stayers<<<apbf_grid(ctx, cap, 256), 256, 0, st>>>(lists_of(sim, false), lists_of(sim, true), perm, M.words, cap);
			APBF_LAUNCHED(ctx);
			k_mgl_unpack_arrivals<<<apbf_grid(ctx, (size_t)world * M.route_cap, 256), 256, 0, st>>>(lists_of(sim, true), M.words, recv_bufs(sim, SR), world, rank, M.route_cap, cap);
			APBF_LAUNCHED(ctx);
			apbf_sim_swap_buffers(sim);
			k_mgl_begin<<<1, 1, 0, st>>>(M.words, f.particle.length, f.particle.hidden_length, misc);
			APBF_LAUNCHED(ctx);
			}
			// ---- HALO: owned particles inside another rank's grown brick go there as ghosts ----------------------------------------
			apbf_prof_scope ps_halo(ctx, PROF_MG_HALO);
			APBF_CUDA(ctx, cudaMemsetAsync(M.words + MGL_HALO_SEND, 0, sizeof(uint32_t) * 8, st));
			k_mgl_halo_lists<<<apbf_grid(ctx, cap, 256), 256, 0, st>>>((const int32_t*)f.particle.position.data, M.words, g, grown_boxes(sim), C, M.send_ids);
			APBF_LAUNCHED(ctx);
			const halo_lists hl{ (const int4*)f.particle.position.data, (const uint32_t*)f.particle.inverse_mass.data, (const uint32_t*)f.particle.radius.data,
			                     (const uint32_t*)f.kernel_width.data, (const uint32_t*)f.target_radius.data, (const uint32_t*)f.boundary_distance.data };
			const mgl_sig SH = mgl_begin_exchange(sim);
			k_mgl_pack_halo<<<apbf_grid(ctx, M.halo_total, 256, 2), 256, 0, st>>>(hl, M.words, M.send_ids, C, send_bufs(sim, SH), M.halo_total, SH);
			APBF_LAUNCHED(ctx);
			APBF_TRY(mgl_exchange(sim, [&](int r) { return 16 + (size_t)M.halo_cap[r] * HALO_INT4 * 16; }));
			k_mgl_halo_plan<<<1, 32, 0, st>>>(M.words, recv_bufs(sim, SH), C, cap, f.particle.length, f.particle.hidden_length, misc, SH);
			APBF_LAUNCHED(ctx);
			const halo_lists_out ho{ (int4*)f.particle.position.data, (uint32_t*)f.particle.inverse_mass.data, (uint32_t*)f.particle.radius.data,
			                         (uint32_t*)f.kernel_width.data, (uint32_t*)f.target_radius.data, (uint32_t*)f.boundary_distance.data,
			                         (uint32_t*)f.particle.index_list.data };
			k_mgl_unpack_halo<<<apbf_grid(ctx, M.halo_total, 256), 256, 0, st>>>(ho, M.words, recv_bufs(sim, SH), C, M.ghost_ids, M.halo_total);
			APBF_LAUNCHED(ctx);
		}
		APBF_TRY(apbf_sim_mg_phase(sim, 1, 0)); // search over owned + ghosts (+ the fused spread_kernel_width); leaves old slot -> new id
		if (world > 1 && M.halo_total > 0u) {
			const uint32_t* inv = (const uint32_t*)ctx->scratch_get(SLOT_MG_INV, sizeof(uint32_t) * (size_t)cap);
			if (!inv) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
			uint32_t* tiles = (uint32_t*)ctx->scratch_get(SLOT_MG_TILES, sizeof(uint32_t) * ((size_t)cap / 32u + 2u));
			if (!tiles) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
			APBF_CUDA(ctx, cudaMemsetAsync(tiles, 0, sizeof(uint32_t) * ((size_t)cap / 32u + 2u), st));
			k_mgl_remap<<<apbf_grid(ctx, M.halo_total, 256), 256, 0, st>>>(M.send_ids, inv, M.words, MGL_HALO_SEND, C, M.halo_total, tiles);
			APBF_LAUNCHED(ctx);
			k_mgl_remap<<<apbf_grid(ctx, M.halo_total, 256), 256, 0, st>>>(M.ghost_ids, inv, M.words, MGL_HALO_RECV, C, M.halo_total, nullptr);
			APBF_LAUNCHED(ctx);
		}
		if (adaptive) {
			APBF_TRY(apbf_sim_mg_phase(sim, 2, 0));
			if (world > 1) APBF_TRY(mgl_refresh(sim, 1)); // the owners' new widths overwrite the ghosts'
		}
		APBF_TRY(apbf_sim_mg_phase(sim, 3, 0));
		// Every iteration needs two exchanges -- the owners' packed positions before the density sweep, their lambdas before the apply
		// sweep -- and each costs a round trip between the GPUs.  Both hide behind work that does not need them: the sweeps run over the
		// INTERIOR tiles (no particle of theirs can have a ghost neighbour, k_mgl_remap) while the messages travel and over the BOUNDARY
		// tiles after the unpack kernel has seen the peers' flags.  OFF by default (APBF_MG_OVERLAP=1 turns it on): at 10^6 particles per
		// GPU the two extra launches per sweep cost more (+27 us each) than the wait they hide -- an exchange is 19 us of pack kernel and
		// 5.6 us of unpack kernel, its wait included, so there is hardly any wait to hide (2.43 ms without, 2.61 ms with, 2 GPUs).
		static const bool overlap = getenv("APBF_MG_OVERLAP") != nullptr && atoi(getenv("APBF_MG_OVERLAP")) != 0;
		const float* bmin = sim->boxes;
		const float* bmax = sim->boxes ? sim->boxes + 4 * (size_t)c.n_boxes : nullptr;
		for (int it = 0; it < c.solver_iterations; it++) {
			const bool last = it + 1 == c.solver_iterations;
			APBF_TRY(apbf_sim_mg_phase(sim, 4, it));
			if (world > 1 && M.halo_total > 0u && overlap) {
				const int t2 = ITER_RUN_T2 | ITER_T2_COMMIT | (last ? 0 : ITER_T2_NEXT_BOX);
				mgl_sig S;
				APBF_TRY(mgl_refresh_part(sim, 2, true, false, S));
				APBF_TRY(apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_T1 | ITER_TILES_INTERIOR, bmin, bmax, c.n_boxes, nullptr, nullptr));
				APBF_TRY(mgl_refresh_part(sim, 2, false, true, S));
				APBF_TRY(apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_T1 | ITER_TILES_BOUNDARY, bmin, bmax, c.n_boxes, nullptr, nullptr));
				APBF_TRY(mgl_refresh_part(sim, 3, true, false, S));
				APBF_TRY(apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, t2 | ITER_TILES_INTERIOR, bmin, bmax, c.n_boxes, nullptr, nullptr));
				APBF_TRY(mgl_refresh_part(sim, 3, false, true, S));
				APBF_TRY(apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, t2 | ITER_TILES_BOUNDARY, bmin, bmax, c.n_boxes, nullptr, nullptr));
				sim->mg_t2_tail_pending = true;
				continue;
			}
			if (world > 1) APBF_TRY(mgl_refresh(sim, 2));
			APBF_TRY(apbf_sim_mg_phase(sim, 5, it));
			if (world > 1) APBF_TRY(mgl_refresh(sim, 3));
			APBF_TRY(apbf_sim_mg_phase(sim, last ? 11 : 10, it));
		}
		APBF_TRY(apbf_sim_mg_phase(sim, 7, 0));
		if (c.update_transfers && !c.basic_pbf) {
			if (world > 1) APBF_TRY(mgl_refresh(sim, 4)); // distances are taken after the solver
			APBF_TRY(apbf_sim_mg_phase(sim, 9, 0));
		}
		// the lists a caller sees (apbf_sim_download) are this rank's own particles: the ghosts sit behind them and are dropped
		k_mgl_begin<<<1, 1, 0, st>>>(M.words, sim->fluid.particle.length, sim->fluid.particle.hidden_length, misc);
		APBF_LAUNCHED(ctx);
	}
	return APBF_OK;
}

} // extern "C"
