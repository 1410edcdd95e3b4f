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
__global__ void k_unpack_lambda(float4* __restrict__ L4, const float4* __restrict__ KG, const uint32_t* __restrict__ ids, uint32_t count,
                                const float* __restrict__ in)
{
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const uint32_t id = ids[k];
		const float4 kg = KG[id];
		L4[id] = make_float4(in[k], kg.x, kg.y, kg.z);
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
	APBF_REQUIRE(ctx, rank >= 0 && rank < world && halo_range >= 0.0f && !sim->cfg.use_binary_search);
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
int apbf_sim_mg_route(apbf_sim* sim, uint32_t* counts_dev)
{
	if (!sim || !counts_dev) return APBF_ERR_INVALID;
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
		                                      (const float4*)ctx->scratch_get(SLOT_KG, sizeof(float4) * (size_t)cap), ids_dev, count, (const float*)in);
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
			if (sim->mg_fused) // search + spread_kernel_width in one pass (pool.cpp:83-89), ghosts included
				APBF_TRY(apbf_green_search(ctx, &sim->fluid, &sim->fluid.kernel_width, &sim->nb, 1.5f, c.min_pos, c.max_pos, c.res_log2, nullptr, true, nullptr, false));
			else
				APBF_TRY(apbf_green_search(ctx, &sim->fluid, &sim->fluid.kernel_width, &sim->nb, unit_scale ? 1.0f : 1.5f, c.min_pos, c.max_pos, c.res_log2, nullptr, false, nullptr, false));
			apbf_sim_swap_buffers(sim);
			// old slot -> new id, for the send lists and the ghost slots (the search's sorted_index is still in scratch)
			const uint32_t cap = c.particle_capacity;
			uint32_t* inv = (uint32_t*)ctx->scratch_get(SLOT_MG_INV, sizeof(uint32_t) * (size_t)cap);
			const uint32_t* sidx = (const uint32_t*)ctx->scratch_get(SLOT_TMP_VALS, sizeof(uint32_t) * (size_t)cap);
			if (!inv || !sidx) return apbf_fail(ctx, APBF_ERR_OOM, "scratch", __FILE__, __LINE__);
			k_inverse_perm<<<apbf_grid(ctx, cap, 256), 256, 0, ctx->stream>>>(sidx, sim->fluid.particle.hidden_length, inv);
			APBF_LAUNCHED(ctx);
			return APBF_OK;
		}
		case 2: return sim->mg_fused ? APBF_OK : apbf_spread_kernel_width_apply(ctx, &sim->fluid, &sim->nb, nullptr);
		case 3: return apbf_solver_prepare(ctx, &sim->fluid);
		case 4: return apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_BEGIN | ITER_BEGIN_BOX | (iteration > 0 ? ITER_BEGIN_COMMIT : 0), bmin, bmax, c.n_boxes, nullptr, nullptr);
		case 5: return apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_T1, bmin, bmax, c.n_boxes, nullptr, nullptr);
		case 6: return apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_T2, bmin, bmax, c.n_boxes, nullptr, nullptr);
		case 7: return apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, ITER_RUN_BEGIN | ITER_BEGIN_COMMIT, nullptr, nullptr, 0u, nullptr, nullptr);
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
	for (int r = 0; r < sim->mg.world; r++) {
		if (r == sim->mg.rank) continue;
		if (L.send_counts[r]) APBF_NCCL(ctx, N->Send(stage_s + so * words, (size_t)L.send_counts[r] * words, NCCL_INT32, r, sim->nccl_comm, ctx->stream));
		if (L.ghost_counts[r]) APBF_NCCL(ctx, N->Recv(stage_r + ro * words, (size_t)L.ghost_counts[r] * words, NCCL_INT32, r, sim->nccl_comm, ctx->stream));
		so += L.send_counts[r]; ro += L.ghost_counts[r];
	}
	APBF_NCCL(ctx, N->GroupEnd());
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
		APBF_TRY(apbf_sim_mg_phase(sim, 6, it));
	}
	return apbf_sim_mg_phase(sim, 7, 0);
}

} // extern "C"
