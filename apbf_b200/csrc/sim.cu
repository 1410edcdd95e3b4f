// sim.cu -- a whole scene resident in HBM and one substep in the reference's order (source/pool.cpp:67-106):
//   [velocity_handling] -> neighbour search -> [spread_kernel_width] -> solverIterations x (box_collision, incompressibility)
// plus host <-> device transfer of the lists (the reference-facing call with host buffers).
#include "common.cuh"
#include "neighbors.cuh"
#include "solver.cuh"

#include "sim.cuh"
#include <stdlib.h>

namespace {

int alloc_array(apbf_sim* sim, apbf_array* a, size_t bytes)
{
	for (void** p : { &a->data, &a->reorder_out }) {
		if (cudaMalloc(p, bytes ? bytes : 16) != cudaSuccess) return apbf_fail(sim->ctx, APBF_ERR_OOM, "cudaMalloc", __FILE__, __LINE__);
		cudaMemsetAsync(*p, 0, bytes ? bytes : 16, sim->ctx->stream);
		sim->owned.push_back(*p);
	}
	return APBF_OK;
}

void swap_array(apbf_array* a)
{
	void* t = a->data; a->data = a->reorder_out; a->reorder_out = t;
}

__global__ void k_set_lengths(uint32_t* a, uint32_t* b, uint32_t n) { *a = n; *b = n; }

} // namespace

void apbf_sim_swap_buffers(apbf_sim* sim)
{
	apbf_fluid& f = sim->fluid;
	apbf_array* all[] = { &f.particle.index_list, &f.particle.position, &f.particle.velocity, &f.particle.inverse_mass,
	                      &f.particle.radius, &f.particle.pos_backup, &f.particle.transferring, &f.target_radius,
	                      &f.kernel_width, &f.boundariness, &f.boundary_distance };
	for (apbf_array* a : all) swap_array(a);
}

extern "C" {

int apbf_sim_create(apbf_ctx* ctx, const apbf_sim_config* cfg, apbf_sim** out_sim)
{
	if (!ctx || !cfg || !out_sim) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, cfg->dims == 2 || cfg->dims == 3);
	APBF_REQUIRE(ctx, cfg->n_boxes <= 64 && (cfg->n_boxes == 0 || (cfg->box_min4_host && cfg->box_max4_host)));
	*out_sim = nullptr;
	apbf_sim* sim = new apbf_sim();
	sim->ctx = ctx;
	sim->cfg = *cfg;
	sim->last_dt = 1.0f;
	sim->no_fuse = getenv("APBF_NO_FUSE") != nullptr; // debugging aid: search and spread_kernel_width as two operators
	if (const char* g = getenv("APBF_SIM_GRAPHS")) sim->graphs_on = atoi(g) != 0; // 0: enqueue every substep launch by launch
	sim->boxes = nullptr;
	memset(&sim->fluid, 0, sizeof sim->fluid);
	memset(&sim->nb, 0, sizeof sim->nb);
	const size_t n = cfg->particle_capacity;
	apbf_particles& p = sim->fluid.particle;
	p.capacity = p.hidden_capacity = cfg->particle_capacity;
	int rc = APBF_OK;
	rc = rc ? rc : alloc_array(sim, &p.index_list, 4 * n);
	rc = rc ? rc : alloc_array(sim, &p.position, 16 * n);
	rc = rc ? rc : alloc_array(sim, &p.velocity, 16 * n);
	rc = rc ? rc : alloc_array(sim, &p.inverse_mass, 4 * n);
	rc = rc ? rc : alloc_array(sim, &p.radius, 4 * n);
	rc = rc ? rc : alloc_array(sim, &p.pos_backup, 16 * n);
	rc = rc ? rc : alloc_array(sim, &p.transferring, 4 * n);
	rc = rc ? rc : alloc_array(sim, &sim->fluid.target_radius, 4 * n);
	rc = rc ? rc : alloc_array(sim, &sim->fluid.kernel_width, 4 * n);
	rc = rc ? rc : alloc_array(sim, &sim->fluid.boundariness, 4 * n);
	rc = rc ? rc : alloc_array(sim, &sim->fluid.boundary_distance, 4 * n);
	void* words = nullptr;
	if (!rc && cudaMalloc(&words, 64) != cudaSuccess) rc = APBF_ERR_OOM;
	if (!rc) {
		sim->owned.push_back(words);
		cudaMemsetAsync(words, 0, 64, ctx->stream);
		p.length = (uint32_t*)words;
		p.hidden_length = (uint32_t*)words + 1;
		sim->nb.length = (uint32_t*)words + 2;
		sim->nb.capacity = cfg->neighbor_capacity;
		void* pairs = nullptr;
		if (cudaMalloc(&pairs, 8 * (size_t)(cfg->neighbor_capacity ? cfg->neighbor_capacity : 1)) != cudaSuccess) rc = APBF_ERR_OOM;
		else { sim->owned.push_back(pairs); sim->nb.pairs = (uint32_t*)pairs; }
	}
	if (!rc && cfg->n_boxes) {
		void* b = nullptr;
		if (cudaMalloc(&b, 32 * (size_t)cfg->n_boxes) != cudaSuccess) rc = APBF_ERR_OOM;
		else {
			sim->owned.push_back(b);
			sim->boxes = (float*)b;
			cudaMemcpyAsync(sim->boxes, cfg->box_min4_host, 16 * (size_t)cfg->n_boxes, cudaMemcpyHostToDevice, ctx->stream);
			cudaMemcpyAsync(sim->boxes + 4 * (size_t)cfg->n_boxes, cfg->box_max4_host, 16 * (size_t)cfg->n_boxes, cudaMemcpyHostToDevice, ctx->stream);
			cudaStreamSynchronize(ctx->stream); // the host box arrays need not outlive this call
		}
	}
	if (!rc && cfg->transfers) {
		const size_t tc = cfg->transfer_capacity ? cfg->transfer_capacity : cfg->particle_capacity;
		sim->cfg.transfer_capacity = (uint32_t)tc;
		sim->tr.capacity = (uint32_t)tc;
		sim->tr.length = (uint32_t*)words + 3;
		rc = rc ? rc : alloc_array(sim, &sim->tr.source, 4 * tc);
		rc = rc ? rc : alloc_array(sim, &sim->tr.target, 4 * tc);
		rc = rc ? rc : alloc_array(sim, &sim->tr.time_left, 4 * tc);
		void* si = nullptr;
		if (!rc && cudaMalloc(&si, 4 * (n ? n : 4)) != cudaSuccess) rc = APBF_ERR_OOM;
		if (!rc) { sim->owned.push_back(si); sim->sorted_index = (uint32_t*)si; }
	}
	sim->cfg.box_min4_host = sim->cfg.box_max4_host = nullptr;
	if (rc) { apbf_sim_destroy(sim); return apbf_fail(ctx, rc, "sim allocation failed", __FILE__, __LINE__); }
	*out_sim = sim;
	return APBF_OK;
}

void apbf_sim_destroy(apbf_sim* sim)
{
	if (!sim) return;
	cudaStreamSynchronize(sim->ctx->stream);
	apbf_sim_mg_comm_destroy(sim); // the library's own NCCL communicator, if one was made
	for (apbf_sim_graph& g : sim->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
	if (sim->capture_stream) cudaStreamDestroy(sim->capture_stream);
	if (sim->copy_stream) cudaStreamDestroy(sim->copy_stream);
	for (cudaEvent_t e : { sim->ev_fork, sim->ev_up, sim->ev_lists, sim->ev_down }) if (e) cudaEventDestroy(e);
	for (void* p : sim->owned) cudaFree(p);
	apbf_nbr_forget(sim->ctx, sim->nb.pairs);
	delete sim;
}

int apbf_sim_upload(apbf_sim* sim, const apbf_host_state* h)
{
	if (!sim || !h) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, h->n <= sim->cfg.particle_capacity);
	cudaStream_t st = ctx->stream;
	apbf_fluid& f = sim->fluid;
	const size_t n = h->n;
	struct { const void* src; void* dst; size_t stride; } items[] = {
		{ h->position, f.particle.position.data, 16 }, { h->velocity, f.particle.velocity.data, 16 },
		{ h->inverse_mass, f.particle.inverse_mass.data, 4 }, { h->radius, f.particle.radius.data, 4 },
		{ h->pos_backup, f.particle.pos_backup.data, 16 }, { h->transferring, f.particle.transferring.data, 4 },
		{ h->target_radius, f.target_radius.data, 4 }, { h->kernel_width, f.kernel_width.data, 4 },
		{ h->boundariness, f.boundariness.data, 4 }, { h->boundary_distance, f.boundary_distance.data, 4 },
		{ h->index_list, f.particle.index_list.data, 4 },
	};
	for (auto& it : items)
		if (it.src && n) APBF_CUDA(ctx, cudaMemcpyAsync(it.dst, it.src, it.stride * n, cudaMemcpyHostToDevice, st));
	k_set_lengths<<<1, 1, 0, st>>>(f.particle.length, f.particle.hidden_length, (uint32_t)n);
	APBF_LAUNCHED(ctx);
	if (sim->tr.length) APBF_CUDA(ctx, cudaMemsetAsync(sim->tr.length, 0, 4, st)); // a fresh scene has no transfers under way
	if (!h->index_list) APBF_TRY(apbf_write_sequence(ctx, (uint32_t*)f.particle.index_list.data, f.particle.length, sim->cfg.particle_capacity, 0u, 1u, 1u));
	return APBF_OK;
}

int apbf_sim_download(apbf_sim* sim, apbf_host_state* h)
{
	if (!sim || !h) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	cudaStream_t st = ctx->stream;
	apbf_fluid& f = sim->fluid;
	uint32_t n = 0;
	APBF_CUDA(ctx, cudaMemcpyAsync(&n, f.particle.length, 4, cudaMemcpyDeviceToHost, st));
	APBF_CUDA(ctx, cudaStreamSynchronize(st));
	h->n = n;
	struct { void* dst; const void* src; size_t stride; } items[] = {
		{ h->position, f.particle.position.data, 16 }, { h->velocity, f.particle.velocity.data, 16 },
		{ h->inverse_mass, f.particle.inverse_mass.data, 4 }, { h->radius, f.particle.radius.data, 4 },
		{ h->pos_backup, f.particle.pos_backup.data, 16 }, { h->transferring, f.particle.transferring.data, 4 },
		{ h->target_radius, f.target_radius.data, 4 }, { h->kernel_width, f.kernel_width.data, 4 },
		{ h->boundariness, f.boundariness.data, 4 }, { h->boundary_distance, f.boundary_distance.data, 4 },
		{ h->index_list, f.particle.index_list.data, 4 },
	};
	for (auto& it : items)
		if (it.dst && n) APBF_CUDA(ctx, cudaMemcpyAsync(it.dst, it.src, it.stride * (size_t)n, cudaMemcpyDeviceToHost, st));
	APBF_CUDA(ctx, cudaStreamSynchronize(st));
	return APBF_OK;
}

} // extern "C"

namespace {

// ---- one substep, enqueued on the context's stream ----------------------------------------------------------------------------------
int substep_once(apbf_sim* sim)
{
	apbf_ctx* ctx = sim->ctx;
	const apbf_sim_config& c = sim->cfg;
	const apbf_settings& s = ctx->settings;
	const bool unit_scale = c.basic_pbf || s.mBaseKernelWidthOnBoundaryDistance;     // pool.cpp:83
	const bool adaptive = !c.basic_pbf && !s.mBaseKernelWidthOnBoundaryDistance;     // pool.cpp:87
	const bool transfers = c.transfers && !c.basic_pbf && (s.mMerge || s.mSplit);    // pool.cpp:73, :99
	apbf_search_debug follow;
	memset(&follow, 0, sizeof follow);
	follow.sorted_index = sim->sorted_index;
	const apbf_search_debug* dbg = transfers ? &follow : nullptr;
	{
		if (c.integrate) {                                                            // pool.cpp:71
			APBF_TRY(apbf_velocity_handling_apply(ctx, &sim->fluid.particle, c.dt, sim->last_dt, c.accel));
			sim->last_dt = c.dt;
		}
		if (transfers) {                                                              // pool.cpp:73-75
			APBF_TRY(apbf_particle_transfer_apply(ctx, &sim->fluid, &sim->tr, c.dt, nullptr));
			apbf_sim_swap_buffers(sim);
			for (apbf_array* a : { &sim->tr.source, &sim->tr.target, &sim->tr.time_left }) swap_array(a);
		}
		if (c.update_transfers && !c.basic_pbf && s.mBaseKernelWidthOnBoundaryDistance)  // pool.cpp:77-80
			APBF_TRY(apbf_kernel_width_from_boundary_distance(ctx, &sim->fluid));
		const float scale = unit_scale ? 1.0f : 1.5f;
		// pool.cpp:83-89.  Green search followed by spread_kernel_width runs as one fused pass (same lists, pair for pair)
		const bool fused = adaptive && !ctx->mg_enabled && !sim->no_fuse;
		// Nothing reads the public (id, idN) list between the search and the sweeps, which work on NB + offsets: the search skips
		// its 8-byte stores (write_public = false) and apbf_sim_neighbors() writes the list on demand.
		if (c.use_binary_search)                                                      // pool.cpp:83-84
			APBF_TRY(apbf_binary_search(ctx, &sim->fluid, &sim->fluid.kernel_width, &sim->nb, scale, dbg, fused, nullptr, false));
		else
			APBF_TRY(apbf_green_search(ctx, &sim->fluid, &sim->fluid.kernel_width, &sim->nb, scale, c.min_pos, c.max_pos, c.res_log2, dbg, fused, nullptr, false));
		apbf_sim_swap_buffers(sim);
		if (transfers) // the transfers' source / target lists share the hidden particle data (pool.cpp:18-19)
			APBF_TRY(apbf_transfers_follow_reorder(ctx, &sim->tr, sim->sorted_index, sim->fluid.particle.hidden_length, c.particle_capacity));
		if (adaptive && !fused) APBF_TRY(apbf_spread_kernel_width_apply(ctx, &sim->fluid, &sim->nb, nullptr)); // pool.cpp:87-89
		// pool.cpp:92-95: solverIterations x (box_collision, incompressibility).  Same results as calling the two
		// operators in turn; the per-particle constants are computed once (kernel widths are fixed from here on) and
		// box collision + the previous iteration's position update ride in the next iteration's prologue.
		if (c.solver_iterations > 0) APBF_TRY(apbf_solver_prepare(ctx, &sim->fluid));
		for (int it = 0; it < c.solver_iterations; it++) {
			const bool last = it == c.solver_iterations - 1;
			const int flags = ITER_BEGIN_BOX | (it > 0 ? ITER_BEGIN_COMMIT | ITER_SKIP_IF_T2_DID : 0) | (last ? ITER_END_COMMIT : ITER_T2_NEXT_BOX) | ITER_T2_COMMIT;
			APBF_TRY(apbf_solver_iteration(ctx, &sim->fluid, &sim->nb, flags, sim->boxes,
			                               sim->boxes ? sim->boxes + 4 * (size_t)c.n_boxes : nullptr, c.n_boxes, nullptr, nullptr));
		}
		if (transfers)                                                                // pool.cpp:99-102
			APBF_TRY(apbf_update_transfers_split_merge_apply(ctx, &sim->fluid, &sim->nb, &sim->tr, c.split_duration, nullptr));
		else if (c.update_transfers && !c.basic_pbf)
			APBF_TRY(apbf_update_transfers_apply(ctx, &sim->fluid, &sim->nb, nullptr));
	}
	return APBF_OK;
}

// ---- substeps as CUDA graphs ----------------------------------------------------------------------------------------------------------
// Every launch of a substep is length-agnostic (lengths are device words, grids come from capacities), so a substep is the same
// sequence of ~36 launches every time -- except that each search swaps the two buffers of every list.  A graph is therefore captured
// per buffer parity, keyed by everything on the host that shapes the launches; any change (settings, a re-allocated scratch slot,
// another pair list made active, profiling switched on) simply misses the key and the substep is enqueued launch by launch again.
uint64_t fnv(uint64_t h, const void* p, size_t n)
{
	const unsigned char* b = (const unsigned char*)p;
	for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}

uint64_t state_key(const apbf_sim* sim)
{
	const apbf_ctx* ctx = sim->ctx;
	const apbf_fluid& f = sim->fluid;
	uint64_t h = 1469598103934665603ull;
	const void* ptrs[] = { f.particle.index_list.data, f.particle.position.data, f.particle.velocity.data, f.particle.inverse_mass.data,
	                       f.particle.radius.data, f.particle.pos_backup.data, f.particle.transferring.data, f.target_radius.data,
	                       f.kernel_width.data, f.boundariness.data, f.boundary_distance.data, sim->nb.pairs, ctx->nbr_struct_pairs, ctx->stream };
	h = fnv(h, ptrs, sizeof ptrs);
	h = fnv(h, &ctx->settings, sizeof ctx->settings);
	h = fnv(h, &sim->cfg, sizeof sim->cfg);
	const uint64_t words[] = { ctx->scratch_epoch, (uint64_t)ctx->dims, (uint64_t)ctx->nbr_struct_n_cap, (uint64_t)ctx->nbr_valid, (uint64_t)ctx->nbr_public,
	                           (uint64_t)ctx->stream_blocks_cap, (uint64_t)ctx->match_grid_min, (uint64_t)sim->no_fuse, (uint64_t)ctx->num_sms };
	h = fnv(h, words, sizeof words);
	h = fnv(h, &sim->last_dt, sizeof sim->last_dt);
	return h ? h : 1ull;
}

void drop_graphs(apbf_sim* sim)
{
	for (apbf_sim_graph& g : sim->graphs) {
		if (g.exec) cudaGraphExecDestroy(g.exec);
		g = apbf_sim_graph();
	}
}

// 1: the substep ran as a graph; 0: not this time (enqueue it the ordinary way); < 0: error
int substep_as_graph(apbf_sim* sim)
{
	apbf_ctx* ctx = sim->ctx;
	cudaStream_t st = ctx->stream;
	const uint64_t key = state_key(sim);
	for (apbf_sim_graph& g : sim->graphs) {
		if (g.exec && g.key == key) {
			if (cudaGraphLaunch(g.exec, st) != cudaSuccess) { cudaGetLastError(); drop_graphs(sim); return 0; }
			// what the captured call did on the host
			apbf_sim_swap_buffers(sim);
			sim->last_dt = sim->cfg.integrate ? sim->cfg.dt : sim->last_dt;
			ctx->nbr_struct_pairs = g.nbr_pairs; ctx->nbr_struct_n_cap = g.nbr_n_cap; ctx->nbr_valid = g.nbr_valid; ctx->nbr_public = g.nbr_public;
			ctx->launches += g.launches;
			sim->graph_replays++;
			return 1;
		}
	}
	if (key == sim->bad_key) return 0;
	bool seen = false;
	for (uint64_t k : sim->seen_keys) seen = seen || k == key;
	if (!seen) { // first visit of this state: run it the ordinary way (scratch slots may still be growing)
		for (int i = 3; i > 0; i--) sim->seen_keys[i] = sim->seen_keys[i - 1];
		sim->seen_keys[0] = key;
		return 0;
	}
	// second visit: capture.  Nothing executes during the capture, the graph is launched afterwards.
	const uint64_t epoch0 = ctx->scratch_epoch, launches0 = ctx->launches;
	const float last_dt0 = sim->last_dt;
	// recorded on a stream of the scene's own (the context's stream may be the legacy default stream, which cannot capture; the
	// nodes do not remember the stream) and launched into the context's stream afterwards
	if (!sim->capture_stream && cudaStreamCreateWithFlags(&sim->capture_stream, cudaStreamNonBlocking) != cudaSuccess) {
		cudaGetLastError(); sim->capture_stream = nullptr; sim->bad_key = key; return 0;
	}
	if (cudaStreamBeginCapture(sim->capture_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); sim->bad_key = key; return 0; }
	ctx->stream = sim->capture_stream;
	const int rc = substep_once(sim);
	ctx->stream = st;
	cudaGraph_t graph = nullptr;
	const cudaError_t ce = cudaStreamEndCapture(sim->capture_stream, &graph);
	cudaGraphExec_t exec = nullptr;
	bool ok = rc == APBF_OK && ce == cudaSuccess && graph && ctx->scratch_epoch == epoch0;
	if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
	if (graph) cudaGraphDestroy(graph);
	if (!ok) { // the host went through a substep, the device did not: put the host back and let the ordinary path run
		cudaGetLastError();
		if (exec) cudaGraphExecDestroy(exec);
		apbf_sim_swap_buffers(sim);
		sim->last_dt = last_dt0;
		ctx->launches = launches0;
		ctx->nbr_valid = false; // (whatever the capture recorded about the pair list did not happen)
		sim->bad_key = key;
		return rc != APBF_OK ? rc : 0;
	}
	apbf_sim_graph* slot = &sim->graphs[0];
	if (slot->exec && !sim->graphs[1].exec) slot = &sim->graphs[1];
	else if (slot->exec && sim->graphs[1].exec) { // both taken by stale states: start over
		drop_graphs(sim);
		slot = &sim->graphs[0];
	}
	slot->key = key; slot->exec = exec; slot->launches = ctx->launches - launches0;
	slot->nbr_pairs = ctx->nbr_struct_pairs; slot->nbr_n_cap = ctx->nbr_struct_n_cap; slot->nbr_valid = ctx->nbr_valid; slot->nbr_public = ctx->nbr_public;
	if (cudaGraphLaunch(exec, st) != cudaSuccess) return apbf_fail(ctx, APBF_ERR_CUDA, "cudaGraphLaunch", __FILE__, __LINE__);
	sim->graph_replays++;
	return 1;
}

} // namespace

extern "C" {

int apbf_sim_set_graphs(apbf_sim* sim, int enable)
{
	if (!sim) return APBF_ERR_INVALID;
	sim->graphs_on = enable ? 1 : 0;
	if (!enable) drop_graphs(sim);
	return APBF_OK;
}

int apbf_sim_graph_replays(const apbf_sim* sim, uint64_t* out)
{
	if (!sim || !out) return APBF_ERR_INVALID;
	*out = sim->graph_replays;
	return APBF_OK;
}

int apbf_sim_substep(apbf_sim* sim, uint32_t n_substeps)
{
	if (!sim) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	const apbf_sim_config& c = sim->cfg;
	APBF_TRY(apbf_ctx_set_dimensions(ctx, c.dims));
	const apbf_settings& s = ctx->settings;
	const bool transfers = c.transfers && !c.basic_pbf && (s.mMerge || s.mSplit);    // pool.cpp:73, :99
	if (transfers && ctx->mg_enabled) return apbf_fail(ctx, APBF_ERR_UNSUPPORTED, "merge / split is not available on slabs", __FILE__, __LINE__);
	// graphs: not while the passes are being timed or counted (events / extra launches), not with merge / split (their lists swap
	// buffers on their own schedule), not on slabs (apbf_sim_mg_substep drives those)
	const bool may_graph = sim->graphs_on && !ctx->prof_on && !ctx->search_stats && !transfers && !ctx->mg_enabled;
	for (uint32_t step = 0; step < n_substeps; step++) {
		if (may_graph) {
			const int g = substep_as_graph(sim);
			if (g < 0) return g;
			if (g == 1) continue;
		}
		APBF_TRY(substep_once(sim));
	}
	return APBF_OK;
}

// Host buffers in, one substep, host buffers out -- the reference-facing call when the scene lives in host memory -- with the
// copies overlapped with the work instead of bracketing it:
//   * positions, velocities and position back-ups go up first on the context's stream (velocity_handling and the hash + sort need
//     nothing else); the seven small lists follow on a second stream and are waited for right before the search re-orders the lists;
//   * from the re-order on, every list but positions, kernel widths and boundariness (and what update_transfers writes) is final:
//     those 56 of 80 bytes per particle come down on the second stream while emit and solver run; the rest follows the last kernel.
// Same results as apbf_sim_upload + apbf_sim_substep(1) + apbf_sim_download (which is what runs for configurations whose particle
// count can change -- merge / split -- or that use the binary search or slabs).  host_in / host_out should be pinned; they may be the
// same buffers.  Synchronises before it returns.
int apbf_sim_step_host(apbf_sim* sim, const apbf_host_state* in, apbf_host_state* out)
{
	if (!sim || !in || !out) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	const apbf_sim_config& c = sim->cfg;
	const apbf_settings& s = ctx->settings;
	const bool transfers = c.transfers && !c.basic_pbf && (s.mMerge || s.mSplit);
	if (transfers || c.use_binary_search || ctx->mg_enabled || ctx->prof_on) {
		APBF_TRY(apbf_sim_upload(sim, in));
		APBF_TRY(apbf_sim_substep(sim, 1));
		return apbf_sim_download(sim, out);
	}
	APBF_REQUIRE(ctx, in->n <= c.particle_capacity);
	APBF_TRY(apbf_ctx_set_dimensions(ctx, c.dims));
	if (!sim->copy_stream) {
		APBF_CUDA(ctx, cudaStreamCreateWithFlags(&sim->copy_stream, cudaStreamNonBlocking));
		for (cudaEvent_t* e : { &sim->ev_fork, &sim->ev_up, &sim->ev_lists, &sim->ev_down }) APBF_CUDA(ctx, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
	}
	cudaStream_t st = ctx->stream, cs = sim->copy_stream;
	const size_t n = in->n;
	struct item { const void* src; void* dst; size_t stride; };
	// ---- up ----
	APBF_CUDA(ctx, cudaEventRecord(sim->ev_fork, st)); // (the second stream starts behind whatever the context's stream still holds)
	APBF_CUDA(ctx, cudaStreamWaitEvent(cs, sim->ev_fork, 0));
	{
		apbf_fluid& f = sim->fluid;
		const item first[] = { { in->position, f.particle.position.data, 16 }, { in->velocity, f.particle.velocity.data, 16 }, { in->pos_backup, f.particle.pos_backup.data, 16 } };
		const item rest[] = { { in->inverse_mass, f.particle.inverse_mass.data, 4 }, { in->radius, f.particle.radius.data, 4 }, { in->transferring, f.particle.transferring.data, 4 },
		                      { in->target_radius, f.target_radius.data, 4 }, { in->kernel_width, f.kernel_width.data, 4 }, { in->boundariness, f.boundariness.data, 4 },
		                      { in->boundary_distance, f.boundary_distance.data, 4 }, { in->index_list, f.particle.index_list.data, 4 } };
		for (const item& it : first) if (it.src && n) APBF_CUDA(ctx, cudaMemcpyAsync(it.dst, it.src, it.stride * n, cudaMemcpyHostToDevice, st));
		for (const item& it : rest) if (it.src && n) APBF_CUDA(ctx, cudaMemcpyAsync(it.dst, it.src, it.stride * n, cudaMemcpyHostToDevice, cs));
		k_set_lengths<<<1, 1, 0, st>>>(f.particle.length, f.particle.hidden_length, (uint32_t)n);
		APBF_LAUNCHED(ctx);
		if (!in->index_list) APBF_TRY(apbf_write_sequence(ctx, (uint32_t*)f.particle.index_list.data, f.particle.length, c.particle_capacity, 0u, 1u, 1u));
		APBF_CUDA(ctx, cudaEventRecord(sim->ev_up, cs));
	}
	// ---- one substep, launch by launch (the two hooks sit inside the search) ----
	ctx->hook_wait_before_reorder = sim->ev_up;
	ctx->hook_record_after_reorder = sim->ev_lists;
	const int rc = substep_once(sim);
	ctx->hook_wait_before_reorder = ctx->hook_record_after_reorder = nullptr;
	APBF_TRY(rc);
	// ---- down ----
	apbf_fluid& f = sim->fluid; // (after the search's buffer swap)
	out->n = (uint32_t)n;
	const bool late_tr = c.update_transfers && !c.basic_pbf; // update_transfers writes these two after the solver
	const item early[] = { { f.particle.velocity.data, out->velocity, 16 }, { f.particle.pos_backup.data, out->pos_backup, 16 }, { f.particle.inverse_mass.data, out->inverse_mass, 4 },
	                       { f.particle.radius.data, out->radius, 4 }, { f.particle.transferring.data, out->transferring, 4 }, { f.particle.index_list.data, out->index_list, 4 },
	                       { late_tr ? nullptr : f.target_radius.data, out->target_radius, 4 }, { late_tr ? nullptr : f.boundary_distance.data, out->boundary_distance, 4 } };
	const item late[] = { { f.particle.position.data, out->position, 16 }, { f.kernel_width.data, out->kernel_width, 4 }, { f.boundariness.data, out->boundariness, 4 },
	                      { late_tr ? f.target_radius.data : nullptr, out->target_radius, 4 }, { late_tr ? f.boundary_distance.data : nullptr, out->boundary_distance, 4 } };
	APBF_CUDA(ctx, cudaStreamWaitEvent(cs, sim->ev_lists, 0));
	for (const item& it : early) if (it.src && it.dst && n) APBF_CUDA(ctx, cudaMemcpyAsync(it.dst, it.src, it.stride * n, cudaMemcpyDeviceToHost, cs));
	for (const item& it : late) if (it.src && it.dst && n) APBF_CUDA(ctx, cudaMemcpyAsync(it.dst, it.src, it.stride * n, cudaMemcpyDeviceToHost, st));
	APBF_CUDA(ctx, cudaEventRecord(sim->ev_down, cs));
	APBF_CUDA(ctx, cudaStreamWaitEvent(st, sim->ev_down, 0)); // the context's stream stays the one ordered stream of work
	APBF_CUDA(ctx, cudaStreamSynchronize(st));
	return APBF_OK;
}

int apbf_sim_fluid(apbf_sim* sim, apbf_fluid* out)
{
	if (!sim || !out) return APBF_ERR_INVALID;
	*out = sim->fluid;
	return APBF_OK;
}

int apbf_sim_neighbors(apbf_sim* sim, apbf_neighbors* out)
{
	if (!sim || !out) return APBF_ERR_INVALID;
	// the whole-scene substep keeps only the grouped structure; whoever asks for the list gets the (id, idN) pairs written now
	APBF_TRY(apbf_nbr_materialize(sim->ctx, sim->fluid.particle.length, sim->cfg.particle_capacity, &sim->nb));
	*out = sim->nb;
	return APBF_OK;
}

int apbf_sim_stats(apbf_sim* sim, uint32_t out[4])
{
	if (!sim || !out) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	uint32_t words[8];
	APBF_CUDA(ctx, cudaMemcpyAsync(words, ctx->misc(), sizeof words, cudaMemcpyDeviceToHost, ctx->stream));
	APBF_CUDA(ctx, cudaMemcpyAsync(&out[0], sim->fluid.particle.length, 4, cudaMemcpyDeviceToHost, ctx->stream));
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	out[1] = words[MW_TOTAL_PAIRS];
	out[2] = words[MW_KEPT_PAIRS] == 0xFFFFFFFFu ? words[MW_TOTAL_PAIRS] : words[MW_KEPT_PAIRS];
	out[3] = words[MW_N_ASYM];
	return APBF_OK;
}

int apbf_sim_transfers(apbf_sim* sim, apbf_transfers* out)
{
	if (!sim || !out) return APBF_ERR_INVALID;
	APBF_REQUIRE(sim->ctx, sim->cfg.transfers);
	*out = sim->tr;
	return APBF_OK;
}

int apbf_sim_download_transfers(apbf_sim* sim, uint32_t* out_n, uint32_t* source_host, uint32_t* target_host, float* time_left_host)
{
	if (!sim || !out_n) return APBF_ERR_INVALID;
	apbf_ctx* ctx = sim->ctx;
	APBF_REQUIRE(ctx, sim->cfg.transfers);
	cudaStream_t st = ctx->stream;
	uint32_t n = 0;
	APBF_CUDA(ctx, cudaMemcpyAsync(&n, sim->tr.length, 4, cudaMemcpyDeviceToHost, st));
	APBF_CUDA(ctx, cudaStreamSynchronize(st));
	if (n > sim->tr.capacity) n = sim->tr.capacity;
	*out_n = n;
	struct { void* dst; const void* src; } items[] = { { source_host, sim->tr.source.data }, { target_host, sim->tr.target.data }, { time_left_host, sim->tr.time_left.data } };
	for (auto& it : items)
		if (it.dst && n) APBF_CUDA(ctx, cudaMemcpyAsync(it.dst, it.src, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
	APBF_CUDA(ctx, cudaStreamSynchronize(st));
	return APBF_OK;
}

int apbf_sim_neighbor_count(apbf_sim* sim, uint32_t* out_count)
{
	if (!sim || !out_count) return APBF_ERR_INVALID;
	APBF_CUDA(sim->ctx, cudaMemcpyAsync(out_count, sim->nb.length, 4, cudaMemcpyDeviceToHost, sim->ctx->stream));
	APBF_CUDA(sim->ctx, cudaStreamSynchronize(sim->ctx->stream));
	return APBF_OK;
}

} // extern "C"
