// ctx.cu -- context, buffer pool, byte copies (replaces shader_provider's recording state, gpu_list_data's pool
// (source/gpu_list_data.cpp:6-45) and algorithms::copy_bytes (source/algorithms.cpp:4-36)).
#include "common.cuh"
#include "neighbors.cuh"

int apbf_fail(apbf_ctx* ctx, int code, const char* what, const char* file, int line)
{
	if (ctx) {
		char buf[512];
		snprintf(buf, sizeof buf, "%s (%s:%d)", what, file, line);
		ctx->last_error = buf;
	}
	return code;
}

static cudaEvent_t prof_event(apbf_ctx* ctx)
{
	if (!ctx->prof_free.empty()) { cudaEvent_t e = ctx->prof_free.back(); ctx->prof_free.pop_back(); return e; }
	cudaEvent_t e;
	cudaEventCreate(&e);
	return e;
}

int apbf_prof_begin(apbf_ctx* ctx, int cat)
{
	apbf_prof_span sp;
	sp.cat = cat;
	sp.beg = prof_event(ctx);
	sp.end = prof_event(ctx);
	cudaEventRecord(sp.beg, ctx->stream);
	ctx->prof_spans.push_back(sp);
	return (int)ctx->prof_spans.size() - 1;
}

void apbf_prof_end(apbf_ctx* ctx, int span)
{
	if (span >= 0 && (size_t)span < ctx->prof_spans.size()) cudaEventRecord(ctx->prof_spans[(size_t)span].end, ctx->stream);
}

static void prof_collect(apbf_ctx* ctx)
{
	for (auto& sp : ctx->prof_spans) {
		float ms = 0.f;
		cudaEventSynchronize(sp.end);
		if (cudaEventElapsedTime(&ms, sp.beg, sp.end) == cudaSuccess) { ctx->prof_ms[sp.cat] += ms; ctx->prof_calls[sp.cat]++; }
		ctx->prof_free.push_back(sp.beg);
		ctx->prof_free.push_back(sp.end);
	}
	ctx->prof_spans.clear();
}

void* apbf_ctx::scratch_get(int slot, size_t bytes)
{
	apbf_scratch& s = scratch[slot];
	if (bytes < 256) bytes = 256;
	if (s.bytes >= bytes) return s.ptr;
	// growth happens only while sizes are still warming up; cudaFree/cudaMalloc synchronise, which keeps ordering safe
	if (s.ptr) cudaFree(s.ptr);
	size_t want = bytes + bytes / 8;
	want = (want + 255) & ~(size_t)255;
	if (cudaMalloc(&s.ptr, want) != cudaSuccess) {
		s.ptr = nullptr;
		s.bytes = 0;
		last_error = "scratch allocation failed";
		return nullptr;
	}
	s.bytes = want;
	scratch_epoch++;
	if (slot == SLOT_MISC_WORDS) {
		cudaMemset(s.ptr, 0, want);
		const uint32_t none = 0xFFFFFFFFu;
		cudaMemcpy((uint32_t*)s.ptr + MW_N_OWNED, &none, 4, cudaMemcpyHostToDevice);
	}
	return s.ptr;
}

extern "C" {

void apbf_default_settings(apbf_settings* s) // source/settings.cpp:5-31
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

int apbf_ctx_create(int device, void* cuda_stream, apbf_ctx** out_ctx)
{
	if (!out_ctx) return APBF_ERR_INVALID;
	*out_ctx = nullptr;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return APBF_ERR_NO_DEVICE; // no CPU fallback
	if (device < 0 || device >= count) return APBF_ERR_INVALID;
	if (cudaSetDevice(device) != cudaSuccess) return APBF_ERR_CUDA;
	apbf_ctx* ctx = new apbf_ctx();
	ctx->device = device;
	ctx->stream = (cudaStream_t)cuda_stream;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
	apbf_default_settings(&ctx->settings);
	if (!ctx->misc()) { delete ctx; return APBF_ERR_OOM; }
	*out_ctx = ctx;
	return APBF_OK;
}

void apbf_ctx_destroy(apbf_ctx* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	prof_collect(ctx);
	for (auto e : ctx->prof_free) cudaEventDestroy(e);
	for (auto& s : ctx->scratch) if (s.ptr) cudaFree(s.ptr);
	for (auto& b : ctx->pool) cudaFree(b.ptr);
	for (auto& e : ctx->nbr_cache) { if (e.offsets.ptr) cudaFree(e.offsets.ptr); if (e.nbl.ptr) cudaFree(e.nbl.ptr); if (e.words) cudaFree(e.words); }
	delete ctx;
}

int apbf_ctx_set_stream(apbf_ctx* ctx, void* cuda_stream)
{
	if (!ctx) return APBF_ERR_INVALID;
	ctx->stream = (cudaStream_t)cuda_stream;
	return APBF_OK;
}

int apbf_ctx_synchronize(apbf_ctx* ctx)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return APBF_OK;
}

const char* apbf_ctx_last_error(apbf_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int apbf_ctx_set_settings(apbf_ctx* ctx, const apbf_settings* s)
{
	if (!ctx || !s) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, s->mHeightKernelId >= 0 && s->mHeightKernelId <= 4);
	APBF_REQUIRE(ctx, s->mGradientKernelId >= 0 && s->mGradientKernelId <= 4);
	APBF_REQUIRE(ctx, s->mBoundarinessCalculationMethod >= 0 && s->mBoundarinessCalculationMethod <= 2);
	ctx->settings = *s;
	return APBF_OK;
}

int apbf_ctx_set_dimensions(apbf_ctx* ctx, int dims)
{
	if (!ctx) return APBF_ERR_INVALID;
	APBF_REQUIRE(ctx, dims == 2 || dims == 3);
	ctx->dims = dims;
	return APBF_OK;
}

uint64_t apbf_ctx_launch_count(apbf_ctx* ctx) { return ctx ? ctx->launches : 0; }
int apbf_ctx_set_stream_blocks(apbf_ctx* ctx, uint32_t max_blocks)
{
	if (!ctx) return APBF_ERR_INVALID;
	ctx->stream_blocks_cap = max_blocks;
	return APBF_OK;
}
int apbf_ctx_set_match_grid_min(apbf_ctx* ctx, uint32_t min_candidates)
{
	if (!ctx) return APBF_ERR_INVALID;
	ctx->match_grid_min = min_candidates;
	return APBF_OK;
}
int apbf_ctx_set_search_stats(apbf_ctx* ctx, int enable)
{
	if (!ctx) return APBF_ERR_INVALID;
	ctx->search_stats = enable != 0;
	return APBF_OK;
}

int apbf_ctx_profile(apbf_ctx* ctx, int enable)
{
	if (!ctx) return APBF_ERR_INVALID;
	if (!enable && ctx->prof_on) prof_collect(ctx);
	if (enable && !ctx->prof_on) {
		for (int i = 0; i < PROF_COUNT; i++) { ctx->prof_ms[i] = 0.0; ctx->prof_calls[i] = 0; }
	}
	ctx->prof_on = enable != 0;
	return APBF_OK;
}

static const char* const k_prof_names[PROF_COUNT] = {
	"hash_sort", "reorder", "cell_ranges", "emit_count", "emit_scan", "emit_fill", "kw_spread", "kw_compact", "kw_misc",
	"box_collision", "density_lambda", "apply_delta", "commit", "velocity", "solver_prepare", "update_transfers",
	"mg_route", "mg_halo", "mg_exchange" // slab sections (mgpu.cu); mg_route and mg_halo contain their own exchange
};

int apbf_ctx_profile_read(apbf_ctx* ctx, int category, const char** out_name, double* out_ms, uint64_t* out_calls)
{
	if (!ctx) return APBF_ERR_INVALID;
	if (category < 0 || category >= PROF_COUNT) return APBF_ERR_INVALID;
	prof_collect(ctx);
	if (out_name) *out_name = k_prof_names[category];
	if (out_ms) *out_ms = ctx->prof_ms[category];
	if (out_calls) *out_calls = ctx->prof_calls[category];
	return APBF_OK;
}

int apbf_ctx_device_flags(apbf_ctx* ctx, uint32_t* out_flags)
{
	if (!ctx || !out_flags) return APBF_ERR_INVALID;
	APBF_CUDA(ctx, cudaMemcpyAsync(out_flags, ctx->misc() + MW_FLAGS, 4, cudaMemcpyDeviceToHost, ctx->stream));
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return APBF_OK;
}

int apbf_ctx_list_state(apbf_ctx* ctx, uint32_t out[4])
{
	if (!ctx || !out) return APBF_ERR_INVALID;
	uint32_t w[MW_WORDS];
	APBF_CUDA(ctx, cudaMemcpyAsync(w, ctx->misc(), sizeof w, cudaMemcpyDeviceToHost, ctx->stream));
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	out[0] = w[MW_N_ASYM];
	out[1] = w[MW_H_NONUNIFORM] == 0u ? 1u : 0u;
	out[2] = 1u;
	for (int k = 0; k < 4; k++) if (w[MW_THR_MIN + k] < w[MW_THR_MAX + k]) out[2] = 0u;
	out[3] = w[MW_OCC_CELLS];
	return APBF_OK;
}

// best-fit pool like gpu_list_data::get_list (source/gpu_list_data.cpp:6-45)
int apbf_buffer_acquire(apbf_ctx* ctx, size_t bytes, void** out)
{
	if (!ctx || !out) return APBF_ERR_INVALID;
	if (bytes == 0) bytes = 4;
	apbf_pool_block* best = nullptr;
	for (auto& b : ctx->pool)
		if (!b.in_use && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
	if (!best) {
		void* p = nullptr;
		size_t want = (bytes + 255) & ~(size_t)255;
		if (cudaMalloc(&p, want) != cudaSuccess) return apbf_fail(ctx, APBF_ERR_OOM, "cudaMalloc failed", __FILE__, __LINE__);
		ctx->pool.push_back({ p, want, false });
		best = &ctx->pool.back();
	}
	best->in_use = true;
	*out = best->ptr;
	apbf_nbr_forget(ctx, best->ptr); // a recycled buffer is a new list: whatever structure was known for this address is gone
	return APBF_OK;
}

int apbf_buffer_release(apbf_ctx* ctx, void* p)
{
	if (!ctx) return APBF_ERR_INVALID;
	for (auto& b : ctx->pool)
		if (b.ptr == p) { b.in_use = false; apbf_nbr_forget(ctx, p); return APBF_OK; }
	return apbf_fail(ctx, APBF_ERR_INVALID, "buffer not from this pool", __FILE__, __LINE__);
}

int apbf_copy_bytes(apbf_ctx* ctx, const void* src, void* dst, size_t bytes)
{
	if (!ctx) return APBF_ERR_INVALID;
	if (bytes == 0) return APBF_OK; // algorithms.cpp:6
	apbf_nbr_touch(ctx, dst);
	APBF_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	return APBF_OK;
}

int apbf_copy_bytes_from_host(apbf_ctx* ctx, const void* src, void* dst, size_t bytes)
{
	if (!ctx) return APBF_ERR_INVALID;
	if (bytes == 0) return APBF_OK;
	apbf_nbr_touch(ctx, dst);
	APBF_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return APBF_OK;
}

int apbf_copy_bytes_to_host(apbf_ctx* ctx, const void* src, void* dst, size_t bytes)
{
	if (!ctx) return APBF_ERR_INVALID;
	if (bytes == 0) return APBF_OK;
	APBF_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	APBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return APBF_OK;
}

int apbf_host_alloc_pinned(size_t bytes, void** out)
{
	if (!out) return APBF_ERR_INVALID;
	return cudaMallocHost(out, bytes ? bytes : 4) == cudaSuccess ? APBF_OK : APBF_ERR_OOM;
}

int apbf_host_free_pinned(void* p) { return cudaFreeHost(p) == cudaSuccess ? APBF_OK : APBF_ERR_CUDA; }

size_t apbf_prefix_sum_calculate_needed_helper_list_length(size_t max_count) // algorithms.cpp:48-58
{
	uint32_t n = (uint32_t)max_count, result = 0u;
	do {
		n = (n + 511u) / 512u;
		result += n;
	} while (n > 1u);
	return result == 0u ? 0u : result + 10u;
}

size_t apbf_sort_calculate_needed_helper_list_length(size_t max_count) // algorithms.cpp:38-46
{
	uint32_t tables = 16u * (((uint32_t)max_count + 511u) / 512u);
	return tables + apbf_prefix_sum_calculate_needed_helper_list_length(tables);
}

} // extern "C"
