// solver.cuh -- internal interface of the incompressibility sweeps (incompress.cu)
#pragma once
#include "common.cuh"

enum apbf_iter_flags {
	ITER_BEGIN_COMMIT = 1, // add the previous iteration's pending position deltas first
	ITER_BEGIN_BOX = 2,    // box_collision fused into the prologue (pool.cpp:93)
	ITER_END_COMMIT = 4,   // add this iteration's deltas to the positions before returning
	// multi-GPU: run only part of the iteration (halo exchanges happen in between); none of the three set = all of them
	ITER_RUN_BEGIN = 8, ITER_RUN_T1 = 16, ITER_RUN_T2 = 32
};

// per-particle constants that only depend on kernel width / radius / inverse mass (exact double-precision pow);
// valid until one of those lists changes
int apbf_solver_prepare(apbf_ctx* ctx, apbf_fluid* fluid);
// one incompressibility::apply(); positions are only final after an iteration run with ITER_END_COMMIT
int apbf_solver_iteration(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, int flags, const float* box_min4,
                          const float* box_max4, uint32_t n_boxes, float* out_lambda, uint32_t* out_incomp);
