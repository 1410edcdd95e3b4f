// solver.cuh -- internal interface of the incompressibility sweeps (incompress.cu)
#pragma once
#include "common.cuh"

enum apbf_iter_flags {
	ITER_BEGIN_COMMIT = 1, // add the previous iteration's pending position deltas first
	ITER_BEGIN_BOX = 2,    // box_collision fused into the prologue (pool.cpp:93)
	ITER_END_COMMIT = 4,   // add this iteration's deltas to the positions before returning
	// multi-GPU: run only part of the iteration (halo exchanges happen in between); none of the three set = all of them
	ITER_RUN_BEGIN = 8, ITER_RUN_T1 = 16, ITER_RUN_T2 = 32,
	// The apply sweep may finish the iteration itself: in its equal-width form (one record per neighbour, nobody reads the
	// packed positions of others) every lane ends up with its particle's complete shift, so it adds it to the position right
	// away (ITER_T2_COMMIT) and, if another iteration follows, runs that iteration's box collision and packs the new position
	// (ITER_T2_NEXT_BOX; pass the boxes).  Whether the sweep takes that form is decided on the device; the prologue / commit
	// launch that would otherwise do the work is told with ITER_SKIP_IF_T2_DID to return at once in that case.
	ITER_T2_COMMIT = 64, ITER_T2_NEXT_BOX = 128, ITER_SKIP_IF_T2_DID = 256,
	// slabs: a sweep over part of the tiles (32 consecutive particles each).  BOUNDARY = tiles that hold a particle some other rank has
	// as a ghost -- the only owned particles that can have ghost neighbours -- or a ghost; INTERIOR = the rest, which needs nothing
	// from the halo exchange that is in flight (mgpu.cu: the exchanges hide behind the interior sweeps).  Flags: SLOT_MG_TILES.
	ITER_TILES_INTERIOR = 512, ITER_TILES_BOUNDARY = 1024
};

// per-particle constants that only depend on kernel width / radius / inverse mass (exact double-precision pow);
// valid until one of those lists changes
int apbf_solver_prepare(apbf_ctx* ctx, apbf_fluid* fluid);
// one incompressibility::apply(); positions are only final after an iteration run with ITER_END_COMMIT
int apbf_solver_iteration(apbf_ctx* ctx, apbf_fluid* fluid, const apbf_neighbors* nb, int flags, const float* box_min4,
                          const float* box_max4, uint32_t n_boxes, float* out_lambda, uint32_t* out_incomp);
