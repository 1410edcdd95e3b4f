// box.cuh -- box_collision.comp:20-60 as device functions (shared by the stand-alone operator and the fused solver prologue)
#pragma once
#include "kernels.cuh"

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

// the loop over the boxes of box_collision.comp:45-57; the caller converts to / from fixed point
__device__ __forceinline__ void box_push(float& px, float& py, float& pz, uint32_t id, float radius, const float4* __restrict__ bmin,
                                         const float4* __restrict__ bmax, uint32_t n_boxes)
{
	for (uint32_t i = 0; i < n_boxes; i++) {
		{
			// A push needs the particle strictly inside the box grown by radius + jitter on every axis, and the jitter is
			// 0.05 * hash with hash in [0, 1).  Outside the box grown by radius + 0.06 nothing can happen: skip the two
			// position hashes (most particles are nowhere near a wall).  0.01 covers the roundings many times over.
			const float4 lo = __ldg(bmin + i), hi = __ldg(bmax + i);
			const float m = radius + 0.06f;
			if (px < lo.x - m || py < lo.y - m || pz < lo.z - m || px > hi.x + m || py > hi.y + m || pz > hi.z + m) continue;
		}
		const vec3f h0 = hash31((float)id * px - py - pz);
		const vec3f h1 = hash31((float)id * py + px + pz);
		const float4 lo = __ldg(bmin + i), hi = __ldg(bmax + i);
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
