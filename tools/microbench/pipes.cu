// pipes.cu -- issue rates of the instruction kinds the pair emit and the solver sweeps are made of, measured on the GPU at hand
// (B200, sm_100a): warp instructions per cycle and SM sub-partition, for independent chains (throughput, not latency).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes   (tools/microbench/README in DESIGN.md 3.5)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 64
#define ITERS 256

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, unsigned long long* cyc, float seed)
{
	float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
	float b = seed * 0.5f + 1.0001f, c = seed + 0.25f;
	unsigned u0 = threadIdx.x + 1, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7, u4 = u0 * 11, u5 = u0 * 13, u6 = u0 * 17, u7 = u0 * 19, ub = (unsigned)seed + 12345u;
	unsigned long long p0, p1, p2, p3, pb;
	asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(a0), "f"(a1));
	asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(a2), "f"(a3));
	asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(a4), "f"(a5));
	asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(a6), "f"(a7));
	asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(c));
	__shared__ float4 sm[64];
	if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed + 1, seed + 2, seed + 3);
	__syncthreads();
	unsigned long long t0 = clock64();
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int r = 0; r < REP / 8; r++) {
			if (OP == 0) { // FADD reg,reg
				asm volatile("add.rn.f32 %0,%0,%8; add.rn.f32 %1,%1,%8; add.rn.f32 %2,%2,%8; add.rn.f32 %3,%3,%8; add.rn.f32 %4,%4,%8; add.rn.f32 %5,%5,%8; add.rn.f32 %6,%6,%8; add.rn.f32 %7,%7,%8;"
				             : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b));
			} else if (OP == 1) { // FMUL reg,reg
				asm volatile("mul.rn.f32 %0,%0,%8; mul.rn.f32 %1,%1,%8; mul.rn.f32 %2,%2,%8; mul.rn.f32 %3,%3,%8; mul.rn.f32 %4,%4,%8; mul.rn.f32 %5,%5,%8; mul.rn.f32 %6,%6,%8; mul.rn.f32 %7,%7,%8;"
				             : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b));
			} else if (OP == 2) { // FFMA 3 reg
				asm volatile("fma.rn.f32 %0,%0,%8,%9; fma.rn.f32 %1,%1,%8,%9; fma.rn.f32 %2,%2,%8,%9; fma.rn.f32 %3,%3,%8,%9; fma.rn.f32 %4,%4,%8,%9; fma.rn.f32 %5,%5,%8,%9; fma.rn.f32 %6,%6,%8,%9; fma.rn.f32 %7,%7,%8,%9;"
				             : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(b), "f"(c));
			} else if (OP == 3) { // FADD2 (4 packed = 8 flops per 4 instr): 8 instr per r
				asm volatile("add.rn.f32x2 %0,%0,%4; add.rn.f32x2 %1,%1,%4; add.rn.f32x2 %2,%2,%4; add.rn.f32x2 %3,%3,%4; add.rn.f32x2 %0,%0,%4; add.rn.f32x2 %1,%1,%4; add.rn.f32x2 %2,%2,%4; add.rn.f32x2 %3,%3,%4;"
				             : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3) : "l"(pb));
			} else if (OP == 4) { // FMUL2
				asm volatile("mul.rn.f32x2 %0,%0,%4; mul.rn.f32x2 %1,%1,%4; mul.rn.f32x2 %2,%2,%4; mul.rn.f32x2 %3,%3,%4; mul.rn.f32x2 %0,%0,%4; mul.rn.f32x2 %1,%1,%4; mul.rn.f32x2 %2,%2,%4; mul.rn.f32x2 %3,%3,%4;"
				             : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3) : "l"(pb));
			} else if (OP == 5) { // IADD3
				asm volatile("add.u32 %0,%0,%1; add.u32 %1,%1,%2; add.u32 %2,%2,%3; add.u32 %3,%3,%4; add.u32 %4,%4,%5; add.u32 %5,%5,%6; add.u32 %6,%6,%7; add.u32 %7,%7,%0;"
				             : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(ub));
			} else if (OP == 6) { // LOP3
				asm volatile("lop3.b32 %0,%0,%8,%1,0x96; lop3.b32 %1,%1,%8,%2,0x96; lop3.b32 %2,%2,%8,%3,0x96; lop3.b32 %3,%3,%8,%4,0x96; lop3.b32 %4,%4,%8,%5,0x96; lop3.b32 %5,%5,%8,%6,0x96; lop3.b32 %6,%6,%8,%7,0x96; lop3.b32 %7,%7,%8,%0,0x96;"
				             : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(ub));
			} else if (OP == 7) { // 4 FMUL + 4 IADD interleaved
				asm volatile("mul.rn.f32 %0,%0,%8; add.u32 %4,%4,%9; mul.rn.f32 %1,%1,%8; add.u32 %5,%5,%9; mul.rn.f32 %2,%2,%8; add.u32 %6,%6,%9; mul.rn.f32 %3,%3,%8; add.u32 %7,%7,%9;"
				             : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "f"(b), "r"(ub));
			} else if (OP == 8) { // setp.gt + predicated or (the emit's test tail), 4 pairs
				asm volatile("{.reg .pred p;\n setp.gt.f32 p,%0,%8; @!p or.b32 %4,%4,%9; setp.gt.f32 p,%1,%8; @!p or.b32 %5,%5,%9; setp.gt.f32 p,%2,%8; @!p or.b32 %6,%6,%9; setp.gt.f32 p,%3,%8; @!p or.b32 %7,%7,%9;}"
				             : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "f"(b), "r"(ub));
			} else if (OP == 9) { // F2I.TRUNC + I2F round trips (both conversions, dependent: nothing for ptxas to drop)
				asm volatile("cvt.rzi.s32.f32 %0,%8; cvt.rzi.s32.f32 %1,%9; cvt.rzi.s32.f32 %2,%10; cvt.rzi.s32.f32 %3,%11; cvt.rzi.s32.f32 %4,%12; cvt.rzi.s32.f32 %5,%13; cvt.rzi.s32.f32 %6,%14; cvt.rzi.s32.f32 %7,%15;"
				             : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3), "=r"(u4), "=r"(u5), "=r"(u6), "=r"(u7) : "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5), "f"(a6), "f"(a7));
				asm volatile("cvt.rn.f32.s32 %0,%8; cvt.rn.f32.s32 %1,%9; cvt.rn.f32.s32 %2,%10; cvt.rn.f32.s32 %3,%11; cvt.rn.f32.s32 %4,%12; cvt.rn.f32.s32 %5,%13; cvt.rn.f32.s32 %6,%14; cvt.rn.f32.s32 %7,%15;"
				             : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7) : "r"(u0), "r"(u1), "r"(u2), "r"(u3), "r"(u4), "r"(u5), "r"(u6), "r"(u7));
			} else if (OP == 10) { // the same conversions by magic numbers: FADD.RZ + IADD (float -> int, 0 <= x < 2^23), IADD + FADD (int -> float)
				asm volatile("add.rz.f32 %0,%0,0f4B000000; add.rz.f32 %1,%1,0f4B000000; add.rz.f32 %2,%2,0f4B000000; add.rz.f32 %3,%3,0f4B000000; add.rz.f32 %4,%4,0f4B000000; add.rz.f32 %5,%5,0f4B000000; add.rz.f32 %6,%6,0f4B000000; add.rz.f32 %7,%7,0f4B000000;"
				             : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));
				asm volatile("add.rn.f32 %0,%0,0fCB000000; add.rn.f32 %1,%1,0fCB000000; add.rn.f32 %2,%2,0fCB000000; add.rn.f32 %3,%3,0fCB000000; add.rn.f32 %4,%4,0fCB000000; add.rn.f32 %5,%5,0fCB000000; add.rn.f32 %6,%6,0fCB000000; add.rn.f32 %7,%7,0fCB000000;"
				             : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));
			} else if (OP == 11) { // MUFU.EX2
				asm volatile("ex2.approx.ftz.f32 %0,%0; ex2.approx.ftz.f32 %1,%1; ex2.approx.ftz.f32 %2,%2; ex2.approx.ftz.f32 %3,%3; ex2.approx.ftz.f32 %4,%4; ex2.approx.ftz.f32 %5,%5; ex2.approx.ftz.f32 %6,%6; ex2.approx.ftz.f32 %7,%7;"
				             : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));
			} else if (OP == 12) { // LDS.128 broadcast
				float4 v;
#pragma unroll
				for (int j = 0; j < 8; j++) {
					asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(r * 8 + j + it) & 63])));
					a0 += v.x; // dependent consumer keeps it alive (1 FADD per load)
				}
			} else if (OP == 13) { // SHFL
				asm volatile("shfl.sync.bfly.b32 %0,%1,1,31,0xffffffff; shfl.sync.bfly.b32 %1,%2,2,31,0xffffffff; shfl.sync.bfly.b32 %2,%3,4,31,0xffffffff; shfl.sync.bfly.b32 %3,%4,8,31,0xffffffff; shfl.sync.bfly.b32 %4,%5,16,31,0xffffffff; shfl.sync.bfly.b32 %5,%6,1,31,0xffffffff; shfl.sync.bfly.b32 %6,%7,2,31,0xffffffff; shfl.sync.bfly.b32 %7,%0,4,31,0xffffffff;"
				             : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7) : "r"(ub & 31u));
			} else if (OP == 14) { // VOTE.ballot of a float compare (setp + vote)
				asm volatile("{.reg .pred p;\n setp.gt.f32 p,%8,%9; vote.sync.ballot.b32 %0,p,0xffffffff; setp.gt.f32 p,%10,%9; vote.sync.ballot.b32 %1,p,0xffffffff; setp.gt.f32 p,%11,%9; vote.sync.ballot.b32 %2,p,0xffffffff; setp.gt.f32 p,%12,%9; vote.sync.ballot.b32 %3,p,0xffffffff;}"
				             : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "f"(a0), "f"(b), "f"(a1), "f"(a2), "f"(a3), "r"(u4), "r"(u5), "r"(u6), "r"(u7));
				a0 += __uint_as_float(u0 & 0x3f800000u); a1 += __uint_as_float(u1 & 0x3f800000u); a2 += __uint_as_float(u2 & 0x3f800000u); a3 += __uint_as_float(u3 & 0x3f800000u); u4 += u0; u5 += u1; u6 += u2; u7 += u3;
			} else if (OP == 15) { // REDUX (min.u32)
				asm volatile("redux.sync.min.u32 %0,%0,0xffffffff; redux.sync.min.u32 %1,%1,0xffffffff; redux.sync.min.u32 %2,%2,0xffffffff; redux.sync.min.u32 %3,%3,0xffffffff; redux.sync.min.u32 %4,%4,0xffffffff; redux.sync.min.u32 %5,%5,0xffffffff; redux.sync.min.u32 %6,%6,0xffffffff; redux.sync.min.u32 %7,%7,0xffffffff;"
				             : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7));
			} else if (OP == 16) { // POPC
				asm volatile("popc.b32 %0,%0; popc.b32 %1,%1; popc.b32 %2,%2; popc.b32 %3,%3; popc.b32 %4,%4; popc.b32 %5,%5; popc.b32 %6,%6; popc.b32 %7,%7;"
				             : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3), "+r"(u4), "+r"(u5), "+r"(u6), "+r"(u7));
			} else if (OP == 17) { // emit inner loop as written today (MODE 2): LDS.128 + LDS.32 + 8 flops + 2 setp + 2 por + shift
				const float4 q = sm[(r * 8) & 63]; const float su = ((float*)sm)[(r * 8 + 3) & 63];
#pragma unroll
				for (int j = 0; j < 8; j++) {
					float4 qv; float uu;
					asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];" : "=f"(qv.x), "=f"(qv.y), "=f"(qv.z), "=f"(qv.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(r * 8 + j + it) & 63])));
					asm volatile("ld.shared.f32 %0,[%1];" : "=f"(uu) : "r"((unsigned)__cvta_generic_to_shared(((float*)sm) + ((r * 8 + j + it) & 63))));
					const float dx = __fsub_rn(qv.x, a0), dy = __fsub_rn(qv.y, a1), dz = __fsub_rn(qv.z, a2);
					const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
					asm volatile("{.reg .pred pk, pu;\n setp.gt.f32 pk,%2,%3; setp.gt.f32 pu,%2,%4; @!pk or.b32 %0,%0,%5; @!pu or.b32 %1,%1,%5;}" : "+r"(u0), "+r"(u1) : "f"(d2), "f"(qv.w), "f"(uu), "r"(u2));
					u2 += u2;
				}
				(void)q; (void)su;
			} else if (OP == 18) { // the same with two queries per step in f32x2 (explicit mul/add, no fma), one threshold, ballot-free
#pragma unroll
				for (int j = 0; j < 8; j += 2) {
					float4 qa, qb;
					asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];" : "=f"(qa.x), "=f"(qa.y), "=f"(qa.z), "=f"(qa.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(r * 8 + j + it) & 63])));
					asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];" : "=f"(qb.x), "=f"(qb.y), "=f"(qb.z), "=f"(qb.w) : "r"((unsigned)__cvta_generic_to_shared(&sm[(r * 8 + j + 1 + it) & 63])));
					unsigned long long X, Y, Z, cx, cy, cz, s;
					asm("mov.b64 %0,{%1,%2};" : "=l"(X) : "f"(qa.x), "f"(qb.x));
					asm("mov.b64 %0,{%1,%2};" : "=l"(Y) : "f"(qa.y), "f"(qb.y));
					asm("mov.b64 %0,{%1,%2};" : "=l"(Z) : "f"(qa.z), "f"(qb.z));
					asm("mov.b64 %0,{%1,%1};" : "=l"(cx) : "f"(a0));
					asm("mov.b64 %0,{%1,%1};" : "=l"(cy) : "f"(a1));
					asm("mov.b64 %0,{%1,%1};" : "=l"(cz) : "f"(a2));
					asm volatile("{.reg .b64 dx,dy,dz,t;\n sub.rn.f32x2 dx,%1,%4; sub.rn.f32x2 dy,%2,%5; sub.rn.f32x2 dz,%3,%6; mul.rn.f32x2 dx,dx,dx; mul.rn.f32x2 dy,dy,dy; mul.rn.f32x2 dz,dz,dz; add.rn.f32x2 t,dx,dy; add.rn.f32x2 %0,t,dz;}"
					             : "=l"(s) : "l"(X), "l"(Y), "l"(Z), "l"(cx), "l"(cy), "l"(cz));
					float s0, s1;
					asm("mov.b64 {%0,%1},%2;" : "=f"(s0), "=f"(s1) : "l"(s));
					asm volatile("{.reg .pred pk, pu;\n setp.gt.f32 pk,%2,%3; setp.gt.f32 pu,%4,%5; @!pk or.b32 %0,%0,%6; @!pu or.b32 %1,%1,%6;}" : "+r"(u0), "+r"(u1) : "f"(s0), "f"(qa.w), "f"(s1), "f"(qb.w), "r"(u2));
					u2 += u2;
				}
			}
		}
	}
	unsigned long long t1 = clock64();
	float acc = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + __uint_as_float(u0 ^ u1 ^ u2 ^ u3 ^ u4 ^ u5 ^ u6 ^ u7);
	float x0, x1;
	asm("mov.b64 {%0,%1},%2;" : "=f"(x0), "=f"(x1) : "l"(p0 ^ p1 ^ p2 ^ p3));
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc + x0 + x1;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_rep8, int warps_per_sm)
{
	int dev_sms = 148;
	float* out; unsigned long long* cyc;
	const int blocks = dev_sms * (warps_per_sm / 8);
	cudaMalloc(&out, sizeof(float) * blocks * 256);
	cudaMalloc(&cyc, sizeof(unsigned long long) * blocks);
	k<OP><<<blocks, 256>>>(out, cyc, 1.0f);
	k<OP><<<blocks, 256>>>(out, cyc, 1.0f);
	cudaDeviceSynchronize();
	unsigned long long* h = new unsigned long long[blocks];
	cudaMemcpy(h, cyc, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost);
	double mean = 0; for (int i = 0; i < blocks; i++) mean += (double)h[i]; mean /= blocks;
	// warp instructions issued per SMSP: warps on the SMSP x ITERS x (REP/8) x instr_per_rep8
	const double wi = (double)(warps_per_sm / 4) * ITERS * (REP / 8) * instr_per_rep8;
	printf("%-44s warps/SM %2d  %7.3f warp-instr/clk/SMSP  (%6.2f clk per instr)  err=%s\n", name, warps_per_sm, wi / mean, mean / wi, cudaGetErrorString(cudaGetLastError()));
	cudaFree(out); cudaFree(cyc); delete[] h;
}

int main()
{
	for (int w : { 8, 32 }) {
		run<0>("FADD r,r", 8, w);
		run<1>("FMUL r,r", 8, w);
		run<2>("FFMA r,r,r", 8, w);
		run<3>("FADD2 (packed f32x2)", 8, w);
		run<4>("FMUL2 (packed f32x2)", 8, w);
		run<5>("IADD3", 8, w);
		run<6>("LOP3", 8, w);
		run<7>("FMUL + IADD3 interleaved", 8, w);
		run<8>("FSETP + predicated LOP3", 8, w);
		run<9>("F2I.TRUNC + I2F (XU conversions)", 16, w);
		run<10>("FADD.RZ + FADD magic conversions", 16, w);
		run<11>("MUFU.EX2", 8, w);
		run<12>("LDS.128 broadcast + FADD", 16, w);
		run<13>("SHFL.IDX", 8, w);
		run<14>("FSETP + VOTE.ballot + LOP + FADD + IADD", 20, w);
		run<15>("REDUX.min", 8, w);
		run<16>("POPC", 8, w);
		run<17>("emit test loop, scalar: clk per query (32 tests)", 8, w);
		run<18>("emit test loop, f32x2 pairs: clk per query", 8, w);
	}
	return 0;
}
