// pipe_microbench.cu -- how the two integer pipes of an sm_100a sub-partition share issue slots.
//
// Question left open by round 1: a front-end variant with 7 % FEWER instructions but an 18-long run of
// IMADs ran 3 % slower.  This measures, per SM and for 1..24 resident warps, the cycles per instruction of
//   A  32 IMAD then 32 LOP3/SHF, each phase fed by the other (same-pipe runs ptxas cannot break up)
//   B  IMAD, LOP3, IMAD, SHF, ...       (alternating pipes)
//   C  64 IMAD                           (FMA pipe only)
//   D  64 LOP3/SHF                       (ALU pipe only)
// every instruction on its own dependency chain of length 8 (so latency is not the limit at >= 2 warps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench pipe_microbench.cu && ./pipe_microbench
// Check the order ptxas kept with:  cuobjdump -sass pipe_microbench | less
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x

template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t *out, long long *cyc, int iters, uint32_t m, uint32_t a)
{
	uint32_t f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
	uint32_t l0 = f0 ^ 9, l1 = f1 ^ 9, l2 = f2 ^ 9, l3 = f3 ^ 9, l4 = f4 ^ 9, l5 = f5 ^ 9, l6 = f6 ^ 9, l7 = f7 ^ 9;
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#define FMA8 asm volatile("mad.lo.u32 %0,%0,%8,%9; mad.lo.u32 %1,%1,%8,%9; mad.lo.u32 %2,%2,%8,%9; mad.lo.u32 %3,%3,%8,%9;" \
                          "mad.lo.u32 %4,%4,%8,%9; mad.lo.u32 %5,%5,%8,%9; mad.lo.u32 %6,%6,%8,%9; mad.lo.u32 %7,%7,%8,%9;" \
                          : "+r"(f0), "+r"(f1), "+r"(f2), "+r"(f3), "+r"(f4), "+r"(f5), "+r"(f6), "+r"(f7) : "r"(m), "r"(a));
#define ALU8 asm volatile("lop3.b32 %0,%0,%8,%9,0x96; shf.l.wrap.b32 %1,%1,%1,%8; lop3.b32 %2,%2,%8,%9,0x96; shf.l.wrap.b32 %3,%3,%3,%8;" \
                          "lop3.b32 %4,%4,%8,%9,0x96; shf.l.wrap.b32 %5,%5,%5,%8; lop3.b32 %6,%6,%8,%9,0x96; shf.l.wrap.b32 %7,%7,%7,%8;" \
                          : "+r"(l0), "+r"(l1), "+r"(l2), "+r"(l3), "+r"(l4), "+r"(l5), "+r"(l6), "+r"(l7) : "r"(m), "r"(a));
#define MIX8 asm volatile("mad.lo.u32 %0,%0,%8,%9; lop3.b32 %4,%4,%8,%9,0x96; mad.lo.u32 %1,%1,%8,%9; shf.l.wrap.b32 %5,%5,%5,%8;" \
                          "mad.lo.u32 %2,%2,%8,%9; lop3.b32 %6,%6,%8,%9,0x96; mad.lo.u32 %3,%3,%8,%9; shf.l.wrap.b32 %7,%7,%7,%8;" \
                          : "+r"(f0), "+r"(f1), "+r"(f2), "+r"(f3), "+r"(l0), "+r"(l1), "+r"(l2), "+r"(l3) : "r"(m), "r"(a));
		// the same with each phase consuming the other pipe's results, so that ptxas cannot interleave them (left alone it
		// turns "32 IMAD then 32 LOP3/SHF" into strict alternation -- it knows)
#define FMA8D asm volatile("mad.lo.u32 %0,%0,%8,%9; mad.lo.u32 %1,%1,%8,%10; mad.lo.u32 %2,%2,%8,%11; mad.lo.u32 %3,%3,%8,%12;" \
                           "mad.lo.u32 %4,%4,%8,%13; mad.lo.u32 %5,%5,%8,%14; mad.lo.u32 %6,%6,%8,%15; mad.lo.u32 %7,%7,%8,%16;" \
                           : "+r"(f0), "+r"(f1), "+r"(f2), "+r"(f3), "+r"(f4), "+r"(f5), "+r"(f6), "+r"(f7) \
                           : "r"(m), "r"(l0), "r"(l1), "r"(l2), "r"(l3), "r"(l4), "r"(l5), "r"(l6), "r"(l7));
#define ALU8D asm volatile("lop3.b32 %0,%0,%8,%9,0x96; lop3.b32 %1,%1,%8,%10,0x96; lop3.b32 %2,%2,%8,%11,0x96; lop3.b32 %3,%3,%8,%12,0x96;" \
                           "lop3.b32 %4,%4,%8,%13,0x96; lop3.b32 %5,%5,%8,%14,0x96; lop3.b32 %6,%6,%8,%15,0x96; lop3.b32 %7,%7,%8,%16,0x96;" \
                           : "+r"(l0), "+r"(l1), "+r"(l2), "+r"(l3), "+r"(l4), "+r"(l5), "+r"(l6), "+r"(l7) \
                           : "r"(m), "r"(f0), "r"(f1), "r"(f2), "r"(f3), "r"(f4), "r"(f5), "r"(f6), "r"(f7));
		if (MODE == 0) { FMA8D FMA8 FMA8 FMA8 ALU8D ALU8 ALU8 ALU8 FMA8D FMA8 FMA8 FMA8 ALU8D ALU8 ALU8 ALU8
		                 FMA8D FMA8 FMA8 FMA8 ALU8D ALU8 ALU8 ALU8 FMA8D FMA8 FMA8 FMA8 ALU8D ALU8 ALU8 ALU8 }
		if (MODE == 1) { REP8(MIX8) REP8(MIX8) REP8(MIX8) REP8(MIX8) }
		if (MODE == 2) { REP8(FMA8) REP8(FMA8) REP8(FMA8) REP8(FMA8) }
		if (MODE == 3) { REP8(ALU8) REP8(ALU8) REP8(ALU8) REP8(ALU8) }
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + l0 + l1 + l2 + l3 + l4 + l5 + l6 + l7;
	if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; }      // warp 0 of the block; all warps of a block share its SM
}

template <int MODE>
static double run(int warps_per_sm, int n_sm, uint32_t *out, long long *cyc)
{
	// ONE block of 32 * warps_per_sm threads per SM (the first version launched single-warp blocks and its rows from
	// 12 warps/SM on showed more than one instruction per clock: the timed blocks were not sharing their SM with the rest)
	const int iters = 2000, blocks = n_sm, threads = 32 * warps_per_sm;
	k<MODE><<<blocks, threads>>>(out, cyc, 10, 3u, 7u);
	k<MODE><<<blocks, threads>>>(out, cyc, iters, 3u, 7u);
	cudaDeviceSynchronize();
	long long h[256];
	cudaMemcpy(h, cyc, sizeof(long long) * (blocks < 256 ? blocks : 256), cudaMemcpyDeviceToHost);
	double worst = 0;
	for (int i = 0; i < (blocks < 256 ? blocks : 256); i++) { if ((double)h[i] > worst) { worst = (double)h[i]; } }
	// 256 instructions per iteration per warp; cycles per instruction PER SUB-PARTITION (4 per SM)
	return worst / ((double)iters * 256.0 * warps_per_sm / 4.0);
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	uint32_t *out; long long *cyc;
	cudaMalloc(&out, sizeof(uint32_t) * 1024 * p.multiProcessorCount);
	cudaMalloc(&cyc, sizeof(long long) * 32 * p.multiProcessorCount);
	printf("%s, %d SMs; cycles per warp instruction per sub-partition (1.0 = one issue per clock)\n", p.name, p.multiProcessorCount);
	printf("warps/SM   runs-of-32   alternating   IMAD-only   ALU-only\n");
	for (int w : {4, 8, 12, 16, 24, 32}) {
		printf("%8d   %10.3f   %11.3f   %9.3f   %8.3f\n", w, run<0>(w, p.multiProcessorCount, out, cyc), run<1>(w, p.multiProcessorCount, out, cyc),
		       run<2>(w, p.multiProcessorCount, out, cyc), run<3>(w, p.multiProcessorCount, out, cyc));
	}
	return 0;
}
