// op_microbench.cu -- latency and issue interval of single sm_100a instructions, per sub-partition.
//
// Question (round 2, fm_back_kernel): is IMAD.HI.U32 (the quotient of the de-emphasis step) a full-rate instruction?
// For each op: LAT = cycles per instruction of ONE dependent chain in one warp; then the cycles per warp instruction
// and sub-partition with 8 independent chains per thread and 1 / 2 / 4 / 8 warps per sub-partition (one block per SM,
// clock64 around the loop of warp 0; all warps of a block share the SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o op_microbench.bin op_microbench.cu && ./op_microbench.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { OP_IMAD, OP_IMADHI, OP_IADD3, OP_LOP3, OP_SHF, OP_PRMT, OP_FFMA, OP_FADD, OP_IDP, OP_IMADWIDE, OP_N };
static const char *NAMES[OP_N] = { "IMAD", "IMAD.HI.U32", "IADD3", "LOP3", "SHF", "PRMT", "FFMA", "FADD", "IDP.2A", "IMAD.WIDE" };

template <int OP>
__device__ __forceinline__ void step(uint32_t &x, uint32_t m, uint32_t a)
{
	if (OP == OP_IMAD) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(a)); }
	if (OP == OP_IMADHI) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(a)); }
	if (OP == OP_IADD3) { asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a)); }
	if (OP == OP_LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(m), "r"(a)); }
	if (OP == OP_SHF) { asm volatile("shf.l.wrap.b32 %0, %0, %0, %1;" : "+r"(x) : "r"(m)); }
	if (OP == OP_PRMT) { asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(a)); }
	if (OP == OP_FFMA) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(a)); }
	if (OP == OP_FADD) { asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(x) : "r"(a)); }
	if (OP == OP_IDP) { asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(a)); }
	if (OP == OP_IMADWIDE) {
		asm volatile("{ .reg .u64 t; mul.wide.u32 t, %0, %1; cvt.u32.u64 %0, t; shr.u64 t, t, 32; }" : "+r"(x) : "r"(m));
	}
}

template <int OP, int CHAINS>
__global__ void __launch_bounds__(1024) k(uint32_t *out, long long *cyc, int iters, uint32_t m, uint32_t a)
{
	uint32_t v[8];
#pragma unroll
	for (int j = 0; j < 8; j++) { v[j] = threadIdx.x + j; }
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int r = 0; r < 8; r++) {
#pragma unroll
			for (int j = 0; j < CHAINS; j++) { step<OP>(v[j], m, a); }
		}
	}
	long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) { s += v[j]; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; }
}

template <int OP, int CHAINS>
static double run(int warps_per_sm, int n_sm, uint32_t *out, long long *cyc)
{
	const int iters = 1000;
	k<OP, CHAINS><<<n_sm, 32 * warps_per_sm>>>(out, cyc, iters, 0x9e3779b1u, 12345u);
	cudaDeviceSynchronize();
	k<OP, CHAINS><<<n_sm, 32 * warps_per_sm>>>(out, cyc, iters, 0x9e3779b1u, 12345u);
	cudaDeviceSynchronize();
	long long h[256];
	cudaMemcpy(h, cyc, n_sm * sizeof(long long), cudaMemcpyDeviceToHost);
	double s = 0;
	for (int i = 0; i < n_sm; i++) { s += (double)h[i]; }
	const double insts_per_warp = (double)iters * 8 * CHAINS;
	const double warps_per_smsp = warps_per_sm / 4.0;
	// cycles per warp instruction and sub-partition (all warps advance together)
	return (s / n_sm) / (insts_per_warp * (warps_per_smsp < 1 ? 1 : warps_per_smsp));
}

template <int OP>
static void report(int n_sm, uint32_t *out, long long *cyc)
{
	const double lat = run<OP, 1>(1, n_sm, out, cyc);
	printf("%-12s latency %6.2f   interval at 1/2/4/8 warps per sub-partition (8 chains): %6.2f %6.2f %6.2f %6.2f\n", NAMES[OP], lat,
	       run<OP, 8>(4, n_sm, out, cyc), run<OP, 8>(8, n_sm, out, cyc), run<OP, 8>(16, n_sm, out, cyc), run<OP, 8>(32, n_sm, out, cyc));
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int n_sm = p.multiProcessorCount;
	uint32_t *out; long long *cyc;
	cudaMalloc(&out, (size_t)n_sm * 1024 * 4);
	cudaMalloc(&cyc, n_sm * sizeof(long long));
	printf("%s, %d SMs; cycles per warp instruction\n", p.name, n_sm);
	report<OP_IMAD>(n_sm, out, cyc); report<OP_IMADHI>(n_sm, out, cyc); report<OP_IMADWIDE>(n_sm, out, cyc);
	report<OP_IADD3>(n_sm, out, cyc); report<OP_LOP3>(n_sm, out, cyc); report<OP_SHF>(n_sm, out, cyc); report<OP_PRMT>(n_sm, out, cyc);
	report<OP_FFMA>(n_sm, out, cyc); report<OP_FADD>(n_sm, out, cyc); report<OP_IDP>(n_sm, out, cyc);
	return 0;
}
