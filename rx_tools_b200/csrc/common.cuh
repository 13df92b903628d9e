// common.cuh — shared helpers of librxb200.so (error plumbing, wrap-safe integer helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/rxb200.h"

namespace rxb {

void set_error(const char *fmt, ...);

#define RXB_CUDA(call)                                                                   \
	do {                                                                                 \
		cudaError_t e__ = (call);                                                        \
		if (e__ != cudaSuccess) {                                                        \
			rxb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),      \
			               __FILE__, __LINE__);                                          \
			return RXB200_ECUDA;                                                         \
		}                                                                                \
	} while (0)

// same, running `cleanup` first (create paths: nothing may leak when a CUDA call fails half way)
#define RXB_CUDA_OR(call, cleanup)                                                       \
	do {                                                                                 \
		cudaError_t e__ = (call);                                                        \
		if (e__ != cudaSuccess) {                                                        \
			rxb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),      \
			               __FILE__, __LINE__);                                          \
			cleanup;                                                                     \
			return RXB200_ECUDA;                                                         \
		}                                                                                \
	} while (0)

// ---- two's-complement helpers: the reference relies on x86 wrap-around for int overflow and
// on truncating stores to int16_t (SURVEY.md §7 hard part 5); unsigned arithmetic makes the
// wrap well-defined here.
__host__ __device__ __forceinline__ int wrap16(int v) { return (int)(int16_t)v; }
__host__ __device__ __forceinline__ int mul_w(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__host__ __device__ __forceinline__ int add_w(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__host__ __device__ __forceinline__ int sub_w(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__host__ __device__ __forceinline__ int neg_w(int a) { return (int)(0u - (unsigned)a); }
// C's truncating division; the two inputs that trap on x86 return a fixed value instead.
__host__ __device__ __forceinline__ int div_c(int n, int d)
{
	if (d == 0) { return 0; }
	if (d == -1) { return neg_w(n); }
	return n / d;
}

// (int16)(x/32767.0*128.0+0.4) for every int16 x, in integer form (src/rtl_fm.c:846; the exact
// identity is checked exhaustively against the oracle in tests/test_host_logic.py and on the GPU
// in tests/test_fm_gpu.py).  Result is in [-127, 128].
__host__ __device__ __forceinline__ int scale_cs16(int x)
{
	// t < 0  <=>  x <= -103  <=>  the real value is negative: floor + 1 == truncation toward zero
	int t = x * 32769 + 3355366;
	return (t >> 23) - (t >> 31);
}

#ifdef __CUDACC__
// C's truncating n / d for d > 0 when |n / d| is small (< 2^20): fp32 estimate (error < 1) and an
// exact integer remainder check.  Falls back to div_c otherwise.  Bit-exact with '/'.
__device__ __forceinline__ int div_small_quotient(int n, int d)
{
	if (d <= 0) { return div_c(n, d); }
	int q = __float2int_rz(__int2float_rn(n) * __frcp_rn(__int2float_rn(d)));
	int r = sub_w(n, mul_w(q, d));           // exact in two's complement whenever |q - n/d| <= 1
	if (n >= 0) {
		if (r < 0) { q--; } else if (r >= d) { q++; }
	} else {
		if (r > 0) { q++; } else if (r <= -d) { q--; }
	}
	return q;
}
#endif

}  // namespace rxb
