// fm_kernels.cu — rx_fm hot path on sm_100a: ONE fused kernel per stream batch covering
//   CS16->8-bit-range scale (src/rtl_fm.c:846) -> rotate16_90 (:309) -> low_pass (:351) |
//   fifth_order x P (:411) + generic_fir (:442) -> fm_demod (:584; std/fast/lut/ale) | am/usb/lsb/raw
//   -> deemph_filter (:667) -> low_pass_real (:389)
// with no intermediate buffer in HBM (algorithmic traffic: 4 B in + 2*rate_out2/rate_capture B out
// per complex sample).
//
// Parallel decomposition (DESIGN.md "rx_fm kernel"): the stream of every channel is cut into
// segments of S complex samples; one THREAD owns one segment and runs the reference's per-sample
// state machine over it with all carry state in registers.  To know the state at its segment
// start without waiting for its left neighbour it first replays `warm` samples before the segment:
//   * decimators / FIRs / discriminator have finite memory, so the replay makes them exact;
//   * deemph_filter is a rounding (non-linear) IIR: the replay runs it from BOTH extreme states
//     (-32768 and +32767).  The step map is monotone in the state, so the true state is bracketed,
//     and once the two trajectories meet the state is exact (SURVEY.md §7 hard part 2).
// Segments whose brackets have not met at the segment start (quiet input: the IIR has a dead
// zone) are flagged and recomputed serially from their neighbour's exact end state by
// fm_fixup_kernel, so the result is bit-exact for every input.
// Per-chunk semantics (rotation phase restart, fifth_order dropping the last sample of a chunk,
// first FM output of a chunk through atan2: SURVEY F7, F8) are reproduced literally: the chunk
// length is a kernel argument.
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <new>
#include <vector>
#include "common.cuh"

namespace rxb {

// ------------------------------------------------------------------------------ device config
struct FmDev {
	int mode, D, P, fir_on, atan_mode, out_scale, post_ds;
	int deemph, a, a_half, a_even;
	unsigned a_magic; int a_K, a_use_magic;   // floor((n)/a) == umulhi(n, a_magic) for n < 2^19
	int resample, fast, slow, lpr_div;
	int offset_tuning;
	int fir[6];
	const int *atan_lut;
};

struct FmCall {
	const int16_t *in;        // [n_ch][n] complex CS16
	int16_t *out;             // [n_ch][out_stride] int16
	long long n;              // complex samples per channel in this call
	long long out_stride;     // int16 per channel
	int chunk;                // complex samples per chunk
	int n_ch;
	int S;                    // segment length (complex)
	int warm;                 // replay length before a segment (complex)
	int dec_exact;            // decimated samples after which the replayed front end is exact
	int nseg;                 // segments per channel
	int state_words;
	const uint32_t *carry_in; // [n_ch][state_words]
	uint32_t *carry_out;      // [n_ch][state_words]
	uint32_t *seg_state;      // [n_ch*nseg][state_words]  end state of every segment
	int *seg_flags;           // [n_ch*nseg]  bit0 start exact, bit1 end exact
	int *fix_count;           // number of segments recomputed by the fix-up kernel
};

enum { ST_BOX_I = 0, ST_BOX_Q, ST_BOX_N, ST_PRE_I, ST_PRE_Q, ST_AVG_LO, ST_AVG_HI, ST_LPR_ACC,
       ST_LPR_PHASE, ST_DIRTY, ST_SQ_HITS, ST_ADC, ST_RDC_I, ST_RDC_Q, ST_RSV0, ST_RSV1, ST_HDR = 16 };

static inline int fm_state_words(int P) { return ST_HDR + 7 * P + 9; }

__device__ __forceinline__ uint32_t pack2(int i, int q) { return ((uint32_t)i & 0xffffu) | ((uint32_t)q << 16); }
__device__ __forceinline__ int lo16(uint32_t w) { return (int)(int16_t)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return (int)(int16_t)(w >> 16); }

// ------------------------------------------------------------------------------ per-thread state
template <int P>
struct FmState {
	int box_i, box_q, box_n;
	int wi[P > 0 ? P : 1][6], wq[P > 0 ? P : 1][6];   // fifth_order windows (a..f) per pass
	int pi[P > 0 ? P : 1], pq[P > 0 ? P : 1];         // odd-indexed sample waiting for its pair
	int di[9], dq[9];                                 // generic_fir history
	int pre_i, pre_q;
	int lo, hi, dirty;
	int lpr_acc, lpr_phase;
};

template <int P>
__device__ __forceinline__ void state_zero(FmState<P> &s)
{
	s.box_i = s.box_q = s.box_n = 0;
#pragma unroll
	for (int l = 0; l < (P > 0 ? P : 1); l++) {
#pragma unroll
		for (int j = 0; j < 6; j++) { s.wi[l][j] = 0; s.wq[l][j] = 0; }
		s.pi[l] = 0; s.pq[l] = 0;
	}
#pragma unroll
	for (int j = 0; j < 9; j++) { s.di[j] = 0; s.dq[j] = 0; }
	s.pre_i = s.pre_q = 0;
	s.lo = -32768; s.hi = 32767; s.dirty = 1;
	s.lpr_acc = 0; s.lpr_phase = 0;
}

template <int P>
__device__ __forceinline__ void state_load(FmState<P> &s, const uint32_t *g)
{
	s.box_i = (int)g[ST_BOX_I]; s.box_q = (int)g[ST_BOX_Q]; s.box_n = (int)g[ST_BOX_N];
	s.pre_i = (int)g[ST_PRE_I]; s.pre_q = (int)g[ST_PRE_Q];
	s.lo = (int)g[ST_AVG_LO]; s.hi = (int)g[ST_AVG_HI];
	s.lpr_acc = (int)g[ST_LPR_ACC]; s.lpr_phase = (int)g[ST_LPR_PHASE];
	s.dirty = (int)g[ST_DIRTY];
#pragma unroll
	for (int l = 0; l < P; l++) {
#pragma unroll
		for (int j = 0; j < 6; j++) { uint32_t w = g[ST_HDR + 7 * l + j]; s.wi[l][j] = lo16(w); s.wq[l][j] = hi16(w); }
		uint32_t w = g[ST_HDR + 7 * l + 6]; s.pi[l] = lo16(w); s.pq[l] = hi16(w);
	}
#pragma unroll
	for (int j = 0; j < 9; j++) { uint32_t w = g[ST_HDR + 7 * P + j]; s.di[j] = lo16(w); s.dq[j] = hi16(w); }
}

template <int P>
__device__ __forceinline__ void state_store(const FmState<P> &s, uint32_t *g)
{
	g[ST_BOX_I] = (uint32_t)s.box_i; g[ST_BOX_Q] = (uint32_t)s.box_q; g[ST_BOX_N] = (uint32_t)s.box_n;
	g[ST_PRE_I] = (uint32_t)s.pre_i; g[ST_PRE_Q] = (uint32_t)s.pre_q;
	g[ST_AVG_LO] = (uint32_t)s.lo; g[ST_AVG_HI] = (uint32_t)s.hi;
	g[ST_LPR_ACC] = (uint32_t)s.lpr_acc; g[ST_LPR_PHASE] = (uint32_t)s.lpr_phase;
	g[ST_DIRTY] = (uint32_t)s.dirty;
#pragma unroll
	for (int l = 0; l < P; l++) {
#pragma unroll
		for (int j = 0; j < 6; j++) { g[ST_HDR + 7 * l + j] = pack2(s.wi[l][j], s.wq[l][j]); }
		g[ST_HDR + 7 * l + 6] = pack2(s.pi[l], s.pq[l]);
	}
#pragma unroll
	for (int j = 0; j < 9; j++) { g[ST_HDR + 7 * P + j] = pack2(s.di[j], s.dq[j]); }
}

// ------------------------------------------------------------------------------ stages
// fifth_order cascade, one input sample at in-chunk index idx of pass L (src/rtl_fm.c:411-440,
// :765-768).  A pass emits at even input indices: index 0 of a chunk slides the window by one
// (a..e = hist[1..5], f = data[0]), every later even index by two (e = data[i-2], f = data[i]);
// an odd-indexed sample waits in p* and is lost if the chunk ends on it.
template <int L, int P>
__device__ __forceinline__ bool cascade_push(FmState<P> &s, int xi, int xq, unsigned idx, int &oi, int &oq)
{
	if constexpr (L >= P) {
		oi = xi; oq = xq;
		return true;
	} else {
		if (idx & 1u) { s.pi[L] = xi; s.pq[L] = xq; return false; }
		if (idx == 0u) {
#pragma unroll
			for (int j = 0; j < 5; j++) { s.wi[L][j] = s.wi[L][j + 1]; s.wq[L][j] = s.wq[L][j + 1]; }
		} else {
#pragma unroll
			for (int j = 0; j < 4; j++) { s.wi[L][j] = s.wi[L][j + 2]; s.wq[L][j] = s.wq[L][j + 2]; }
			s.wi[L][4] = s.pi[L]; s.wq[L][4] = s.pq[L];
		}
		s.wi[L][5] = xi; s.wq[L][5] = xq;
		int yi = wrap16((s.wi[L][0] + (s.wi[L][1] + s.wi[L][4]) * 5 + (s.wi[L][2] + s.wi[L][3]) * 10 + s.wi[L][5]) >> 4);
		int yq = wrap16((s.wq[L][0] + (s.wq[L][1] + s.wq[L][4]) * 5 + (s.wq[L][2] + s.wq[L][3]) * 10 + s.wq[L][5]) >> 4);
		return cascade_push<L + 1, P>(s, yi, yq, idx >> 1, oi, oq);
	}
}

// generic_fir (src/rtl_fm.c:442-465): output from the previous nine samples, then shift in x.
__device__ __forceinline__ int droop9(int (&h)[9], const int (&c)[6], int x)
{
	int acc = mul_w(h[0] + h[8], c[1]);
	acc = add_w(acc, mul_w(h[1] + h[7], c[2]));
	acc = add_w(acc, mul_w(h[2] + h[6], c[3]));
	acc = add_w(acc, mul_w(h[3] + h[5], c[4]));
	acc = add_w(acc, mul_w(h[4], c[5]));
#pragma unroll
	for (int j = 0; j < 8; j++) { h[j] = h[j + 1]; }
	h[8] = x;
	return wrap16(acc >> 15);
}

// polar_discriminant (src/rtl_fm.c:476-483); note 3.14159.
__device__ __noinline__ int disc_std(int cr, int cj)
{
	double angle = atan2((double)cj, (double)cr);
	return (int)(angle / 3.14159 * (double)(1 << 14));
}

// fast_atan2 (src/rtl_fm.c:485-506), int32 wrap-around preserved.
__device__ __forceinline__ int fast_atan2_i(int y, int x)
{
	const int q1 = 1 << 12, q3 = 3 * (1 << 12);
	if (x == 0 && y == 0) { return 0; }
	int ya = y < 0 ? neg_w(y) : y;
	int ang;
	if (x >= 0) { ang = sub_w(q1, div_c(mul_w(q1, sub_w(x, ya)), add_w(x, ya))); }
	else        { ang = sub_w(q3, div_c(mul_w(q1, add_w(x, ya)), sub_w(ya, x))); }
	return y < 0 ? neg_w(ang) : ang;
}

// polar_disc_lut (src/rtl_fm.c:528-564)
__device__ __forceinline__ int disc_lut(const int *__restrict__ lut, int cr, int cj)
{
	if (cr == 0 || cj == 0) {
		if (cr == 0 && cj == 0) { return 0; }
		if (cr == 0) { return cj > 0 ? (1 << 13) : -(1 << 13); }
		return cr > 0 ? 0 : (1 << 14);
	}
	int x = div_c((int)((unsigned)cj << 8), cr);
	int xa = x < 0 ? neg_w(x) : x;
	if (xa >= 131072 || xa < 0) { return cj > 0 ? (1 << 13) : -(1 << 13); }
	if (x > 0) { return cj > 0 ? __ldg(lut + x) : __ldg(lut + x) - (1 << 14); }
	return cj > 0 ? (1 << 14) - __ldg(lut + xa) : -__ldg(lut + xa);
}

// esbensen (src/rtl_fm.c:566-582)
__device__ __forceinline__ int disc_ale(int ar, int aj, int br, int bj)
{
	int dr = mul_w(sub_w(br, ar), 2), dj = mul_w(sub_w(bj, aj), 2);
	int cj = sub_w(mul_w(bj, dr), mul_w(br, dj));
	return div_c(mul_w(2608, cj), add_w(add_w(mul_w(ar, ar), mul_w(aj, aj)), 1));
}

// One deemph_filter step (src/rtl_fm.c:673-680): avg += trunc((d +- a/2)/a).
//   d > 0 : trunc((d + h)/a)          = floor((d + h)/a)
//   d <= 0: trunc((d - h)/a)          = floor((d - h + a - 1)/a)       (h = a/2)
// i.e. floor((d + c)/a) with c = h for odd a, and c = h - (d <= 0) for even a.  With a bias of
// K*a the numerator is non-negative and the floor is one umulhi by a host-verified reciprocal.
__device__ __forceinline__ int deemph_step(const FmDev &c, int avg, int x)
{
	int d = x - avg;
	if (c.a_use_magic) {
		int n = d + c.a_half + c.a_K * c.a;
		if (c.a_even) { n -= (d <= 0) ? 1 : 0; }
		return avg + (int)__umulhi((unsigned)n, c.a_magic) - c.a_K;
	}
	return avg + ((d > 0) ? (d + c.a_half) / c.a : (d - c.a_half) / c.a);
}

// ------------------------------------------------------------------------------ segment runner
template <int P>
struct SegCtx {
	FmState<P> st;
	long long out_idx;      // next output slot (int16 index within the channel's output)
	int nd;                 // decimated samples produced since the replay start
	int first_in_chunk;     // next decimated sample is the first of its chunk (F8)
	bool exact_start;       // state came from an exact carry, no bracket needed
	bool emit;              // outputs are owned (t >= s0)
};

// Everything after the decimator for one decimated sample (di,dq).
template <int P>
__device__ __forceinline__ void back_end(const FmDev &c, const FmCall &k, SegCtx<P> &x, int16_t *__restrict__ out,
                                         int di, int dq)
{
	FmState<P> &s = x.st;
	if (c.fir_on) {
		di = droop9(s.di, c.fir, di);
		dq = droop9(s.dq, c.fir, dq);
	}
	const bool pcm_valid = x.exact_start || x.nd >= k.dec_exact;
	if (!x.exact_start && x.nd == k.dec_exact) { s.lo = -32768; s.hi = 32767; s.dirty = 1; }
	x.nd++;
	int pcm;
	if (c.mode == RXB200_MODE_FM) {
		int br = s.pre_i, bj = s.pre_q;
		int cr = add_w(mul_w(di, br), mul_w(dq, bj));       // x[n] * conj(x[n-1]) (src/rtl_fm.c:470-474)
		int cj = sub_w(mul_w(dq, br), mul_w(di, bj));
		if (x.first_in_chunk || c.atan_mode == RXB200_ATAN_STD) { pcm = disc_std(cr, cj); }
		else if (c.atan_mode == RXB200_ATAN_FAST) { pcm = fast_atan2_i(cj, cr); }
		else if (c.atan_mode == RXB200_ATAN_LUT) { pcm = disc_lut(c.atan_lut, cr, cj); }
		else { pcm = disc_ale(di, dq, br, bj); }
		pcm = wrap16(pcm);
		s.pre_i = di; s.pre_q = dq;
	} else if (c.mode == RXB200_MODE_AM) {
		int e = add_w(mul_w(di, di), mul_w(dq, dq));
		pcm = wrap16(mul_w(wrap16((int)sqrt((double)e)), c.out_scale));
	} else if (c.mode == RXB200_MODE_USB) {
		pcm = wrap16(mul_w(wrap16(di + dq), c.out_scale));
	} else if (c.mode == RXB200_MODE_LSB) {
		pcm = wrap16(mul_w(wrap16(di - dq), c.out_scale));
	} else {   // raw: lowpassed copied out, nothing after (src/rtl_fm.c:658-665, :809-811)
		if (x.emit) { out[x.out_idx] = (int16_t)di; out[x.out_idx + 1] = (int16_t)dq; }
		x.out_idx += 2;
		x.first_in_chunk = 0;
		s.dirty = pcm_valid ? 0 : 1;
		return;
	}
	x.first_in_chunk = 0;
	bool inexact = !pcm_valid;
	if (c.deemph) {
		bool same = (s.lo == s.hi);
		s.lo = deemph_step(c, s.lo, pcm);
		s.hi = same ? s.lo : deemph_step(c, s.hi, pcm);
		pcm = wrap16(s.lo);
		inexact = inexact || (s.lo != s.hi);
	}
	if (c.resample) {     // low_pass_real (src/rtl_fm.c:389-409)
		if (inexact) { s.dirty = 1; }
		s.lpr_acc = add_w(s.lpr_acc, pcm);
		s.lpr_phase += c.slow;
		if (s.lpr_phase >= c.fast) {
			if (x.emit) { out[x.out_idx] = (int16_t)(s.lpr_acc / c.lpr_div); }
			x.out_idx++;
			s.lpr_phase -= c.fast;
			s.lpr_acc = 0;
			s.dirty = 0;
		}
	} else {
		s.dirty = inexact ? 1 : 0;
		if (x.emit) { out[x.out_idx] = (int16_t)pcm; }
		x.out_idx++;
	}
}

// One input sample at in-chunk index u.
template <int P>
__device__ __forceinline__ void front_end(const FmDev &c, const FmCall &k, SegCtx<P> &x, int16_t *__restrict__ out,
                                          uint32_t w, unsigned u)
{
	int xi = scale_cs16(lo16(w));
	int xq = scale_cs16(hi16(w));
	if (!c.offset_tuning) {     // rotate16_90: sample n of the chunk times j^n (src/rtl_fm.c:309-327)
		int ri, rq;
		switch (u & 3u) {
		case 1: ri = -xq; rq = xi; break;
		case 2: ri = -xi; rq = -xq; break;
		case 3: ri = xq; rq = -xi; break;
		default: ri = xi; rq = xq; break;
		}
		xi = ri; xq = rq;
	}
	FmState<P> &s = x.st;
	if constexpr (P == 0) {    // low_pass boxcar (src/rtl_fm.c:351-371)
		s.box_i += xi; s.box_q += xq;
		if (++s.box_n >= c.D) {
			int di = wrap16(s.box_i), dq = wrap16(s.box_q);
			s.box_i = 0; s.box_q = 0; s.box_n = 0;
			back_end<P>(c, k, x, out, di, dq);
		}
	} else {
		int di, dq;
		if (cascade_push<0, P>(s, xi, xq, u, di, dq)) {
			if ((u >> P) == 0u) { x.first_in_chunk = 1; }
			back_end<P>(c, k, x, out, di, dq);
		}
	}
}

// Number of decimated samples the reference has produced after t input samples of this call.
__device__ __forceinline__ long long dec_before(const FmDev &c, int P, long long t, int box_n0)
{
	if (P > 0) { return t >> P; }
	return (t + box_n0) / c.D;
}

// Output slot of the first sample produced at/after decimated index m.
__device__ __forceinline__ long long out_before(const FmDev &c, long long m, int phase0)
{
	if (c.mode == RXB200_MODE_RAW) { return 2 * m; }
	if (!c.resample) { return m; }
	return ((long long)phase0 + m * (long long)c.slow) / (long long)c.fast;
}

// Runs [w0, s1) of channel ch; outputs are stored for t >= s0.  If from_exact the state in x.st is
// the exact state at w0 (carry or a neighbour's end state).  Returns flags: bit0 = state was exact
// when the owned part began, bit1 = state exact at the end.
template <int P>
__device__ int run_segment(const FmDev &c, const FmCall &k, SegCtx<P> &x, int ch, long long w0, long long s0,
                           long long s1, int box_n0, int phase0, bool from_exact)
{
	const uint4 *__restrict__ in4 = reinterpret_cast<const uint4 *>(k.in + 2 * (size_t)ch * (size_t)k.n);
	int16_t *__restrict__ out = k.out + (size_t)ch * (size_t)k.out_stride;
	FmState<P> &s = x.st;
	x.exact_start = from_exact;
	x.nd = 0;
	// position bookkeeping at w0
	unsigned u = (unsigned)(w0 % k.chunk);
	long long chunk_base = w0 - u;
	long long m0 = dec_before(c, P, w0, box_n0);
	if (!from_exact) {
		if (P == 0) { s.box_n = (int)((w0 + box_n0) % c.D); }
		if (c.resample) { s.lpr_phase = (int)(((long long)phase0 + m0 * (long long)c.slow) % (long long)c.fast); }
	}
	x.out_idx = out_before(c, m0, phase0);
	// boxcar: is the next decimated sample the first of its chunk?
	if (P == 0) { x.first_in_chunk = (dec_before(c, P, chunk_base, box_n0) == m0) ? 1 : 0; }
	else { x.first_in_chunk = 0; }
	int flags = from_exact ? 1 : 0;
	x.emit = false;
	for (long long t = w0; t < s1; t += 8) {
		if (t == s0) {
			x.emit = true;
			if (!from_exact) {
				bool ok = (x.nd >= k.dec_exact) && !s.dirty && (!c.deemph || s.lo == s.hi);
				flags = ok ? 1 : 0;
			}
		}
		uint4 a = __ldg(in4 + (t >> 2));
		uint4 b = __ldg(in4 + (t >> 2) + 1);
		if (P == 0 && u == 0u) { x.first_in_chunk = 1; }
		const unsigned ub = u;     // multiple of 8
		front_end<P>(c, k, x, out, a.x, ub | 0u);
		front_end<P>(c, k, x, out, a.y, ub | 1u);
		front_end<P>(c, k, x, out, a.z, ub | 2u);
		front_end<P>(c, k, x, out, a.w, ub | 3u);
		front_end<P>(c, k, x, out, b.x, ub | 4u);
		front_end<P>(c, k, x, out, b.y, ub | 5u);
		front_end<P>(c, k, x, out, b.z, ub | 6u);
		front_end<P>(c, k, x, out, b.w, ub | 7u);
		u += 8u;
		if (u >= (unsigned)k.chunk) { u = 0u; }
	}
	bool end_ok = from_exact || ((flags & 1) != 0) ||
	              ((x.nd >= k.dec_exact) && !s.dirty && (!c.deemph || s.lo == s.hi));
	if (end_ok) { flags |= 2; }
	return flags;
}

template <int P>
__global__ void __launch_bounds__(128) fm_main_kernel(const FmDev c, const FmCall k)
{
	long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (long long)k.n_ch * k.nseg) { return; }
	int ch = (int)(gid / k.nseg);
	int seg = (int)(gid % k.nseg);
	long long s0 = (long long)seg * k.S;
	long long s1 = s0 + k.S < k.n ? s0 + k.S : k.n;
	long long w0 = s0 - k.warm;
	const uint32_t *carry = k.carry_in + (size_t)ch * k.state_words;
	const int box_n0 = (int)carry[ST_BOX_N];
	const int phase0 = (int)carry[ST_LPR_PHASE];
	SegCtx<P> x;
	bool from_exact = (w0 <= 0);
	if (from_exact) { w0 = 0; state_load<P>(x.st, carry); }
	else { state_zero<P>(x.st); }
	int flags = run_segment<P>(c, k, x, ch, w0, s0, s1, box_n0, phase0, from_exact);
	uint32_t *dst = k.seg_state + (size_t)gid * k.state_words;
	state_store<P>(x.st, dst);
	k.seg_flags[gid] = flags;
	if (seg == k.nseg - 1) { state_store<P>(x.st, k.carry_out + (size_t)ch * k.state_words); }
}

// Serial repair of segments whose de-emphasis bracket had not closed at their start.  Thread g
// acts only if segment g failed and segment g-1 ended exact; it then walks right until it has
// re-run a segment whose ORIGINAL end state was already exact (its right neighbour is either
// fine or has its own repair thread).
template <int P>
__global__ void __launch_bounds__(128) fm_fixup_kernel(const FmDev c, const FmCall k)
{
	long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (long long)k.n_ch * k.nseg) { return; }
	int ch = (int)(gid / k.nseg);
	int seg = (int)(gid % k.nseg);
	if (seg == 0) { return; }
	if (k.seg_flags[gid] & 1) { return; }
	if (!(k.seg_flags[gid - 1] & 2)) { return; }
	const uint32_t *carry = k.carry_in + (size_t)ch * k.state_words;
	const int box_n0 = (int)carry[ST_BOX_N];
	const int phase0 = (int)carry[ST_LPR_PHASE];
	SegCtx<P> x;
	for (int sgi = seg; sgi < k.nseg; sgi++) {
		long long g = (long long)ch * k.nseg + sgi;
		int orig = k.seg_flags[g];
		if (sgi != seg && (orig & 1)) { break; }
		long long s0 = (long long)sgi * k.S;
		long long s1 = s0 + k.S < k.n ? s0 + k.S : k.n;
		state_load<P>(x.st, k.seg_state + (size_t)(g - 1) * k.state_words);
		run_segment<P>(c, k, x, ch, s0, s0, s1, box_n0, phase0, true);
		state_store<P>(x.st, k.seg_state + (size_t)g * k.state_words);
		if (sgi == k.nseg - 1) { state_store<P>(x.st, k.carry_out + (size_t)ch * k.state_words); }
		atomicAdd(k.fix_count, 1);
		if (orig & 2) { break; }
	}
}

typedef void (*fm_kernel_fn)(const FmDev, const FmCall);
template <int P> struct KernelPair { static fm_kernel_fn main_k() { return fm_main_kernel<P>; } static fm_kernel_fn fix_k() { return fm_fixup_kernel<P>; } };

static void pick_kernels(int P, fm_kernel_fn *mk, fm_kernel_fn *fk)
{
	switch (P) {
#define RXB_CASE(N) case N: *mk = KernelPair<N>::main_k(); *fk = KernelPair<N>::fix_k(); break;
	RXB_CASE(0) RXB_CASE(1) RXB_CASE(2) RXB_CASE(3) RXB_CASE(4) RXB_CASE(5)
	RXB_CASE(6) RXB_CASE(7) RXB_CASE(8) RXB_CASE(9) RXB_CASE(10)
#undef RXB_CASE
	default: *mk = nullptr; *fk = nullptr;
	}
}

}  // namespace rxb

// ================================================================================ host side
using namespace rxb;

static const int k_droop9_host[11][10] = {
	// droop-compensation FIR rows (cic_9_tables, src/rtl_fm.c:287-300): {taps, c1..c9} x 2^15
	{0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
	{9, -156,  -97, 2798, -15489, 61019, -15489, 2798,  -97, -156},
	{9, -128, -568, 5593, -24125, 74126, -24125, 5593, -568, -128},
	{9, -129, -639, 6187, -26281, 77511, -26281, 6187, -639, -129},
	{9, -122, -612, 6082, -26353, 77818, -26353, 6082, -612, -122},
	{9, -120, -602, 6015, -26269, 77757, -26269, 6015, -602, -120},
	{9, -120, -582, 5951, -26128, 77542, -26128, 5951, -582, -120},
	{9, -119, -580, 5931, -26094, 77505, -26094, 5931, -580, -119},
	{9, -119, -578, 5921, -26077, 77484, -26077, 5921, -578, -119},
	{9, -119, -577, 5917, -26067, 77473, -26067, 5917, -577, -119},
	{9, -199, -362, 5303, -25505, 77489, -25505, 5303, -362, -199},
};

struct rxb200_fm {
	rxb200_fm_params p;
	int device;
	int n_channels;
	FmDev dev;
	int state_words;
	cudaStream_t stream;
	uint32_t *d_carry[2];
	int cur;                       // which carry buffer holds the current state
	uint32_t *d_seg_state; size_t seg_state_cap;   // words
	int *d_seg_flags; size_t seg_flags_cap;
	int *d_fix_count;
	int *d_atan_lut;
	int16_t *d_in; size_t d_in_cap;                // int16 elements
	int16_t *d_out; size_t d_out_cap;
	// host mirror of the closed-form counters
	int h_box_n;                   // demod.prev_index
	int h_lpr_phase;               // demod.prev_lpr_index
	int tune_seg, tune_warm;
	rxb200_fm_stats stats;
	fm_kernel_fn main_k, fix_k;
	cudaEvent_t ev0, ev1;
};

static int fm_validate(const rxb200_fm_params *p)
{
	if (p->mode < RXB200_MODE_FM || p->mode > RXB200_MODE_RAW) { set_error("mode %d", p->mode); return RXB200_EINVAL; }
	if (p->downsample_passes < 0 || p->downsample_passes > 10) { set_error("downsample_passes %d", p->downsample_passes); return RXB200_EINVAL; }
	if (p->downsample_passes == 0 && (p->downsample < 1 || p->downsample > 4096)) { set_error("downsample %d", p->downsample); return RXB200_EINVAL; }
	if (p->custom_atan < 0 || p->custom_atan > 3) { set_error("custom_atan %d", p->custom_atan); return RXB200_EINVAL; }
	if (p->deemph && p->deemph_a < 1) { set_error("deemph_a %d", p->deemph_a); return RXB200_EINVAL; }
	if (p->rate_out2 > 0 && (p->rate_out < p->rate_out2 || p->rate_out <= 0)) {
		set_error("low_pass_real needs rate_out >= rate_out2 > 0 (the reference divides by rate_out/rate_out2)");
		return RXB200_EINVAL;
	}
	if (p->post_downsample > 1 || p->squelch_level || p->dc_block_audio || p->dc_block_raw) {
		set_error("post_downsample / squelch / dc blocks: per-chunk reductions not implemented yet");
		return RXB200_EUNSUPPORTED;
	}
	return RXB200_OK;
}

static void fm_fill_dev(rxb200_fm *h)
{
	const rxb200_fm_params &p = h->p;
	FmDev &d = h->dev;
	memset(&d, 0, sizeof d);
	d.mode = p.mode; d.P = p.downsample_passes; d.D = p.downsample_passes ? (1 << p.downsample_passes) : p.downsample;
	d.fir_on = (p.downsample_passes > 0 && p.comp_fir_size == 9) ? 1 : 0;
	d.atan_mode = p.custom_atan; d.out_scale = p.output_scale; d.post_ds = p.post_downsample;
	d.deemph = (p.deemph && p.mode != RXB200_MODE_RAW) ? 1 : 0;
	d.a = p.deemph ? p.deemph_a : 1; d.a_half = d.a / 2; d.a_even = (d.a % 2 == 0) ? 1 : 0;
	// reciprocal for floor(n/a), n in [0, 2^19): verified exhaustively, else fall back to '/'
	d.a_use_magic = 0;
	if (d.a >= 1 && d.a < 16384) {
		unsigned magic = (unsigned)(0x100000000ULL / (unsigned)d.a) + 1u;
		int K = (65536 + 32768 + d.a) / d.a + 1;
		bool ok = true;
		unsigned nmax = (unsigned)(K * d.a + 65536 + 32768 + d.a);
		for (unsigned n = 0; n <= nmax && ok; n++) {
			if ((unsigned)(((unsigned long long)n * magic) >> 32) != n / (unsigned)d.a) { ok = false; }
		}
		if (ok) { d.a_use_magic = 1; d.a_magic = magic; d.a_K = K; }
	}
	d.resample = (p.rate_out2 > 0 && p.mode != RXB200_MODE_RAW) ? 1 : 0;
	d.fast = p.rate_out; d.slow = p.rate_out2; d.lpr_div = d.resample ? (p.rate_out / p.rate_out2) : 1;
	d.offset_tuning = p.offset_tuning;
	for (int j = 0; j < 6; j++) { d.fir[j] = k_droop9_host[d.P][j]; }
	d.atan_lut = h->d_atan_lut;
}

extern "C" int rxb200_fm_create(const rxb200_fm_params *params, int device, int n_channels, rxb200_fm **out)
{
	if (!params || !out || n_channels < 1) { set_error("null argument"); return RXB200_EINVAL; }
	*out = nullptr;
	int rc = fm_validate(params);
	if (rc != RXB200_OK) { return rc; }
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device: librxb200 has no CPU fallback"); return RXB200_ENODEV; }
	if (device < 0 || device >= ndev) { set_error("device %d out of range (%d)", device, ndev); return RXB200_ENODEV; }
	RXB_CUDA(cudaSetDevice(device));
	rxb200_fm *h = new (std::nothrow) rxb200_fm();
	if (!h) { return RXB200_ENOMEM; }
	memset(h, 0, sizeof *h);
	h->p = *params; h->device = device; h->n_channels = n_channels;
	h->state_words = fm_state_words(params->downsample_passes);
	pick_kernels(params->downsample_passes, &h->main_k, &h->fix_k);
	RXB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	RXB_CUDA(cudaEventCreate(&h->ev0));
	RXB_CUDA(cudaEventCreate(&h->ev1));
	size_t cbytes = (size_t)n_channels * h->state_words * sizeof(uint32_t);
	RXB_CUDA(cudaMalloc(&h->d_carry[0], cbytes));
	RXB_CUDA(cudaMalloc(&h->d_carry[1], cbytes));
	RXB_CUDA(cudaMalloc(&h->d_fix_count, sizeof(int)));
	if (params->custom_atan == RXB200_ATAN_LUT && params->mode == RXB200_MODE_FM) {
		// atan_lut_init (src/rtl_fm.c:515-526): host libm, uploaded once
		std::vector<int> lut(131072);
		for (int i = 0; i < 131072; i++) { lut[i] = (int)(atan((double)i / (double)(1 << 8)) / 3.14159 * (double)(1 << 14)); }
		RXB_CUDA(cudaMalloc(&h->d_atan_lut, lut.size() * sizeof(int)));
		RXB_CUDA(cudaMemcpy(h->d_atan_lut, lut.data(), lut.size() * sizeof(int), cudaMemcpyHostToDevice));
	}
	fm_fill_dev(h);
	*out = h;
	return rxb200_fm_reset(h);
}

extern "C" int rxb200_fm_reset(rxb200_fm *h)
{
	if (!h) { return RXB200_EINVAL; }
	RXB_CUDA(cudaSetDevice(h->device));
	size_t cbytes = (size_t)h->n_channels * h->state_words * sizeof(uint32_t);
	RXB_CUDA(cudaMemsetAsync(h->d_carry[0], 0, cbytes, h->stream));
	RXB_CUDA(cudaMemsetAsync(h->d_carry[1], 0, cbytes, h->stream));
	// squelch_hits starts at 11 (demod_init, src/rtl_fm.c:1091)
	std::vector<uint32_t> init((size_t)h->n_channels * h->state_words, 0u);
	for (int c = 0; c < h->n_channels; c++) { init[(size_t)c * h->state_words + ST_SQ_HITS] = 11u; }
	RXB_CUDA(cudaMemcpyAsync(h->d_carry[0], init.data(), cbytes, cudaMemcpyHostToDevice, h->stream));
	RXB_CUDA(cudaStreamSynchronize(h->stream));
	h->cur = 0; h->h_box_n = 0; h->h_lpr_phase = 0;
	return RXB200_OK;
}

extern "C" void rxb200_fm_destroy(rxb200_fm *h)
{
	if (!h) { return; }
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	cudaFree(h->d_carry[0]); cudaFree(h->d_carry[1]); cudaFree(h->d_seg_state); cudaFree(h->d_seg_flags);
	cudaFree(h->d_fix_count); cudaFree(h->d_atan_lut); cudaFree(h->d_in); cudaFree(h->d_out);
	cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
	cudaStreamDestroy(h->stream);
	delete h;
}

extern "C" int rxb200_fm_kernel_ms(rxb200_fm *h, float *ms)
{
	if (!h || !ms) { return RXB200_EINVAL; }
	RXB_CUDA(cudaSetDevice(h->device));
	RXB_CUDA(cudaEventSynchronize(h->ev1));
	RXB_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
	return RXB200_OK;
}

// closed-form per-chunk result_len; advances the host mirrors when commit is set
static size_t fm_count_outputs(rxb200_fm *h, size_t n_int16, size_t chunk_int16, int *chunk_result_len, bool commit)
{
	const rxb200_fm_params &p = h->p;
	long long box_n = h->h_box_n, phase = h->h_lpr_phase;
	size_t total = 0, pos = 0, c = 0;
	while (pos < n_int16) {
		size_t len16 = n_int16 - pos < chunk_int16 ? n_int16 - pos : chunk_int16;
		long long L = (long long)(len16 / 2), dec;
		if (p.downsample_passes) { dec = L >> p.downsample_passes; }
		else { dec = (box_n + L) / p.downsample; box_n = (box_n + L) % p.downsample; }
		long long res;
		if (p.mode == RXB200_MODE_RAW) { res = 2 * dec; }
		else if (p.rate_out2 > 0) {
			long long tot = phase + dec * (long long)p.rate_out2;
			res = tot / p.rate_out; phase = tot % p.rate_out;
		} else { res = dec; }
		if (chunk_result_len) { chunk_result_len[c] = (int)res; }
		total += (size_t)res; pos += len16; c++;
	}
	if (commit) { h->h_box_n = (int)box_n; h->h_lpr_phase = (int)phase; }
	return total;
}

extern "C" size_t rxb200_fm_max_output(const rxb200_fm *h, size_t n_int16, size_t chunk_int16)
{
	if (!h || chunk_int16 == 0) { return 0; }
	(void)chunk_int16;
	const rxb200_fm_params &p = h->p;
	size_t L = n_int16 / 2;
	size_t D = p.downsample_passes ? ((size_t)1 << p.downsample_passes) : (size_t)p.downsample;
	size_t dec = L / D + 2;
	if (p.mode == RXB200_MODE_RAW) { return 2 * dec; }
	if (p.rate_out2 > 0) { return (size_t)(((unsigned long long)dec * (unsigned)p.rate_out2) / (unsigned)p.rate_out) + 2; }
	return dec;
}

static int fm_check_shape(const rxb200_fm *h, size_t n_int16, size_t chunk_int16)
{
	size_t g16 = 16;                                  // 8 complex per vector step
	size_t p16 = (size_t)2 << h->p.downsample_passes; // 2^P complex
	if (p16 > g16) { g16 = p16; }
	if (chunk_int16 == 0 || chunk_int16 > 262144 || chunk_int16 % g16 != 0) {
		set_error("chunk_int16=%zu must be a multiple of %zu and <= 262144", chunk_int16, g16);
		return RXB200_EUNSUPPORTED;
	}
	if (n_int16 % g16 != 0) {
		set_error("n_int16=%zu must be a multiple of %zu (last chunk included)", n_int16, g16);
		return RXB200_EUNSUPPORTED;
	}
	return RXB200_OK;
}

static int fm_launch(rxb200_fm *h, const int16_t *d_in, size_t n_int16, size_t chunk_int16, int16_t *d_out,
                     size_t out_stride)
{
	const rxb200_fm_params &p = h->p;
	const long long n = (long long)(n_int16 / 2);
	const int P = p.downsample_passes;
	const long long Dtot = P ? (1LL << P) : p.downsample;
	const long long G = (1LL << P) > 8 ? (1LL << P) : 8;
	// replay length in decimated samples: front-end flush, de-emphasis bracket, one resampler group
	int dec_exact = P ? 26 : 3;
	long long wd = 0;
	if (h->dev.deemph) { wd = h->tune_warm > 0 ? h->tune_warm : 16LL * p.deemph_a + 64; }
	long long warm_dec = dec_exact + wd + (h->dev.resample ? (p.rate_out / p.rate_out2 + 2) : 0) + 2;
	long long warm = ((warm_dec * Dtot + G - 1) / G) * G;
	long long S;
	if (h->tune_seg > 0) { S = h->tune_seg; }
	else {
		const char *e = getenv("RXB200_FM_SEG");
		S = e ? atoll(e) : 0;
		if (S <= 0) {
			long long want = (n * h->n_channels) / (148LL * 384LL);
			S = want > 3 * warm ? want : 3 * warm;
		}
	}
	S = ((S + G - 1) / G) * G;
	if (S < G) { S = G; }
	long long nseg = (n + S - 1) / S;
	if (nseg < 1) { nseg = 1; }
	size_t total_seg = (size_t)nseg * h->n_channels;
	size_t need_words = total_seg * h->state_words;
	if (need_words > h->seg_state_cap) {
		cudaFree(h->d_seg_state); h->d_seg_state = nullptr; h->seg_state_cap = 0;
		RXB_CUDA(cudaMalloc(&h->d_seg_state, need_words * sizeof(uint32_t)));
		h->seg_state_cap = need_words;
	}
	if (total_seg > h->seg_flags_cap) {
		cudaFree(h->d_seg_flags); h->d_seg_flags = nullptr; h->seg_flags_cap = 0;
		RXB_CUDA(cudaMalloc(&h->d_seg_flags, total_seg * sizeof(int)));
		h->seg_flags_cap = total_seg;
	}
	FmCall k;
	k.in = d_in; k.out = d_out; k.n = n; k.out_stride = (long long)out_stride; k.chunk = (int)(chunk_int16 / 2);
	k.n_ch = h->n_channels; k.S = (int)S; k.warm = (int)warm; k.dec_exact = dec_exact; k.nseg = (int)nseg;
	k.state_words = h->state_words; k.carry_in = h->d_carry[h->cur]; k.carry_out = h->d_carry[h->cur ^ 1];
	k.seg_state = h->d_seg_state; k.seg_flags = h->d_seg_flags; k.fix_count = h->d_fix_count;
	RXB_CUDA(cudaMemsetAsync(h->d_fix_count, 0, sizeof(int), h->stream));
	unsigned blocks = (unsigned)((total_seg + 127) / 128);
	RXB_CUDA(cudaEventRecord(h->ev0, h->stream));
	h->main_k<<<blocks, 128, 0, h->stream>>>(h->dev, k);
	RXB_CUDA(cudaGetLastError());
	RXB_CUDA(cudaEventRecord(h->ev1, h->stream));
	int launches = 1;
	if (nseg > 1) {
		h->fix_k<<<blocks, 128, 0, h->stream>>>(h->dev, k);
		RXB_CUDA(cudaGetLastError());
		launches++;
	}
	h->cur ^= 1;
	h->stats.launches = launches; h->stats.segments = (int)total_seg; h->stats.segment_len = (int)S;
	h->stats.warmup_len = (int)warm; h->stats.fixup_segments = -1;
	return RXB200_OK;
}

extern "C" int rxb200_fm_process_device(rxb200_fm *h, const int16_t *d_cs16, size_t n_int16, size_t chunk_int16,
                                        int16_t *d_pcm, size_t pcm_stride, size_t *n_pcm, int sync)
{
	if (!h || !d_cs16 || !d_pcm) { set_error("null argument"); return RXB200_EINVAL; }
	if (((uintptr_t)d_cs16 & 15u) != 0) { set_error("d_cs16 must be 16-byte aligned"); return RXB200_EINVAL; }
	int rc = fm_check_shape(h, n_int16, chunk_int16);
	if (rc != RXB200_OK) { return rc; }
	RXB_CUDA(cudaSetDevice(h->device));
	if (n_int16 == 0) { if (n_pcm) { *n_pcm = 0; } return RXB200_OK; }
	size_t total = fm_count_outputs(h, n_int16, chunk_int16, nullptr, false);
	if (total > pcm_stride) { set_error("pcm_stride %zu < %zu outputs", pcm_stride, total); return RXB200_ECAPACITY; }
	rc = fm_launch(h, d_cs16, n_int16, chunk_int16, d_pcm, pcm_stride);
	if (rc != RXB200_OK) { return rc; }
	fm_count_outputs(h, n_int16, chunk_int16, nullptr, true);
	if (n_pcm) { *n_pcm = total; }
	if (sync) {
		RXB_CUDA(cudaStreamSynchronize(h->stream));
		RXB_CUDA(cudaMemcpy(&h->stats.fixup_segments, h->d_fix_count, sizeof(int), cudaMemcpyDeviceToHost));
	}
	return RXB200_OK;
}

extern "C" int rxb200_fm_process(rxb200_fm *h, const int16_t *cs16, size_t n_int16, size_t chunk_int16,
                                 int16_t *pcm, size_t pcm_stride, size_t *n_pcm, int *chunk_result_len)
{
	if (!h || !cs16 || !pcm) { set_error("null argument"); return RXB200_EINVAL; }
	int rc = fm_check_shape(h, n_int16, chunk_int16);
	if (rc != RXB200_OK) { return rc; }
	RXB_CUDA(cudaSetDevice(h->device));
	if (n_int16 == 0) { if (n_pcm) { *n_pcm = 0; } return RXB200_OK; }
	size_t total = fm_count_outputs(h, n_int16, chunk_int16, chunk_result_len, false);
	if (total > pcm_stride) { set_error("pcm_stride %zu < %zu outputs", pcm_stride, total); return RXB200_ECAPACITY; }
	size_t in_elems = n_int16 * (size_t)h->n_channels;
	size_t out_elems = (total + 8) * (size_t)h->n_channels;
	if (in_elems > h->d_in_cap) {
		cudaFree(h->d_in); h->d_in = nullptr; h->d_in_cap = 0;
		RXB_CUDA(cudaMalloc(&h->d_in, in_elems * sizeof(int16_t)));
		h->d_in_cap = in_elems;
	}
	if (out_elems > h->d_out_cap) {
		cudaFree(h->d_out); h->d_out = nullptr; h->d_out_cap = 0;
		RXB_CUDA(cudaMalloc(&h->d_out, out_elems * sizeof(int16_t)));
		h->d_out_cap = out_elems;
	}
	RXB_CUDA(cudaMemcpyAsync(h->d_in, cs16, in_elems * sizeof(int16_t), cudaMemcpyHostToDevice, h->stream));
	rc = fm_launch(h, h->d_in, n_int16, chunk_int16, h->d_out, total + 8);
	if (rc != RXB200_OK) { return rc; }
	fm_count_outputs(h, n_int16, chunk_int16, nullptr, true);
	if (pcm_stride == total + 8 || h->n_channels == 1) {
		RXB_CUDA(cudaMemcpyAsync(pcm, h->d_out, (h->n_channels == 1 ? total : out_elems) * sizeof(int16_t),
		                         cudaMemcpyDeviceToHost, h->stream));
	} else {
		RXB_CUDA(cudaMemcpy2DAsync(pcm, pcm_stride * sizeof(int16_t), h->d_out, (total + 8) * sizeof(int16_t),
		                           total * sizeof(int16_t), (size_t)h->n_channels, cudaMemcpyDeviceToHost, h->stream));
	}
	RXB_CUDA(cudaMemcpyAsync(&h->stats.fixup_segments, h->d_fix_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	RXB_CUDA(cudaStreamSynchronize(h->stream));
	if (n_pcm) { *n_pcm = total; }
	return RXB200_OK;
}

extern "C" int rxb200_fm_squelch_hits(rxb200_fm *h, int *hits)
{
	if (!h || !hits) { return RXB200_EINVAL; }
	RXB_CUDA(cudaSetDevice(h->device));
	std::vector<uint32_t> st((size_t)h->n_channels * h->state_words);
	RXB_CUDA(cudaMemcpyAsync(st.data(), h->d_carry[h->cur], st.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
	RXB_CUDA(cudaStreamSynchronize(h->stream));
	for (int c = 0; c < h->n_channels; c++) { hits[c] = (int)st[(size_t)c * h->state_words + ST_SQ_HITS]; }
	return RXB200_OK;
}

extern "C" void *rxb200_fm_stream(rxb200_fm *h) { return h ? (void *)h->stream : nullptr; }

extern "C" int rxb200_fm_last_stats(rxb200_fm *h, rxb200_fm_stats *out)
{
	if (!h || !out) { return RXB200_EINVAL; }
	if (h->stats.fixup_segments < 0) {
		RXB_CUDA(cudaSetDevice(h->device));
		RXB_CUDA(cudaStreamSynchronize(h->stream));
		RXB_CUDA(cudaMemcpy(&h->stats.fixup_segments, h->d_fix_count, sizeof(int), cudaMemcpyDeviceToHost));
	}
	*out = h->stats;
	return RXB200_OK;
}

extern "C" int rxb200_fm_tune(rxb200_fm *h, int segment_len, int deemph_warmup)
{
	if (!h || segment_len < 0 || deemph_warmup < 0) { return RXB200_EINVAL; }
	h->tune_seg = segment_len; h->tune_warm = deemph_warmup;
	return RXB200_OK;
}
