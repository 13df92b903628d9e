// host_plan.cpp — the host-side arithmetic that surrounds the kernels: parameter derivation
// (rx_fm main() + optimal_settings()), the rx_power hop planner (frequency_range()), the window /
// sine tables and the CSV row formatter (csv_dbm()).  Pure C++, no CUDA, no reference code: each
// function restates what the cited reference lines compute so the drop-in host shells and the
// Python mirror get identical numbers (checked against the reference in tests/test_oracle_pin.py
// and tests/test_host_logic.py).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <cuda_runtime.h>
#include "../../include/rxb200.h"

namespace rxb {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
}
}  // namespace rxb
using rxb::set_error;

extern "C" const char *rxb200_last_error(void) { return rxb::g_err; }
extern "C" int rxb200_abi_version(void) { return RXB200_ABI_VERSION; }
extern "C" int rxb200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

// ---------------------------------------------------------------------------------- rx_fm
// main(): defaults (demod_init, src/rtl_fm.c:1084-1115), the -M wbfm preset (:1331-1341), -s/-r/-F/-A
// (:1255-1262, :1305-1319), rate_in *= post_downsample (:1371), output.rate default (:1373),
// optimal_settings() (:960-997) and deemph_a (:1410-1412).
extern "C" int rxb200_fm_derive(const rxb200_fm_cli *cli, rxb200_fm_derived *out)
{
	if (!cli || !out) { set_error("null argument"); return RXB200_EINVAL; }
	memset(out, 0, sizeof *out);
	rxb200_fm_params &p = out->params;
	int rate_in = 24000, rate_out = 24000, rate_out2 = -1, output_rate = 0;
	int custom_atan = RXB200_ATAN_STD, deemph = 0, squelch = cli->squelch_level;
	int mode = cli->mode;
	if (cli->wbfm) {
		mode = RXB200_MODE_FM;
		rate_in = 170000; rate_out = 170000; rate_out2 = 32000; output_rate = 32000;
		// the preset zeroes squelch_level WHEN -M is parsed (:1339), so a later -l survives: the caller applies that
		// parse-order rule to cli->squelch_level (host/rx_fm_b200.c does), derive keeps what it is given
		custom_atan = RXB200_ATAN_FAST; deemph = 1;
	}
	if (cli->rate_s > 0) { rate_in = cli->rate_s; rate_out = cli->rate_s; }
	if (cli->rate_r > 0) { output_rate = cli->rate_r; rate_out2 = cli->rate_r; }
	if (cli->custom_atan >= 0) { custom_atan = cli->custom_atan; }
	if (cli->deemph >= 0) { deemph = cli->deemph; }
	int post = cli->post_downsample > 0 ? cli->post_downsample : 1;
	int passes = cli->use_F ? 1 : 0;
	rate_in *= post;
	if (!output_rate) { output_rate = rate_out; }
	if (rate_in <= 0) { set_error("rate_in %d", rate_in); return RXB200_EINVAL; }
	int downsample = (1000000 / rate_in) + 1;
	if (passes) {
		passes = (int)log2((double)downsample) + 1;
		downsample = 1 << passes;
	}
	int capture_rate = downsample * rate_in;
	int capture_off = cli->offset_tuning ? 0 : capture_rate / 4;
	int output_scale = (1 << 15) / (128 * downsample);
	if (output_scale < 1) { output_scale = 1; }
	if (mode == RXB200_MODE_FM) { output_scale = 1; }
	int deemph_a = 0;
	if (deemph) {
		double tc = (double)(cli->time_constant_us > 0 ? cli->time_constant_us : 75) * 1e-6;
		deemph_a = (int)round(1.0 / ((1.0 - exp(-1.0 / (rate_out * tc)))));
	}
	p.mode = mode; p.downsample = downsample; p.downsample_passes = passes;
	p.comp_fir_size = cli->use_F ? cli->comp_fir_size : 0;
	p.custom_atan = custom_atan; p.output_scale = output_scale; p.post_downsample = post;
	p.deemph = deemph; p.deemph_a = deemph_a; p.rate_out = rate_out; p.rate_out2 = rate_out2;
	p.squelch_level = squelch; p.dc_block_audio = cli->dc_block_audio; p.adc_block_const = 9;
	p.dc_block_raw = cli->dc_block_raw; p.rdc_block_const = cli->rdc_block_const > 0 ? cli->rdc_block_const : 9;
	p.offset_tuning = cli->offset_tuning;
	out->capture_rate = capture_rate;
	out->capture_freq_offset = capture_off;
	out->output_rate = output_rate;
	return RXB200_OK;
}

// ---------------------------------------------------------------------------------- rx_power
#define RXB_MAXIMUM_RATE 2800000      /* src/rtl_power.c:74 */
#define RXB_MINIMUM_RATE 1000000      /* :75 */
#define RXB_DEFAULT_BUF  16384        /* :71 */
#define RXB_MAX_TUNES    10000        /* :111 */

// frequency_range() (src/rtl_power.c:431-543) on already-parsed numbers.
extern "C" int rxb200_power_plan_range(int64_t lower, int64_t upper, int64_t max_size, double crop,
                                       int boxcar, int comp_fir_size, int peak_hold, rxb200_power_plan *out)
{
	if (!out) { set_error("null argument"); return RXB200_EINVAL; }
	if (crop < 0.0 || crop >= 1.0 || upper <= lower || max_size <= 0) { set_error("bad range/crop"); return RXB200_EINVAL; }
	memset(out, 0, sizeof *out);
	int64_t bw_seen = 0, bw_used = 0, downsample = 1, passes = 0;
	int tune_count = 0, bin_e = 0;
	for (int i = 1; i < 1500; i++) {
		bw_seen = (upper - lower) / i;
		bw_used = (int64_t)((double)bw_seen / (1.0 - crop));
		if (bw_used > RXB_MAXIMUM_RATE) { continue; }
		tune_count = i;
		break;
	}
	if (bw_used <= 0) { set_error("unsupported bandwidth"); return RXB200_EINVAL; }
	if (bw_used < RXB_MINIMUM_RATE) {
		tune_count = 1;
		downsample = RXB_MAXIMUM_RATE / bw_used;
		if (downsample <= 0) { set_error("unsupported bandwidth"); return RXB200_EINVAL; }
		bw_used = bw_used * downsample;
	}
	if (!boxcar && downsample > 1) {
		passes = (int)log2((double)downsample);
		downsample = (int64_t)1 << passes;
		bw_used = (int)((double)(bw_seen * downsample) / (1.0 - crop));
	}
	double bin_size = 0.0;
	for (int i = 1; i <= 21; i++) {
		bin_e = i;
		bin_size = (double)bw_used / (double)(((int64_t)1 << i) * downsample);
		if (bin_size <= (double)max_size) { break; }
	}
	if (max_size >= RXB_MINIMUM_RATE) {
		bw_seen = max_size; bw_used = max_size;
		tune_count = (int)((upper - lower) / bw_seen);
		bin_e = 0; crop = 0;
	}
	if (tune_count <= 0) { set_error("no tuning ranges"); return RXB200_EINVAL; }
	if (tune_count > RXB_MAX_TUNES) { set_error("bandwidth too wide"); return RXB200_EINVAL; }
	int64_t buf_len = 2 * ((int64_t)1 << bin_e) * downsample;
	if (buf_len < RXB_DEFAULT_BUF) { buf_len = RXB_DEFAULT_BUF; }
	out->params.n_hops = tune_count; out->params.bin_e = bin_e; out->params.buf_len = (int)buf_len;
	out->params.downsample = (int)downsample; out->params.downsample_passes = (int)passes;
	out->params.comp_fir_size = comp_fir_size; out->params.boxcar = boxcar; out->params.peak_hold = peak_hold;
	out->rate = (int)bw_used; out->crop = crop;
	out->first_freq = lower + bw_seen / 2; out->freq_step = bw_seen;
	out->bin_size_hz = bin_size;
	return RXB200_OK;
}

// window shapes, src/rtl_power.c:322-401
static double window_value(int id, int i, int length)
{
	const double n1 = (double)(length - 1);
	double w;
	switch (id) {
	case RXB200_WIN_HAMMING:
		return 25.0 / 46.0 - (21.0 / 46.0) * cos(2 * i * M_PI / n1);
	case RXB200_WIN_BLACKMAN:
		return 7938.0 / 18608.0 - (9240.0 / 18608.0) * cos(2 * i * M_PI / n1) + (1430.0 / 18608.0) * cos(4 * i * M_PI / n1);
	case RXB200_WIN_BLACKMAN_HARRIS:
		return 0.35875 - 0.48829 * cos(2 * i * M_PI / n1) + 0.14128 * cos(4 * i * M_PI / n1) - 0.01168 * cos(6 * i * M_PI / n1);
	case RXB200_WIN_HANN_POISSON:
		return 0.5 * (1 - cos(2 * M_PI * i / n1)) * pow(M_E, (-2.0 * (double)abs((int)(n1 - 1 - 2 * i))) / n1);
	case RXB200_WIN_YOUSSEF:
		w = 0.35875 - 0.48829 * cos(2 * i * M_PI / n1) + 0.14128 * cos(4 * i * M_PI / n1) - 0.01168 * cos(6 * i * M_PI / n1);
		return w * pow(M_E, (-0.0025 * (double)abs((int)(n1 - 1 - 2 * i))) / n1);
	case RXB200_WIN_BARTLETT:
		w = (i - n1 / 2) / ((double)length / 2);
		return 1 - (w < 0 ? -w : w);
	case RXB200_WIN_HANN:
		return 0.5 * (1 - cos(2 * M_PI * i / n1));
	default:
		return 1.0;   // rectangle, kaiser (src/rtl_power.c:322, :385)
	}
}

extern "C" int rxb200_window_table(int window, int length, int *coefs)
{
	if (!coefs || length < 1 || window < 0 || window > RXB200_WIN_HANN) { set_error("bad window request"); return RXB200_EINVAL; }
	for (int i = 0; i < length; i++) { coefs[i] = (int)(256 * window_value(window, i, length)); }   // :1036
	return RXB200_OK;
}

extern "C" int rxb200_sine_table(int log2_n, int16_t *sine)
{
	if (!sine || log2_n < 0 || log2_n > 21) { set_error("bad sine table request"); return RXB200_EINVAL; }
	const int n = 1 << log2_n;
	for (int i = 0; i < n * 3 / 4; i++) {
		double d = (double)i * 2.0 * M_PI / n;
		sine[i] = (int16_t)(int)round(32767 * sin(d));   // src/rtl_power.c:250-251
	}
	return RXB200_OK;
}

// csv_dbm() (src/rtl_power.c:774-817) for one row, without the date/time prefix and without the
// zeroing (the device accumulators are cleared by rxb200_power_reset).
extern "C" int rxb200_power_row_len(int bin_e, double crop)
{
	if (bin_e < 0 || bin_e > 30) { return RXB200_EINVAL; }
	const int len = 1 << bin_e;
	const int i1 = 0 + (int)((double)len * crop * 0.5);
	const int i2 = (len - 1) - (int)((double)len * crop * 0.5);
	return (i2 >= i1 ? i2 - i1 + 1 : 0) + 1;
}

extern "C" int rxb200_power_format_db_row(const double *db, int bin_e, int64_t freq, int rate, int downsample,
                                          double crop, int samples, char *dst, size_t cap)
{
	if (!db || !dst) { set_error("null argument"); return RXB200_EINVAL; }
	const int len = 1 << bin_e, ds = downsample;
	const int n = rxb200_power_row_len(bin_e, crop);
	size_t w = 0;
#define RXB_EMIT(...)                                                        \
	do {                                                                     \
		int n__ = snprintf(dst + w, w < cap ? cap - w : 0, __VA_ARGS__);     \
		if (n__ < 0 || w + (size_t)n__ >= cap) { return RXB200_ECAPACITY; }  \
		w += (size_t)n__;                                                    \
	} while (0)
	int bin_count = (int)((double)len * (1.0 - crop));
	int bw2 = (int)(((double)rate * (double)bin_count) / (len * 2 * ds));
	RXB_EMIT("%lli, %lli, %.2f, %i, ", (long long)freq - bw2, (long long)freq + bw2,
	         (double)rate / (double)(len * ds), samples);
	for (int i = 0; i < n - 1; i++) { RXB_EMIT("%.2f, ", db[i]); }
	RXB_EMIT("%.2f\n", db[n - 1]);
#undef RXB_EMIT
	return (int)w;
}

extern "C" int rxb200_power_format_row(int64_t *avg, int bin_e, int64_t freq, int rate, int downsample,
                                       double crop, int samples, char *dst, size_t cap)
{
	if (!avg || !dst) { set_error("null argument"); return RXB200_EINVAL; }
	const int len = 1 << bin_e, ds = downsample;
	size_t w = 0;
#define RXB_EMIT(...)                                                        \
	do {                                                                     \
		int n__ = snprintf(dst + w, w < cap ? cap - w : 0, __VA_ARGS__);     \
		if (n__ < 0 || w + (size_t)n__ >= cap) { return RXB200_ECAPACITY; }  \
		w += (size_t)n__;                                                    \
	} while (0)
	if (bin_e > 0) {
		avg[0] = avg[1];                       // DC bin copied from its neighbour (:784)
		for (int i = 0; i < len / 2; i++) {    // half-swap (:786-790)
			int64_t t = avg[i]; avg[i] = avg[i + len / 2]; avg[i + len / 2] = t;
		}
	}
	int bin_count = (int)((double)len * (1.0 - crop));
	int bw2 = (int)(((double)rate * (double)bin_count) / (len * 2 * ds));
	RXB_EMIT("%lli, %lli, %.2f, %i, ", (long long)freq - bw2, (long long)freq + bw2,
	         (double)rate / (double)(len * ds), samples);
	int i1 = 0 + (int)((double)len * crop * 0.5);
	int i2 = (len - 1) - (int)((double)len * crop * 0.5);
	double dbm;
	for (int i = i1; i <= i2; i++) {
		dbm = (double)avg[i];
		dbm /= (double)rate;
		dbm /= (double)samples;
		dbm = 10 * log10(dbm);
		RXB_EMIT("%.2f, ", dbm);
	}
	dbm = (double)avg[i2] / ((double)rate * (double)samples);   // last bin once more (:807)
	if (bin_e == 0) { dbm = ((double)avg[0] / ((double)rate * (double)samples)); }
	dbm = 10 * log10(dbm);
	RXB_EMIT("%.2f\n", dbm);
#undef RXB_EMIT
	return (int)w;
}
