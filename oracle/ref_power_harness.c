/* ref_power_harness.c — synchronous driver around the UNMODIFIED reference rx_power scanner.
 *
 * TEST INFRASTRUCTURE.  Built only when /root/reference is present (oracle/Makefile,
 * target _ref/libref_power.so).  #includes src/rtl_power.c where it lies, renames main(),
 * and runs frequency_range() (src/rtl_power.c:431), sine_table() (:240), the window-table
 * build of main() (:1032-1037) and scanner() (:670) against hop buffers served from memory
 * by the fake device.  Test inputs are defined as "what lands in ts->buf16[0..buf_len)"
 * (SURVEY.md F10): readStream asks for buf_len COMPLEX elements but only the first buf_len
 * int16 are used, so the hook below copies buf_len int16 per hop and zero-fills the rest.
 * Nothing here is copied from the reference.
 */
#define main rtl_power_reference_main
#include "rtl_power.c"
#undef main

#include <time.h>

/* provided by soapy_fake.c */
void soapy_fake_set_read_hook(int (*hook)(void * const *buffs, size_t numElems));

static const int16_t *g_hops = NULL;   /* [n_pass][tune_count][buf_len] int16 */
static size_t g_hop_idx = 0, g_hop_total = 0;
static int g_buf_len = 0;

static int hop_read_hook(void * const *buffs, size_t numElems)
{
	if (buffs[0] == (void *)dump) {              /* retune() flush read, src/rtl_power.c:568 */
		return (int)numElems;
	}
	if (g_hop_idx >= g_hop_total) { return -2; }
	memcpy(buffs[0], g_hops + g_hop_idx * (size_t)g_buf_len, (size_t)g_buf_len * 2);
	if (numElems * 2 > (size_t)g_buf_len) {
		memset((int16_t *)buffs[0] + g_buf_len, 0, (numElems * 2 - (size_t)g_buf_len) * 2);
	}
	g_hop_idx++;
	return (int)numElems;
}

static void free_plan(void)
{
	int i;
	for (i = 0; i < tune_count; i++) {
		free(tunes[i].avg); tunes[i].avg = NULL;
		free(tunes[i].buf16); tunes[i].buf16 = NULL;
	}
	tune_count = 0;
	free(fft_buf); fft_buf = NULL;
	free(window_coefs); window_coefs = NULL;
	free(Sinewave); Sinewave = NULL;
	free(power_table); power_table = NULL;
}

typedef struct {
	int tune_count, bin_e, buf_len, downsample, downsample_passes, rate;
	double crop;
} ref_power_plan_t;

/* window: 0 rectangle 1 hamming 2 blackman 3 blackman-harris 4 hann-poisson 5 youssef
 * 6 kaiser 7 bartlett (src/rtl_power.c:881-898); custom_window != NULL overrides the table
 * ("Hann" is just another host-built table). */
int ref_power_setup(const char *freq_arg, double crop_frac, int use_boxcar, int fir_size, int peak,
                    int window, const int *custom_window, ref_power_plan_t *plan)
{
	static double (*fns[8])(int, int) = { rectangle, hamming, blackman, blackman_harris,
	                                      hann_poisson, youssef, kaiser, bartlett };
	char *arg = strdup(freq_arg);
	int i, length;
	free_plan();
	boxcar = use_boxcar; comp_fir_size = fir_size; peak_hold = peak;
	tuner_sleep_usec = 0;
	frequency_range(arg, crop_frac);
	free(arg);
	if (tune_count == 0) { return -1; }
	dev = SoapySDRDevice_makeStrArgs("");
	stream = NULL;
	sine_table(tunes[0].bin_e);
	fft_buf = malloc(tunes[0].buf_len * sizeof(int16_t) * 2);
	length = 1 << tunes[0].bin_e;
	window_coefs = malloc(length * sizeof(int));
	for (i = 0; i < length; i++) {
		window_coefs[i] = custom_window ? custom_window[i] : (int)(256 * fns[window & 7](i, length));
	}
	plan->tune_count = tune_count;
	plan->bin_e = tunes[0].bin_e;
	plan->buf_len = tunes[0].buf_len;
	plan->downsample = tunes[0].downsample;
	plan->downsample_passes = tunes[0].downsample_passes;
	plan->rate = tunes[0].rate;
	plan->crop = tunes[0].crop;
	soapy_fake_set_read_hook(hop_read_hook);
	return 0;
}

int ref_power_tables(int *window_out, int16_t *sine_out)
{
	int i, length = 1 << tunes[0].bin_e;
	if (window_out) { for (i = 0; i < length; i++) { window_out[i] = window_coefs[i]; } }
	if (sine_out) { for (i = 0; i < N_WAVE * 3 / 4; i++) { sine_out[i] = Sinewave[i]; } }
	return N_WAVE;
}

long long ref_power_hop_freq(int i) { return (long long)tunes[i].freq; }

/* n_pass sweeps of scanner(); hop_bufs = int16[n_pass][tune_count][buf_len]. */
int ref_power_scan(const int16_t *hop_bufs, int n_pass)
{
	int p;
	g_hops = hop_bufs; g_hop_idx = 0; g_buf_len = tunes[0].buf_len;
	g_hop_total = (size_t)n_pass * (size_t)tune_count;
	do_exit = 0;
	for (p = 0; p < n_pass; p++) { scanner(0); }
	return (int)g_hop_idx;
}

double ref_power_time(const int16_t *hop_bufs, int n_pass, int repeats)
{
	struct timespec t0, t1;
	int r;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (r = 0; r < repeats; r++) { ref_power_scan(hop_bufs, n_pass); }
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void ref_power_get(int64_t *avg_out /*[tune_count][2^bin_e]*/, int *samples_out)
{
	int i, j, len = 1 << tunes[0].bin_e;
	for (i = 0; i < tune_count; i++) {
		for (j = 0; j < len; j++) { avg_out[(size_t)i * len + j] = tunes[i].avg[j]; }
		samples_out[i] = tunes[i].samples;
	}
}

void ref_power_reset(void)
{
	int i, j, len = 1 << tunes[0].bin_e;
	for (i = 0; i < tune_count; i++) {
		for (j = 0; j < len; j++) { tunes[i].avg[j] = 0; }
		tunes[i].samples = 0;
	}
}

/* csv_dbm() over every hop into a file (consumes and zeroes avg, like the reference main
 * loop, src/rtl_power.c:1047-1050); the date/time prefix is supplied by the caller. */
int ref_power_csv(const char *path, const char *tstr)
{
	int i;
	file = fopen(path, "wb");
	if (!file) { return -1; }
	for (i = 0; i < tune_count; i++) {
		fprintf(file, "%s, ", tstr);
		csv_dbm(&tunes[i]);
	}
	fclose(file);
	return 0;
}

/* direct leaf access for unit-level pinning of the port */
int ref_fix_fft(int16_t *iq, int m, int log2_n_wave)
{
	if (!Sinewave || LOG2_N_WAVE != log2_n_wave) {
		free(Sinewave); free(power_table); Sinewave = NULL; power_table = NULL;
		sine_table(log2_n_wave);
	}
	return fix_fft(iq, m);
}
