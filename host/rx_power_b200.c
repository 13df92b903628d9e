/* rx_power_b200 — drop-in for rx_tools' rx_power whose per-hop DSP runs on a B200 through librxb200.so.
 *
 * Same command line (option string of src/rtl_power.c:848), same CSV rows
 * ("date, time, Hz low, Hz high, Hz step, samples, dB, dB, ..."), same SoapySDR CS16 stream surface.
 * Host logic is our own: plan (rxb200_power_plan_range == frequency_range()), retune + flush read per
 * hop, one readStream per hop buffer; the window x fix_fft x power-accumulate of every hop of a sweep is
 * ONE rxb200_power_accumulate() call, the report is rxb200_power_read_db() (csv_dbm's arithmetic on the
 * device) + rxb200_power_format_db_row().
 *
 * Several GPUs: RXB200_GPUS=n (environment, so the reference's option string stays as it is) shards the hops over
 * n GPUs of this box through rxb200_power_group_* -- every GPU transforms the hops it owns, ONE NCCL all-gather
 * collates the spectrum rows in hop order before the report loop (src/rtl_power.c:1047-1050).
 */
#include <math.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "sdr_util.h"
#include "rxb200.h"

#define BUFFER_DUMP 16384            /* elements flushed after a retune, src/rtl_power.c:71-72, :568 */

static volatile sig_atomic_t g_stop = 0;
static void on_signal(int s) { (void)s; g_stop++; }

static void usage(void)
{
	fprintf(stderr,
		"rx_power_b200, rx_power with the FFT on a B200 GPU\n\n"
		"Use:\trx_power_b200 -f freq_range [-options] [filename]\n"
		"\t-f lower:upper:bin_size [Hz]\n"
		"\t[-i integration_interval (default: 10 seconds)] [-1 single-shot] [-e exit_timer]\n"
		"\t[-d device] [-C channel] [-a antenna] [-g gain] [-p ppm_error]\n"
		"\t[-w rectangle|hamming|blackman|blackman-harris|hann-poisson|youssef|kaiser|bartlett]\n"
		"\t[-c crop_percent] [-F fir_size] [-P peak hold] [-D direct_sampling] [-O offset tuning]\n"
		"\t[-S tuner_sleep_usec] [-R tuner_retry_max]\n"
		"\tfilename ('-' means stdout)\n");
	exit(1);
}

static int window_id(const char *name)
{
	static const char *names[] = {"rectangle", "hamming", "blackman", "blackman-harris", "hann-poisson", "youssef", "kaiser", "bartlett"};
	for (int i = 0; i < 8; i++) { if (!strcmp(name, names[i])) { return i; } }
	return -1;
}

int main(int argc, char **argv)
{
	const char *dev_query = "", *gain_str = NULL, *antenna = NULL, *filename = "-", *freq_arg = NULL;
	int opt, interval = 10, single = 0, ppm = 0, direct_sampling = 0, offset_tuning = 0, window = RXB200_WIN_RECTANGLE;
	int boxcar = 1, comp_fir_size = 0, peak_hold = 0, tuner_sleep_usec = 5000, tuner_retry_max = 3;
	size_t channel = 0;
	double crop = 0.0;
	time_t exit_after = 0;
	while ((opt = getopt(argc, argv, "a:C:f:i:s:t:d:g:p:e:w:c:F:1PD:OS:R:h")) != -1) {
		switch (opt) {
		case 'a': antenna = optarg; break;
		case 'C': channel = (size_t)atoi(optarg); break;
		case 'f': freq_arg = optarg; break;
		case 'd': dev_query = optarg; break;
		case 'g': gain_str = optarg; break;
		case 'c': crop = parse_fraction(optarg); break;
		case 'i': interval = (int)round(parse_seconds(optarg)); break;
		case 'e': exit_after = (time_t)((int)round(parse_seconds(optarg))); break;
		case 's': case 't': break;                    /* parsed but unused by the reference as well (:875-880, :899-901) */
		case 'w': { int w = window_id(optarg); if (w >= 0) { window = w; } } break;
		case 'p': ppm = atoi(optarg); break;
		case '1': single = 1; break;
		case 'P': peak_hold = 1; break;
		case 'D': direct_sampling = atoi(optarg); break;
		case 'O': offset_tuning = 1; break;
		case 'F': boxcar = 0; comp_fir_size = atoi(optarg); break;
		case 'S': tuner_sleep_usec = atoi(optarg); break;
		case 'R': tuner_retry_max = atoi(optarg); break;
		default: usage();
		}
	}
	if (!freq_arg) { fprintf(stderr, "No frequency range provided.\n"); usage(); }
	if (crop < 0.0 || crop > 1.0) { fprintf(stderr, "Crop value outside of 0 to 1.\n"); return 1; }
	if (argc > optind) { filename = argv[optind]; }
	if (interval < 1) { interval = 1; }

	/* -f lower:upper:bin */
	char *fa = strdup(freq_arg), *c1 = strchr(fa, ':'), *c2 = c1 ? strchr(c1 + 1, ':') : NULL;
	if (!c1 || !c2) { fprintf(stderr, "Bad frequency range.\n"); return 1; }
	*c1++ = 0; *c2++ = 0;
	rxb200_power_plan plan;
	if (rxb200_power_plan_range((int64_t)parse_scaled(fa), (int64_t)parse_scaled(c1), (int64_t)parse_scaled(c2), crop,
	                            boxcar, comp_fir_size, peak_hold, &plan) != RXB200_OK) {
		fprintf(stderr, "rxb200: %s\n", rxb200_last_error());
		return 1;
	}
	free(fa);
	const rxb200_power_params *pp = &plan.params;
	const int n_hops = pp->n_hops, N = 1 << pp->bin_e, buf_len = pp->buf_len;
	fprintf(stderr, "Number of frequency hops: %i\n", n_hops);
	fprintf(stderr, "Dongle bandwidth: %iHz\n", plan.rate);
	fprintf(stderr, "Downsampling by: %ix\n", pp->downsample);
	fprintf(stderr, "Cropping by: %0.2f%%\n", plan.crop * 100);
	fprintf(stderr, "Total FFT bins: %i\n", n_hops * N);
	fprintf(stderr, "Logged FFT bins: %i\n", (int)((double)(n_hops * N) * (1.0 - plan.crop)));
	fprintf(stderr, "FFT bin size: %0.2fHz\n", plan.bin_size_hz);
	fprintf(stderr, "Buffer size: %i bytes (%0.2fms)\n", buf_len, 1000 * 0.5 * (float)buf_len / (float)plan.rate);
	fprintf(stderr, "Reporting every %i seconds\n", interval);

	int *window_coefs = (int *)malloc(sizeof(int) * (size_t)N);
	rxb200_window_table(window, N, window_coefs);
	int n_gpus = getenv("RXB200_GPUS") ? atoi(getenv("RXB200_GPUS")) : 1;
	if (n_gpus < 1) { n_gpus = 1; }
	if (n_gpus > n_hops) { n_gpus = n_hops; }
	rxb200_power_group *grp = NULL;
	if (rxb200_power_group_create(pp, window_coefs, NULL, n_gpus, NULL, &grp) != RXB200_OK) { fprintf(stderr, "rxb200: %s\n", rxb200_last_error()); return 1; }
	rxb200_power *pw = rxb200_power_group_member(grp, 0);     /* after the gather member 0 holds every row */
	if (n_gpus > 1) { fprintf(stderr, "Hops sharded over %d GPUs\n", n_gpus); }

	SoapySDRDevice *dev = NULL; SoapySDRStream *stream = NULL;
	if (sdr_open(dev_query, channel, &dev, &stream) != 0) { fprintf(stderr, "Failed to open sdr device matching '%s'.\n", dev_query); return 1; }
	if (antenna && SoapySDRDevice_setAntenna(dev, SOAPY_SDR_RX, channel, antenna) != 0) { fprintf(stderr, "Failed to set antenna"); }
	SoapySDRDevice_activateStream(dev, stream, 0, 0, 0);
	struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_handler = on_signal;
	sigaction(SIGINT, &sa, NULL); sigaction(SIGTERM, &sa, NULL); sigaction(SIGQUIT, &sa, NULL);
	signal(SIGPIPE, SIG_IGN);
	if (direct_sampling) { SoapySDRDevice_writeSetting(dev, "direct_samp", direct_sampling == 1 ? "1" : "2"); }
	if (offset_tuning) { SoapySDRDevice_writeSetting(dev, "offset_tune", "true"); }
	sdr_set_gain(dev, channel, gain_str);
	SoapySDRDevice_setFrequencyCorrection(dev, SOAPY_SDR_RX, channel, (double)ppm);
	FILE *out = !strcmp(filename, "-") ? stdout : fopen(filename, "wb");
	if (!out) { fprintf(stderr, "Failed to open %s\n", filename); return 1; }
	SoapySDRDevice_setSampleRate(dev, SOAPY_SDR_RX, channel, (double)plan.rate);

	/* buffers: one read of buf_len COMPLEX elements per hop (the reference asks for that many, src/rtl_power.c:694),
	 * of which the first buf_len int16 are the hop buffer (:715-720) */
	int16_t *rd = (int16_t *)malloc((size_t)buf_len * 4);
	int16_t *dump = (int16_t *)malloc((size_t)BUFFER_DUMP * 4);
	int16_t *stage = (int16_t *)malloc((size_t)n_hops * (size_t)buf_len * 2);
	const int row_len = rxb200_power_row_len(pp->bin_e, plan.crop);
	double *db = (double *)malloc((size_t)n_hops * (size_t)row_len * sizeof(double));
	int *samples = (int *)malloc((size_t)n_hops * sizeof(int));
	char *row = (char *)malloc(64 + 16 * ((size_t)N + 8));
	const char *hook = getenv("RXB200_MAX_SWEEPS");       /* test hook: report after this many sweeps and exit */
	const long max_sweeps = hook ? atol(hook) : 0;
	long sweeps = 0;
	time_t next_tick = time(NULL) + interval, exit_time = exit_after ? time(NULL) + exit_after : 0;
	SoapySDRKwargs none = {0, NULL, NULL};
	int done = 0;
	while (!done && !g_stop) {
		/* ---- one sweep: scanner()'s device side (:679-708) */
		int first_ok = -1, n_ok = 0;
		for (int i = 0; i < n_hops && g_stop < 2; i++) {
			const int64_t f = plan.first_freq + (int64_t)i * plan.freq_step;
			if ((int64_t)SoapySDRDevice_getFrequency(dev, SOAPY_SDR_RX, channel) != f) {       /* retune() :548-580 */
				if (SoapySDRDevice_setFrequency(dev, SOAPY_SDR_RX, channel, (double)f, &none) != 0) {
					fprintf(stderr, "Error: failed to set frequency %lli Hz\n", (long long)f);
				} else {
					usleep((useconds_t)tuner_sleep_usec);
					int r = -1, flags = 0; long long tn = 0; void *db[] = {dump};
					for (int a = 0; a < tuner_retry_max && r < 0; a++) { r = SoapySDRDevice_readStream(dev, stream, db, BUFFER_DUMP, &flags, &tn, 1000000); }
					if (r < 0) { fprintf(stderr, "Error: bad retune at %lli Hz, r=%d (try increasing -S or -R).\n", (long long)f, r); }
				}
			}
			void *buffs[] = {rd};
			int flags = 0; long long tn = 0;
			int r = SoapySDRDevice_readStream(dev, stream, buffs, (size_t)buf_len, &flags, &tn, 1000000);
			if (r < 0) { fprintf(stderr, "Error: reading stream %d\n", r); continue; }           /* :700-703 */
			if (first_ok < 0) { first_ok = i; }
			if (i != first_ok + n_ok) {
				/* a hop in the middle failed: flush the contiguous run gathered so far, start a new one */
				rxb200_power_group_accumulate(grp, stage + (size_t)first_ok * buf_len, 1, first_ok, first_ok + n_ok);
				first_ok = i; n_ok = 0;
			}
			memcpy(stage + (size_t)i * buf_len, rd, (size_t)buf_len * 2);
			n_ok++;
		}
		if (n_ok > 0) {
			if (rxb200_power_group_accumulate(grp, stage + (size_t)first_ok * buf_len, 1, first_ok, first_ok + n_ok) != RXB200_OK) {
				fprintf(stderr, "rxb200: %s\n", rxb200_last_error());
				break;
			}
		}
		sweeps++;
		const int stream_dry = (first_ok < 0);              /* replay device ran out of samples */
		time_t now = time(NULL);
		const int hooked = max_sweeps && sweeps >= max_sweeps;
		if (now < next_tick && !hooked && !stream_dry) { continue; }
		/* ---- report (:1044-1058) */
		struct tm cal; char tstr[50];
		localtime_r(&now, &cal);
		strftime(tstr, sizeof tstr, "%Y-%m-%d, %H:%M:%S", &cal);
		if (getenv("RXB200_FIXED_TIME")) { snprintf(tstr, sizeof tstr, "%s", getenv("RXB200_FIXED_TIME")); }
		if (rxb200_power_group_gather(grp) != RXB200_OK) { fprintf(stderr, "rxb200: %s\n", rxb200_last_error()); break; }
		if (rxb200_power_read_db(pw, plan.rate, plan.crop, db, (size_t)row_len, samples) != RXB200_OK) { fprintf(stderr, "rxb200: %s\n", rxb200_last_error()); break; }
		int have = 0;
		for (int i = 0; i < n_hops; i++) { have |= samples[i]; }
		if (have) {
			for (int i = 0; i < n_hops; i++) {
				int n = rxb200_power_format_db_row(db + (size_t)i * row_len, pp->bin_e, plan.first_freq + (int64_t)i * plan.freq_step, plan.rate,
				                                pp->downsample, plan.crop, samples[i], row, 64 + 16 * ((size_t)N + 8));
				if (n < 0) { fprintf(stderr, "rxb200: %s\n", rxb200_last_error()); break; }
				fprintf(out, "%s, ", tstr);
				fwrite(row, 1, (size_t)n, out);
			}
			fflush(out);
		}
		rxb200_power_group_reset(grp);
		while (time(NULL) >= next_tick) { next_tick += interval; }
		if (single || hooked || stream_dry) { done = 1; }
		if (exit_time && time(NULL) >= exit_time) { done = 1; }
	}
	fprintf(stderr, g_stop ? "\nUser cancel, exiting...\n" : "\nDone, exiting...\n");
	if (out != stdout) { fclose(out); }
	rxb200_power_group_destroy(grp);
	sdr_close(dev, stream);
	return 0;
}
