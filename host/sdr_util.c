#include "sdr_util.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double with_suffix(const char *s, const char *units, const double *mult)
{
	size_t n = strlen(s);
	if (n == 0) { return 0.0; }
	const char *u = strchr(units, s[n - 1]);
	if (u && *u) {
		char tmp[64];
		if (n - 1 >= sizeof tmp) { return 0.0; }
		memcpy(tmp, s, n - 1); tmp[n - 1] = 0;
		return atof(tmp) * mult[u - units];
	}
	return atof(s);
}

double parse_scaled(const char *s)
{
	static const double m[] = {1e3, 1e3, 1e6, 1e6, 1e9, 1e9};
	return with_suffix(s, "kKmMgG", m);
}
double parse_seconds(const char *s)
{
	static const double m[] = {1, 1, 60, 60, 3600, 3600};
	return with_suffix(s, "sSmMhH", m);
}
double parse_fraction(const char *s)
{
	static const double m[] = {0.01};
	return with_suffix(s, "%", m);
}

int sdr_open(const char *query, size_t channel, SoapySDRDevice **dev, SoapySDRStream **stream)
{
	SoapySDRKwargs none = {0, NULL, NULL};
	*dev = SoapySDRDevice_makeStrArgs(query ? query : "");
	if (!*dev) { fprintf(stderr, "SoapySDRDevice_make failed: %s\n", SoapySDRDevice_lastError()); return -1; }
	char *hw = SoapySDRDevice_getHardwareKey(*dev);
	fprintf(stderr, "Using device %s\n", hw ? hw : "?");
	free(hw);
	if (channel >= SoapySDRDevice_getNumChannels(*dev, SOAPY_SDR_RX)) {
		fprintf(stderr, "Invalid channel %d selected\n", (int)channel);
		return -3;
	}
	*stream = SoapySDRDevice_setupStream(*dev, SOAPY_SDR_RX, SOAPY_SDR_CS16, &channel, 1, &none);
	if (!*stream) { fprintf(stderr, "SoapySDRDevice_setupStream failed: %s\n", SoapySDRDevice_lastError()); return -3; }
	return 0;
}

void sdr_close(SoapySDRDevice *dev, SoapySDRStream *stream)
{
	if (dev && stream) { SoapySDRDevice_deactivateStream(dev, stream, 0, 0); SoapySDRDevice_closeStream(dev, stream); }
	if (dev) { SoapySDRDevice_unmake(dev); }
}

void sdr_set_gain(SoapySDRDevice *dev, size_t channel, const char *gain_str)
{
	if (!gain_str) {
		if (SoapySDRDevice_setGainMode(dev, SOAPY_SDR_RX, channel, true) != 0) { fprintf(stderr, "WARNING: Failed to enable automatic gain.\n"); }
		return;
	}
	if (strchr(gain_str, '=')) {
		SoapySDRKwargs kw = SoapySDRKwargs_fromString(gain_str);
		for (size_t i = 0; i < kw.size; i++) {
			if (SoapySDRDevice_setGainElement(dev, SOAPY_SDR_RX, channel, kw.keys[i], atof(kw.vals[i])) != 0) {
				fprintf(stderr, "WARNING: setGainElement(%s) failed: %s\n", kw.keys[i], SoapySDRDevice_lastError());
			}
		}
		SoapySDRKwargs_clear(&kw);
	} else if (SoapySDRDevice_setGain(dev, SOAPY_SDR_RX, channel, atof(gain_str)) != 0) {
		fprintf(stderr, "WARNING: Failed to set tuner gain: %s\n", SoapySDRDevice_lastError());
	}
}
