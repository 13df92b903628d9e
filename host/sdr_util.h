/* sdr_util.h — small host helpers shared by the drop-in shells rx_fm_b200 / rx_power_b200.
 * Own implementation of what rx_tools keeps in src/convenience/ (suffix parsing, device/stream
 * bring-up); behaviourally equivalent, not copied. */
#ifndef RXB_SDR_UTIL_H
#define RXB_SDR_UTIL_H
#include <stddef.h>
#include <stdint.h>
#include <SoapySDR/Device.h>
#include <SoapySDR/Formats.h>

double parse_scaled(const char *s);    /* 100M, 24k, 1.2G   (convenience.c:65-90)  */
double parse_seconds(const char *s);   /* 10s, 5m, 1h       (convenience.c:92-117) */
double parse_fraction(const char *s);  /* 28.5% -> 0.285    (convenience.c:119-136) */

/* makeStrArgs + CS16 RX stream on `channel`; prints like the reference's verbose_* helpers. Returns 0 on success. */
int sdr_open(const char *query, size_t channel, SoapySDRDevice **dev, SoapySDRStream **stream);
void sdr_close(SoapySDRDevice *dev, SoapySDRStream *stream);
/* gain: NULL = automatic, "30" overall, or "LNA=20,VGA=10" per element */
void sdr_set_gain(SoapySDRDevice *dev, size_t channel, const char *gain_str);
#endif
