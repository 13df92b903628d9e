/* Declarations-only stand-in for <SoapySDR/Version.h> (SoapySDR is not installed in
 * this image).  TEST INFRASTRUCTURE: lets the unmodified reference sources compile
 * for the oracle (oracle/Makefile) and lets the drop-in host shells link against the
 * file-replay fake device (host/soapy_replay/soapy_fake.c).  Written from the public SoapySDR
 * 0.8 C API; contains no reference code. */
#pragma once
#define SOAPY_SDR_API_VERSION 0x00080000
#define SOAPY_SDR_ABI_VERSION "0.8"
