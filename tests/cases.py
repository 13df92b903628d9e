"""Shared parity-test cases.  Each case = derived DSP parameters + a seeded
input + the chunk length (chunk length is part of every vector: SURVEY F7)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import numpy as np

from oracle import ATAN_ALE, ATAN_FAST, ATAN_LUT, ATAN_STD, MODE_AM, MODE_FM, MODE_LSB, MODE_RAW, MODE_USB, FmParams
from rx_tools_b200 import synth


@dataclass
class FmCase:
    name: str
    params: FmParams
    make_input: Callable[[], np.ndarray]
    chunk_int16: int = 262144
    exact: bool = True          # False: std atan2 path, compare within tolerance
    tags: tuple = field(default_factory=tuple)


def _wb(n, seed=2, amp=0.45):
    return lambda: synth.fm_iq(n, fs=2.4e6, deviation_hz=75e3, tones=[(400.0, 1.0), (3000.0, 0.7), (11000.0, 0.4)],
                               amplitude=amp * 32767, noise_lsb=64, seed=seed)


def _nb(n, fs, seed, amp=16000.0):
    return lambda: synth.fm_iq(n, fs=fs, deviation_hz=5e3, tones=[(1000.0, 1.0)], amplitude=amp, noise_lsb=64, seed=seed)


def fm_cases(scale: int = 1) -> list[FmCase]:
    """scale multiplies the input length (1 = CPU-seconds sized)."""
    n = (1 << 19) * scale
    C = 262144
    cs = []
    # cfg1 family: -M fm -s 1024000 -r 24000, D=1
    p1 = dict(downsample=1, rate_out=1_024_000, rate_out2=24_000)
    cs.append(FmCase("cfg1_std", FmParams(custom_atan=ATAN_STD, **p1), _nb(n, 1.024e6, 12345), C, exact=False))
    cs.append(FmCase("cfg1_fast", FmParams(custom_atan=ATAN_FAST, **p1), _nb(n, 1.024e6, 12345), C))
    cs.append(FmCase("cfg1_lut", FmParams(custom_atan=ATAN_LUT, **p1), _nb(n, 1.024e6, 12345), C))
    cs.append(FmCase("cfg1_ale", FmParams(custom_atan=ATAN_ALE, **p1), _nb(n, 1.024e6, 12345), C))
    # cfg2A: -M wbfm -s 2400000 -r 48000 : D=1, fast, deemph a=181, /50
    cs.append(FmCase("cfg2A", FmParams(downsample=1, custom_atan=ATAN_FAST, deemph=1, deemph_a=181,
                                       rate_out=2_400_000, rate_out2=48_000), _wb(n), C))
    # cfg2B: -M wbfm -s 300k -F 9 -r 48k : P=3, droop FIR, fast, deemph a=23, /6
    cs.append(FmCase("cfg2B", FmParams(downsample=8, downsample_passes=3, comp_fir_size=9, custom_atan=ATAN_FAST,
                                       deemph=1, deemph_a=23, rate_out=300_000, rate_out2=48_000), _wb(n), C))
    # same with a smaller chunk (F7: result differs from chunk 262144)
    cs.append(FmCase("cfg2B_chunk65536", FmParams(downsample=8, downsample_passes=3, comp_fir_size=9,
                                                  custom_atan=ATAN_FAST, deemph=1, deemph_a=23, rate_out=300_000,
                                                  rate_out2=48_000), _wb(n), 65536))
    # a full-scale tone that lands on the decimated Nyquist frequency after the fs/4 rotation (-fs/4 + fs/16), where the droop
    # FIR has its largest gain (4.4) and consecutive decimated samples alternate in sign: the largest conjugate products a
    # tone can give the row front end's FP32 discriminator (fm_rows.cuh); second half ordinary signal
    def _overshoot():
        k = np.arange(n // 2)
        ph = 2.0 * np.pi * (-3.0 / 16.0) * k
        x = np.empty(n, dtype=np.int16)              # n // 2 complex samples of the tone
        x[0::2] = np.clip(np.round(32767 * np.cos(ph)), -32768, 32767).astype(np.int16)
        x[1::2] = np.clip(np.round(32767 * np.sin(ph)), -32768, 32767).astype(np.int16)
        return np.concatenate([x, _wb(n // 2, seed=61)()])
    cs.append(FmCase("cfg2B_fir_overshoot", FmParams(downsample=8, downsample_passes=3, comp_fir_size=9, custom_atan=ATAN_FAST,
                                                     deemph=1, deemph_a=23, rate_out=300_000, rate_out2=48_000), _overshoot, C))
    # wbfm with one and two passes + droop FIR (-s 1200k / -s 600k -F 9): the row front end's P = 1, 2 instantiations
    cs.append(FmCase("wbfm_P1_fir", FmParams(downsample=2, downsample_passes=1, comp_fir_size=9, custom_atan=ATAN_FAST,
                                             deemph=1, deemph_a=91, rate_out=1_200_000, rate_out2=48_000), _wb(n, seed=41), C))
    cs.append(FmCase("wbfm_P2_fir", FmParams(downsample=4, downsample_passes=2, comp_fir_size=9, custom_atan=ATAN_FAST,
                                             deemph=1, deemph_a=46, rate_out=600_000, rate_out2=48_000), _wb(n, seed=42), C))
    cs.append(FmCase("wbfm_P2_nofir", FmParams(downsample=4, downsample_passes=2, comp_fir_size=0, custom_atan=ATAN_FAST,
                                               deemph=1, deemph_a=46, rate_out=600_000, rate_out2=48_000), _wb(n, seed=43), 65536))
    # -F 0: half-band passes without droop FIR, P=1 and P=5
    cs.append(FmCase("F0_P1", FmParams(downsample=2, downsample_passes=1, comp_fir_size=0, custom_atan=ATAN_FAST,
                                       rate_out=1_200_000, rate_out2=48_000), _wb(n, seed=3), C))
    cs.append(FmCase("F9_P5_lut", FmParams(downsample=32, downsample_passes=5, comp_fir_size=9, custom_atan=ATAN_LUT,
                                           deemph=1, deemph_a=13, rate_out=75_000, rate_out2=25_000),
                     _wb(n, seed=4, amp=0.05), C))
    # default wbfm preset: -s 170k, D=6, fast, deemph a=13, -r 32k (/5)
    cs.append(FmCase("wbfm_default", FmParams(downsample=6, custom_atan=ATAN_FAST, deemph=1, deemph_a=13,
                                              rate_out=170_000, rate_out2=32_000), _wb(n, seed=5), C))
    # default nbfm: -s 24k, D=42 (not a divisor of the chunk), lut and std
    cs.append(FmCase("nbfm_D42_lut", FmParams(downsample=42, custom_atan=ATAN_LUT, rate_out=24_000),
                     _nb(n, 1.008e6, 6), C))
    cs.append(FmCase("nbfm_D42_std", FmParams(downsample=42, custom_atan=ATAN_STD, rate_out=24_000),
                     _nb(n, 1.008e6, 6), C, exact=False))
    # cfg5A: D=100 boxcar, lut, 2.4 Msps -> 24 k
    cs.append(FmCase("cfg5A", FmParams(downsample=100, custom_atan=ATAN_LUT, rate_out=24_000),
                     lambda: synth.cfg5_iq(n, 3), C))
    # cfg5B: -F 0 P=6 (/64 -> 37.5 k) + resample to 24 k... integer ratio 37500/24000 = 1
    cs.append(FmCase("cfg5B", FmParams(downsample=64, downsample_passes=6, comp_fir_size=0, custom_atan=ATAN_FAST,
                                       rate_out=37_500, rate_out2=24_000), lambda: synth.cfg5_iq(n, 4), C))
    # offset tuning (no rotation), eu de-emphasis (a even), no resampler
    cs.append(FmCase("offset_deemph_even_a", FmParams(downsample=4, custom_atan=ATAN_FAST, deemph=1, deemph_a=16,
                                                      rate_out=300_000, offset_tuning=1),
                     lambda: synth.fm_iq(n, fs=1.2e6, deviation_hz=75e3, tones=[(1000.0, 1.0)],
                                         amplitude=12000, noise_lsb=64, seed=7, center_hz=0.0), C))
    # full-range uniform noise: pins int16 wrap / int32 wrap behaviour (fast path overflows)
    cs.append(FmCase("fullscale_noise_P3", FmParams(downsample=8, downsample_passes=3, comp_fir_size=9,
                                                    custom_atan=ATAN_FAST, deemph=1, deemph_a=23,
                                                    rate_out=300_000, rate_out2=48_000),
                     lambda: synth.uniform_iq(n, -32768, 32767, 8), C))
    cs.append(FmCase("fullscale_noise_D1_lut", FmParams(downsample=1, custom_atan=ATAN_LUT, rate_out=1_000_000,
                                                        rate_out2=50_000),
                     lambda: synth.uniform_iq(n, -32768, 32767, 9), C))
    # quiet input: all zeros, and a constant -> de-emphasis dead zone (lo/hi never meet)
    cs.append(FmCase("zeros_deemph", FmParams(downsample=1, custom_atan=ATAN_FAST, deemph=1, deemph_a=181,
                                              rate_out=2_400_000, rate_out2=48_000),
                     lambda: np.zeros(2 * n, dtype=np.int16), C))
    cs.append(FmCase("burst_then_silence", FmParams(downsample=8, downsample_passes=3, comp_fir_size=9,
                                                    custom_atan=ATAN_FAST, deemph=1, deemph_a=23,
                                                    rate_out=300_000, rate_out2=48_000),
                     lambda: np.concatenate([_wb(n // 4, seed=11)(), np.zeros(2 * (n - n // 4), dtype=np.int16)]), C))
    # a murmur: PCM wanders by a few tens of LSB, so brackets stay open and only sometimes move or meet
    # (exercises every piece kind of the back end's chain: exact / merged / pass-through / open)
    cs.append(FmCase("murmur_deemph", FmParams(downsample=8, downsample_passes=3, comp_fir_size=9,
                                               custom_atan=ATAN_FAST, deemph=1, deemph_a=23, rate_out=300_000,
                                               rate_out2=48_000),
                     lambda: synth.fm_iq(n, fs=2.4e6, deviation_hz=260.0, tones=[(31.0, 1.0), (5.0, 0.6)],
                                         amplitude=14000, noise_lsb=0, seed=17), C))
    # ragged: stream not a multiple of the chunk; short last chunk
    cs.append(FmCase("ragged_tail", FmParams(downsample=8, downsample_passes=3, comp_fir_size=9,
                                             custom_atan=ATAN_FAST, deemph=1, deemph_a=23, rate_out=300_000,
                                             rate_out2=48_000), _wb(n - 40 * 1024, seed=12), C))
    # other demodulators
    cs.append(FmCase("am_D4", FmParams(mode=MODE_AM, downsample=4, output_scale=64, rate_out=250_000,
                                       rate_out2=50_000), _wb(n, seed=13), C))
    cs.append(FmCase("usb_D8", FmParams(mode=MODE_USB, downsample=8, output_scale=32, rate_out=125_000),
                     _wb(n, seed=14), C))
    cs.append(FmCase("lsb_P2", FmParams(mode=MODE_LSB, downsample=4, downsample_passes=2, comp_fir_size=9,
                                        output_scale=64, rate_out=250_000), _wb(n, seed=15), C))
    cs.append(FmCase("raw_D4", FmParams(mode=MODE_RAW, downsample=4, output_scale=64, rate_out=250_000),
                     _wb(n, seed=16), C))
    return cs


def fm_optional_cases(scale: int = 1) -> list[FmCase]:
    """Optional stages (squelch, DC blocks, -o): per-chunk reductions."""
    n = (1 << 19) * scale
    C = 262144
    base = dict(downsample=8, downsample_passes=3, comp_fir_size=9, custom_atan=ATAN_FAST, deemph=1, deemph_a=23,
                rate_out=300_000, rate_out2=48_000)
    cs = []
    cs.append(FmCase("rdc", FmParams(dc_block_raw=1, **base), _wb(n, seed=21), C))
    cs.append(FmCase("adc", FmParams(dc_block_audio=1, **base), _wb(n, seed=22), C))
    cs.append(FmCase("squelch", FmParams(squelch_level=60, **base),
                     lambda: np.concatenate([_wb(n // 2, seed=23)(), _wb(n // 2, seed=24, amp=0.002)()]), C))
    cs.append(FmCase("post_ds4", FmParams(downsample=2, custom_atan=ATAN_FAST, post_downsample=4, deemph=1,
                                          deemph_a=13, rate_out=150_000, rate_out2=50_000), _wb(n, seed=25), C))
    return cs


# ------------------------------------------------------------------------- rx_power
@dataclass
class PowerCase:
    name: str
    freq_arg: str              # -f lower:upper:bin
    crop: float = 0.0          # -c
    boxcar: int = 1            # 0 with -F
    comp_fir_size: int = 0     # -F arg
    peak_hold: int = 0         # -P
    window: str = "rectangle"  # -w (or "hann": host-built custom table)
    n_pass: int = 2
    noise: int = 100
    tone_amp: float = 50.0
    seed: int = 777
    fullscale: bool = False


def power_cases() -> list[PowerCase]:
    return [
        PowerCase("cfg3_hann_poisson", "100M:101M:1k", window="hann-poisson", n_pass=6),
        PowerCase("cfg3_hamming", "100M:101M:1k", window="hamming", n_pass=6),
        PowerCase("cfg3_hann_custom", "100M:101M:1k", window="hann", n_pass=6),
        PowerCase("cfg3_rect_fullscale", "100M:101M:1k", window="rectangle", n_pass=3, fullscale=True, seed=778),
        PowerCase("cfg3_blackman_fullscale", "100M:101M:1k", window="blackman", n_pass=3, fullscale=True, seed=779),
        PowerCase("cfg3_peak_hold", "100M:101M:1k", window="hamming", n_pass=4, peak_hold=1),
        PowerCase("cfg4_small", "24M:60M:1k", crop=0.285, window="hamming", n_pass=2, seed=4000),
        PowerCase("n256_youssef", "100M:102M:8k", window="youssef", n_pass=3),
        PowerCase("n16384", "100M:102M:100", window="blackman-harris", n_pass=2),
        PowerCase("boxcar_ds28", "100M:100.1M:100", window="bartlett", n_pass=3),
        PowerCase("F9_passes4", "100M:100.1M:100", boxcar=0, comp_fir_size=9, window="hamming", n_pass=3),
        PowerCase("F0_passes2", "100M:100.5M:1k", boxcar=0, comp_fir_size=0, window="kaiser", n_pass=3),
        PowerCase("rms_bins", "100M:110M:1M", n_pass=4),
        PowerCase("rms_bins_peak", "100M:110M:1M", n_pass=4, peak_hold=1, fullscale=True),
    ]


def power_input(case: PowerCase, n_hops: int, buf_len: int) -> np.ndarray:
    if case.fullscale:
        rng = np.random.default_rng(case.seed)
        return rng.integers(-20000, 20001, size=(case.n_pass, n_hops, buf_len), dtype=np.int32).astype(np.int16)
    return synth.power_hops(case.n_pass, n_hops, buf_len, seed=case.seed, noise=case.noise,
                            tones=((0.11, case.tone_amp), (-0.27, case.tone_amp)))
