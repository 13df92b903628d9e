#!/usr/bin/env python
"""rx_fm shapes against the CTA width (RXB200_FM_THREADS, honoured by the boxcar kernels) on a device-resident stream:
calibrates fm_cta_threads() in csrc/fm_kernels.cu.  Also times a front-end-only shape (no de-emphasis, no resampler:
the kernel stores the discriminator output itself) next to the full wbfm chain -- the gap is what the serial stages cost."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rx_tools_b200 import fm, synth  # noqa: E402


def run(name, params, d_in, n16, width=None, chunk16=262144, reps=3):
    if width:
        os.environ["RXB200_FM_THREADS"] = str(width)
    else:
        os.environ.pop("RXB200_FM_THREADS", None)
    dem = fm.FmDemod(params)
    cap = dem.max_output(n16, chunk16) + 8
    out = torch.zeros(cap, dtype=torch.int16, device="cuda")
    best = 1e9
    for _ in range(reps):
        dem.reset()
        dem.process_device(d_in.data_ptr(), n16, chunk16, out.data_ptr(), cap, sync=True)
        best = min(best, dem.kernel_ms())
    st = dem.stats()
    print(f"{name:44s} width {width or 'rule':>4}  {n16 / 2 / best / 1e6:8.1f} Gsamples/s  {best:8.4f} ms  seg {st['segment_len']}", flush=True)
    dem.close()


if __name__ == "__main__":
    period = torch.from_numpy(synth.cfg2_iq(1 << 24)).cuda()
    d_in = period.repeat(8).contiguous()                 # 512 MiB
    n16 = d_in.numel()
    # the full wbfm chain and its front end alone (P = 3, droop FIR, fast atan)
    full = fm.FmParams(downsample=8, downsample_passes=3, comp_fir_size=9, custom_atan=fm.ATAN_FAST, deemph=1, deemph_a=23,
                       rate_out=300_000, rate_out2=48_000)
    fe_only = fm.FmParams(downsample=8, downsample_passes=3, comp_fir_size=9, custom_atan=fm.ATAN_FAST, deemph=0, deemph_a=23,
                          rate_out=300_000, rate_out2=0)
    run("wbfm P=3 full chain", full, d_in, n16)
    run("wbfm P=3 front end only", fe_only, d_in, n16)
    # boxcar shapes: D sweep, wbfm-like back end (de-emphasis at rate_in, resampler to 48 k)
    for D in (1, 2, 4, 6, 8, 16, 42, 100):
        rate = 2_400_000 // D
        a = max(3, int(round(1.0 / (1.0 - pow(2.718281828459045, -1.0 / (rate * 75e-6))))))
        # fast_atan2 leaves its no-overflow range once the boxcar gain is large: the LUT discriminator from D = 16 on
        atan = fm.ATAN_FAST if D <= 8 else fm.ATAN_LUT
        p = fm.FmParams(downsample=D, downsample_passes=0, comp_fir_size=0, custom_atan=atan, deemph=1, deemph_a=a,
                        rate_out=rate, rate_out2=min(48_000, rate))
        for w in (128, 256):
            run(f"boxcar D={D} a={a} {'fast' if D <= 8 else 'lut'} + deemph + resample", p, d_in, n16, w)
    # NBFM without serial stages (fm5a-like, one channel): lut atan, boxcar
    for D in (42, 100):
        p = fm.FmParams(downsample=D, downsample_passes=0, comp_fir_size=0, custom_atan=fm.ATAN_LUT, deemph=0, deemph_a=1,
                        rate_out=2_400_000 // D, rate_out2=0)
        for w in (128, 256):
            run(f"boxcar D={D} lut direct", p, d_in, n16, w)
