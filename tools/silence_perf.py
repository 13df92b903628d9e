#!/usr/bin/env python
"""Time rx_fm (fm2b shape) on inputs that keep the de-emphasis brackets open: all zeros (a closed squelch),
a murmur, and the normal wbfm test signal for comparison.  Prints Gsamples/s and the number of pieces
the back end's chain had to hand a start state to."""
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rx_tools_b200 import fm, synth  # noqa: E402


def run(name, period, reps, mib=256):
    p = fm.FmParams(downsample=8, downsample_passes=3, comp_fir_size=9, custom_atan=fm.ATAN_FAST, deemph=1, deemph_a=23,
                    rate_out=300_000, rate_out2=48_000)
    d_in = torch.from_numpy(period).cuda().repeat(reps).contiguous()
    n16 = d_in.numel()
    dem = fm.FmDemod(p)
    cap = dem.max_output(n16, 262144) + 8
    out = torch.empty(cap, dtype=torch.int16, device="cuda")
    for _ in range(2):
        dem.reset()
        dem.process_device(d_in.data_ptr(), n16, 262144, out.data_ptr(), cap, sync=True)
    ms = dem.kernel_ms()
    print(f"{name:10s} {n16 / 2 / ms / 1e6:8.1f} Gsamples/s  kernel {ms:8.3f} ms  chained pieces {dem.stats()['fixup_segments']}")
    dem.close()


if __name__ == "__main__":
    n = 1 << 24                                            # 64 MiB period
    run("wbfm", synth.cfg2_iq(n), 4)
    run("zeros", np.zeros(2 * n, dtype=np.int16), 4)
    run("murmur", synth.fm_iq(n, fs=2.4e6, deviation_hz=260.0, tones=[(31.0, 1.0), (5.0, 0.6)], amplitude=14000,
                              noise_lsb=0, seed=17), 4)
