#!/usr/bin/env python
"""fm2b shape, device-resident 1 GiB: kernel time against the two run-time knobs of the back end -- the de-emphasis
replay length (rxb200_fm_tune) and the number of back-end lanes (RXB200_FM_BE_LANES) -- with the number of pieces
whose bracket stayed open.  Output must not change with either knob; the first run's PCM is the yardstick."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rx_tools_b200 import fm, synth  # noqa: E402

A = 23


def run(d_in, n16, warm, lanes, ref):
    if lanes:
        os.environ["RXB200_FM_BE_LANES"] = str(lanes)
    else:
        os.environ.pop("RXB200_FM_BE_LANES", None)
    p = fm.FmParams(downsample=8, downsample_passes=3, comp_fir_size=9, custom_atan=fm.ATAN_FAST, deemph=1, deemph_a=A,
                    rate_out=300_000, rate_out2=48_000)
    dem = fm.FmDemod(p)
    dem.tune(0, warm)
    cap = dem.max_output(n16, 262144) + 8
    out = torch.zeros(cap, dtype=torch.int16, device="cuda")
    best = 1e9
    for _ in range(4):
        dem.reset()
        dem.process_device(d_in.data_ptr(), n16, 262144, out.data_ptr(), cap, sync=True)
        best = min(best, dem.kernel_ms())
    same = True if ref is None else bool(torch.equal(out, ref))
    print(f"warm {warm or 16 * A + 64:4d} lanes {lanes or 'auto':>4}  {n16 / 2 / best / 1e6:8.1f} Gsamples/s  {best:7.4f} ms  "
          f"open pieces {dem.stats()['fixup_segments']:6d}  same output {same}", flush=True)
    dem.close()
    return out


if __name__ == "__main__":
    period = torch.from_numpy(synth.cfg2_iq(1 << 24)).cuda()
    d_in = period.repeat(16).contiguous()                 # 1 GiB
    n16 = d_in.numel()
    ref = run(d_in, n16, 0, 0, None)
    for warm in (14 * A + 56, 12 * A + 48, 10 * A + 40, 8 * A + 32):
        run(d_in, n16, warm, 0, ref)
    for lanes in (64, 96, 128, 160, 192):
        run(d_in, n16, 0, lanes, ref)
