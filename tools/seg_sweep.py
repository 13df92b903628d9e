#!/usr/bin/env python
"""Segment-length sweep of the fused kernel's per-thread front end (rxb200_fm_tune): kernel time against Sf for a
multi-channel boxcar shape.  usage: python tools/seg_sweep.py n_channels lo hi step"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rx_tools_b200 import fm, synth
import bench

n_ch, lo, hi, step = (int(a) for a in sys.argv[1:5])
p = bench.fm_params("fm5a")
n_per = bench.FM5A_PER
period = n_per - n_per % (bench.CHUNK // 2)
x = torch.from_numpy(synth.cfg5_iq(period, 0)).cuda()
d_in = x.repeat(n_ch).contiguous()
dem = fm.FmDemod(p, n_channels=n_ch)
n16 = 2 * period
cap = dem.max_output(n16, bench.CHUNK) + 8
d_out = torch.empty(n_ch * cap, dtype=torch.int16, device="cuda")
for seg in range(lo, hi + 1, step):
    dem.tune(segment_len=seg)
    ms = []
    for _ in range(4):
        dem.process_device(d_in.data_ptr(), n16, bench.CHUNK, d_out.data_ptr(), cap, sync=False)
        ms.append(dem.kernel_ms())
    st = dem.stats()
    print(seg, st["segment_len"], st["segments"] // 128, "%.4f" % min(ms[1:]), "%.0f" % (n_ch * period / min(ms[1:]) / 1e3), flush=True)
