"""BASELINE.json's full sizes (1 GiB per step) through size-independent properties: the oracle cannot run 268 M
samples in test time, so the big run is pinned by (a) its prefix equalling the oracle on the first chunks,
(b) split-call idempotence (state carry across calls == one call), (c) rx_power: exact additivity over passes
and a prefix of hops/passes equalling the oracle."""
import numpy as np
import pytest

import oracle
from rx_tools_b200 import fm, power, synth

pytestmark = pytest.mark.gpu
CHUNK16 = 262144


def _torch():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA")
    return torch


def test_fm_one_gib_prefix_and_split_calls(port):
    torch = _torch()
    p = fm.derive_params(wbfm=1, rate_s=300000, rate_r=48000, use_F=1, comp_fir_size=9).params
    period = synth.cfg2_iq(1 << 24)                       # 64 MiB, tiled 16x -> 1 GiB
    d_in = torch.from_numpy(period).cuda().repeat(16).contiguous()
    n16 = d_in.numel()
    assert n16 * 2 == 1 << 30
    dem = fm.FmDemod(p)
    cap = dem.max_output(n16, CHUNK16) + 8
    out1 = torch.empty(cap, dtype=torch.int16, device="cuda")
    n1 = dem.process_device(d_in.data_ptr(), n16, CHUNK16, out1.data_ptr(), cap, sync=True)
    assert n1 == 5368709                                  # 2^28 samples -> /8 -> *48/300
    full = out1[:n1].cpu().numpy()
    # (a) prefix == oracle on the first 8 chunks
    k = 8 * CHUNK16
    want = port.fm_run(oracle.FmParams(**p.reference_fields()), period[:k], CHUNK16)
    assert np.array_equal(full[:want.size], want)
    # (b) two calls (split on a chunk boundary) == one call
    dem.reset()
    half = (n16 // 2 // CHUNK16) * CHUNK16
    out2 = torch.empty(cap, dtype=torch.int16, device="cuda")
    na = dem.process_device(d_in.data_ptr(), half, CHUNK16, out2.data_ptr(), cap, sync=True)
    nb = dem.process_device(d_in.data_ptr() + half * 2, n16 - half, CHUNK16, out2.data_ptr() + na * 2, cap - na, sync=True)
    assert na + nb == n1
    assert torch.equal(out1[:n1], out2[:n1])
    # the tiled period has a discontinuity every 2^24 samples; nothing but exactness is assumed about it
    dem.close()


def test_power_one_gib_additivity_and_prefix(port):
    torch = _torch()
    plan = power.plan_range("24M:1766M:1k", 0.285)
    assert plan.n_hops == 871
    win = power.window_table("hamming", 4096)
    n_pass = 36                                           # 36 x 871 x 32 KiB = 1.03 GB
    base = synth.power_hops(2, plan.n_hops, plan.buf_len, seed=4000)
    d_base = torch.from_numpy(base.reshape(-1)).cuda()
    d_in = d_base.repeat(n_pass // 2).contiguous()
    sc = power.PowerScanner(plan, win)
    sc.scanner_device(d_in.data_ptr(), n_pass, sync=True)
    avg_all, smp_all = sc.read()
    assert int(smp_all[0]) == n_pass * 2 and np.all(smp_all == smp_all[0])
    # additivity: the same passes in two batches accumulate to the same int64 rows
    sc.reset()
    sc.scanner_device(d_in.data_ptr(), 10, sync=True)
    sc.scanner_device(d_in.data_ptr() + 10 * plan.n_hops * plan.buf_len * 2, n_pass - 10, sync=True)
    avg_two, _ = sc.read()
    assert np.array_equal(avg_all, avg_two)
    # the input is the same 2 sweeps repeated 18 times: rows must be 18 x the oracle's rows for those 2 sweeps
    pp = oracle.PowerParams(bin_e=plan.bin_e, buf_len=plan.buf_len)
    hops = [0, 1, 435, 870]
    sub = np.ascontiguousarray(base[:, hops])
    want, _ = port.power_scan(pp, win, sub, 2, len(hops))
    assert np.array_equal(avg_all[hops], want * (n_pass // 2))
    sc.close()
