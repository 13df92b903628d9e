"""Host-side arithmetic of librxb200 (planner, tables, CSV formatter, scale identity) against the
port oracle and the golden files.  CPU only."""
import json
import os

import numpy as np
import pytest

from cases import power_cases
from rx_tools_b200 import power
from rx_tools_b200.synth import digest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PW_GOLD = json.load(open(os.path.join(G, "power_golden.json")))


@pytest.mark.parametrize("case", power_cases(), ids=lambda c: c.name)
def test_planner_matches_reference_plan(case):
    g = PW_GOLD[case.name]
    plan = power.plan_range(case.freq_arg, case.crop, case.boxcar, case.comp_fir_size, case.peak_hold)
    assert (plan.n_hops, plan.bin_e, plan.buf_len, plan.downsample, plan.downsample_passes, plan.rate) == \
        (g["tune_count"], g["bin_e"], g["buf_len"], g["downsample"], g["downsample_passes"], g["rate"])
    if plan.bin_e:
        assert digest(power.window_table(case.window, 1 << plan.bin_e)) == g["window_sha256"]


def test_cfg4_plan_is_871_hops():
    plan = power.plan_range("24M:1766M:1k", 0.285)
    assert (plan.n_hops, plan.bin_e, plan.buf_len, plan.rate) == (871, 12, 16384, 2797202)


@pytest.mark.parametrize("log2n", [1, 4, 10, 12, 15])
def test_sine_table_matches_port(log2n, port):
    assert np.array_equal(power.sine_table(log2n), port.sine_table(log2n))


@pytest.mark.parametrize("name", sorted(power.WINDOWS))
def test_window_tables_match_port(name, port):
    for n in (2, 256, 1024, 4096):
        assert np.array_equal(power.window_table(name, n), port.window_table(name, n))


@pytest.mark.ref
def test_planner_and_csv_match_reference(ref_power, port):
    for arg, crop, boxcar, fir in [("24M:1766M:1k", 0.285, 1, 0), ("24M:1766M:1k", 0.0, 1, 0), ("88M:108M:10k", 0.1, 1, 0),
                                   ("100M:100.2M:50", 0.0, 0, 9), ("433M:434M:100", 0.2, 1, 0), ("100M:120M:2M", 0, 1, 0)]:
        rp = ref_power.setup(arg, crop, boxcar, fir, 0, "hamming")
        plan = power.plan_range(arg, crop, boxcar, fir, 0)
        assert (plan.n_hops, plan.bin_e, plan.buf_len, plan.downsample, plan.downsample_passes, plan.rate) == \
            (rp.tune_count, rp.bin_e, rp.buf_len, rp.downsample, rp.downsample_passes, rp.rate), arg
        assert abs(plan.crop - rp.crop) < 1e-15
        freqs = ref_power.hop_freqs()
        assert [plan.hop_freq(i) for i in range(plan.n_hops)] == [int(f) for f in freqs]
    # CSV text identical for the same accumulators
    rp = ref_power.setup("24M:60M:1k", 0.285, 1, 0, 0, "hamming")
    rng = np.random.default_rng(5)
    x = rng.integers(-100, 101, size=(2, rp.tune_count, rp.buf_len), dtype=np.int32).astype(np.int16)
    avg, smp = ref_power.scan(x, 2)
    plan = power.plan_range("24M:60M:1k", 0.285)
    mine = power.csv_rows(plan, avg, smp)
    theirs = ref_power.csv("/tmp/_csv_ref.txt")
    assert mine == theirs


def test_wbfm_preset_keeps_a_later_squelch_level():
    """-M wbfm zeroes squelch_level when it is parsed (src/rtl_fm.c:1331-1341); `-M wbfm -l 50` keeps 50.  The parse-order
    rule lives in the shell (host/rx_fm_b200.c), rxb200_fm_derive passes cli.squelch_level through."""
    from rx_tools_b200 import fm
    assert fm.derive_params(wbfm=1, squelch_level=50).params.squelch_level == 50
    assert fm.derive_params(wbfm=1).params.squelch_level == 0


def test_deemph_fp32_identity():
    """The CUDA back end runs deemph_filter (src/rtl_fm.c:667-682) for odd a as  U' = fma(X - U, fl(1/a), U)  in FP32
    with U = 2^23 + 32768 + avg (csrc/fm_kernels.cu, DeemphOp<false, true>).  That is exact iff rounding d * fl(1/a) to
    the nearest integer equals the reference's trunc((d +- a/2) / a) for every difference d of two int16 values: checked
    here for every d and a spread of odd a (the kernels' own parity tests cover a = 23 and a = 181 on the GPU)."""
    d = np.arange(-65535, 65536, dtype=np.int64)
    for a in list(range(1, 400, 2)) + [1023, 1801, 4095, 18001, 32767]:
        inv = np.float64(np.float32(1.0) / np.float32(a))
        h = a // 2
        num = np.where(d > 0, d + h, d - h)
        ref = np.sign(num) * (np.abs(num) // a)            # C's truncating division
        y = d.astype(np.float64) * inv                     # exact: 17 x 24 significant bits
        assert np.array_equal(np.rint(y).astype(np.int64), ref), a
        # no tie in sight: the nearest half-integer is further away than any rounding of the fma could reach
        assert np.min(np.abs((y - np.floor(y)) - 0.5)) > 65536 * 2.0 ** -24 / a
