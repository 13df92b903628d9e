"""Port oracle vs the committed golden vectors (minted from the unmodified reference by
tests/golden/make_golden.py).  Runs anywhere (no GPU, no /root/reference)."""
import json
import os

import numpy as np
import pytest

import oracle
from cases import fm_cases, fm_optional_cases, power_cases, power_input
from rx_tools_b200.synth import digest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FM_GOLD = json.load(open(os.path.join(G, "fm_golden.json")))
PW_GOLD = json.load(open(os.path.join(G, "power_golden.json")))


@pytest.mark.parametrize("case", fm_cases() + fm_optional_cases(), ids=lambda c: c.name)
def test_fm_port_matches_golden(case, port):
    g = FM_GOLD[case.name]
    x = case.make_input()
    assert digest(x) == g["input_sha256"], "synthetic generator drifted"
    y, lens, hits = port.fm_run(case.params, x, case.chunk_int16, return_chunks=True)
    assert [int(v) for v in lens] == g["chunk_result_len"]
    assert [int(v) for v in hits] == g["squelch_hits"]
    assert y.size == g["n_out"]
    assert [int(v) for v in y[:16]] == g["head"]
    assert digest(y) == g["output_sha256"]


@pytest.mark.parametrize("case", power_cases(), ids=lambda c: c.name)
def test_power_port_matches_golden(case, port):
    g = PW_GOLD[case.name]
    n = 1 << g["bin_e"]
    x = power_input(case, g["tune_count"], g["buf_len"])
    assert digest(x) == g["input_sha256"], "synthetic generator drifted"
    win = port.window_table(case.window, n)
    assert digest(win.astype(np.int32)) == g["window_sha256"]
    p = oracle.PowerParams(bin_e=g["bin_e"], buf_len=g["buf_len"], downsample=g["downsample"],
                           downsample_passes=g["downsample_passes"], comp_fir_size=case.comp_fir_size,
                           boxcar=case.boxcar, peak_hold=case.peak_hold)
    avg, smp = port.power_scan(p, win, x, case.n_pass, g["tune_count"])
    assert [int(v) for v in smp[:4]] == g["samples"]
    assert digest(avg) == g["avg_sha256"]


def test_literal_vectors(port):
    z = np.load(os.path.join(G, "literal_vectors.npz"))
    p = oracle.FmParams(downsample=8, downsample_passes=3, comp_fir_size=9, custom_atan=1, deemph=1, deemph_a=23,
                        rate_out=300000, rate_out2=48000)
    assert np.array_equal(port.fm_run(p, z["fm_in"], int(z["fm_chunk"])), z["fm_out"])
    pp = oracle.PowerParams(bin_e=10, buf_len=16384)
    avg, smp = port.power_scan(pp, port.window_table("hamming", 1024), z["pw_in"], 2, 1)
    assert np.array_equal(avg, z["pw_avg"]) and np.array_equal(smp, z["pw_samples"])
