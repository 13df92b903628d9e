"""The port oracle (oracle/rx_oracle.c) against the reference-minted FULL-SIZE goldens (tests/golden/full_golden.json):
the restatement is pinned at BASELINE.json's sizes too, not only on the 2^19-sample cases.  CPU only."""
import json
import os

import numpy as np
import pytest

import full_inputs as FI
import oracle
from rx_tools_b200.synth import digest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = json.load(open(os.path.join(G, "full_golden.json")))
VEC = np.load(os.path.join(G, "full_std_vectors.npz"))


def test_port_fm2b_one_gib(port):
    g = GOLD["fm2b"]
    period = FI.fm_one_gib_period()
    assert digest(period) == g["period_sha256"]
    y, lens, _ = port.fm_run(oracle.FmParams(**g["params"]), np.tile(period, FI.FM_TILES), FI.CHUNK16, return_chunks=True)
    assert y.size == g["n_out"]
    assert digest(lens.astype(np.int32)) == g["chunk_result_len_sha256"]
    assert digest(y) == g["output_sha256"]


@pytest.mark.parametrize("nm", ["std", "fast", "lut"])
def test_port_cfg1(nm, port):
    g = GOLD[f"cfg1_{nm}"]
    y = port.fm_run(oracle.FmParams(**g["params"]), FI.cfg1_input(), FI.CHUNK16)
    assert y.size == 24576 and digest(y) == g["output_sha256"]        # same libm as the reference run: exact, atan2 included
    if nm == "std":
        assert np.array_equal(y, VEC["cfg1_std_full"])


def test_port_cfg5a_some_channels(port):
    g = GOLD["cfg5A"]
    for ch, sha in g["channel_sha256"].items():
        y = port.fm_run(oracle.FmParams(**g["params"]), FI.cfg5_channel(int(ch)), FI.CHUNK16)
        assert y.size == g["n_out_per_channel"] and digest(y) == sha


def test_port_cfg3(port):
    g = GOLD["cfg3"]
    hb = FI.cfg3_hops(16384)
    assert digest(hb) == g["input_sha256"]
    avg, smp = port.power_scan(oracle.PowerParams(bin_e=10, buf_len=16384), port.window_table("hann", 1024), hb, FI.CFG3_BUFFERS, 1)
    assert int(smp[0]) == 4880 and digest(avg) == g["avg_sha256"]
