"""The BASELINE.json full-size inputs, shared by tests/golden/make_full_golden.py (which runs the
UNMODIFIED reference over them in the authoring container) and tests/test_full_size_gpu.py (which runs the CUDA path
over them on the GPU box and compares hashes).  Pure numpy, seeded, no reference needed."""
import numpy as np

from rx_tools_b200 import synth

CHUNK16 = 262144                          # MAXIMUM_BUF_LENGTH int16 = 131072 complex (src/rtl_fm.c:80-82)
FM_PERIOD = 1 << 24                       # complex samples of the seeded period that is tiled to 1 GiB
FM_TILES = 16
CFG5_CHANNELS = 256
CFG5_BASE = 300_000                       # complex samples generated per channel, tiled 8x -> 2.4 M
CFG5_TILES = 8
CFG4_PASSES = 36
CFG3_BUFFERS = 610


def fm_one_gib_period() -> np.ndarray:
    """cfg2: the 64 MiB period; the 1 GiB stream is this tiled FM_TILES times (2 048 chunks)."""
    return synth.cfg2_iq(FM_PERIOD)


def cfg1_input() -> np.ndarray:
    return synth.cfg1_iq(1 << 20)          # 2^20 complex -> 24 576 PCM


def cfg5_channel(ch: int) -> np.ndarray:
    return np.tile(synth.cfg5_iq(CFG5_BASE, ch), CFG5_TILES)


def cfg4_hops(plan_n_hops: int, buf_len: int) -> np.ndarray:
    """int16[36][871][16384], every hop and every sweep different (seed 4000 + hop)."""
    return synth.power_hops(CFG4_PASSES, plan_n_hops, buf_len, seed=4000)


def cfg3_hops(buf_len: int) -> np.ndarray:
    """int16[610][1][16384]: 10 s of hop buffers at 1 Msps, samples = 4 880."""
    return synth.power_hops(CFG3_BUFFERS, 1, buf_len, seed=777)
