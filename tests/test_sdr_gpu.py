"""rx_sdr sample-format conversions against the port oracle: all 65 536 int16 inputs for the
CS16 targets, random bytes for the CS12 unpack, ragged lengths."""
import ctypes as C

import numpy as np
import pytest

import oracle
from rx_tools_b200 import fm

pytestmark = pytest.mark.gpu


def _port(name, src, dst):
    L = oracle.port().L
    f = getattr(L, name)
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    f.restype = None
    f(src.ctypes.data, src.size if name != "orx_sdr_cs12_to_cs16" else src.size // 3, dst.ctypes.data)
    return dst


@pytest.mark.parametrize("n_complex", [32768, 32767, 5, 1])
def test_cs16_targets(n_complex):
    v = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    rng = np.random.default_rng(n_complex)
    x = np.concatenate([v, rng.integers(-32768, 32768, size=65536, dtype=np.int32).astype(np.int16)])[: 2 * n_complex]
    x = np.ascontiguousarray(x)
    assert np.array_equal(fm.sdr_convert(fm.CVT_CS16_CS8, x), _port("orx_sdr_cs16_to_cs8", x, np.empty(x.size, np.uint8)))
    assert np.array_equal(fm.sdr_convert(fm.CVT_CS16_CU8, x), _port("orx_sdr_cs16_to_cu8", x, np.empty(x.size, np.uint8)))
    got = fm.sdr_convert(fm.CVT_CS16_CF32, x)
    want = _port("orx_sdr_cs16_to_cf32", x, np.empty(x.size, np.float32))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))        # bit-identical floats


def test_cs12_unpack():
    rng = np.random.default_rng(12)
    b = rng.integers(0, 256, size=3 * 10007, dtype=np.int32).astype(np.uint8)
    assert np.array_equal(fm.sdr_convert(fm.CVT_CS12_CS16, b), _port("orx_sdr_cs12_to_cs16", b, np.empty(2 * 10007, np.int16)))


def test_against_the_reference_executable_golden():
    """tests/golden/sdr_golden.json: what the reference's own rx_sdr (oracle/_ref/rx_sdr_ref, src/rtl_sdr.c's main()
    recording from the replay device) wrote for these captures -- every int16 value, and random packed CS12."""
    import json
    import os
    import sdr_inputs as SI
    from rx_tools_b200.synth import digest
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdr_golden.json")))
    x = np.ascontiguousarray(SI.cs16_capture()[: 2 * SI.N_ELEMS])
    for kind, key in ((fm.CVT_CS16_CS8, "CS16_CS8"), (fm.CVT_CS16_CU8, "CS16_CU8"), (fm.CVT_CS16_CF32, "CS16_CF32")):
        got = fm.sdr_convert(kind, x)
        assert got.nbytes == g[key]["n_bytes"]
        assert digest(got.view(np.uint8)) == g[key]["output_sha256"], key
    y = np.ascontiguousarray(SI.cs12_capture()[: 3 * SI.N_ELEMS_12])
    got = fm.sdr_convert(fm.CVT_CS12_CS16, y)
    assert digest(got.view(np.uint8)) == g["CS12_CS16"]["output_sha256"]
