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
