"""Captures for the rx_sdr conversion pin (src/rtl_sdr.c:348-391): shared by tests/golden/make_golden.py, which runs the
reference's own rx_sdr executable (oracle/_ref/rx_sdr_ref) over them, and by the tests.  The recorder only stops on a
read that over-delivers (:341-346), so a capture holds more elements than are recorded and N is not a multiple of the
block (16384)."""
import numpy as np

N_ELEMS = 66536            # complex elements recorded from the CS16 capture: every int16 value at least once
N_ELEMS_12 = 20500


def cs16_capture() -> np.ndarray:
    v = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    r = np.random.default_rng(77).integers(-32768, 32768, size=65536, dtype=np.int32).astype(np.int16)
    return np.ascontiguousarray(np.concatenate([v, r, v[::-1]]))        # 98 304 complex elements


def cs12_capture() -> np.ndarray:
    return np.random.default_rng(12).integers(0, 256, size=3 * 40000, dtype=np.int32).astype(np.uint8)
