"""Pin the port (oracle/rx_oracle.c) against the UNMODIFIED reference compiled into oracle/_ref.

The reference ships no tests or golden vectors, so the reference code itself,
executed, is the pin.  These tests need oracle/_ref (built here from /root/reference)."""
import dataclasses
import os

import numpy as np
import pytest

import oracle
from cases import fm_cases, fm_optional_cases

pytestmark = pytest.mark.ref


def test_tables_match_reference(port, ref_fm):
    assert np.array_equal(port.atan_table(), ref_fm.atan_table())
    for row in range(11):
        assert np.array_equal(port.droop9(row), ref_fm.cic9(row))


def test_scale_integer_form(port):
    # (int16)(x/32767.0*128.0+0.4) == trunc((1280x+131068)/327670) for all x
    tab = port.scale_table()
    x = np.arange(-32768, 32768, dtype=np.int64)
    num = 1280 * x + 131068
    q = np.where(num >= 0, num // 327670, -((-num) // 327670))
    assert np.array_equal(q.astype(np.int16), tab)
    assert tab.min() == -127 and tab.max() == 128


@pytest.mark.parametrize("case", fm_cases() + fm_optional_cases(), ids=lambda c: c.name)
def test_fm_port_equals_reference(case, port, ref_fm):
    x = case.make_input()
    a, la, ha = port.fm_run(case.params, x, case.chunk_int16, return_chunks=True)
    b, lb, hb = ref_fm.run(case.params, x, case.chunk_int16, return_chunks=True)
    assert np.array_equal(la, lb)
    assert np.array_equal(ha, hb)
    assert a.size == b.size
    assert np.array_equal(a, b)          # same libm, same machine: bit-exact incl. atan2 path


@pytest.mark.parametrize("case", (fm_cases() + fm_optional_cases())[::2], ids=lambda c: c.name)
def test_fm_levels_port_equals_reference(case, port, ref_fm):
    # the per-chunk `sr` behind the -L statistics (src/rtl_fm.c:792-806)
    x = case.make_input()
    a = port.fm_levels(case.params, x, case.chunk_int16)
    b = ref_fm.levels(case.params, x, case.chunk_int16)
    assert a.size == b.size and a.size >= 1
    assert np.array_equal(a, b)
    assert b.max() > 0 or case.name.startswith("zeros")      # an all-zero capture has level 0


def test_derivation_matches_optimal_settings(ref_fm):
    from rx_tools_b200 import fm
    combos = [dict(rate_s=1024000, rate_r=24000), dict(wbfm=1), dict(wbfm=1, rate_s=2400000, rate_r=48000),
              dict(wbfm=1, rate_s=300000, rate_r=48000, use_F=1, comp_fir_size=9), dict(rate_s=24000),
              dict(rate_s=24000, custom_atan=2), dict(mode=oracle.MODE_AM, rate_s=12000),
              dict(mode=oracle.MODE_USB, rate_s=48000, use_F=1, comp_fir_size=0),
              dict(wbfm=1, time_constant_us=50), dict(rate_s=170000, post_downsample=4, deemph=1)]
    for kw in combos:
        want, cap_rate, cap_off = ref_fm.derive(**kw)
        got = fm.derive_params(**kw)
        mine = dataclasses.asdict(got.params)
        assert mine.pop("report_levels") == 0          # library-only switch, not a reference field
        assert mine == dataclasses.asdict(want), (kw, got.params, want)
        assert got.capture_rate == cap_rate and got.capture_freq_offset == cap_off, kw


# ------------------------------------------------------------------------- rx_power
from cases import power_cases, power_input  # noqa: E402


def _window_table(port, case, n):
    return port.window_table(case.window, n)


@pytest.mark.parametrize("case", power_cases(), ids=lambda c: c.name)
def test_power_port_equals_reference(case, port, ref_power):
    n_guess = None
    custom = None
    if case.window == "hann":
        # need N first: plan once with any window
        plan0 = ref_power.setup(case.freq_arg, case.crop, case.boxcar, case.comp_fir_size, case.peak_hold, "rectangle")
        custom = port.window_table("hann", 1 << plan0.bin_e)
    plan = ref_power.setup(case.freq_arg, case.crop, case.boxcar, case.comp_fir_size, case.peak_hold,
                           case.window if custom is None else "rectangle", custom)
    n = 1 << plan.bin_e
    win_ref, sine_ref = ref_power.tables()
    win = _window_table(port, case, n)
    assert np.array_equal(win, win_ref)
    if plan.bin_e > 0:
        assert np.array_equal(port.sine_table(plan.bin_e), sine_ref[: n * 3 // 4])
    x = power_input(case, plan.tune_count, plan.buf_len)
    avg_r, smp_r = ref_power.scan(x, case.n_pass)
    p = oracle.PowerParams(bin_e=plan.bin_e, buf_len=plan.buf_len, downsample=plan.downsample,
                           downsample_passes=plan.downsample_passes, comp_fir_size=case.comp_fir_size,
                           boxcar=case.boxcar, peak_hold=case.peak_hold)
    avg_p, smp_p = port.power_scan(p, win, x, case.n_pass, plan.tune_count)
    assert np.array_equal(smp_p, smp_r)
    assert np.array_equal(avg_p, avg_r)
    assert avg_r.any()


@pytest.mark.parametrize("m", [1, 2, 5, 10, 12])
def test_fix_fft_port_equals_reference(m, port, ref_power):
    rng = np.random.default_rng(m)
    iq = rng.integers(-32768, 32768, size=2 << m, dtype=np.int32).astype(np.int16)
    assert np.array_equal(port.fix_fft(iq, m), ref_power.fix_fft(iq, m))
    # smaller transform inside a larger sine table (fix_fft allows n < N_WAVE)
    if m > 2:
        iq2 = iq[: 2 << (m - 2)]
        assert np.array_equal(port.fix_fft(iq2, m - 2, m), ref_power.fix_fft(iq2, m - 2, m))


# ---- rx_sdr conversions (src/rtl_sdr.c:348-391): they live inline in main(), so the pin is the reference's own
# executable (oracle/_ref/rx_sdr_ref) recording from the replay device
def _port_sdr(port, name, src, dst, count):
    import ctypes as C
    f = getattr(port.L, name)
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    f.restype = None
    f(src.ctypes.data, count, dst.ctypes.data)
    return dst


@pytest.mark.ref
def test_sdr_conversions_port_equals_reference_executable(port):
    import oracle
    import sdr_inputs as SI
    if not os.path.exists(oracle.REF_SDR_BIN):
        pytest.skip("oracle/_ref/rx_sdr_ref not built (no /root/reference here)")
    x = SI.cs16_capture()
    n = SI.N_ELEMS
    for fmt, name, dt in (("CS8", "orx_sdr_cs16_to_cs8", np.uint8), ("CU8", "orx_sdr_cs16_to_cu8", np.uint8),
                          ("CF32", "orx_sdr_cs16_to_cf32", np.float32)):
        ref = np.frombuffer(oracle.ref_rx_sdr(x, "CS16", fmt, n), dtype=np.uint8)
        mine = _port_sdr(port, name, np.ascontiguousarray(x[:2 * n]), np.empty(2 * n, dt), 2 * n)
        assert np.array_equal(ref, mine.view(np.uint8)), fmt
    y = SI.cs12_capture()
    n12 = SI.N_ELEMS_12
    ref = np.frombuffer(oracle.ref_rx_sdr(y, "CS12", "CS16", n12), dtype=np.int16)
    assert np.array_equal(ref, _port_sdr(port, "orx_sdr_cs12_to_cs16", y, np.empty(2 * n12, np.int16), n12))


def test_sdr_conversions_port_equals_golden(port):
    import json
    import sdr_inputs as SI
    from rx_tools_b200.synth import digest
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdr_golden.json")))
    x, n = SI.cs16_capture(), SI.N_ELEMS
    assert digest(x) == g["CS16_CS8"]["input_sha256"]
    for fmt, name, dt in (("CS8", "orx_sdr_cs16_to_cs8", np.uint8), ("CU8", "orx_sdr_cs16_to_cu8", np.uint8),
                          ("CF32", "orx_sdr_cs16_to_cf32", np.float32)):
        mine = _port_sdr(port, name, np.ascontiguousarray(x[:2 * n]), np.empty(2 * n, dt), 2 * n)
        assert digest(mine.view(np.uint8)) == g["CS16_" + fmt]["output_sha256"], fmt
    y, n12 = SI.cs12_capture(), SI.N_ELEMS_12
    mine = _port_sdr(port, "orx_sdr_cs12_to_cs16", y, np.empty(2 * n12, np.int16), n12)
    assert digest(mine.view(np.uint8)) == g["CS12_CS16"]["output_sha256"]


@pytest.mark.ref
@pytest.mark.parametrize("freq", ["100M:102.8M:40", "100M:100.4M:2"])
def test_power_port_equals_reference_beyond_65536_bins(freq, port, ref_power):
    """bin_e 17 / 18 (hop buffers of 0.5 / 7 MB, the second one boxcar-decimated by 7): the shapes the library serves from
    global memory (tests/test_power_gpu.py::test_hop_buffers_beyond_shared_memory)."""
    import oracle
    rp = ref_power.setup(freq, 0.0, 1, 0, 0, "blackman")
    assert rp.bin_e >= 17
    rng = np.random.default_rng(rp.bin_e)
    x = rng.integers(-3000, 3001, size=(2, rp.tune_count, rp.buf_len), dtype=np.int32).astype(np.int16)
    avg, smp = ref_power.scan(x, 2)
    win, _ = ref_power.tables()
    pp = oracle.PowerParams(bin_e=rp.bin_e, buf_len=rp.buf_len, downsample=rp.downsample, downsample_passes=rp.downsample_passes)
    a2, s2 = port.power_scan(pp, win, x, 2, rp.tune_count)
    assert np.array_equal(smp, s2) and np.array_equal(avg, a2)
