"""rx_power path — host-side mirror of the reference's interface (src/rtl_power.c).

``plan_range`` is ``frequency_range()`` (:431-543), ``window_table``/``sine_table`` are the tables
``main()`` builds (:1028, :1034-1037), ``PowerScanner.scanner`` is n passes of ``scanner()`` (:670-772)
over hop buffers, executed by the batched sm_100a kernel behind ``rxb200_power_accumulate``
(include/rxb200.h); ``csv_rows`` is ``csv_dbm()`` (:774-817).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib

WINDOWS = {"rectangle": 0, "hamming": 1, "blackman": 2, "blackman-harris": 3, "hann-poisson": 4,
           "youssef": 5, "kaiser": 6, "bartlett": 7, "hann": 8}


def atofs(s: str) -> float:
    """'k'/'M'/'G' suffixed number (convenience.c:65-90)."""
    s = s.strip()
    mult = {"g": 1e9, "G": 1e9, "m": 1e6, "M": 1e6, "k": 1e3, "K": 1e3}.get(s[-1:] if s else "", None)
    if mult is not None:
        return float(s[:-1]) * mult
    return float(s)


@dataclass
class Plan:
    n_hops: int
    bin_e: int
    buf_len: int
    downsample: int
    downsample_passes: int
    comp_fir_size: int
    boxcar: int
    peak_hold: int
    rate: int
    crop: float
    first_freq: int
    freq_step: int
    bin_size_hz: float

    def to_c(self) -> _lib.PowerParamsC:
        return _lib.PowerParamsC(self.n_hops, self.bin_e, self.buf_len, self.downsample, self.downsample_passes,
                                 self.comp_fir_size, self.boxcar, self.peak_hold)

    def hop_freq(self, i: int) -> int:
        return self.first_freq + i * self.freq_step


def plan_range(freq_arg: str, crop: float = 0.0, boxcar: int = 1, comp_fir_size: int = 0, peak_hold: int = 0) -> Plan:
    lo, hi, step = freq_arg.split(":")
    out = _lib.PowerPlanC()
    _lib.check(_lib.lib().rxb200_power_plan_range(int(atofs(lo)), int(atofs(hi)), int(atofs(step)), crop, boxcar,
                                                  comp_fir_size, peak_hold, C.byref(out)))
    p = out.params
    return Plan(p.n_hops, p.bin_e, p.buf_len, p.downsample, p.downsample_passes, p.comp_fir_size, p.boxcar,
                p.peak_hold, out.rate, out.crop, out.first_freq, out.freq_step, out.bin_size_hz)


def window_table(name_or_id, length: int) -> np.ndarray:
    wid = WINDOWS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
    t = np.empty(length, dtype=np.int32)
    _lib.check(_lib.lib().rxb200_window_table(wid, length, t.ctypes.data_as(C.POINTER(C.c_int))))
    return t


def sine_table(log2_n: int) -> np.ndarray:
    t = np.empty(max((1 << log2_n) * 3 // 4, 1), dtype=np.int16)
    _lib.check(_lib.lib().rxb200_sine_table(log2_n, t.ctypes.data_as(C.POINTER(C.c_int16))))
    return t


class PowerScanner:
    def __init__(self, plan: Plan, window: Sequence[int] | str = "rectangle", device: int = 0,
                 sinewave: Optional[np.ndarray] = None):
        self.plan = plan
        n = 1 << plan.bin_e
        self.window = window_table(window, n) if isinstance(window, str) else np.ascontiguousarray(window, np.int32)
        assert self.window.size == n
        self._h = C.c_void_p()
        pc = plan.to_c()
        sw = None if sinewave is None else np.ascontiguousarray(sinewave, np.int16).ctypes.data_as(C.POINTER(C.c_int16))
        _lib.check(_lib.lib().rxb200_power_create(C.byref(pc), self.window.ctypes.data_as(C.POINTER(C.c_int)), sw,
                                                  device, C.byref(self._h)))

    def close(self) -> None:
        if _lib is not None and getattr(self, "_h", None) is not None and self._h:
            _lib.lib().rxb200_power_destroy(self._h)
            self._h = None

    __del__ = close

    def scanner(self, hop_bufs: np.ndarray, n_pass: int, hop_begin: int = 0, hop_end: Optional[int] = None) -> None:
        """hop_bufs: int16[n_pass][hop_end-hop_begin][buf_len] (host)."""
        hop_end = self.plan.n_hops if hop_end is None else hop_end
        hb = np.ascontiguousarray(hop_bufs, dtype=np.int16).reshape(-1)
        assert hb.size == n_pass * (hop_end - hop_begin) * self.plan.buf_len
        _lib.check(_lib.lib().rxb200_power_accumulate(self._h, hb.ctypes.data, n_pass, hop_begin, hop_end))

    def scanner_device(self, d_ptr: int, n_pass: int, hop_begin: int = 0, hop_end: Optional[int] = None,
                       sync: bool = False) -> None:
        hop_end = self.plan.n_hops if hop_end is None else hop_end
        _lib.check(_lib.lib().rxb200_power_accumulate_device(self._h, d_ptr, n_pass, hop_begin, hop_end,
                                                             1 if sync else 0))

    def read(self):
        n = 1 << self.plan.bin_e
        avg = np.zeros((self.plan.n_hops, n), dtype=np.int64)
        samples = np.zeros(self.plan.n_hops, dtype=np.int32)
        _lib.check(_lib.lib().rxb200_power_read(self._h, avg.ctypes.data_as(C.POINTER(C.c_int64)),
                                                samples.ctypes.data_as(C.POINTER(C.c_int))))
        return avg, samples

    def read_db(self):
        """csv_dbm's numbers computed on the device (src/rtl_power.c:783-811): float64[n_hops][row_len], samples."""
        row_len = _lib.check(_lib.lib().rxb200_power_row_len(self.plan.bin_e, self.plan.crop))
        db = np.zeros((self.plan.n_hops, row_len), dtype=np.float64)
        samples = np.zeros(self.plan.n_hops, dtype=np.int32)
        _lib.check(_lib.lib().rxb200_power_read_db(self._h, self.plan.rate, self.plan.crop,
                                                   db.ctypes.data_as(C.POINTER(C.c_double)), row_len,
                                                   samples.ctypes.data_as(C.POINTER(C.c_int))))
        return db, samples

    def csv_rows_device(self, tstr: str = "2026-01-01, 00:00:00") -> str:
        """The CSV text of csv_dbm() over every hop with the dB values computed on the device."""
        db, samples = self.read_db()
        buf = C.create_string_buffer(64 + 16 * (db.shape[1] + 8))
        lines = []
        for i in range(self.plan.n_hops):
            row = np.ascontiguousarray(db[i])
            r = _lib.check(_lib.lib().rxb200_power_format_db_row(row.ctypes.data_as(C.POINTER(C.c_double)), self.plan.bin_e,
                                                                 self.plan.hop_freq(i), self.plan.rate, self.plan.downsample,
                                                                 self.plan.crop, int(samples[i]), buf, len(buf)))
            lines.append(tstr + ", " + buf.raw[:r].decode())
        return "".join(lines)

    def reset(self) -> None:
        _lib.check(_lib.lib().rxb200_power_reset(self._h))

    def kernel_ms(self) -> float:
        ms = C.c_float(0)
        _lib.check(_lib.lib().rxb200_power_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def gather(self, comm: "Comm", sync: bool = True) -> None:
        """The collation before the report (src/rtl_power.c:1047-1050 walks every hop): ONE in-place NCCL all-gather of
        the accumulator rows inside the library; afterwards read()/read_db() see every rank's hops."""
        _lib.check(_lib.lib().rxb200_power_gather(self._h, comm._c, 1 if sync else 0))

    @property
    def device_avg_ptr(self) -> int:
        return int(_lib.lib().rxb200_power_device_avg(self._h) or 0)

    @property
    def stream(self) -> int:
        return int(_lib.lib().rxb200_power_stream(self._h) or 0)

    def csv_rows(self, avg: np.ndarray, samples: np.ndarray, tstr: str = "2026-01-01, 00:00:00") -> str:
        """csv_dbm() over every hop (src/rtl_power.c:1047-1050): returns the CSV text."""
        return csv_rows(self.plan, avg, samples, tstr)


def csv_rows(plan: Plan, avg: np.ndarray, samples: np.ndarray, tstr: str = "2026-01-01, 00:00:00") -> str:
    n = 1 << plan.bin_e
    buf = C.create_string_buffer(64 + 16 * (n + 8))
    lines = []
    a = np.array(avg, dtype=np.int64, copy=True).reshape(plan.n_hops, n)
    for i in range(plan.n_hops):
        row = np.ascontiguousarray(a[i])
        r = _lib.check(_lib.lib().rxb200_power_format_row(row.ctypes.data_as(C.POINTER(C.c_int64)), plan.bin_e,
                                                          plan.hop_freq(i), plan.rate, plan.downsample, plan.crop,
                                                          int(samples[i]), buf, len(buf)))
        lines.append(tstr + ", " + buf.raw[:r].decode())
    return "".join(lines)


def shard(n_hops: int, n_ranks: int, rank: int):
    """Contiguous hop range [begin, end) of `rank` (rxb200_power_shard)."""
    b, e = C.c_int(0), C.c_int(0)
    _lib.check(_lib.lib().rxb200_power_shard(n_hops, n_ranks, rank, C.byref(b), C.byref(e)))
    return b.value, e.value


class Comm:
    """One rank's NCCL communicator, owned by librxb200 (include/rxb200.h: rxb200_comm_*)."""

    def __init__(self, n_ranks: int, rank: int, unique_id: bytes, device: int):
        assert len(unique_id) == 128
        self._c = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        _lib.check(_lib.lib().rxb200_comm_create(n_ranks, rank, buf, device, C.byref(self._c)))
        self.n_ranks, self.rank = n_ranks, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(_lib.lib().rxb200_comm_unique_id(buf))
        return buf.raw

    def close(self) -> None:
        if _lib is not None and getattr(self, "_c", None) is not None and self._c:
            _lib.lib().rxb200_comm_destroy(self._c)
            self._c = None

    __del__ = close


class PowerGroup:
    """All ranks in one process: n_dev GPUs, hops sharded, one all-gather before the report (what the drop-in
    rx_power_b200 shell uses when RXB200_GPUS > 1)."""

    def __init__(self, plan: Plan, window: Sequence[int] | str = "rectangle", n_dev: int = 1):
        self.plan = plan
        n = 1 << plan.bin_e
        self.window = window_table(window, n) if isinstance(window, str) else np.ascontiguousarray(window, np.int32)
        self._g = C.c_void_p()
        pc = plan.to_c()
        _lib.check(_lib.lib().rxb200_power_group_create(C.byref(pc), self.window.ctypes.data_as(C.POINTER(C.c_int)), None,
                                                        n_dev, None, C.byref(self._g)))
        self.n_dev = n_dev

    def close(self) -> None:
        if _lib is not None and getattr(self, "_g", None) is not None and self._g:
            _lib.lib().rxb200_power_group_destroy(self._g)
            self._g = None

    __del__ = close

    def scanner(self, hop_bufs: np.ndarray, n_pass: int, hop_begin: int = 0, hop_end: Optional[int] = None) -> None:
        hop_end = self.plan.n_hops if hop_end is None else hop_end
        hb = np.ascontiguousarray(hop_bufs, dtype=np.int16).reshape(-1)
        assert hb.size == n_pass * (hop_end - hop_begin) * self.plan.buf_len
        _lib.check(_lib.lib().rxb200_power_group_accumulate(self._g, hb.ctypes.data, n_pass, hop_begin, hop_end))

    def gather(self) -> None:
        _lib.check(_lib.lib().rxb200_power_group_gather(self._g))

    def read(self, member: int = 0):
        n = 1 << self.plan.bin_e
        avg = np.zeros((self.plan.n_hops, n), dtype=np.int64)
        samples = np.zeros(self.plan.n_hops, dtype=np.int32)
        h = _lib.lib().rxb200_power_group_member(self._g, member)
        _lib.check(_lib.lib().rxb200_power_read(h, avg.ctypes.data_as(C.POINTER(C.c_int64)),
                                                samples.ctypes.data_as(C.POINTER(C.c_int))))
        return avg, samples

    def reset(self) -> None:
        _lib.check(_lib.lib().rxb200_power_group_reset(self._g))
