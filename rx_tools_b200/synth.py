"""Deterministic synthetic CS16 I/Q generators.

The reference ships no sample captures, so every test vector and bench workload is minted
here from a fixed seed.  All generators return interleaved little-endian int16 (I0,Q0,I1,Q1..)
exactly as SoapySDR's CS16 ``readStream`` delivers it (src/rtl_fm.c:894, src/rtl_power.c:694).
Phase is computed in closed form so a long stream can be produced block by block with no state.
"""
from __future__ import annotations

from typing import Iterable, Sequence, Tuple

import numpy as np

Tone = Tuple[float, float]  # (audio frequency Hz, relative weight)


def digest(data) -> str:
    """sha256 over the raw little-endian bytes — the hash stored in tests/golden/*.json."""
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(data).tobytes()).hexdigest()


def fm_iq(n_complex: int, fs: float, deviation_hz: float, tones: Sequence[Tone], amplitude: float,
          noise_lsb: int, seed: int, center_hz: float | None = None, start: int = 0) -> np.ndarray:
    """Constant-envelope FM at ``center_hz`` (default -fs/4, which rotate16_90 brings to DC,
    SURVEY §9) modulated by a sum of audio tones, plus uniform integer noise in
    [-noise_lsb, +noise_lsb] on I and Q.  ``start`` lets a long stream be generated piecewise."""
    if center_hz is None:
        center_hz = -fs / 4.0
    n = np.arange(start, start + n_complex, dtype=np.float64)
    t = n / fs
    wsum = float(sum(w for _, w in tones)) or 1.0
    phase = 2.0 * np.pi * center_hz * t
    for f, w in tones:
        # integral of (w/wsum) * cos(2 pi f t) * 2 pi dev
        phase += (deviation_hz * (w / wsum) / f) * np.sin(2.0 * np.pi * f * t)
    rng = np.random.default_rng([seed, start])
    out = np.empty(2 * n_complex, dtype=np.int16)
    i = amplitude * np.cos(phase)
    q = amplitude * np.sin(phase)
    if noise_lsb > 0:
        nz = rng.integers(-noise_lsb, noise_lsb + 1, size=2 * n_complex, dtype=np.int32)
        i = i + nz[0::2]
        q = q + nz[1::2]
    out[0::2] = np.clip(np.rint(i), -32768, 32767).astype(np.int16)
    out[1::2] = np.clip(np.rint(q), -32768, 32767).astype(np.int16)
    return out


def fm_iq_stream(n_complex: int, block: int = 1 << 22, **kw) -> Iterable[np.ndarray]:
    pos = 0
    while pos < n_complex:
        m = min(block, n_complex - pos)
        yield fm_iq(m, start=pos, **kw)
        pos += m


def uniform_iq(n_complex: int, lo: int, hi: int, seed: int) -> np.ndarray:
    """Full-range uniform noise, used to pin wrap/overflow semantics."""
    rng = np.random.default_rng(seed)
    return rng.integers(lo, hi + 1, size=2 * n_complex, dtype=np.int32).astype(np.int16)


def power_hops(n_pass: int, n_hops: int, buf_len: int, seed: int, noise: int = 100,
               tones: Sequence[Tuple[float, float]] = ((0.11, 50.0), (-0.27, 50.0)),
               per_hop_seed: bool = True) -> np.ndarray:
    """Hop buffers int16[n_pass][n_hops][buf_len]: uniform noise in [-noise, noise] plus complex
    tones given as (cycles/sample, amplitude).  With per_hop_seed every hop gets seed+hop so
    rows differ and a mis-ordered gather is detectable."""
    out = np.empty((n_pass, n_hops, buf_len), dtype=np.int16)
    k = np.arange(buf_len // 2, dtype=np.float64)
    for h in range(n_hops):
        rng = np.random.default_rng([seed + (h if per_hop_seed else 0), 7])
        base_i = np.zeros(buf_len // 2)
        base_q = np.zeros(buf_len // 2)
        for f, a in tones:
            ff = f + 0.013 * (h % 17)
            base_i += a * np.cos(2 * np.pi * ff * k)
            base_q += a * np.sin(2 * np.pi * ff * k)
        for p in range(n_pass):
            nz = rng.integers(-noise, noise + 1, size=buf_len, dtype=np.int32)
            out[p, h, 0::2] = np.clip(np.rint(base_i + nz[0::2]), -32768, 32767).astype(np.int16)
            out[p, h, 1::2] = np.clip(np.rint(base_q + nz[1::2]), -32768, 32767).astype(np.int16)
    return out


# ---- the named BASELINE.json workloads -----------------------------------
def cfg1_iq(n_complex: int = 1 << 20, seed: int = 12345) -> np.ndarray:
    """cfg1: NBFM at 1.024 Msps, 1 kHz tone, +-5 kHz deviation, amplitude 16000, noise +-64."""
    return fm_iq(n_complex, fs=1_024_000.0, deviation_hz=5000.0, tones=[(1000.0, 1.0)],
                 amplitude=16000.0, noise_lsb=64, seed=seed)


def cfg2_iq(n_complex: int, seed: int = 2, start: int = 0) -> np.ndarray:
    """cfg2: WBFM-like at 2.4 Msps, +-75 kHz deviation, 400 Hz + 3 kHz + 11 kHz audio,
    amplitude 0.45 FS (keeps fast_atan2 inside int32 after 3 half-band passes), noise +-64."""
    return fm_iq(n_complex, fs=2_400_000.0, deviation_hz=75000.0,
                 tones=[(400.0, 1.0), (3000.0, 0.7), (11000.0, 0.4)],
                 amplitude=0.45 * 32767.0, noise_lsb=64, seed=seed, start=start)


def cfg5_iq(n_complex: int, channel: int, seed: int = 5000) -> np.ndarray:
    """cfg5: NBFM channel at 2.4 Msps, +-5 kHz deviation, amplitude 0.2 FS, seed = 5000+channel."""
    return fm_iq(n_complex, fs=2_400_000.0, deviation_hz=5000.0,
                 tones=[(700.0 + 37.0 * (channel % 16), 1.0)],
                 amplitude=0.2 * 32767.0, noise_lsb=64, seed=seed + channel)
