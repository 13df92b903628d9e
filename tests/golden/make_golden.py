"""Mint tests/golden/*.json from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the authoring container only:  python tests/golden/make_golden.py
Each entry pins: the seeded input (sha256), the reference output (sha256, length, head samples,
per-chunk result_len).  The GPU box has no /root/reference: tests there check the port and the
CUDA path against these files."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from cases import fm_cases, fm_optional_cases, power_cases, power_input  # noqa: E402
from rx_tools_b200.synth import digest  # noqa: E402


def main():
    oracle.build()
    assert oracle.have_ref(), "needs /root/reference"
    rf, rp, port = oracle.RefFm(), oracle.RefPower(), oracle.port()
    fm = {}
    for c in fm_cases() + fm_optional_cases():
        x = c.make_input()
        y, lens, hits = rf.run(c.params, x, c.chunk_int16, return_chunks=True)
        fm[c.name] = dict(input_sha256=digest(x), n_in=int(x.size), chunk_int16=c.chunk_int16,
                          output_sha256=digest(y), n_out=int(y.size), head=[int(v) for v in y[:16]],
                          tail=[int(v) for v in y[-8:]], chunk_result_len=[int(v) for v in lens],
                          squelch_hits=[int(v) for v in hits], exact=c.exact)
    with open(os.path.join(HERE, "fm_golden.json"), "w") as f:
        json.dump(fm, f, indent=1, sort_keys=True)
    pw = {}
    for c in power_cases():
        custom = None
        if c.window == "hann":
            p0 = rp.setup(c.freq_arg, c.crop, c.boxcar, c.comp_fir_size, c.peak_hold, "rectangle")
            custom = port.window_table("hann", 1 << p0.bin_e)
        plan = rp.setup(c.freq_arg, c.crop, c.boxcar, c.comp_fir_size, c.peak_hold,
                        c.window if custom is None else "rectangle", custom)
        x = power_input(c, plan.tune_count, plan.buf_len)
        avg, smp = rp.scan(x, c.n_pass)
        win, _ = rp.tables()
        pw[c.name] = dict(input_sha256=digest(x), tune_count=plan.tune_count, bin_e=plan.bin_e, buf_len=plan.buf_len,
                          downsample=plan.downsample, downsample_passes=plan.downsample_passes, rate=plan.rate,
                          window_sha256=digest(win.astype(np.int32)), avg_sha256=digest(avg),
                          avg_head=[int(v) for v in avg.reshape(-1)[:8]], samples=[int(v) for v in smp[:4]],
                          csv_sha256=digest(np.frombuffer(rp.csv("/tmp/_golden.csv").encode(), dtype=np.uint8)))
    with open(os.path.join(HERE, "power_golden.json"), "w") as f:
        json.dump(pw, f, indent=1, sort_keys=True)
    # two small literal vectors (input AND output) so the pin does not depend on numpy's RNG
    rng = np.random.default_rng(99)
    x = rng.integers(-32768, 32768, size=2 * 4096, dtype=np.int32).astype(np.int16)
    p = oracle.FmParams(downsample=8, downsample_passes=3, comp_fir_size=9, custom_atan=1, deemph=1, deemph_a=23,
                        rate_out=300000, rate_out2=48000)
    y = rf.run(p, x, 2048)
    hb = rng.integers(-3000, 3001, size=(2, 1, 16384), dtype=np.int32).astype(np.int16)
    rp.setup("100M:101M:1k", 0.0, 1, 0, 0, "hamming")
    avg, smp = rp.scan(hb, 2)
    np.savez_compressed(os.path.join(HERE, "literal_vectors.npz"), fm_in=x, fm_out=y, fm_chunk=np.int64(2048),
                        pw_in=hb, pw_avg=avg, pw_samples=smp)
    # rx_sdr (src/rtl_sdr.c:348-391): the reference's own executable over a capture that holds every int16 value
    import sdr_inputs as SI
    sd = {}
    x = SI.cs16_capture()
    for fmt in ("CS8", "CU8", "CF32"):
        b = oracle.ref_rx_sdr(x, "CS16", fmt, SI.N_ELEMS)
        sd["CS16_" + fmt] = dict(input_sha256=digest(x), n_elems=SI.N_ELEMS, output_sha256=digest(np.frombuffer(b, dtype=np.uint8)), n_bytes=len(b))
    y = SI.cs12_capture()
    b = oracle.ref_rx_sdr(y, "CS12", "CS16", SI.N_ELEMS_12)
    sd["CS12_CS16"] = dict(input_sha256=digest(y), n_elems=SI.N_ELEMS_12, output_sha256=digest(np.frombuffer(b, dtype=np.uint8)), n_bytes=len(b))
    with open(os.path.join(HERE, "sdr_golden.json"), "w") as f:
        json.dump(sd, f, indent=1, sort_keys=True)
    print("wrote", len(fm), "fm,", len(pw), "power and", len(sd), "sdr golden entries")


if __name__ == "__main__":
    main()
