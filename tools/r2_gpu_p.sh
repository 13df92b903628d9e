#!/bin/bash
# round 2, session P: compute-sanitizer (memcheck, racecheck) over the final build: rows kernel (P = 3 with FIR, P = 2, P = 1 without
# FIR), the fused kernel, the 512-thread rx_power kernel, the global-memory FFT path, the multi-item look-back (small items).
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2p; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date
cat > /tmp/san.py <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import oracle
from rx_tools_b200 import fm, power, synth
from cases import fm_cases
port = oracle.port()
for name, seg in (("cfg2B", 0), ("cfg2B", 4096), ("wbfm_P2_fir", 2048), ("F0_P1", 0), ("wbfm_default", 0), ("cfg2A", 0)):
    c = next(x for x in fm_cases() if x.name == name)
    x = c.make_input()[:2 * 131072]
    d = fm.FmDemod(c.params)
    if seg: d.tune(segment_len=seg)
    got = d.full_demod(x, 2 * 32768); want = port.fm_run(c.params, x, 2 * 32768)
    print(name, seg, d.stats()["kernel"], bool(np.array_equal(got, want))); d.close()
plan = power.plan_range("24M:60M:1k", 0.285); win = power.window_table("hamming", 1 << plan.bin_e)
hb = synth.power_hops(3, plan.n_hops, plan.buf_len, seed=1); sc = power.PowerScanner(plan, win); sc.scanner(hb, 3); a, s = sc.read()
wa, ws = port.power_scan(oracle.PowerParams(bin_e=plan.bin_e, buf_len=plan.buf_len), win, hb, 3, plan.n_hops); print("power 4096 x18", bool(np.array_equal(a, wa))); sc.close()
plan = power.plan_range("100M:102.8M:40"); win = power.window_table("blackman", 1 << plan.bin_e)
hb = np.random.default_rng(3).integers(-3000, 3001, size=(1, 1, plan.buf_len), dtype=np.int32).astype(np.int16)
sc = power.PowerScanner(plan, win); sc.scanner(hb, 1); a, s = sc.read()
wa, ws = port.power_scan(oracle.PowerParams(bin_e=plan.bin_e, buf_len=plan.buf_len), win, hb, 1, 1); print("power bin_e 17", bool(np.array_equal(a, wa))); sc.close()
PY
timeout 600 compute-sanitizer --tool memcheck python /tmp/san.py > $OUT/memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -12 $OUT/memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python /tmp/san.py > $OUT/racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -12 $OUT/racecheck.txt
timeout 600 compute-sanitizer --tool synccheck python /tmp/san.py > $OUT/synccheck.txt 2>&1; echo "synccheck rc=$?"; tail -4 $OUT/synccheck.txt
timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu > $OUT/bench_clockcheck.json 2>$OUT/bench_clockcheck.err; python -c "
import json; d=json.loads(open('$OUT/bench_clockcheck.json').read().strip().splitlines()[-1]); print('clocks', d['clocks'], 'value', round(d['value']))"
date
