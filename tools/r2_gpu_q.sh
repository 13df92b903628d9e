#!/bin/bash
# round 2, session Q: SPEC 3 (multi-channel NBFM boxcar shape specialised at compile time) A/B on fm5a
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2q; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  %-12s %8.0f Msamples/s  frac %.4f  kernel_ms %.4f  %s %s" % (sys.argv[2], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["kernel"], d.get("detail", "")))
except Exception as e:
    print("  %-12s no line: %s" % (sys.argv[2], e))
PY
}
B="--no-extras --no-cpu --no-e2e --steps 20 --warmup 5"
timeout 400 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_golden.py tests/test_full_size_gpu.py tests/test_dropin.py -x -q -m gpu > $OUT/tests.txt 2>&1; echo "tests rc=$?"; tail -3 $OUT/tests.txt
timeout 200 python bench.py $B --workload fm5a > $OUT/bench_fm5a.json 2> $OUT/bench_fm5a.err; line $OUT/bench_fm5a.json fm5a-spec3
RXB200_FM_NOSPEC3=1 timeout 200 python bench.py $B --workload fm5a > $OUT/bench_fm5a_spec0.json 2> $OUT/bench_fm5a_spec0.err; line $OUT/bench_fm5a_spec0.json fm5a-spec0
timeout 200 python bench.py $B --workload fm5a --size-mib 293 > $OUT/bench_fm5a_32.json 2> $OUT/bench_fm5a_32.err; line $OUT/bench_fm5a_32.json fm5a-32ch
for t in 128 256; do RXB200_FM_THREADS=$t timeout 200 python bench.py $B --workload fm5a > $OUT/bench_fm5a_t$t.json 2>/dev/null; line $OUT/bench_fm5a_t$t.json fm5a-T$t; done
timeout 300 python tools/seg_sweep.py 256 1000 3000 200 > $OUT/seg_sweep_256.txt 2>&1; cat $OUT/seg_sweep_256.txt
timeout 300 ncu --set full --clock-control none -k regex:fm_fused -c 1 -o /tmp/prof_fm5a -f \
	python bench.py --workload fm5a --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fm5a.log 2>&1; echo "ncu fm5a rc=$?"
ncu -i /tmp/prof_fm5a.ncu-rep --page raw --csv > $OUT/raw_fm5a.csv 2>/dev/null
date
