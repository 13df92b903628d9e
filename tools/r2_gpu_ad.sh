#!/bin/bash
# session AD: FP32 discriminator in the row front end (warp-uniform range check, integer form as the other branch) --
# parity, fm2b three times (noise), the fused kernel, fm2a
OUT=gpurun_out/r2ad; mkdir -p $OUT
exec > $OUT/session.log 2>&1
date
timeout 900 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_full_size_gpu.py tests/test_dropin.py -m gpu -x -q > $OUT/tests.txt 2>&1
echo "tests rc=$?"; tail -4 $OUT/tests.txt
run() { # name, workload, env...
	local name=$1 wl=$2; shift; shift
	env "$@" timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-extras > $OUT/bench_$name.json 2> $OUT/bench_$name.err
	python - $OUT/bench_$name.json $name <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  %-18s %8.0f Msamples/s  frac %.4f  kernel_ms %.4f  %s" % (sys.argv[2], r["value"], r["roofline"]["frac"], r["roofline"]["kernel_ms"], r["roofline"]["kernel"]))
except Exception as e:
    print("  %-18s FAILED %s" % (sys.argv[2], e))
PY
}
run fm2b_1 fm2b X=1
run fm2b_2 fm2b X=1
run fm2b_3 fm2b X=1
run fm2b_fused fm2b RXB200_FM_NOROWS=1
run fm2a fm2a X=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_split -c 1 -o /tmp/prof_fm2b -f python bench.py --workload fm2b --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
ncu -i /tmp/prof_fm2b.ncu-rep --page raw --csv > $OUT/raw_fm2b.csv 2>/dev/null
echo "ncu rc=$?"
date
