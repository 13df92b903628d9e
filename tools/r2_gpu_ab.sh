#!/bin/bash
# session AB: loops back to one quad of read-ahead (session Z), quads as single 8-byte shared-memory loads -- parity, fm2a, fm2b, ncu
OUT=gpurun_out/r2ab; mkdir -p $OUT
exec > $OUT/session.log 2>&1
date
timeout 900 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > $OUT/tests.txt 2>&1
echo "tests rc=$?"; tail -4 $OUT/tests.txt
run() { # name, workload, env...
	local name=$1 wl=$2; shift; shift
	env "$@" timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras > $OUT/bench_$name.json 2> $OUT/bench_$name.err
	python - $OUT/bench_$name.json $name <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  %-18s %8.0f Msamples/s  frac %.4f  kernel_ms %.4f  fixups %s  %s" % (sys.argv[2], r["value"], r["roofline"]["frac"], r["roofline"]["kernel_ms"], r["detail"].get("fixup_segments"), r["roofline"]["kernel"]))
except Exception as e:
    print("  %-18s FAILED %s" % (sys.argv[2], e))
PY
}
run fm2a_default fm2a X=1
run fm2b fm2b X=1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fm_back -c 1 -o /tmp/fm2a_back python bench.py --workload fm2a --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
ncu -i /tmp/fm2a_back.ncu-rep --page raw --csv > $OUT/raw_fm2a_back.csv 2>/dev/null
ncu -i /tmp/fm2a_back.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/source_fm2a_back.csv.gz
echo "ncu rc=$?"
date
