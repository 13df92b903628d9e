#!/bin/bash
# session R: stream path (front kernel + back kernel) on the undecimated wbfm shape -- parity, then fm2a A/B and a piece sweep
OUT=gpurun_out/r2r; mkdir -p $OUT
exec > $OUT/session.log 2>&1
date
timeout 900 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_golden_gpu.py tests/test_full_size_gpu.py tests/test_dropin.py -m gpu -x -q > $OUT/tests.txt 2>&1
echo "tests rc=$?"; tail -4 $OUT/tests.txt
run() { # name, env...
	local name=$1; shift
	env "$@" timeout 600 python bench.py --workload fm2a --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras > $OUT/bench_$name.json 2> $OUT/bench_$name.err
	python - $OUT/bench_$name.json $name <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  %-14s %8.0f Msamples/s  frac %.4f  kernel_ms %.4f  %s %s" % (sys.argv[2], r["value"], r["roofline"]["frac"], r["roofline"]["kernel_ms"], r["roofline"]["kernel"], r["config"].get("geometry")))
except Exception as e:
    print("  %-14s FAILED %s" % (sys.argv[2], e))
PY
}
run stream X=1
run fused RXB200_FM_NOSTREAM=1
for p in 1000 2000 3000 6000 12000 24000; do run piece$p RXB200_FM_STREAM_PIECE=$p; done
# per-kernel times of the stream path
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fm_ -c 6 --csv --log-file $OUT/launches_fm2a.csv python bench.py --workload fm2a --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
grep -o '"fm_[a-z_]*[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*$' $OUT/launches_fm2a.csv | tail -4
cut -d, -f5,12- $OUT/launches_fm2a.csv | tail -4
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fm_ -s 2 -c 2 -o /tmp/fm2a_stream python bench.py --workload fm2a --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
ncu -i /tmp/fm2a_stream.ncu-rep --page raw --csv > $OUT/raw_fm2a_stream.csv 2>/dev/null
ncu -i /tmp/fm2a_stream.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/source_fm2a_stream.csv.gz
echo "ncu rc=$?"
date
