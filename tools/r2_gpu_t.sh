#!/bin/bash
# session T: back kernel with cp.async double-buffered windows -- parity, then fm2a A/B over window size x lanes per item x piece
OUT=gpurun_out/r2t; mkdir -p $OUT
exec > $OUT/session.log 2>&1
date
timeout 900 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_golden_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > $OUT/tests.txt 2>&1
echo "tests rc=$?"; tail -4 $OUT/tests.txt
run() { # name, env...
	local name=$1; shift
	env "$@" timeout 600 python bench.py --workload fm2a --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras > $OUT/bench_$name.json 2> $OUT/bench_$name.err
	python - $OUT/bench_$name.json $name <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  %-18s %8.0f Msamples/s  frac %.4f  kernel_ms %.4f  %s" % (sys.argv[2], r["value"], r["roofline"]["frac"], r["roofline"]["kernel_ms"], r["roofline"]["kernel"]))
except Exception as e:
    print("  %-18s FAILED %s" % (sys.argv[2], e))
PY
}
for ws in 128 256; do for t in 32 64 128; do run w${ws}_t${t} RXB200_FM_STREAM_WIN=$ws RXB200_FM_STREAM_T=$t; done; done
for p in 1500 2200 3000 4500 6000; do run w128_t128_p$p RXB200_FM_STREAM_PIECE=$p; done
for p in 1500 3000 6000; do run w128_t32_p$p RXB200_FM_STREAM_T=32 RXB200_FM_STREAM_PIECE=$p; done
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras > $OUT/bench_fm2b.json 2>&1; python -c "import json;r=json.loads(open(\"$OUT/bench_fm2b.json\").read().strip().splitlines()[-1]);print(\"  fm2b\", r[\"value\"], r[\"roofline\"][\"frac\"])"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fm_back -c 1 -o /tmp/fm2a_back python bench.py --workload fm2a --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
ncu -i /tmp/fm2a_back.ncu-rep --page raw --csv > $OUT/raw_fm2a_back.csv 2>/dev/null
ncu -i /tmp/fm2a_back.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/source_fm2a_back.csv.gz
echo "ncu rc=$?"
date
