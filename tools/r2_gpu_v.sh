#!/bin/bash
# session V: back kernel after the straggler fix (longer stream replay, windowed probe, one warp per item, 64-sample
# windows) -- parity, then fm2a A/B; then session U (pipe-balance variants of the rows kernel, fm1 ncu)
OUT=gpurun_out/r2v; mkdir -p $OUT
exec > $OUT/session.log 2>&1
date
timeout 900 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > $OUT/tests.txt 2>&1
echo "tests rc=$?"; tail -4 $OUT/tests.txt
run() { # name, env...
	local name=$1; shift
	env "$@" timeout 600 python bench.py --workload fm2a --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras > $OUT/bench_$name.json 2> $OUT/bench_$name.err
	python - $OUT/bench_$name.json $name <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  %-18s %8.0f Msamples/s  frac %.4f  kernel_ms %.4f  fixups %s" % (sys.argv[2], r["value"], r["roofline"]["frac"], r["roofline"]["kernel_ms"], r["detail"].get("fixup_segments")))
except Exception as e:
    print("  %-18s FAILED %s" % (sys.argv[2], e))
PY
}
for ws in 64 128; do for wa in 16 20 24; do run w${ws}_a${wa} RXB200_FM_STREAM_WIN=$ws RXB200_FM_STREAM_WARM_A=$wa; done; done
for p in 2000 3000 4500; do run w64_p$p RXB200_FM_STREAM_WIN=64 RXB200_FM_STREAM_PIECE=$p; done
for p in 3000 4500 6000; do run w128_p$p RXB200_FM_STREAM_PIECE=$p; done
run w64_t128 RXB200_FM_STREAM_WIN=64 RXB200_FM_STREAM_T=128
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fm_back -c 1 -o /tmp/fm2a_back python bench.py --workload fm2a --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
ncu -i /tmp/fm2a_back.ncu-rep --page raw --csv > $OUT/raw_fm2a_back.csv 2>/dev/null
ncu -i /tmp/fm2a_back.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/source_fm2a_back.csv.gz
echo "ncu rc=$?"
date
bash tools/r2_gpu_u.sh
