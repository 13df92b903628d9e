#!/bin/bash
# round 2, session D (1 GPU): rows kernel after the second ncu pass (no spills, no 64-bit modulo, loads issued after
# level 0), CTA-shape A/B, segment-length sweep of the fused kernel on fm5a.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2d; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv
T0=$SECONDS
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  %-10s %8.0f Msamples/s  frac %.4f  kernel_ms %.4f  %s %s" % (sys.argv[2], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["kernel"], d["detail"]))
except Exception as e:
    print("  %-10s no line: %s" % (sys.argv[2], e))
PY
}
B="--no-extras --no-cpu --no-e2e --steps 20 --warmup 5"
timeout 200 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_golden.py -x -q -m gpu > $OUT/fm_tests.txt 2>&1; echo "fm tests rc=$? t=$((SECONDS-T0))"; tail -3 $OUT/fm_tests.txt
timeout 200 python bench.py $B > $OUT/bench_base.json 2> $OUT/bench_base.err; line $OUT/bench_base.json base
for so in rx_tools_b200/variants/librxb200_*.so; do
	v=$(basename $so .so); v=${v#librxb200_}
	timeout 200 env RXB200_LIB=$PWD/$so python -m pytest tests/test_fm_gpu.py -x -q -m gpu -k "cfg2B or burst or murmur or fullscale_noise_P3 or ragged_tail" > $OUT/test_$v.log 2>&1; rc=$?
	timeout 120 env RXB200_LIB=$PWD/$so python bench.py $B > $OUT/bench_$v.json 2> $OUT/bench_$v.err
	echo "$v tests rc=$rc ($(tail -1 $OUT/test_$v.log)) t=$((SECONDS-T0))"; line $OUT/bench_$v.json $v
done
timeout 300 python tools/seg_sweep.py 128 896 3104 8 > $OUT/seg_sweep_128.txt 2>&1; echo "sweep rc=$? t=$((SECONDS-T0))"
timeout 300 python tools/seg_sweep.py 32 512 2560 8 > $OUT/seg_sweep_32.txt 2>&1; echo "sweep32 rc=$? t=$((SECONDS-T0))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_split -c 1 -o $OUT/prof_fm2b_rows -f \
	python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fm2b.log 2>&1; echo "ncu full fm2b rc=$? t=$((SECONDS-T0))"
date
