#!/bin/bash
# round 2, session C (1 GPU): rows kernel after the first ncu pass (2 back-end warps, prefetch fence, conflict-free
# exchange), A/B of CTA shapes, full-size reference goldens, fm5a segment sweep.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2c; mkdir -p $OUT
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
timeout 200 python bench.py $B > $OUT/bench_base.json 2> $OUT/bench_base.err; line $OUT/bench_base.json base
RXB200_FM_ROWS_BE=32 timeout 200 python bench.py $B > $OUT/bench_be32.json 2> $OUT/bench_be32.err; line $OUT/bench_be32.json be32
RXB200_FM_NOROWS=1 timeout 200 python bench.py $B > $OUT/bench_norows.json 2> $OUT/bench_norows.err; line $OUT/bench_norows.json norows
FM_TESTS="tests/test_fm_gpu.py tests/test_fuzz_gpu.py"
for so in rx_tools_b200/variants/librxb200_*.so; do
	v=$(basename $so .so); v=${v#librxb200_}
	timeout 200 env RXB200_LIB=$PWD/$so python -m pytest $FM_TESTS -x -q -m gpu -k "cfg2B or burst or murmur or fullscale_noise_P3 or ragged or fuzz" > $OUT/test_$v.log 2>&1; rc=$?
	timeout 120 env RXB200_LIB=$PWD/$so python bench.py $B > $OUT/bench_$v.json 2> $OUT/bench_$v.err
	echo "$v tests rc=$rc ($(tail -1 $OUT/test_$v.log)) t=$((SECONDS-T0))"; line $OUT/bench_$v.json $v
done
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_multi_gpu.py > $OUT/gpu_tests.txt 2>&1; echo "gpu suite rc=$? t=$((SECONDS-T0))"; tail -12 $OUT/gpu_tests.txt
# fm5a: 128 channels, segment length sweep (the fused kernel's per-thread segments)
for seg in 0 1000 1528 2040 2048 2056 3000; do
	RXB200_FM_SEG=$seg timeout 120 python bench.py --workload fm5a --size-mib 1172 --no-extras --no-cpu --no-e2e --steps 10 > $OUT/bench_fm5a_seg$seg.json 2> $OUT/bench_fm5a_seg$seg.err; line $OUT/bench_fm5a_seg$seg.json "5a/128/$seg"
done
RXB200_FM_SEG=1024 timeout 120 python bench.py --workload fm5a --size-mib 293 --no-extras --no-cpu --no-e2e --steps 10 > $OUT/bench_fm5a_32ch.json 2>/dev/null; line $OUT/bench_fm5a_32ch.json "5a/32/1024"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_split -c 1 -o $OUT/prof_fm2b_rows -f \
	python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fm2b.log 2>&1; echo "ncu full fm2b rc=$? t=$((SECONDS-T0))"
date
