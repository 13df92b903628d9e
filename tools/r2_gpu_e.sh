#!/bin/bash
# round 2, session E (1 GPU): fresh base build of the rows kernel + ncu, split kernel with segment front end (fm2a),
# boxcar segments aligned to the decimation period (fm5a), CTA-shape A/B.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2e; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv
T0=$SECONDS
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  %-12s %8.0f Msamples/s  frac %.4f  kernel_ms %.4f  %s %s" % (sys.argv[2], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["kernel"], d["detail"]))
except Exception as e:
    print("  %-12s no line: %s" % (sys.argv[2], e))
PY
}
B="--no-extras --no-cpu --no-e2e --steps 20 --warmup 5"
timeout 300 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_golden.py -x -q -m gpu > $OUT/fm_tests.txt 2>&1; echo "fm tests rc=$? t=$((SECONDS-T0))"; tail -5 $OUT/fm_tests.txt
timeout 200 python bench.py $B > $OUT/bench_base.json 2> $OUT/bench_base.err; line $OUT/bench_base.json base
timeout 200 python bench.py $B --workload fm2a > $OUT/bench_fm2a.json 2> $OUT/bench_fm2a.err; line $OUT/bench_fm2a.json fm2a-split
RXB200_FM_NOSPLIT=1 timeout 200 python bench.py $B --workload fm2a > $OUT/bench_fm2a_fused.json 2> $OUT/bench_fm2a_fused.err; line $OUT/bench_fm2a_fused.json fm2a-fused
timeout 200 python bench.py $B --workload fm5a > $OUT/bench_fm5a.json 2> $OUT/bench_fm5a.err; line $OUT/bench_fm5a.json fm5a-256
timeout 200 python bench.py $B --workload fm5a --size-mib 1172 > $OUT/bench_fm5a_128.json 2> $OUT/bench_fm5a_128.err; line $OUT/bench_fm5a_128.json fm5a-128
timeout 200 python bench.py $B --workload fm5a --size-mib 293 > $OUT/bench_fm5a_32.json 2> $OUT/bench_fm5a_32.err; line $OUT/bench_fm5a_32.json fm5a-32
timeout 200 python bench.py $B --workload fm1 > $OUT/bench_fm1.json 2> $OUT/bench_fm1.err; line $OUT/bench_fm1.json fm1
for so in rx_tools_b200/variants/librxb200_*.so; do
	v=$(basename $so .so); v=${v#librxb200_}
	timeout 200 env RXB200_LIB=$PWD/$so python -m pytest tests/test_fm_gpu.py -x -q -m gpu -k "cfg2B or burst or murmur or fullscale_noise_P3 or ragged_tail" > $OUT/test_$v.log 2>&1; rc=$?
	timeout 120 env RXB200_LIB=$PWD/$so python bench.py $B > $OUT/bench_$v.json 2> $OUT/bench_$v.err
	echo "$v tests rc=$rc ($(tail -1 $OUT/test_$v.log)) t=$((SECONDS-T0))"; line $OUT/bench_$v.json $v
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_split -c 1 -o $OUT/prof_fm2b_rows -f \
	python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fm2b.log 2>&1; echo "ncu full fm2b rc=$? t=$((SECONDS-T0))"
timeout 300 ncu --set full --clock-control none -k regex:fm_split -c 1 -o $OUT/prof_fm2a_split -f \
	python bench.py --workload fm2a --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fm2a.log 2>&1; echo "ncu full fm2a rc=$? t=$((SECONDS-T0))"
date
