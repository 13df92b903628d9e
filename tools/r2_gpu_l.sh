#!/bin/bash
# round 2, session L: chunk-start rows as their own instantiation; 512-thread rx_power fast path; A/B of CTA shapes;
# ncu captures with the CSV pages produced on the box (the .ncu-rep stays there: gpurun_out is capped at 64 MiB)
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2l; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv
T0=$SECONDS
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
timeout 400 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_golden.py tests/test_power_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > $OUT/tests.txt 2>&1; echo "fm+power tests rc=$? t=$((SECONDS-T0))"; tail -5 $OUT/tests.txt
timeout 200 python bench.py $B > $OUT/bench_base.json 2> $OUT/bench_base.err; line $OUT/bench_base.json base
for w in fm2a fm5a fm1; do
	timeout 200 python bench.py $B --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err; line $OUT/bench_$w.json $w
done
RXB200_FM_NOROWS=1 timeout 200 python bench.py $B > $OUT/bench_norows.json 2> $OUT/bench_norows.err; line $OUT/bench_norows.json fused
for w in power3 power4; do
	timeout 200 python bench.py $B --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err; line $OUT/bench_$w.json $w-512
	RXB200_POWER_THREADS=1024 timeout 200 python bench.py $B --workload $w > $OUT/bench_${w}_1024.json 2> $OUT/bench_${w}_1024.err; line $OUT/bench_${w}_1024.json $w-1024
done
for so in rx_tools_b200/variants/librxb200_*.so; do
	v=$(basename $so .so); v=${v#librxb200_}
	timeout 200 env RXB200_LIB=$PWD/$so python -m pytest tests/test_fm_gpu.py -x -q -m gpu -k "cfg2B or burst or murmur or fullscale_noise_P3 or ragged_tail" > $OUT/test_$v.log 2>&1; rc=$?
	timeout 120 env RXB200_LIB=$PWD/$so python bench.py $B > $OUT/bench_$v.json 2> $OUT/bench_$v.err
	echo "$v tests rc=$rc ($(tail -1 $OUT/test_$v.log)) t=$((SECONDS-T0))"; line $OUT/bench_$v.json $v
done
for w in fm2b power3; do
	K=fm_split; [ $w = power3 ] && K=power_fft8
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o /tmp/prof_$w -f \
		python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_$w.log 2>&1; echo "ncu $w rc=$? t=$((SECONDS-T0))"
	ncu -i /tmp/prof_$w.ncu-rep --page raw --csv > $OUT/raw_$w.csv 2>/dev/null
	ncu -i /tmp/prof_$w.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/src_$w.csv.gz
done
ls -la $OUT | head -40; date
