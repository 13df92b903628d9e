#!/bin/bash
# The final session once more after the FP32 form of polar_discriminant (disc_std_f32), without the parts that do not depend
# on the library build (reference arm, instruction microbenchmark): the `-m gpu` suite, smoke(), DRAM traffic per workload,
# the driver's bench command, one bench line per workload, ncu of the fm2b / fm2a / fm1 / power3 kernels, the launch list.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2final4; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv
T0=$SECONDS
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/gpu_tests.txt 2>&1; echo "gpu suite rc=$? t=$((SECONDS-T0))"; tail -6 $OUT/gpu_tests.txt
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.txt
timeout 400 python tools/measure_traffic.py > $OUT/traffic.txt 2>&1; echo "traffic rc=$? t=$((SECONDS-T0))"; cat $OUT/traffic.txt; cp profiles/traffic_*.json $OUT/
timeout 300 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$? t=$((SECONDS-T0))"
for w in fm1 fm2a fm5a power3 power4; do
	timeout 200 python bench.py --workload $w --no-extras --no-cpu > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "$w rc=$? t=$((SECONDS-T0))"
done
RXB200_FM_NOROWS=1 timeout 200 python bench.py --no-extras --no-cpu --no-e2e > $OUT/bench_fm2b_fused.json 2> $OUT/bench_fm2b_fused.err
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/r2final4/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print("  %-26s %9.0f Msamples/s  frac %s  traffic %s  e2e %s" % (os.path.basename(f), d["value"], ("%.4f" % r["frac"]) if r else "-", r.get("traffic"), (d.get("e2e") or {}).get("value")))
    except Exception as e:
        print("  ", f, "no line", e)
PY
for w in fm1 fm2b power3; do
	K=fm_; [ $w = power3 ] && K=power_fft8
	timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o /tmp/prof_$w -f \
		python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_$w.log 2>&1; echo "ncu $w rc=$? t=$((SECONDS-T0))"
	ncu -i /tmp/prof_$w.ncu-rep --page raw --csv > $OUT/raw_$w.csv 2>/dev/null
done
timeout 200 ncu --set full --clock-control none -k regex:fm_ -s 2 -c 2 -o /tmp/prof_fm2a -f \
	python bench.py --workload fm2a --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fm2a.log 2>&1; echo "ncu fm2a rc=$? t=$((SECONDS-T0))"
ncu -i /tmp/prof_fm2a.ncu-rep --page raw --csv > $OUT/raw_fm2a.csv 2>/dev/null
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_bench.csv \
	python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$? t=$((SECONDS-T0))"
date
