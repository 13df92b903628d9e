#!/bin/bash
# Round-end evidence from one GPU-box session on the default build: the `-m gpu` suite, smoke(), one bench line per
# workload and the ncu launch lists / full captures that profiles/ summarises.  Output: gpurun_out/final/.
cd "$(dirname "$0")/.."
OUT=gpurun_out/final; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
T0=$SECONDS
timeout 300 python -m pytest tests -x -q -m gpu > $OUT/gpu_tests.txt 2>&1; echo "gpu suite rc=$? t=$((SECONDS-T0))"; tail -2 $OUT/gpu_tests.txt
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.txt
timeout 240 python bench.py > $OUT/bench_fm2b.json 2> $OUT/bench_fm2b.err; echo "fm2b rc=$? t=$((SECONDS-T0))"
for w in fm5a fm2a fm1 power3 power4; do
	timeout 150 python bench.py --workload $w --no-cpu > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "$w rc=$? t=$((SECONDS-T0))"
done
for w in fm2b power3; do
	timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$w.csv \
		python bench.py --workload $w --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_launch_$w.log 2>&1; echo "ncu launches $w rc=$? t=$((SECONDS-T0))"
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fm_fused -c 1 -o $OUT/prof_fm2b -f \
	python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_full_fm2b.log 2>&1; echo "ncu full fm2b rc=$? t=$((SECONDS-T0))"
# without --import-source: a report with the fused kernel's SASS is ~30 MB and gpurun brings back 64 MiB in all
timeout 200 ncu --set full --clock-control none -k regex:fm_fused -c 1 -o $OUT/prof_fm5a -f \
	python bench.py --workload fm5a --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_full_fm5a.log 2>&1; echo "ncu full fm5a rc=$? t=$((SECONDS-T0))"
ls -la $OUT; du -sh $OUT
date
