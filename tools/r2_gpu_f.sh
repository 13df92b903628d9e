#!/bin/bash
# round 2, session F: ncu captures of the split kernels (rows: fm2b, segments: fm2a); CSV pages written on the box
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2f; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date
for w in fm2b fm2a; do
	timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_split -c 1 -o /tmp/prof_$w -f \
		python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_$w.log 2>&1; echo "ncu $w rc=$?"
	ncu -i /tmp/prof_$w.ncu-rep --page raw --csv > $OUT/raw_$w.csv 2>/dev/null
	ncu -i /tmp/prof_$w.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/src_$w.csv.gz
done
RXB200_FM_NOROWS=1 timeout 300 ncu --set full --clock-control none -k regex:fm_fused -c 1 -o /tmp/prof_fused -f \
	python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fused.log 2>&1
ncu -i /tmp/prof_fused.ncu-rep --page raw --csv > $OUT/raw_fused.csv 2>/dev/null
ls -la $OUT; date
