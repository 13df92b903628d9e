#!/bin/bash
# round 2, session B (1 GPU): first contact of the split/rows kernel -- parity, sanitizer, A/B against the fused kernel.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2b; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv
T0=$SECONDS
timeout 120 python tests/sanitize_smoke.py cfg2B,wbfm_P2_fir,F0_P1 > $OUT/smoke_rows.txt 2>&1; echo "rows smoke rc=$? t=$((SECONDS-T0))"; cat $OUT/smoke_rows.txt | tail -8
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_multi_gpu.py > $OUT/gpu_tests.txt 2>&1; echo "gpu suite rc=$? t=$((SECONDS-T0))"; tail -15 $OUT/gpu_tests.txt
timeout 300 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_golden.py tests/test_full_size_gpu.py -q -m gpu > $OUT/gpu_tests_fm_all.txt 2>&1; echo "fm tests (no -x) rc=$? t=$((SECONDS-T0))"; tail -25 $OUT/gpu_tests_fm_all.txt
for v in rows norows; do
	if [ $v = norows ]; then export RXB200_FM_NOROWS=1; else unset RXB200_FM_NOROWS; fi
	timeout 200 python bench.py --no-extras --no-cpu --steps 20 --warmup 5 > $OUT/bench_fm2b_$v.json 2> $OUT/bench_fm2b_$v.err; echo "bench fm2b $v rc=$? t=$((SECONDS-T0))"
	python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_fm2b_$v.json").read().strip().splitlines()[-1]); print("  $v", round(d["value"]), "Msamples/s frac", round(d["roofline"]["frac"],4), d["detail"], d["roofline"]["kernel"])
except Exception as e: print("  $v: no line", e)
PY
done
unset RXB200_FM_NOROWS
RXB200_FM_ROWS_BE=64 timeout 200 python bench.py --no-extras --no-cpu --no-e2e --steps 20 --warmup 5 > $OUT/bench_fm2b_be64.json 2> $OUT/bench_fm2b_be64.err; echo "bench be64 rc=$?"; python -c "
import json; d=json.loads(open('$OUT/bench_fm2b_be64.json').read().strip().splitlines()[-1]); print('  be64', round(d['value']), d['detail'])"
timeout 300 compute-sanitizer --tool memcheck python tests/sanitize_smoke.py cfg2B > $OUT/memcheck.txt 2>&1; echo "memcheck rc=$? t=$((SECONDS-T0))"; tail -4 $OUT/memcheck.txt
timeout 400 compute-sanitizer --tool racecheck python tests/sanitize_smoke.py cfg2B > $OUT/racecheck.txt 2>&1; echo "racecheck rc=$? t=$((SECONDS-T0))"; tail -4 $OUT/racecheck.txt
# fm5a: why do 128 channels run at half the rate of 256?
for mib in 2344 1172 586; do
	timeout 120 python bench.py --workload fm5a --size-mib $mib --no-extras --no-cpu --no-e2e --steps 10 > $OUT/bench_fm5a_$mib.json 2> $OUT/bench_fm5a_$mib.err
	python -c "
import json; d=json.loads(open('$OUT/bench_fm5a_$mib.json').read().strip().splitlines()[-1]); print('  fm5a $mib MiB', d['config']['channels_per_gpu'], 'ch', round(d['value']), d['detail'])"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_split -c 1 -o $OUT/prof_fm2b_rows -f \
	python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fm2b.log 2>&1; echo "ncu full fm2b rc=$? t=$((SECONDS-T0))"
ls -la $OUT; date
