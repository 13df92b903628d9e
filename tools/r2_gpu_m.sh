#!/bin/bash
# round 2, session M (1 GPU): the build with the ticket hand-off without a front-end barrier, taps reordered around the
# exchange, the slice choice for few hops and shorter segments for little work -- full suite, one bench line per
# workload, DRAM traffic per workload, ncu pages of the fm2b kernel.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2m; mkdir -p $OUT
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
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_multi_gpu.py > $OUT/gpu_tests.txt 2>&1; echo "gpu suite rc=$? t=$((SECONDS-T0))"; tail -6 $OUT/gpu_tests.txt
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.txt
B="--no-extras --no-cpu --steps 20 --warmup 5"
for w in fm2b fm2a fm1 fm5a power3 power4; do
	timeout 300 python bench.py $B --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err; line $OUT/bench_$w.json $w
done
timeout 200 python bench.py $B --no-e2e --workload fm5a --size-mib 293 > $OUT/bench_fm5a_32.json 2> $OUT/bench_fm5a_32.err; line $OUT/bench_fm5a_32.json fm5a-32ch
RXB200_FM_NOROWS=1 timeout 200 python bench.py $B --no-e2e > $OUT/bench_fused.json 2> $OUT/bench_fused.err; line $OUT/bench_fused.json fm2b-fused
python - <<'PY'
# a rank's share of cfg4 at eight GPUs: 109 of the 871 hops, 36 sweeps
import numpy as np, torch
from rx_tools_b200 import power, synth
plan = power.plan_range("24M:1766M:1k", 0.285)
sc = power.PowerScanner(plan, "hamming")
base = torch.from_numpy(synth.power_hops(2, 109, plan.buf_len, seed=4000).reshape(-1)).cuda()
d = base.repeat(18).contiguous()
ms = []
for _ in range(6):
    sc.scanner_device(d.data_ptr(), 36, 0, 109, sync=False); ms.append(sc.kernel_ms())
print("  power4 shard 109 hops x 36 sweeps: kernel %.4f ms (%.0f Msamples/s per GPU)" % (min(ms[1:]), 109 * 36 * 8192 / min(ms[1:]) / 1e3))
PY
timeout 600 python tools/measure_traffic.py > $OUT/traffic.txt 2>&1; echo "traffic rc=$? t=$((SECONDS-T0))"; cat $OUT/traffic.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_split -c 1 -o /tmp/prof_fm2b -f \
	python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fm2b.log 2>&1; echo "ncu fm2b rc=$? t=$((SECONDS-T0))"
ncu -i /tmp/prof_fm2b.ncu-rep --page raw --csv > $OUT/raw_fm2b.csv 2>/dev/null
ncu -i /tmp/prof_fm2b.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/src_fm2b.csv.gz
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_fm2b.csv \
	python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > $OUT/ncu_launch_fm2b.log 2>&1; echo "ncu launches rc=$?"
cp profiles/traffic_*.json $OUT/ 2>/dev/null
date
