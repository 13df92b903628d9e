#!/bin/bash
# round 2, session N (1 GPU): two item sizes (short tail), tap order A/B, corrected slice choice of rx_power
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2n; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date
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
timeout 300 python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py tests/test_power_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > $OUT/tests.txt 2>&1; echo "tests rc=$?"; tail -3 $OUT/tests.txt
timeout 200 python bench.py $B > $OUT/bench_base.json 2> $OUT/bench_base.err; line $OUT/bench_base.json base
RXB200_FM_NOTAIL=1 timeout 200 python bench.py $B > $OUT/bench_notail.json 2> $OUT/bench_notail.err; line $OUT/bench_notail.json notail
so=rx_tools_b200/variants/librxb200_inorder.so
timeout 120 env RXB200_LIB=$PWD/$so python bench.py $B > $OUT/bench_inorder.json 2> $OUT/bench_inorder.err; line $OUT/bench_inorder.json inorder
RXB200_FM_NOTAIL=1 timeout 120 env RXB200_LIB=$PWD/$so python bench.py $B > $OUT/bench_inorder_notail.json 2>/dev/null; line $OUT/bench_inorder_notail.json inorder-notail
for w in power3 power4 fm5a; do
	timeout 200 python bench.py $B --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err; line $OUT/bench_$w.json $w
done
python - <<'PY'
import torch
from rx_tools_b200 import power, synth
plan = power.plan_range("24M:1766M:1k", 0.285)
for nh in (109, 218, 436):
    sc = power.PowerScanner(plan, "hamming")
    base = torch.from_numpy(synth.power_hops(2, nh, plan.buf_len, seed=4000).reshape(-1)).cuda()
    d = base.repeat(18).contiguous()
    ms = []
    for _ in range(6):
        sc.scanner_device(d.data_ptr(), 36, 0, nh, sync=False); ms.append(sc.kernel_ms())
    print("  power4 shard %d hops x 36 sweeps: kernel %.4f ms (%.0f Msamples/s per GPU)" % (nh, min(ms[1:]), nh * 36 * 8192 / min(ms[1:]) / 1e3))
    sc.close()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fm_split -c 1 -o /tmp/prof_fm2b -f \
	python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > $OUT/ncu_full_fm2b.log 2>&1; echo "ncu fm2b rc=$?"
ncu -i /tmp/prof_fm2b.ncu-rep --page raw --csv > $OUT/raw_fm2b.csv 2>/dev/null
ncu -i /tmp/prof_fm2b.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/src_fm2b.csv.gz
date
