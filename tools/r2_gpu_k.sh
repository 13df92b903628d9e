#!/bin/bash
# round 2, session K (8 GPUs): multi-GPU parity (3/4/8-GPU groups, 4 ranks one process per GPU, 2-GPU shell) and the
# bench at N = 8, 4, 2, 1 on the same box; the sharded configs ride in extra.power4 / extra.fm5a.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2k; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
T0=$SECONDS
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_dropin.py -q -m gpu -k "gpu or two_gpus or gather" > $OUT/multi_gpu_tests.txt 2>&1; echo "multi-gpu tests rc=$? t=$((SECONDS-T0))"; tail -4 $OUT/multi_gpu_tests.txt
for n in 8 4 2; do
	timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
		bench.py --gpus $n --steps 10 --warmup 3 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err; echo "bench n$n rc=$? t=$((SECONDS-T0))"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench n1 rc=$? t=$((SECONDS-T0))"
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r2k/bench_n{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    ex = d.get("extra", {})
    print("N=%d fm2b %.0f  e2e %.0f | power4 %.0f (step %.3f ms, kernel %.3f, allgather %s) | fm5a %.0f | fm2a %.0f" % (
        n, d["value"], (d.get("e2e") or {}).get("value", 0), ex.get("power4", {}).get("value", 0), ex.get("power4", {}).get("ms_per_step", 0),
        ex.get("power4", {}).get("kernel_ms_max_over_ranks", 0), ex.get("power4", {}).get("allgather_ms"), ex.get("fm5a", {}).get("value", 0), ex.get("fm2a", {}).get("value", 0)))
PY
date
