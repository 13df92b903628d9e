#!/bin/bash
# round 2, session A (2 GPUs): the -m gpu suite incl. the multi-GPU parity tests, bench at N=1 and N=2 (extras carry
# the sharded configs).  Output: gpurun_out/r2a/.
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2a; mkdir -p $OUT
exec > >(tee $OUT/session.log) 2>&1
date; nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv
T0=$SECONDS
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/gpu_tests.txt 2>&1; echo "gpu suite rc=$? t=$((SECONDS-T0))"; tail -5 $OUT/gpu_tests.txt
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.txt
timeout 400 python bench.py --steps 10 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench n1 rc=$? t=$((SECONDS-T0))"; tail -3 $OUT/bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
	bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$? t=$((SECONDS-T0))"; tail -3 $OUT/bench_n2.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$? t=$((SECONDS-T0))"
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
date
