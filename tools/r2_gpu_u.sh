#!/bin/bash
# session U: pipe-balance variants of the rows kernel (HB_MAD / SC_DP builds under rx_tools_b200/variants/) -- parity of
# each, fm2b / fm5a / fm2a throughput; then an ncu pass of the fm1 kernel (never profiled since the lean atan2)
OUT=gpurun_out/r2u; mkdir -p $OUT
exec > $OUT/session.log 2>&1
date
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  %-22s %8.0f Msamples/s  frac %.4f  kernel_ms %.4f  %s" % (sys.argv[2], r["value"], r["roofline"]["frac"], r["roofline"]["kernel_ms"], r["roofline"]["kernel"]))
except Exception as e:
    print("  %-22s FAILED %s" % (sys.argv[2], e))
PY
}
bench() { # name workload env...
	local name=$1 wl=$2; shift; shift
	env "$@" timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-extras > $OUT/bench_${name}_$wl.json 2> $OUT/bench_${name}_$wl.err
	line $OUT/bench_${name}_$wl.json ${name}_$wl
}
bench base fm2b X=1; bench base fm5a X=1
for so in rx_tools_b200/variants/librxb200_*.so; do
	v=$(basename $so .so); v=${v#librxb200_}
	timeout 300 env RXB200_LIB=$PWD/$so python -m pytest tests/test_fm_gpu.py tests/test_fuzz_gpu.py -x -q -m gpu > $OUT/test_$v.log 2>&1; echo "$v tests rc=$? $(tail -1 $OUT/test_$v.log)"
	bench $v fm2b RXB200_LIB=$PWD/$so
	bench $v fm5a RXB200_LIB=$PWD/$so
done
bench base2 fm2b X=1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fm_fused -c 1 -o /tmp/fm1 python bench.py --workload fm1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
ncu -i /tmp/fm1.ncu-rep --page raw --csv > $OUT/raw_fm1.csv 2>/dev/null
ncu -i /tmp/fm1.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/source_fm1.csv.gz
echo "ncu rc=$?"
date
