#!/bin/bash
# A/B builds of librxb200.so: the same sources with different -D switches of fm_kernels.cu, one .so per variant
# under rx_tools_b200/variants/ (git-ignored, shipped to the GPU box by gpurun; selected with RXB200_LIB=...).
#   tools/build_variants.sh name1="-DX=1 -DY=2" name2="..."
set -e
cd "$(dirname "$0")/../rx_tools_b200/csrc"
make -s -j4 >/dev/null            # the shared objects (power, sdr, host plan) come from the default build
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p ../variants
pids=()
for spec in "$@"; do
	name="${spec%%=*}"; flags="${spec#*=}"
	(
		nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $flags -c -o ../variants/fm_$name.o fm_kernels.cu 2> ../variants/fm_$name.ptxas.log
		nvcc $ARCH -shared -o ../variants/librxb200_$name.so ../variants/fm_$name.o power_kernels.o sdr_kernels.o host_plan.o nccl_dyn.o -ldl
		echo "$flags" > ../variants/$name.flags
		echo "built $name: $flags"
	) &
	pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
