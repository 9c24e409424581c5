#!/bin/bash
# Build an experimental libgt4gpu variant: scripts/build_variant.sh NAME [-DFOO=1 ...] -> build/variants/libgt4gpu_NAME.so
# Use it with GT4GPU_LIB=build/variants/libgt4gpu_NAME.so (A/B runs inside one gpurun call).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/genometester4_b200/csrc
out=$root/build/variants; mkdir -p $out/$name
flags="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas -diag-suppress 177,128 $*"
for f in gt4gpu_kernels gt4gpu_stream_kernel gt4gpu_kway_kernel gt4gpu_fused_kernel gt4gpu_sort_kernel gt4gpu_fasta_kernel gt4gpu_api; do
  nvcc $flags -c -o $out/$name/$f.o $src/$f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $out/libgt4gpu_$name.so $out/$name/*.o
echo built $out/libgt4gpu_$name.so
