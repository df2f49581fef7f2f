#!/bin/bash
# build_variant.sh NAME [extra nvcc flags]  ->  gym_sbr2_b200/_variants/libsbr_NAME.so (A/B builds; git-ignored, travels with gpurun)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
mkdir -p $ROOT/gym_sbr2_b200/_variants
cd $ROOT/gym_sbr2_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" -o $ROOT/gym_sbr2_b200/_variants/libsbr_$name.so sbr_kernels.cu
echo built $name
