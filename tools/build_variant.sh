#!/bin/bash
# build_variant.sh <name> [-Dmacro ...]  -> tools/ab/<name>   (A/B timing, see tools/gpu_ab.sh)
name=$1; shift
mkdir -p tools/ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O3,-ffp-contract=off,-fno-fast-math "$@" --shared -cudart static cs121-softbodysim_b200/csrc/pbd_plan.cpp cs121-softbodysim_b200/csrc/pbd_tileplan.cpp cs121-softbodysim_b200/csrc/pbd_placement.cpp cs121-softbodysim_b200/csrc/pbd_stream.cu cs121-softbodysim_b200/csrc/pbd_tile.cu cs121-softbodysim_b200/csrc/pbd_batch.cu cs121-softbodysim_b200/csrc/pbd_capi.cu -o tools/ab/$name 2>&1 | grep -E " error"
exit 0
