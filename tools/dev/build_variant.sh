#!/bin/bash
# dev: build the library with extra -D flags into tools/dev/libs/<name>.so   (usage: build_variant.sh name -DFOO=1 ...)
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
mkdir -p tools/dev/libs
C=reslic_tcm_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared --fmad=true --prec-div=true --prec-sqrt=true --ftz=false "$@" \
  -o tools/dev/libs/$name.so $C/cabi.cu $C/gc_fused.cu $C/gc_bwd.cu $C/eb_fused.cu $C/eb_bwd.cu $C/stanh_fused.cu $C/rans_slots.cu $C/rate_reduce.cu $C/rate_exchange.cu $C/cdf_tables.cpp $C/rans.cpp
cuobjdump -res-usage tools/dev/libs/$name.so 2>/dev/null | grep -A1 "eb_bwd_kernelILb1" | tail -1
