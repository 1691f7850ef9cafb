#!/bin/bash
# builds the development micro-benchmarks (tools/dev/*.cu) against the in-tree kernel sources; binaries are git-ignored
set -e
cd "$(dirname "$0")/../.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
nvcc $FLAGS -o tools/dev/gcbench tools/dev/gcbench.cu reslic_tcm_b200/csrc/*.cu reslic_tcm_b200/csrc/*.cpp
nvcc $FLAGS -o tools/dev/gcbench_copy tools/dev/copybench.cu
# timeline build: per-CTA %globaltimer stamps (gcbench_trace prints a per-launch table after the timing line)
nvcc $FLAGS -DRESLIC_TRACE -o tools/dev/gcbench_trace tools/dev/gcbench.cu reslic_tcm_b200/csrc/*.cu reslic_tcm_b200/csrc/*.cpp
