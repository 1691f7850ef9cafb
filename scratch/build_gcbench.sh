#!/bin/bash
# builds scratch/gcbench against the in-tree kernel sources
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o scratch/gcbench scratch/gcbench.cu reslic_tcm_b200/csrc/*.cu reslic_tcm_b200/csrc/*.cpp
