#!/bin/bash
# chains / steps-per-graph matrix through bench.py: the full config-3 batch and rank 0's shard of an 8-GPU job (development)
cd "$(dirname "$0")/../.."
for ch in 3 4 6; do for spg in 12 24 48; do
for sh in 0 8; do
timeout 200 python bench.py --steps 192 --warmup 48 --chains $ch --nbuf $ch --steps-per-graph $spg --exchange peer --shard-of $sh --no-e2e --no-cpu-baseline --no-training-kernels --no-whole-y --legs none > gpurun_out/bc_${ch}_${spg}_${sh}.json 2> gpurun_out/bc_${ch}_${spg}_${sh}.err
done; done; done
