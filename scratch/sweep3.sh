#!/bin/bash
cd "$(dirname "$0")/.."
for d in 0 1; do
./scratch/gcbench 24 98304 1 0 40 $d
./scratch/gcbench 24 98304 0 0 40 $d
./scratch/gcbench 24 491520 1 0 10 $d
./scratch/gcbench 64 98304 0 0 20 $d
./scratch/gcbench 256 16384 0 1 20 $d
./scratch/gcbench 16 720896 1 0 10 $d
done
