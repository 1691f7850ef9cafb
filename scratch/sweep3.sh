#!/bin/bash
cd "$(dirname "$0")/.."
./scratch/gcbench 24 98304 1 0 40
./scratch/gcbench 24 98304 0 0 40
./scratch/gcbench 24 491520 1 0 10
./scratch/gcbench 64 98304 0 0 20
./scratch/gcbench 256 16384 0 1 20
./scratch/gcbench 16 720896 1 0 10
python -m pytest tests/test_gc_parity.py tests/test_pipeline_gpu.py tests/test_guard_bands.py -m gpu -x -q 2>&1 | tail -3
