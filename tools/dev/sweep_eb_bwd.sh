#!/bin/bash
for lib in t256m1 t256m2 t128m2 t128m3 t128m4; do
  for sp in 0 2 3 4 6 8; do
    if [ $sp = 0 ]; then unset RESLIC_EBB_SPLITS; else export RESLIC_EBB_SPLITS=$sp; fi
    RESLIC_B200_LIB=$PWD/tools/dev/libs/$lib.so timeout 120 python tools/dev/eb_bwd_sweep.py 2>&1 | tail -1
  done
done
