#!/bin/bash
cd "$(dirname "$0")/../.."
export RESLIC_GC_BALANCE=1
for v in v2 v4 v5 v6; do
for mc in 0 4; do
export RESLIC_GC_MIN_CTAS=$mc
G=./tools/dev/gcbench_$v
for c in 1 3; do
echo "== $v min_ctas=$mc chains=$c"
$G B=24 n=98304 idx=1 chains=$c steps=12 reps=30
$G B=64 n=98304 idx=0 chains=$c steps=12 reps=20
$G B=8 n=98304 idx=0 chains=$c steps=12 reps=100
$G B=256 n=16384 idx=0 noise=1 chains=$c steps=12 reps=30
$G B=32 n=16384 idx=0 noise=1 chains=$c steps=12 reps=100
done
done
done
