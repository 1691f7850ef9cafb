#!/bin/bash
cd "$(dirname "$0")/../.."
G=./tools/dev/gcbench
for bal in 0 1; do
export RESLIC_GC_BALANCE=$bal
for c in 1 3; do
echo "== balance=$bal chains=$c"
$G B=24 n=98304 idx=1 chains=$c steps=12 reps=30
$G B=3 n=98304 idx=1 chains=$c steps=12 reps=100
$G B=64 n=98304 idx=0 chains=$c steps=12 reps=20
$G B=8 n=98304 idx=0 chains=$c steps=12 reps=100
$G B=256 n=16384 idx=0 noise=1 chains=$c steps=12 reps=30
$G B=32 n=16384 idx=0 noise=1 chains=$c steps=12 reps=100
$G B=16 n=720896 idx=1 chains=$c steps=6 reps=6
$G B=2 n=720896 idx=1 chains=$c steps=12 reps=30
$G B=24 n=491520 smul=1 idx=1 chains=$c steps=12 reps=10
done
done
