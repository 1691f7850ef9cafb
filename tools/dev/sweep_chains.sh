#!/bin/bash
# chain-overlap sweep of the GC slice loop (development): BASELINE shapes, full batch and the 8-GPU shard batch
cd "$(dirname "$0")/../.."
G=./tools/dev/gcbench
for c in 1 2 3; do
$G B=24 n=98304 idx=1 chains=$c steps=12 reps=30
$G B=3 n=98304 idx=1 chains=$c steps=12 reps=100
$G B=64 n=98304 idx=0 chains=$c steps=12 reps=20
$G B=8 n=98304 idx=0 chains=$c steps=12 reps=100
$G B=256 n=16384 idx=0 noise=1 chains=$c steps=12 reps=30
$G B=32 n=16384 idx=0 noise=1 chains=$c steps=12 reps=100
$G B=16 n=720896 idx=1 chains=$c steps=6 reps=6
$G B=2 n=720896 idx=1 chains=$c steps=12 reps=30
done
$G B=8 n=98304 idx=0 chains=4 nset=4 steps=12 reps=100
$G B=8 n=98304 idx=0 chains=6 nset=6 steps=12 reps=100
$G B=8 n=98304 idx=0 chains=1 steps=12 reps=100 rate=0
$G B=8 n=98304 idx=0 chains=1 steps=12 reps=100 prefetch=0
$G B=1 n=98304 idx=0 chains=1 steps=12 reps=100
$G B=8 n=491520 smul=1 idx=0 chains=1 steps=12 reps=100
$G B=8 n=491520 smul=1 idx=0 chains=3 steps=12 reps=100
