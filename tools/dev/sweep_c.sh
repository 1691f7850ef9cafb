#!/bin/bash
cd "$(dirname "$0")/../.."
G=./tools/dev/gcbench
export RESLIC_GC_BALANCE=1
for c in 3 4 6; do
$G B=8 n=98304 idx=0 chains=$c nset=6 steps=12 reps=100
$G B=32 n=16384 idx=0 noise=1 chains=$c nset=6 steps=12 reps=100
$G B=3 n=98304 idx=1 chains=$c nset=6 steps=12 reps=100
done
$G B=8 n=98304 idx=0 chains=3 steps=24 reps=50
$G B=8 n=98304 idx=0 chains=3 steps=48 reps=25
$G B=8 n=98304 idx=0 chains=3 steps=48 reps=25 prefetch=0
RESLIC_GC_MIN_CTAS=4 $G B=8 n=98304 idx=0 chains=3 steps=24 reps=50
RESLIC_GC_MIN_CTAS=4 $G B=24 n=98304 idx=1 chains=3 steps=12 reps=30
RESLIC_GC_MIN_CTAS=4 $G B=24 n=98304 idx=1 chains=1 steps=12 reps=30
RESLIC_PDL=0 $G B=8 n=98304 idx=0 chains=3 steps=24 reps=50
RESLIC_PDL=0 $G B=24 n=98304 idx=1 chains=3 steps=12 reps=30
for B in 8 64; do
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fma.sum --clock-control none -k regex:gc_fwd -c 6 --csv $G B=$B n=98304 idx=0 chains=1 steps=3 reps=1 2>&1 | grep -E "gc_fwd|Metric" | tail -8
done
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:gc_fwd -c 6 --csv $G B=24 n=98304 idx=1 chains=1 steps=3 reps=1 2>&1 | grep -E "gc_fwd" | tail -4
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:gc_fwd -c 6 --csv $G B=3 n=98304 idx=1 chains=1 steps=3 reps=1 2>&1 | grep -E "gc_fwd" | tail -4
