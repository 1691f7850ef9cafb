#!/bin/bash
# Round-2 profiling recipe (run on the GPU box through gpurun; everything lands in gpurun_out/):
#   1. the bench command without a profiler (must exit 0), 2. its ncu launch list (per-launch durations, serialised and
#   cold: the SHARE of each kernel is what is compared with bench.py), 3. ncu --set full captures of the slice launches of
#   configs 3, 2 and 5 (dram bytes -> profiles/r02_traffic.json, stall / pipe breakdown -> profiles/r02_ncu_*.txt).
cd "$(dirname "$0")/../.."
O=gpurun_out
BENCH="python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline --no-training-kernels --no-whole-y --legs none"
$BENCH > $O/prof_r2_plain.json 2> $O/prof_r2_plain.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2.csv $BENCH > $O/prof_r2_ncu.log 2>&1
G=./tools/dev/gcbench
ncu --set full --clock-control none --import-source on -k regex:gc_fwd -s 6 -c 2 -f -o $O/prof_gc_r2_cfg3 $G B=64 n=98304 idx=0 chains=1 steps=3 reps=1 > $O/prof_r2_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gc_fwd -s 6 -c 2 -f -o $O/prof_gc_r2_cfg2 $G B=24 n=98304 idx=1 chains=1 steps=3 reps=1 > $O/prof_r2_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gc_fwd -s 6 -c 2 -f -o $O/prof_gc_r2_cfg5 $G B=256 n=16384 idx=0 noise=1 chains=1 steps=3 reps=1 > $O/prof_r2_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gc_fwd -s 6 -c 2 -f -o $O/prof_gc_r2_cfg3_shard $G B=8 n=98304 idx=0 chains=1 steps=3 reps=1 > $O/prof_r2_d.log 2>&1
ls -la $O/prof_gc_r2_*.ncu-rep
