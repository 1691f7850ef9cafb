#!/bin/bash
cd "$(dirname "$0")/.."
for pf in 0 1; do for c in -1 3 4 5 6 8; do
  echo -n "pf=$pf ctas_per_sm=$c : "; RESLIC_GC_PREFETCH=$pf RESLIC_GC_CTAS_PER_SM=$c ./scratch/gcbench 24 98304 1 0 40
done; done
echo "--- no idx (cfg3 slice B=64), noise (cfg5 slice B=256)"
for pf in 0 1; do RESLIC_GC_PREFETCH=$pf ./scratch/gcbench 64 98304 0 0 20; RESLIC_GC_PREFETCH=$pf ./scratch/gcbench 256 16384 0 1 20; RESLIC_GC_PREFETCH=$pf ./scratch/gcbench 16 720896 1 0 10; done
