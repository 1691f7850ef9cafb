#!/bin/bash
# dev: one ncu --set full capture per training-path kernel (run under gpurun; reports land in gpurun_out/)
for spec in "gap:stanh_act" "stanh_soft:stanh_gc_vec" "stanh_bwd:stanh_gc_bwd" "eb_bwd:eb_bwd" "eb_fwd:eb_fwd_fast"; do
  op=${spec%%:*}; k=${spec##*:}
  if [ -n "$1" ] && [[ " $* " != *" $op "* ]]; then continue; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_train_$op \
    python tools/dev/train_kernels_probe.py $op > gpurun_out/ncu_train_$op.log 2>&1
  tail -1 gpurun_out/ncu_train_$op.log
done
