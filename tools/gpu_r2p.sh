#!/bin/bash
# round 2, pass p (1 GPU): staging-tile variants (swizzled 64-row vs 128-row) under ncu
mkdir -p gpurun_out
for v in "" "_mo128"; do
  echo "=== make_operands, library variant '$v'"
  CLIBD_B200_LIB=$PWD/clibd_b200/lib/libclibd_b200$v.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:make_operands -s 4 -c 8 --csv python tools/sim_rank_step.py 32768 2 1 2>/dev/null | grep make_operands | awk -F'","' '{print $10, $NF}' | tr -d '"'
done
