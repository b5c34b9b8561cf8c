#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loss_bwd_tc -s 2 -c 1 -o gpurun_out/prof_bwd -f python bench.py --steps 1 --warmup 1 --no-knn --no-cpu > gpurun_out/ncu_bwd.log 2>&1; tail -2 gpurun_out/ncu_bwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_screen -c 1 -o gpurun_out/prof_knn -f python tools/knn_once.py 20000 200000 > gpurun_out/ncu_knn.log 2>&1; tail -2 gpurun_out/ncu_knn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loss_fwd_tc -s 1 -c 1 -o gpurun_out/prof_fwd -f python bench.py --steps 1 --warmup 1 --no-knn --no-cpu > gpurun_out/ncu_fwd.log 2>&1; tail -2 gpurun_out/ncu_fwd.log
ls -la gpurun_out/*.ncu-rep
