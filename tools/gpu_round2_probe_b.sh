#!/bin/bash
# round 2: per-kernel launch list of rank 0 of an 8-rank step, simulated on one GPU (tools/sim_rank_step.py)
mkdir -p gpurun_out
TAG=${1:-r2u}
timeout 300 python tools/sim_rank_step.py 32768 8 20 > gpurun_out/${TAG}_simrank.log 2>&1; cut -c1-700 gpurun_out/${TAG}_simrank.log | tail -2
CLIBD_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_simrank_w8_launches.csv python tools/sim_rank_step.py 32768 8 1 > gpurun_out/${TAG}_simrank_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_simrank_ncu.log | cut -c1-300
