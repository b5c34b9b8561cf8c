#!/bin/bash
# round 2, pass k (1 GPU): in-process A/B of the side stream; ncu --set full of one 3-modality step and of the kNN
mkdir -p gpurun_out
TAG=${1:-r2k}
echo "=== A/B side stream (1 GPU, N=32768)"
timeout 600 python tools/ab_step.py CLIBD_SIDE_STREAM 0 1 32768 6 10 2>&1 | grep ABSTEP | tee gpurun_out/${TAG}_ab_side_stream.log
echo "=== ncu full: loss kernels of one step (second iteration)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'loss_(fwd_pair|bwd_pair|grad_gemm)' -s 8 -c 8 \
  -f -o gpurun_out/prof_loss_$TAG env N=32768 ITERS=2 MODS=3 python tools/pair_once.py > gpurun_out/prof_loss_$TAG.log 2>&1
tail -2 gpurun_out/prof_loss_$TAG.log
echo "=== ncu full: knn screen + rerank"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_(screen|rerank)' -c 2 \
  -f -o gpurun_out/prof_knn_$TAG python tools/knn_once.py 50000 500000 > gpurun_out/prof_knn_$TAG.log 2>&1
tail -1 gpurun_out/prof_knn_$TAG.log
echo "=== ncu full: support kernels of the sharded step (class sums, make_operands, normalize_bwd, push)"
timeout 900 ncu --set full --clock-control none -k regex:'class_sums_jobs|make_operands|normalize_bwd|shard_push_rows' -s 12 -c 8 \
  -f -o gpurun_out/prof_support_$TAG python tools/sim_rank_step.py 32768 8 1 > gpurun_out/prof_support_$TAG.log 2>&1
tail -1 gpurun_out/prof_support_$TAG.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -4
