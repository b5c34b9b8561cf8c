#!/bin/bash
# first end-to-end GPU session: every stage under its own timeout so a hung kernel cannot hold the box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "=== debug N=512"; timeout 120 python tools/debug_tc.py 512 768 2>&1 | tail -20
echo "=== debug N=1000 d=200"; timeout 120 python tools/debug_tc.py 1000 200 2>&1 | tail -20
echo "=== fp32 golden"; timeout 300 python -m pytest tests/test_loss_gpu.py -q -k "fp32_path or errors" --timeout 120 2>&1 | tail -15
echo "=== tc oracle"; timeout 400 python -m pytest tests/test_loss_gpu.py -q -k "tensor_core or bf16_fed" --timeout 150 2>&1 | tail -25
echo "=== knn small"; timeout 300 python -m pytest tests/test_knn_gpu.py -q -k "small or identical or shard or make_prediction" --timeout 120 2>&1 | tail -25
echo "=== big"; timeout 500 python -m pytest tests/test_loss_gpu.py tests/test_knn_gpu.py -q -k "full_size or medium or large" --timeout 240 2>&1 | tail -25
echo "=== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -5
