#!/bin/bash
# round 2: CUDA-core (fp32 exact) path after the register-prefetch rewrite + geometric key blocks of the host-key kNN
mkdir -p gpurun_out
TAG=${1:-r2v}
echo "=== parity of the paths touched"
timeout 900 python -m pytest tests/test_loss_gpu.py tests/test_infonce_gpu.py tests/test_knn_gpu.py -q -m gpu --timeout 600 -x 2>&1 | tail -3
echo "=== small batches (config 1 = N 256 fp32)"
timeout 300 python tools/small_batch_probe.py 256 512 1024 > gpurun_out/${TAG}_small_batch.log 2>&1; cut -c1-400 gpurun_out/${TAG}_small_batch.log | tail -8
echo "=== host-key kNN blocks: 0 = geometric (default), 4 = equal quarters"
timeout 600 python tools/knn_blocks.py 0 4 > gpurun_out/${TAG}_knn_blocks.log 2>&1; tail -6 gpurun_out/${TAG}_knn_blocks.log
