#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-knn}
timeout 900 python -m pytest tests/test_knn_gpu.py -q -m gpu --timeout 600 -x 2>&1 | tail -5
timeout 600 python tools/knn_bench.py 2>&1 | tail -3 | tee gpurun_out/knn_bench_$TAG.json
