#!/bin/bash
# round 2, pass o (1 GPU): fp32-path error probe, launch list of the final bench command
mkdir -p gpurun_out
TAG=${1:-r2o}
echo "=== fp32 error probe"
timeout 300 python tools/fp32_error_probe.py 2>&1 | grep FP32ERR | tee gpurun_out/${TAG}_fp32_error.log | cut -c1-400
echo "=== ncu launch list of the loss bench (N=32768, 1 GPU)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-knn > gpurun_out/${TAG}_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.log | cut -c1-100
