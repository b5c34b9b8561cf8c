#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -8
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-knn --no-cpu > gpurun_out/ncu_bench.log 2>&1; tail -3 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_r1.csv
