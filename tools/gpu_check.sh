#!/bin/bash
# Quick validation pass: GPU parity tests + one bench line.
mkdir -p gpurun_out
TAG=${1:-chk}
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x --durations=8 2>&1 | tail -16
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e'], 'clk', j['clocks'])
print('roof', j['roofline'])
print('roof_fwd', j.get('roofline_fwd'))
print('roof_grad', j.get('roofline_grad'))
if 'knn' in j: print('knn', j['knn'])
print('cpu', j.get('cpu_baseline'))
PY
