#!/bin/bash
# Measurement pass r1d: parity tests at HEAD, bench, step time at small batches (host/launch floor), config-5 sweep.
mkdir -p gpurun_out
TAG=${1:-r1d}
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -6
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 400 gpurun_out/bench_$TAG.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'], 'clk', j['clocks'])
print('roof', j['roofline'])
print('roof_fwd', j.get('roofline_fwd'))
if 'knn' in j: print('knn', j['knn'])
print('cpu', j.get('cpu_baseline'))
PY
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cat gpurun_out/bench_ref_$TAG.json | cut -c1-600
echo "=== sweep"; timeout 900 python tools/sweep_batch.py 4096 8192 16384 32768 65536 131072 262144 > gpurun_out/sweep_$TAG.jsonl 2> gpurun_out/sweep_$TAG.err
cat gpurun_out/sweep_$TAG.jsonl | cut -c1-330; tail -c 600 gpurun_out/sweep_$TAG.err
