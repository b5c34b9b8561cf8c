#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -4
for cfg in 0 1 2; do
  echo "=== bwd cfg $cfg"; CLIBD_BWD_CFG=$cfg timeout 300 python bench.py --steps 6 --warmup 2 --no-knn --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step',j['ms_per_step'],'bwd_ms',j['roofline']['avg_launch_ms'],'fwd_ms',j['roofline_fwd']['avg_launch_ms'],'clk',j['clocks'])"
done
echo "=== bench full"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; python -c "
import json
j=json.loads(open('gpurun_out/bench2.json').read().strip().splitlines()[-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'])
print('knn',{k:j['knn'][k] for k in ('value','ms_per_step','queries_redone_exhaustively','rerank_ms')}, j['knn']['roofline'], j['knn']['e2e'])
print('cpu',j['cpu_baseline']['value'])"
