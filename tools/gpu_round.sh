#!/bin/bash
# One measurement pass on the GPU box: parity tests, bench, ncu launch list, ncu --set full of the tensor
# kernels. Outputs go to gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
TAG=${1:-r1n}
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -4
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 300 gpurun_out/bench_$TAG.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'], 'clk', j['clocks'])
for k in ('roofline','roofline_fwd','roofline_grad'):
    r=j.get(k) or {}
    print(k, r.get('kernel'), 'ms', r.get('avg_launch_ms'), 'TF', r.get('achieved'), 'frac', r.get('frac'), 'share', r.get('share_of_step'))
print('step frac', j.get('step_tensor_frac_algorithmic'))
if 'knn' in j: print('knn', j['knn'].get('value'), j['knn'].get('ms_per_step'), j['knn'].get('e2e'))
print('cpu', j.get('cpu_baseline'))
PY
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_$TAG.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-knn > gpurun_out/launches_$TAG.log 2>&1
echo "=== ncu full: loss kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'loss_(fwd_pair|bwd_pair|grad_gemm)' -s 3 -c 3 \
  -f -o gpurun_out/prof_loss_$TAG env N=32768 python tools/pair_once.py > gpurun_out/prof_loss_$TAG.log 2>&1
tail -2 gpurun_out/prof_loss_$TAG.log
echo "=== ncu full: knn screen"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_(screen|rerank)' -c 2 \
  -f -o gpurun_out/prof_knn_$TAG python tools/knn_once.py 50000 500000 > gpurun_out/prof_knn_$TAG.log 2>&1
ls -la gpurun_out | tail -12
