#!/bin/bash
# round 2, pass d (1 GPU): parity after the batched support kernels, simulated-rank timing incl. column-split variants
mkdir -p gpurun_out
TAG=${1:-r2d}
echo "=== gpu tests"
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -6
echo "=== sim rank timing"
( for w in 8 4 2; do timeout 300 python tools/sim_rank_step.py 32768 $w 10 2>&1 | grep SIMRANK; done
  for js in 9 12 16 -4; do echo "CLIBD_JSPLIT_MAX=$js"; CLIBD_JSPLIT_MAX=$js timeout 300 python tools/sim_rank_step.py 32768 8 10 2>&1 | grep SIMRANK; done ) | tee gpurun_out/${TAG}_simrank.log
echo "=== sim rank ncu launch list (world 8)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_simrank_w8_launches.csv python tools/sim_rank_step.py 32768 8 2 > gpurun_out/${TAG}_simrank_ncu.log 2>&1
echo "=== bench n=1 (loss only)"
timeout 900 python bench.py --steps 10 --warmup 3 --no-knn --no-cpu > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -c 600 gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n1.json') if l.startswith('{')][-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print(k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
    print('cfg1', j.get('config1_latency_us'))
except Exception as e:
    print('parse fail', e)
PY
