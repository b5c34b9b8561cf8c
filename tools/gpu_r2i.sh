#!/bin/bash
# round 2, pass i (1 GPU): parity after class-sums / normalise fusion / shared-S threshold; simulated-rank launch list
mkdir -p gpurun_out
TAG=${1:-r2i}
echo "=== gpu tests"
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -4
echo "=== sim rank timing world 8"
timeout 300 python tools/sim_rank_step.py 32768 8 10 2>&1 | grep SIMRANK | tee gpurun_out/${TAG}_simrank.log
echo "=== sim rank ncu launch list (world 8)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_simrank_w8_launches.csv python tools/sim_rank_step.py 32768 8 2 > gpurun_out/${TAG}_simrank_ncu.log 2>&1
echo "=== small batch"
timeout 300 python tools/small_batch_probe.py 4096 2>&1 | grep SMALLBATCH
echo "=== bench n=1 loss only"
timeout 900 python bench.py --steps 10 --warmup 3 --no-knn --no-cpu > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n1.json') if l.startswith('{')][-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'launches', j['gpu_launches'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print(k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
except Exception as e:
    print('parse fail', e)
PY
