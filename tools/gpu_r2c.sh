#!/bin/bash
# round 2, pass c (1 GPU): simulated-rank profile of the 8-GPU configuration, 1-GPU bench with the new legs
mkdir -p gpurun_out
TAG=${1:-r2c}
echo "=== sim rank timing"
for w in 8 4 2; do timeout 300 python tools/sim_rank_step.py 32768 $w 10 2>&1 | grep SIMRANK; done | tee gpurun_out/${TAG}_simrank.log
echo "=== sim rank ncu launch list (world 8)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_simrank_w8_launches.csv python tools/sim_rank_step.py 32768 8 2 > gpurun_out/${TAG}_simrank_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_simrank_ncu.log
echo "=== bench n=1"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -c 600 gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n1.json') if l.startswith('{')][-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
    print('loss_check', j['loss_check'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print(k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
    print('cfg1', j.get('config1_latency_us'))
    k=j.get('knn',{})
    print('knn', {a:b for a,b in k.items() if a not in ('config',)})
    print('cpu', j.get('cpu_baseline'))
except Exception as e:
    print('parse fail', e)
PY
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cut -c1-400 gpurun_out/${TAG}_bench_ref.json; tail -c 300 gpurun_out/${TAG}_bench_ref.err
