#!/bin/bash
# round 2, pass l (1 GPU): parity after un-merging the 1-GPU GEMMs / 128-row staging tiles; bench; simulated rank
mkdir -p gpurun_out
TAG=${1:-r2l}
echo "=== gpu tests"
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -3
echo "=== sim rank timing world 8 / 4 / 2"
for w in 8 4 2; do timeout 300 python tools/sim_rank_step.py 32768 $w 10 2>&1 | grep SIMRANK; done | tee gpurun_out/${TAG}_simrank.log
echo "=== bench n=1 (full)"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n1.json') if l.startswith('{')][-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'], 'frac', j['step_tensor_frac_algorithmic'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print('  ', k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'), 'traffic', r.get('traffic'))
    print('cfg1', j.get('config1_latency_us'))
    k=j.get('knn',{})
    print('knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'))
except Exception as e:
    print('parse fail', e)
PY
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
