#!/bin/bash
# round 2, last 1-GPU pass: whole parity suite, smoke, bench (ours + reference arm)
mkdir -p gpurun_out
TAG=${1:-r2zz}
echo "=== gpu tests"
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -3
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench n=1 (full)"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n1.json') if l.startswith('{')][-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'], 'frac', j['step_tensor_frac_algorithmic'], j['step_tensor_frac_algorithmic_vs_burst'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print('  ', k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'), 'traffic', r.get('traffic'))
    print('cfg1', j.get('config1_latency_us'), 'clocks', j.get('clocks'), 'loss_check', j['loss_check']['ok'])
    k=j.get('knn',{})
    print('knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'), k.get('accuracy_ms'), (k.get('cpu_baseline') or {}).get('value'))
    print('cpu', (j.get('cpu_baseline') or {}).get('value'), (j.get('cpu_baseline') or {}).get('kind'), (j.get('cpu_baseline') or {}).get('cores'))
except Exception as e:
    print('parse fail', e)
PY
tail -3 gpurun_out/${TAG}_bench_n1.err | cut -c1-300
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cut -c1-260 gpurun_out/${TAG}_bench_ref.json
