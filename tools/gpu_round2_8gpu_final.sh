#!/bin/bash
# round 2, last 8-GPU pass with the final code: parity at 8 ranks, bench at 8, configs 2 (N = 4096) at 8 GPUs
mkdir -p gpurun_out
TAG=${1:-r2w}
run() { local name=$1 np=$2 port=$3; shift 3
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port "$@" > gpurun_out/${TAG}_${name}.log 2> gpurun_out/${TAG}_${name}.err
  echo "--- $name rc=$?"; }
run multigpu_check_n8 8 29701 tools/multigpu_check.py
grep -c " OK" gpurun_out/${TAG}_multigpu_check_n8.log; grep "FAIL\|MULTIGPU_CHECK_OK" gpurun_out/${TAG}_multigpu_check_n8.log | head -3
run bench_n8 8 29702 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu
run sweep_n8 8 29704 tools/sweep_batch.py 4096 8192
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n8.log') if l.startswith('{')][-1])
    print('n', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'graphs', j['cuda_graphs'], 'loss_check', j['loss_check']['ok'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print('  ', k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
    k=j.get('knn') or {}
    print('   knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'), k.get('error'))
except Exception as e:
    print('parse fail', e)
for l in open('gpurun_out/${TAG}_sweep_n8.log'):
    if l.startswith('{'):
        j=json.loads(l); print('sweep 8', j['N'], j['modalities'], j['labels'], round(j['ms_per_step'],3), 'ms')
PY
