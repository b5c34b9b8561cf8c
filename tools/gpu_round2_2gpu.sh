#!/bin/bash
# round 2, final 2-GPU pass: parity log, configs 2 and 5 at 2 GPUs, bench at 2
mkdir -p gpurun_out
TAG=${1:-r2y}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_n2.log 2>&1
grep -c " OK" gpurun_out/${TAG}_multigpu_check_n2.log; grep "FAIL\|MULTIGPU_CHECK_OK\|Error" gpurun_out/${TAG}_multigpu_check_n2.log | head
timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu --timeout 800 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29604 tools/sweep_batch.py 4096 8192 32768 65536 131072 262144 > gpurun_out/${TAG}_sweep_n2.log 2> gpurun_out/${TAG}_sweep_n2.err
python - <<PY
import json
for l in open('gpurun_out/${TAG}_sweep_n2.log'):
    if l.startswith('{'):
        j=json.loads(l); print('sweep 2', j['N'], j['modalities'], j['labels'], round(j['ms_per_step'],3), 'ms', round(j['frac_of_sustained_peak'],3), round(j['peak_extra_mem_gb'],2), 'GB')
PY
tail -c 300 gpurun_out/${TAG}_sweep_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29605 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n2.json') if l.startswith('{')][-1])
    print('n', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'loss_check', j['loss_check']['ok'])
    k=j.get('knn',{})
    print('knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'), k.get('error'))
except Exception as e:
    print('parse fail', e)
PY
