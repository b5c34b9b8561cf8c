#!/bin/bash
# round 2, pass b (2 GPUs): exchange-mode parity (simulated ranks on one GPU + real 2-rank run), phase timing, bench
mkdir -p gpurun_out
TAG=${1:-r2b}
echo "=== simulated-rank exchange tests (1 GPU)"
timeout 600 python -m pytest tests/test_loss_exchange_gpu.py -x -q --timeout 300 2>&1 | tail -15
echo "=== multigpu check (2 ranks, all exchange forms)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_n2.log 2>&1
grep -v "^\*\|OMP_NUM\|UserWarning\|Consider using\|e_loss = " gpurun_out/${TAG}_multigpu_check_n2.log | tail -40
echo "=== phase timing"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 tools/phase_timing.py > gpurun_out/${TAG}_phase_timing_n2.log 2>&1
grep "PHASES\|FAILED\|Error" gpurun_out/${TAG}_phase_timing_n2.log | tail -8
echo "=== full gpu test suite (1 GPU part)"
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x --deselect tests/test_multigpu.py --deselect tests/test_loss_exchange_gpu.py 2>&1 | tail -8
echo "=== bench n=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
tail -c 400 gpurun_out/${TAG}_bench_n2.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n2.json') if l.startswith('{')][-1])
    print('n', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'loss_check', j['loss_check'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print(k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
    k=j.get('knn',{})
    print('knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'), k.get('error'))
except Exception as e:
    print('parse fail', e)
PY
