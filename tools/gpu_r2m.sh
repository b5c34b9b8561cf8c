#!/bin/bash
# round 2, pass m (2 GPUs): labels-first overlap of the push with the label statistics: parity, timing A/B, bench
mkdir -p gpurun_out
TAG=${1:-r2m}
echo "=== simulated-rank exchange tests"
timeout 900 python -m pytest tests/test_loss_exchange_gpu.py tests/test_loss_gpu.py -x -q --timeout 600 2>&1 | tail -3
echo "=== multigpu check"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_n2.log 2>&1
grep -c " OK" gpurun_out/${TAG}_multigpu_check_n2.log; grep "FAIL\|MULTIGPU_CHECK_OK\|Error" gpurun_out/${TAG}_multigpu_check_n2.log | head
echo "=== phase timing n2: overlap on / off"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 tools/phase_timing.py > gpurun_out/${TAG}_phase_timing_n2.log 2>&1
grep "PHASES\|FAILED\|Error" gpurun_out/${TAG}_phase_timing_n2.log | tail -4
CLIBD_OVERLAP_PUSH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/phase_timing.py > gpurun_out/${TAG}_phase_timing_n2_nooverlap.log 2>&1
grep "PHASES.*peer" gpurun_out/${TAG}_phase_timing_n2_nooverlap.log | tail -2
echo "=== bench n=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
tail -c 300 gpurun_out/${TAG}_bench_n2.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n2.json') if l.startswith('{')][-1])
    print('n', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'graphs', j['cuda_graphs'], 'loss_check', j['loss_check']['ok'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print('  ', k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
    k=j.get('knn',{})
    print('knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'), k.get('error'))
except Exception as e:
    print('parse fail', e)
PY
