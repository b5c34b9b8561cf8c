#!/bin/bash
# round 2, pass j (1 GPU): side-stream overlap of the staging work: parity (twice), simulated rank, bench
mkdir -p gpurun_out
TAG=${1:-r2j}
echo "=== gpu tests (loss families twice: the side stream must not race)"
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -3
timeout 1800 python -m pytest tests/test_loss_gpu.py tests/test_loss_exchange_gpu.py tests/test_infonce_gpu.py -q -m gpu --timeout 900 -x 2>&1 | tail -2
echo "=== sim rank timing world 8: side stream on / off"
timeout 300 python tools/sim_rank_step.py 32768 8 10 2>&1 | grep SIMRANK | tee gpurun_out/${TAG}_simrank.log
CLIBD_SIDE_STREAM=0 timeout 300 python tools/sim_rank_step.py 32768 8 10 2>&1 | grep SIMRANK | tee -a gpurun_out/${TAG}_simrank.log
echo "=== bench n=1 loss only: side stream on / off"
for ss in 1 0; do
CLIBD_SIDE_STREAM=$ss timeout 900 python bench.py --steps 10 --warmup 3 --no-knn --no-cpu > gpurun_out/${TAG}_bench_n1_ss$ss.json 2> gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n1_ss$ss.json') if l.startswith('{')][-1])
    print('side=$ss value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'launches', j['gpu_launches'], 'loss_check', j['loss_check']['rel_err'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print('  ', k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
except Exception as e:
    print('parse fail', e)
PY
done
echo "=== small batch"
timeout 300 python tools/small_batch_probe.py 256 4096 2>&1 | grep SMALLBATCH
