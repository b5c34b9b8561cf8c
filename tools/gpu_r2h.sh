#!/bin/bash
# round 2, pass h (1 GPU): full test suite after the CUDA-core split / default-path / NVTX changes, launch list of the
# 1-GPU bench, small-batch probes (shared-S crossover), compute-sanitizer on a small parity case
mkdir -p gpurun_out
TAG=${1:-r2h}
echo "=== gpu tests"
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -6
echo "=== small batch probe"
timeout 300 python tools/small_batch_probe.py 256 1024 4096 2>&1 | grep SMALLBATCH | tee gpurun_out/${TAG}_small_batch.log
echo "=== shared-S crossover at N=4096 / 2048 (CLIBD_SHARED_S_MIN_N=1)"
CLIBD_SHARED_S_MIN_N=1 timeout 300 python tools/small_batch_probe.py 2048 4096 2>&1 | grep "SMALLBATCH.*bfloat16" | tee -a gpurun_out/${TAG}_small_batch.log
echo "=== ncu launch list of the loss bench (N=32768, 1 GPU)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-knn > gpurun_out/${TAG}_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.log | cut -c1-200
echo "=== compute-sanitizer memcheck: sharded exchange step (simulated ranks), small shapes"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_loss_exchange_gpu.py" -q -x -k "768-200 or 520-768" --timeout 800 > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.log
echo "=== bench n=1 full (with cpu baselines)"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -c 300 gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n1.json') if l.startswith('{')][-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'], 'frac', j['step_tensor_frac_algorithmic'], j['step_tensor_frac_algorithmic_vs_burst'])
    for k in ('roofline','roofline_fwd','roofline_grad'):
        r=j.get(k) or {}
        print(k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
    print('cfg1', j.get('config1_latency_us'))
    k=j.get('knn',{})
    print('knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'), k.get('accuracy_ms'), k.get('cpu_baseline',{}).get('value'))
    print('cpu', (j.get('cpu_baseline') or {}).get('value'), (j.get('cpu_baseline') or {}).get('kind'))
except Exception as e:
    print('parse fail', e)
PY
