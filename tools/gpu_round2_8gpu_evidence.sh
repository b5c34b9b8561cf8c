#!/bin/bash
# round 2, pass g (8 GPUs): parity log at 8 ranks, bench 8 / 4, configs 2 and 5 at 8 / 4 GPUs, phase timing,
# retrieval verification at 8 GPUs, config 3 (train step, 500 per GPU x 8)
mkdir -p gpurun_out
TAG=${1:-r2g}
NG=$(nvidia-smi -L | wc -l)
echo "gpus visible: $NG"
run() { # name nproc port script args...
  local name=$1 np=$2 port=$3; shift 3
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port "$@" > gpurun_out/${TAG}_${name}.log 2> gpurun_out/${TAG}_${name}.err
  echo "--- $name rc=$?"
}
run multigpu_check_n8 8 29701 tools/multigpu_check.py
grep -c " OK" gpurun_out/${TAG}_multigpu_check_n8.log; grep "FAIL\|MULTIGPU_CHECK_OK" gpurun_out/${TAG}_multigpu_check_n8.log | head -5; tail -3 gpurun_out/${TAG}_multigpu_check_n8.err
run bench_n8 8 29702 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu
run phase_timing_n8 8 29703 tools/phase_timing.py
grep PHASES gpurun_out/${TAG}_phase_timing_n8.log
run sweep_n8 8 29704 tools/sweep_batch.py 4096 32768 131072 262144
run knn_verify_n8 8 29705 tools/knn_verify.py 100000 1000000 256
grep KNNVERIFY gpurun_out/${TAG}_knn_verify_n8.log | cut -c1-900
run train_step_n8 8 29706 tools/train_step.py --per-gpu 500 --steps 3 --warmup 2
grep TRAINSTEP gpurun_out/${TAG}_train_step_n8.log | cut -c1-1800
# two 4-GPU jobs side by side on disjoint GPUs
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29707 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_n4.log 2> gpurun_out/${TAG}_bench_n4.err ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29708 tools/sweep_batch.py 4096 32768 131072 262144 > gpurun_out/${TAG}_sweep_n4.log 2> gpurun_out/${TAG}_sweep_n4.err ) &
wait
python - <<PY
import json
for n in (8, 4):
    try:
        j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n%d.log' % n) if l.startswith('{')][-1])
        print('n', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'graphs', j['cuda_graphs'], 'ms_prof', j['ms_per_step_with_kernel_events'], 'loss_check', j['loss_check']['ok'], j['loss_check']['rel_err'])
        for k in ('roofline','roofline_fwd','roofline_grad'):
            r=j.get(k) or {}
            print('  ', k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
        k=j.get('knn',{})
        print('   knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'), k.get('error'))
    except Exception as e:
        print('parse fail', n, e)
for n in (8, 4):
    try:
        for l in open('gpurun_out/${TAG}_sweep_n%d.log' % n):
            if l.startswith('{'):
                j=json.loads(l); print('sweep', n, j['N'], j['modalities'], j['labels'], round(j['ms_per_step'],3), 'ms', round(j['frac_of_sustained_peak'],3), round(j['peak_extra_mem_gb'],2), 'GB')
    except Exception as e:
        print('sweep parse fail', n, e)
PY
tail -c 400 gpurun_out/${TAG}_bench_n8.err; tail -c 300 gpurun_out/${TAG}_sweep_n8.err
