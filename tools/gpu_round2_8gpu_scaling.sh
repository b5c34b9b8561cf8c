#!/bin/bash
# round 2, pass n (8 GPUs): strong scaling 8 / 4 / 2 / 1 on ONE box, phase timing at 8 with / without the push overlap,
# config 2 at 8 GPUs
mkdir -p gpurun_out
TAG=${1:-r2n}
run() { local name=$1 np=$2 port=$3; shift 3
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port "$@" > gpurun_out/${TAG}_${name}.log 2> gpurun_out/${TAG}_${name}.err
  echo "--- $name rc=$?"; }
run bench_n8 8 29702 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu
run phase_timing_n8 8 29703 tools/phase_timing.py
grep PHASES gpurun_out/${TAG}_phase_timing_n8.log
CLIBD_OVERLAP_PUSH=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29713 tools/phase_timing.py 2>/dev/null | grep "PHASES.*peer" | sed 's/^/CLIBD_OVERLAP_PUSH=0 /' | tee -a gpurun_out/${TAG}_phase_timing_n8.log
run sweep_n8 8 29704 tools/sweep_batch.py 4096 8192
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29707 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu --no-knn > gpurun_out/${TAG}_bench_n4.log 2> gpurun_out/${TAG}_bench_n4.err ) &
( CUDA_VISIBLE_DEVICES=4,5 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-knn > gpurun_out/${TAG}_bench_n2.log 2> gpurun_out/${TAG}_bench_n2.err ) &
( CUDA_VISIBLE_DEVICES=6 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-knn > gpurun_out/${TAG}_bench_n1.log 2> gpurun_out/${TAG}_bench_n1.err ) &
wait
python - <<PY
import json
v={}
for n in (8, 4, 2, 1):
    try:
        j=json.loads([l for l in open('gpurun_out/${TAG}_bench_n%d.log' % n) if l.startswith('{')][-1])
        v[n]=j['value']
        print('n', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'fixed', j['step_fixed_ms'], 'e2e', j['e2e']['value'], 'graphs', j['cuda_graphs'], 'loss_check', j['loss_check']['ok'])
        for k in ('roofline','roofline_fwd','roofline_grad'):
            r=j.get(k) or {}
            print('  ', k, 'ms', r.get('avg_launch_ms'), 'n', r.get('launches'), 'frac', r.get('frac'))
        k=j.get('knn') or {}
        if k: print('   knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'), k.get('error'))
    except Exception as e:
        print('parse fail', n, e)
if 1 in v:
    print('efficiency', {n: round(v[n]/(n*v[1]),3) for n in v})
for l in open('gpurun_out/${TAG}_sweep_n8.log'):
    if l.startswith('{'):
        j=json.loads(l); print('sweep 8', j['N'], j['modalities'], j['labels'], round(j['ms_per_step'],3), 'ms')
PY
tail -c 300 gpurun_out/${TAG}_bench_n8.err
