#!/bin/bash
# multi-GPU pass: NCCL parity + bench at N = 1, 2, (4, 8 when visible)
mkdir -p gpurun_out
TAG=${1:-r1}
NG=$(nvidia-smi -L | wc -l)
echo "gpus visible: $NG"
timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu --timeout 600 -x 2>&1 | tail -15
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    echo "=== bench --gpus $n"
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
    fi
    tail -c 400 gpurun_out/scale_${TAG}_n$n.err
    python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/scale_${TAG}_n$n.json') if l.startswith('{')][-1])
    print('n', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], 'bwd', j['roofline']['avg_launch_ms'], 'fwd', j['roofline_fwd']['avg_launch_ms'])
    k=j.get('knn',{})
    print('knn', k.get('value'), k.get('ms_per_step'), k.get('e2e'), k.get('error'))
except Exception as e:
    print('parse fail', e)
PY
  fi
done
