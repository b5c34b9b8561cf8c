#!/bin/bash
# scaling pass on an 8-GPU box: bench at N = 4 and 8 (N = 1, 2 are measured on cheaper boxes)
mkdir -p gpurun_out
TAG=${1:-r1k}
NG=$(nvidia-smi -L | wc -l)
echo "gpus visible: $NG"
for n in 8 4; do
  if [ $n -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
    python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/scale_${TAG}_n$n.json') if l.startswith('{')][-1])
    print('n', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], 'bwd', j['roofline']['avg_launch_ms'], 'fwd', j['roofline_fwd']['avg_launch_ms'])
    k=j.get('knn',{})
    print('knn', k.get('value'), k.get('ms_per_step'), (k.get('e2e') or {}).get('value'), k.get('error'))
except Exception as e:
    print('parse fail', e)
PY
    tail -c 300 gpurun_out/scale_${TAG}_n$n.err
  fi
done
