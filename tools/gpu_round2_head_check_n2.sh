#!/bin/bash
# round 2: bench.py on 2 GPUs at HEAD (both arms, as the driver launches them)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29771 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2head_bench_n2.json 2> gpurun_out/r2head_bench_n2.err
echo "rc=$?"; grep '^{' gpurun_out/r2head_bench_n2.json | cut -c1-260
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2head_bench_n2.json') if l.startswith('{')][-1])
print({k: j.get(k) for k in ('value','ms_per_step','step_fixed_ms','tensor_kernels_overlap','cuda_graphs','shard_exchange','gpu_launches')}, j['e2e']['value'], j['loss_check']['ok'], (j.get('knn') or {}).get('value'))
PY
tail -2 gpurun_out/r2head_bench_n2.err | cut -c1-200
