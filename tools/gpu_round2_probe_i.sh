#!/bin/bash
# round 2: BASELINE config 2 (N = 4096) and N = 8192 on 1 and 2 GPUs with enough warm-up for the CUDA-graph replay
mkdir -p gpurun_out
TAG=${1:-r2n2}
timeout 300 python tools/sweep_batch.py 4096 8192 > gpurun_out/${TAG}_sweep_small_n1.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29761 tools/sweep_batch.py 4096 8192 > gpurun_out/${TAG}_sweep_small_n2.log 2>&1
python - <<PY
import json
for w in (1, 2):
    for l in open(f'gpurun_out/${TAG}_sweep_small_n{w}.log'):
        if l.startswith('{'):
            j=json.loads(l); print('gpus', w, 'N', j['N'], j['modalities'], j['labels'], round(j['ms_per_step'],3), 'ms')
PY
