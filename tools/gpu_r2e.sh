#!/bin/bash
# round 2, pass e (2 GPUs): NVLink load of the 8-GPU job emulated on 2 GPUs; parity of the sharded entry points
mkdir -p gpurun_out
TAG=${1:-r2e}
echo "=== sim rank, remote peers (2 procs): world 8, 4"
for w in 8 4; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/sim_rank_step.py 32768 $w 10 2>&1 | grep "SIMRANK\|Error\|error" ; done | tee gpurun_out/${TAG}_simrank_remote.log
echo "=== sim rank local (1 proc) world 8"
timeout 300 python tools/sim_rank_step.py 32768 8 10 2>&1 | grep SIMRANK | tee -a gpurun_out/${TAG}_simrank_remote.log
echo "=== multigpu check"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_n2.log 2>&1
grep -v "^\*\|OMP_NUM\|UserWarning\|Consider using\|e_loss = " gpurun_out/${TAG}_multigpu_check_n2.log | tail -42
echo "=== gpu tests (loss + knn)"
timeout 1500 python -m pytest tests/test_loss_gpu.py tests/test_loss_exchange_gpu.py tests/test_knn_gpu.py -q -m gpu --timeout 600 -x 2>&1 | tail -4
echo "=== phase timing n2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 tools/phase_timing.py > gpurun_out/${TAG}_phase_timing_n2.log 2>&1
grep "PHASES\|FAILED\|Error" gpurun_out/${TAG}_phase_timing_n2.log | tail -8
