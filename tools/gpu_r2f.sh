#!/bin/bash
# round 2, pass f (2 GPUs): coalesced GEMM epilogue under the emulated 8-GPU NVLink load, CUDA graphs, parity
mkdir -p gpurun_out
TAG=${1:-r2f}
echo "=== gpu tests"
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -6
echo "=== sim rank, remote peers (2 procs): world 8, 4"
for w in 8 4; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/sim_rank_step.py 32768 $w 10 2>&1 | grep "SIMRANK\[0\|Error\|error" ; done | tee gpurun_out/${TAG}_simrank_remote.log
echo "=== sim rank local (1 proc) world 8, graphs on / off"
timeout 300 python tools/sim_rank_step.py 32768 8 10 2>&1 | grep SIMRANK | tee -a gpurun_out/${TAG}_simrank_remote.log
CLIBD_GRAPHS=0 timeout 300 python tools/sim_rank_step.py 32768 8 10 2>&1 | grep SIMRANK | tee -a gpurun_out/${TAG}_simrank_remote.log
echo "=== small batch probe: graphs on"
timeout 300 python tools/small_batch_probe.py 256 1024 4096 8192 2>&1 | grep SMALLBATCH | tee gpurun_out/${TAG}_small_batch.log
echo "=== small batch probe: graphs off"
CLIBD_GRAPHS=0 timeout 300 python tools/small_batch_probe.py 256 1024 4096 8192 2>&1 | grep SMALLBATCH | tee -a gpurun_out/${TAG}_small_batch.log
echo "=== phase timing n2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 tools/phase_timing.py > gpurun_out/${TAG}_phase_timing_n2.log 2>&1
grep "PHASES\|FAILED\|Error" gpurun_out/${TAG}_phase_timing_n2.log | tail -8
echo "=== multigpu check"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_n2.log 2>&1
grep -c " OK" gpurun_out/${TAG}_multigpu_check_n2.log; grep "FAIL\|MULTIGPU_CHECK_OK\|Error" gpurun_out/${TAG}_multigpu_check_n2.log | head
echo "=== train step (config 3) on 2 GPUs, 250 per GPU"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29605 tools/train_step.py --per-gpu 250 --steps 3 --warmup 2 > gpurun_out/${TAG}_train_step_n2.log 2>&1
grep "TRAINSTEP\|Error\|error" gpurun_out/${TAG}_train_step_n2.log | cut -c1-1500 | tail -5
