#!/bin/bash
# round 2, pass a (2 GPUs): peer-memory probe, multi-GPU parity log, phase timing of the sharded step
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2a_topo.txt 2>&1
echo "=== peer probe"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 tools/peer_probe.py > gpurun_out/r2a_peer_probe.log 2>&1
grep -v "NCCL INFO" gpurun_out/r2a_peer_probe.log | tail -60
grep -E "NVLS|P2P|via|Channel 00" gpurun_out/r2a_peer_probe.log | head -12
echo "=== multigpu check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/multigpu_check.py > gpurun_out/r2a_multigpu_check_n2.log 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/r2a_multigpu_check_n2.log | tail -12
echo "=== phase timing"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 tools/phase_timing.py > gpurun_out/r2a_phase_timing_n2.log 2>&1
tail -30 gpurun_out/r2a_phase_timing_n2.log
