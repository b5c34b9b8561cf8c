#!/bin/bash
# round 2 (2 GPUs): A/B of the side-stream gradient GEMM with the per-GPU NVLink load of the 8-GPU job emulated
# (tools/sim_rank_step.py under torchrun: the 7 simulated peers live on the other GPU); multi-GPU parity incl. gather_features;
# host-side cost of a small sharded step
mkdir -p gpurun_out
TAG=${1:-r2s2}
run() { local name=$1 port=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port "$@" > gpurun_out/${TAG}_${name}.log 2> gpurun_out/${TAG}_${name}.err
  echo "--- $name rc=$?"; }
for OV in 0 1 0 1; do
  CLIBD_OVERLAP_GEMM=$OV run simrank_ov$OV 2971$OV tools/sim_rank_step.py 32768 8 20
  grep SIMRANK gpurun_out/${TAG}_simrank_ov$OV.log | cut -c1-330
done
run multigpu_check 29721 tools/multigpu_check.py
grep -c " OK" gpurun_out/${TAG}_multigpu_check.log; grep "FAIL\|MULTIGPU_CHECK_OK\|gather_features" gpurun_out/${TAG}_multigpu_check.log | head -12
tail -3 gpurun_out/${TAG}_multigpu_check.err
run host_probe 29722 tools/host_overhead_probe.py 512 300
head -45 gpurun_out/${TAG}_host_probe.log | cut -c1-200
