#!/bin/bash
# round 2: gradient GEMM of the first group on the side stream, next to the following sweeps (sharded step), A/B on a
# simulated rank of 8 / 4 / 2 ranks; parity of the exchange step
mkdir -p gpurun_out
TAG=${1:-r2t}
timeout 900 python -m pytest tests/test_loss_exchange_gpu.py -q -m gpu --timeout 600 -x 2>&1 | tail -3
for W in 8 4 2; do
  for OV in 0 1; do
    echo "--- world $W overlap $OV"
    CLIBD_OVERLAP_GEMM=$OV timeout 300 python tools/sim_rank_step.py 32768 $W 20 2>&1 | grep SIMRANK | cut -c1-420 | tee -a gpurun_out/${TAG}_simrank_overlap_gemm.log
  done
done
