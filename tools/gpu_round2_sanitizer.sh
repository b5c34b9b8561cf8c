#!/bin/bash
# round 2: compute-sanitizer memcheck over (a) the rewritten CUDA-core loss kernels on every reference golden (fp32 exact
# path), (b) the barrier kernel, (c) the sharded step with the side-stream gradient GEMM and three strip buffers
mkdir -p gpurun_out
TAG=${1:-r2zz}
run() { local name=$1 limit=$2; shift 2
  timeout $limit compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest "$@" -q -m gpu -x --timeout $limit > gpurun_out/${TAG}_sanitizer_${name}.log 2>&1
  echo "$name rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_sanitizer_${name}.log | tail -2; }
if [ "$2" != "exchange_only" ]; then
run simt 280 tests/test_loss_gpu.py -k "fp32_path_matches_reference_golden"
run barrier 120 tests/test_loss_exchange_gpu.py -k "barrier"
fi
run exchange 150 tests/test_loss_exchange_gpu.py -k "gradient_gemm_on_the_side_stream or unequal or pair_filters"
