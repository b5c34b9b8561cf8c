#!/bin/bash
# round 2: compute-sanitizer racecheck (shared-memory hazards) over the CUDA-core loss kernels on the reference goldens
mkdir -p gpurun_out
timeout 250 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_loss_gpu.py -q -m gpu -x -k "fp32_path_matches_reference_golden" --timeout 240 > gpurun_out/r2zz_racecheck_simt.log 2>&1
echo "rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/r2zz_racecheck_simt.log | tail -3
