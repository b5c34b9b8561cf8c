#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu loss"; timeout 900 python -m pytest tests/test_loss_gpu.py -q -m gpu --timeout 300 -x 2>&1 | tail -4
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-knn --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',j['value'],'ms_per_step',j['ms_per_step'],'bwd_ms',j['roofline']['avg_launch_ms'],'frac',j['roofline']['frac'],'fwd_ms',j['roofline_fwd']['avg_launch_ms'],'clk',j['clocks'])"
