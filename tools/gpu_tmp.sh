mkdir -p gpurun_out
for m in 0 1 2 3; do
CLIBD_GT_MODE=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-knn --no-cpu > gpurun_out/bench_m$m.json 2>/dev/null
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_m$m.json').read().strip().splitlines()[-1])
print('mode $m: ms', j['ms_per_step'], 'bwd', j['roofline']['avg_launch_ms'], 'grad', j['roofline_grad']['avg_launch_ms'], 'fwd', j['roofline_fwd']['avg_launch_ms'], 'clk', j['clocks']['sm_mhz'])
PY
done
CLIBD_GT_MODE=3 timeout 600 python -m pytest tests/test_loss_gpu.py -q -m gpu -x 2>&1 | tail -2
