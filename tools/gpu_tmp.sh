mkdir -p gpurun_out
for i in 1 2; do
CLIBD_GT_DBG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-knn --no-cpu > gpurun_out/bench_x.json 2>/dev/null
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1])
print('run: ms', j['ms_per_step'], 'bwd', j['roofline']['avg_launch_ms'], 'grad', j['roofline_grad']['avg_launch_ms'], 'fwd', j['roofline_fwd']['avg_launch_ms'], 'clk', j['clocks']['sm_mhz'])
PY
done
