#!/bin/bash
# round 2: ncu launch list of the final bench.py step (gpu__time_duration.sum, no clock control) + parity of the loss paths
mkdir -p gpurun_out
TAG=${1:-r2zz}
timeout 600 python -m pytest tests/test_loss_gpu.py -q -m gpu --timeout 300 -x 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
grep -c loss_bwd_pair gpurun_out/${TAG}_launches_bench_steps2.csv
timeout 300 python - <<'PY'
import torch, statistics
import clibd_b200 as cb
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
a = torch.randn(256, 768, generator=gen).to(dev); b = torch.randn(256, 768, generator=gen).to(dev)
labels = torch.arange(256, device=dev)
mod = cb.ContrastiveLoss(None, 1 / 0.07)
for grad in (True, False):
    scale = torch.tensor(1 / 0.07, device=dev, requires_grad=grad)
    ts = []
    for it in range(80):
        la, lb = a.detach().requires_grad_(True), b.detach().requires_grad_(True)
        scale.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mod(la, lb, None, labels, scale).backward(); e1.record(); e1.synchronize()
        if it >= 20: ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"CONFIG1 learnable_scale={grad}: median {statistics.median(ts):.1f} us", flush=True)
PY
