"""HBM roofline of the seam kernels (csrc/seam.cu) next to the eager PyTorch expressions they replace.
    python tools/seam_bench.py [n] -> JSON lines"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clibd_b200 import seam  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
T, C = 133, 768
dev = torch.device("cuda:0")
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except Exception:  # noqa: BLE001
    pass
peak = peaks.get("hbm_gbs", 6500.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(iters):
        flush.zero_()  # L2 flush between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


for dtype in (torch.float32, torch.bfloat16):
    b = 4 if dtype == torch.float32 else 2
    logits = (torch.randn(n, T, C, device=dev) * 3).to(dtype)
    gout = torch.randn(n, C, device=dev).to(dtype)
    x = logits.clone().requires_grad_(True)

    def ours_fwd():
        return seam.softmax_mean(logits)

    def eager_fwd():
        return logits.softmax(dim=-1).mean(dim=1)

    def ours_fb():
        x.grad = None
        seam.softmax_mean(x).backward(gout)

    def eager_fb():
        x.grad = None
        x.softmax(dim=-1).mean(dim=1).backward(gout)

    fwd_bytes = n * T * C * b + n * C * b
    fb_bytes = fwd_bytes + 2 * n * T * C * b + n * C * b
    t_f, t_ef, t_fb, t_efb = timed(ours_fwd), timed(eager_fwd), timed(ours_fb), timed(eager_fb)
    print(json.dumps({"kernel": "softmax_mean", "dtype": str(dtype), "n": n, "tokens": T, "classes": C,
                      "fwd_ms": t_f, "fwd_GBs": fwd_bytes / t_f / 1e6, "fwd_frac_hbm": fwd_bytes / t_f / 1e6 / peak,
                      "fwd_bwd_ms": t_fb, "fwd_bwd_GBs": fb_bytes / t_fb / 1e6,
                      "fwd_bwd_frac_hbm": fb_bytes / t_fb / 1e6 / peak,
                      "eager_fwd_ms": t_ef, "eager_fwd_bwd_ms": t_efb, "hbm_peak_GBs": peak}))
    del logits, x

# embedding hand-off: one 4000 x 768 batch appended to the device store vs the reference's host round trip
feat = torch.randn(n, C, device=dev)
store = seam.EmbeddingStore(capacity=64 * n)


def ours_append():
    store.rows = 0
    store.append(feat)


def ref_roundtrip():
    return torch.nn.functional.normalize(feat, dim=-1).cpu().tolist()


t_a = timed(ours_append)
import time  # noqa: E402
t0 = time.perf_counter()
ref_roundtrip()
t_r = (time.perf_counter() - t0) * 1e3
ab = n * C * 8
print(json.dumps({"kernel": "embed_append", "n": n, "d": C, "ms": t_a, "GBs": ab / t_a / 1e6, "frac_hbm": ab / t_a / 1e6 / peak,
                  "reference_normalize_cpu_tolist_ms": t_r}))
