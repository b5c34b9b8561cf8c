"""kNN block of bench.py alone (development aid): python tools/knn_bench.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from clibd_b200 import _lib  # noqa: E402
from clibd_b200 import retrieval as R  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
lib = _lib.load()


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


peaks = json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))
print(json.dumps(bench.bench_knn(torch, None, R, lib, dev, 0, 1, timed, peaks.get("bf16_tflops_sustained", 1400.0))))
