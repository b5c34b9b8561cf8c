"""kNN end-to-end time (host-resident keys, block-pipelined copy) for several block counts (0 = the default:
geometrically growing blocks)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clibd_b200 import retrieval as R  # noqa: E402

dev = torch.device("cuda:0")
Q, K, d = 100_000, 1_000_000, 768
gen = torch.Generator(device=dev).manual_seed(1)
cent = torch.randn(50_000, d, device=dev, generator=gen) / d ** 0.5
keys = (cent[torch.randint(0, 50_000, (K,), device=dev, generator=gen)] + 0.02 * torch.randn(K, d, device=dev, generator=gen)).cpu().pin_memory()
q = (cent[torch.randint(0, 50_000, (Q,), device=dev, generator=gen)] + 0.02 * torch.randn(Q, d, device=dev, generator=gen)).cpu().pin_memory()
del cent
for blocks in [int(a) for a in sys.argv[1:]] or [0, 4, 8]:
    R._search_host_keys_pipelined.__defaults__ = (blocks or None, 0)
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s, i = R.knn_search(q, keys, 5, mode="fp16", device=dev)
        ih = i.cpu()
        e1.record()
        torch.cuda.synchronize()
        if it:
            print("blocks", blocks, "ms", round(e0.elapsed_time(e1), 2))
