"""Launch-bound regime probe: for small global batches, GPU time per fwd+bwd step (CUDA events) next to the host time
needed to ENQUEUE a step (wall clock of the loop before the final synchronize).  When the two agree the step is bound
by the host (Python + launches), not by the kernels.

    python tools/small_batch_probe.py [N ...]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clibd_b200 as cb  # noqa: E402
from clibd_b200 import _lib  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 1024, 4096, 8192]
    dev = torch.device("cuda:0")
    lib = _lib.load()
    for N in sizes:
        for nmod, dtype, operands in ((2, torch.float32, None), (3, torch.bfloat16, None)):
            gen = torch.Generator().manual_seed(0)
            feats = [torch.randn(N, 768, generator=gen).to(dtype).to(dev) for _ in range(nmod)] + [None] * (3 - nmod)
            labels = torch.randint(0, max(1, N // 8), (N,), generator=gen).to(dev)
            mod = cb.ContrastiveLoss(None, 1 / 0.07, tensor_core_operands=operands)
            scale = torch.tensor(1 / 0.07, device=dev)

            def step():
                leaves = [None if f is None else f.detach().requires_grad_(True) for f in feats]
                mod(leaves[0], leaves[1], leaves[2], labels, scale).backward()

            for _ in range(10):
                step()
            torch.cuda.synchronize()
            steps = 100
            l0 = lib.clibd_kernel_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            t_host = time.perf_counter() - t0
            torch.cuda.synchronize()
            print("SMALLBATCH " + json.dumps({
                "N": N, "modalities": nmod, "dtype": str(dtype)[6:], "gpu_us_per_step": round(e0.elapsed_time(e1) / steps * 1e3, 1),
                "host_enqueue_us_per_step": round(t_host / steps * 1e6, 1),
                "launches_per_step": (lib.clibd_kernel_launch_count() - l0) / steps}), flush=True)


if __name__ == "__main__":
    main()
