"""Per-rank cost of the row-sharded loss step WITHOUT the collectives: runs rank 0's share of a W-rank job on one
GPU with torch.distributed's all-gather / all-reduce replaced by local stand-ins (other ranks' rows are random
data already in place).  The gap between this and bench.py --gpus W is the cost of the collectives.
    python tools/shard_probe.py 8 [N]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clibd_b200 as cb  # noqa: E402
from clibd_b200 import loss as L  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
d, n = 768, N // W
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
pool = {}


class FakeDist:
    @staticmethod
    def is_initialized():
        return True

    @staticmethod
    def all_gather_into_tensor(out, inp, group=None):
        key = (tuple(out.shape), out.dtype)
        if key not in pool:  # other ranks' data: generated once
            if out.dtype == torch.int64:
                pool[key] = torch.randint(0, N // 8, out.shape, device=dev)
            elif out.dim() == 1:
                pool[key] = torch.full(out.shape, float(1 / d ** 0.5), device=dev, dtype=out.dtype)
            else:
                pool[key] = torch.randn(out.shape, device=dev).to(out.dtype)
        out.copy_(pool[key])       # stands in for the data movement of the gather (device-local copy)
        out[: inp.shape[0]] = inp

    @staticmethod
    def all_reduce(t, group=None, op=None):
        return None


L.dist = FakeDist
feats = [torch.randn(n, d, generator=gen).bfloat16().to(dev) for _ in range(3)]
labels = torch.randint(0, N // 8, (n,), generator=gen).to(dev)
mod = cb.ClipLoss(local_loss=False, gather_with_grad=True, rank=0, world_size=W)
scale = torch.tensor(1 / 0.07, device=dev)


def step():
    leaves = [f.detach().requires_grad_(True) for f in feats]
    loss = mod(leaves[0], leaves[1], leaves[2], labels, scale)
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
t0 = time.perf_counter()
e0.record()
for _ in range(K):
    step()
e1.record()
t_cpu = (time.perf_counter() - t0) / K * 1e3
torch.cuda.synchronize()
print(f"W={W} N={N} n={n}: {e0.elapsed_time(e1) / K:.3f} ms per step on the device, {t_cpu:.3f} ms of host time to enqueue it; "
      f"ideal (1-GPU step / W) would be {20.4 * (N / 32768) ** 2 / W:.3f} ms")
