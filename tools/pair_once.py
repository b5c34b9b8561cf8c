"""a few loss fwd+bwd steps at bench size (for ncu captures of the tensor kernels); MODS=2 image+DNA, MODS=3 (default)
image+DNA+text -- one step then launches 3 forward kernels, 3 row sweeps and 2 gradient GEMMs (the second one contracts
the concatenated strips of the two text-column pairs)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clibd_b200 as cb  # noqa: E402

N = int(os.environ.get("N", 32768))
d = 768
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
MODS = int(os.environ.get("MODS", 3))
feats = [torch.randn(N, d, generator=gen).bfloat16().to(dev) for _ in range(MODS)] + [None] * (3 - MODS)
labels = torch.randint(0, N // 8, (N,), generator=gen).to(dev)
mod = cb.ContrastiveLoss(None, 1 / 0.07)
scale = torch.tensor(1 / 0.07, device=dev)
for it in range(int(os.environ.get("ITERS", 3))):
    leaves = [None if f is None else f.detach().requires_grad_(True) for f in feats]
    loss = mod(leaves[0], leaves[1], leaves[2], labels, scale)
    loss.backward()
torch.cuda.synchronize()
print("done", float(loss))
