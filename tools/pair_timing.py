"""Development aid: per-role wait-time breakdown of loss_bwd_pair_kernel (needs the CLIBD_BWD_TIMING build:
python -c "from clibd_b200 import _build; _build.build_variant('timing', ['CLIBD_BWD_TIMING'])", then run with
CLIBD_B200_LIB=clibd_b200/lib/libclibd_b200_timing.so).  Prints mean cycles per CTA spent in each mbarrier wait."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clibd_b200 as cb  # noqa: E402
from clibd_b200 import _lib  # noqa: E402

N = int(os.environ.get("N", 32768))
d = 768
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
feats = [torch.randn(N, d, generator=gen).bfloat16().to(dev) for _ in range(2)]
labels = torch.randint(0, N // 8, (N,), generator=gen).to(dev)
mod = cb.ContrastiveLoss(None, 1 / 0.07)
scale = torch.tensor(1 / 0.07, device=dev)
for it in range(3):
    leaves = [f.detach().requires_grad_(True) for f in feats]
    loss = mod(leaves[0], leaves[1], None, labels, scale)
    loss.backward()
torch.cuda.synchronize()
lib = _lib.load()
n = 1024 * 16
buf = (ctypes.c_ulonglong * n)()
lib.clibd_debug_pair_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.clibd_debug_pair_timing(buf, n) == 0
import numpy as np  # noqa: E402
a = np.array(buf[:], dtype=np.float64).reshape(1024, 16)
ncta = min(1024, 2 * ((N + 127) // 128))
a = a[:ncta]
names = ["prod wait empty (S)", "prod wait empty (G)", "mma wait g_full", "mma wait full (G)", "mma wait st_empty",
         "mma wait full (S)", "epi wait st_full", "epi wait g_empty", "epi wait acc_full", "prod total", "mma total",
         "epi total", "epi tmem ld", "epi compute", "epi write+arrive", "epi cc+bar"]
lead = a[0::2]
peer = a[1::2]
for i, nm in enumerate(names):
    print(f"{nm:24s} leader {lead[:, i].mean():12.0f}   peer {peer[:, i].mean():12.0f}")
tiles = (N // 256) / 2 if N >= 32768 else N // 256
print("tiles per item ~", tiles, " cycles per tile (epi total / tiles):", lead[:, 11].mean() / tiles)
