"""one kNN search at bench size (for ncu captures)"""
import sys
import torch
sys.path.insert(0, ".")
from clibd_b200 import retrieval as R
dev = torch.device("cuda:0")
Q, K, d = int(sys.argv[1]) if len(sys.argv) > 1 else 20000, int(sys.argv[2]) if len(sys.argv) > 2 else 200000, 768
g = torch.Generator(device=dev).manual_seed(0)
q = R.normalize_rows(torch.randn(Q, d, device=dev, generator=g), dev)
k = R.normalize_rows(torch.randn(K, d, device=dev, generator=g), dev)
s, i, n = R.search_normalized(q, k, 5, mode="fp16")
torch.cuda.synchronize()
print("done", int(n))
