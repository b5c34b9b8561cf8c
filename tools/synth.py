"""Seeded synthetic inputs of the loss benchmark, identical for every way of slicing them.

bench.py (any number of ranks), the full-size parity test and the golden generator (oracle/gen_golden_fullsize.py) must
see the SAME global batch, so the batch is defined block-wise: rows [b*1024, (b+1)*1024) of modality m come from
torch.Generator().manual_seed(SEED + 3*b + m); labels from one generator over the whole batch.  A rank that owns rows
[r0, r1) generates only the blocks it needs."""
import torch

SEED = 1234
BLOCK = 1024


def feature_rows(N, d, m, r0, r1, dtype=torch.bfloat16):
    """rows [r0, r1) of modality m of the N-row global batch (CPU tensor in `dtype`)."""
    out = []
    b0, b1 = r0 // BLOCK, (r1 + BLOCK - 1) // BLOCK
    for b in range(b0, b1):
        gen = torch.Generator().manual_seed(SEED + 3 * b + m)
        blk = torch.randn(BLOCK, d, generator=gen)
        lo, hi = max(r0, b * BLOCK), min(r1, (b + 1) * BLOCK, N)
        out.append(blk[lo - b * BLOCK:hi - b * BLOCK])
    return torch.cat(out, 0).to(dtype)


def labels_all(N):
    """label-matched multi-positive targets: labels ~ randint(0, N/8) (about 8 positives per row)."""
    gen = torch.Generator().manual_seed(SEED - 1)
    return torch.randint(0, max(1, N // 8), (N,), generator=gen)
