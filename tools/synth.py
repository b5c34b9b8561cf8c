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


def knn_data(dev, Q, K, d, world=1, rank=0, n_species=50_000):
    """Retrieval benchmark data (bench.py, tools/knn_verify.py): class-centroid + noise embeddings, un-normalised
    float32 on `dev`; this rank's contiguous key shard [lo, hi) of the K keys (1000 exact duplicates per shard: the
    tie-break is exercised), all Q queries, the species ids of keys and queries."""
    per = (K + world - 1) // world
    lo, hi = min(K, rank * per), min(K, (rank + 1) * per)
    gen = torch.Generator(device=dev).manual_seed(77)
    cent = torch.randn(n_species, d, device=dev, generator=gen) / d ** 0.5
    # every rank draws the same global species assignment, then keeps its shard
    sp_k = torch.randint(0, n_species, (K,), device=dev, generator=gen)
    sp_q = torch.randint(0, n_species, (Q,), device=dev, generator=gen)
    gen2 = torch.Generator(device=dev).manual_seed(1000 + rank)
    keys = cent[sp_k[lo:hi]] + 0.02 * torch.randn(hi - lo, d, device=dev, generator=gen2)
    if hi - lo > 4000:
        keys[2000:3000] = keys[0:1000]
    genq = torch.Generator(device=dev).manual_seed(2000)
    queries = cent[sp_q] + 0.02 * torch.randn(Q, d, device=dev, generator=genq)
    return queries, keys, lo, hi, sp_q, sp_k
