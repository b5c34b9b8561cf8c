"""Parity evidence for the retrieval at BENCHMARK size (100k queries x 1M keys, d = 768, k = 5), 1..W GPUs.

    python tools/knn_verify.py [Q] [K] [samples]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P tools/knn_verify.py

(1) `samples` evenly spaced queries of the (sharded + merged) result are compared with the C oracle
    (oracle/knn_oracle.c: exhaustive search, float64 sims summed in d order, order (-sim, index)) over ALL keys:
    indices and float64 similarities must be bit-identical.
(2) faiss is absent (parity of the search is unpinned, oracle/knn_oracle.py); what IndexFlatIP computes is restated
    as float32 `torch.mm` (TF32 off) + `topk` per key block; the report states how often that float32 ranking
    agrees with the exact (-sim, index) ranking on this data -- disagreements are float32 rounding of near-ties.
Prints one JSON line (rank 0).  TEST / EVIDENCE TOOL: uses the oracle as the checker."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clibd_b200 import retrieval as R  # noqa: E402
from tools import synth  # noqa: E402


def run(Q, K, samples, dev, world=1, rank=0, d=768, k=5):
    from oracle import knn_oracle as ko
    queries, keys, lo, hi, _, _ = synth.knn_data(dev, Q, K, d, world, rank)
    q32 = R.normalize_rows(queries, dev)
    k32 = R.normalize_rows(keys, dev)
    del queries, keys
    s64, idx, nex = R.search_normalized(q32, k32, k, key_offset=lo, mode="fp16")
    s64, _, idx = R._merge_over_ranks(s64, idx, world, None)
    # ---- (2) float32 IndexFlatIP restatement on the GPU: sgemm blocks + topk, merged over shards the same way
    torch.backends.cuda.matmul.allow_tf32 = False
    bs, bi = None, None
    for c0 in range(0, hi - lo, 131072):
        s = q32 @ k32[c0:c0 + 131072].T
        ts, ti = s.topk(min(k, s.shape[1]), dim=1)
        ti = ti + (lo + c0)
        if bs is None:
            bs, bi = ts, ti
        else:
            cs, ci = torch.cat([bs, ts], 1), torch.cat([bi, ti], 1)
            bs, sel = cs.topk(k, dim=1)
            bi = ci.gather(1, sel)
        del s
    if world > 1:
        all_s = [torch.empty_like(bs) for _ in range(world)]
        all_i = [torch.empty_like(bi) for _ in range(world)]
        dist.all_gather(all_s, bs)
        dist.all_gather(all_i, bi)
        cs, ci = torch.cat(all_s, 1), torch.cat(all_i, 1)
        bs, sel = cs.topk(k, dim=1)
        bi = ci.gather(1, sel)
    same_ordered = float((bi == idx).all(dim=1).float().mean())
    same_set = float((torch.sort(bi, 1).values == torch.sort(idx, 1).values).all(dim=1).float().mean())
    same_top1 = float((bi[:, 0] == idx[:, 0]).float().mean())
    # ---- (1) sampled queries against the C oracle over ALL keys (rank 0 collects the normalised shards)
    pick = torch.linspace(0, Q - 1, samples, device=dev).long()
    if world > 1:
        parts = [torch.empty((min(K, (r + 1) * ((K + world - 1) // world)) - min(K, r * ((K + world - 1) // world)), d),
                             dtype=torch.float32, device=dev) for r in range(world)] if rank == 0 else None
        # shards may differ in length: gather with point-to-point copies
        if rank == 0:
            parts[0] = k32
            for r in range(1, world):
                dist.recv(parts[r], src=r)
            k_all = torch.cat(parts, 0).cpu().numpy()
        else:
            dist.send(k32, dst=0)
            k_all = None
    else:
        k_all = k32.cpu().numpy()
    out = None
    if rank == 0:
        t0 = time.time()
        _, ref_i, ref_s = ko.search(q32[pick].cpu().numpy(), k_all, k)
        t_oracle = time.time() - t0
        got_i, got_s = idx[pick].cpu().numpy(), s64[pick].cpu().numpy()
        out = {"queries": Q, "keys": K, "dim": d, "k": k, "n_gpus": world, "sampled_queries": int(samples),
               "oracle": "oracle/knn_oracle.c (exhaustive, float64 sims in d order, (-sim, index))",
               "oracle_seconds": round(t_oracle, 1),
               "indices_bit_exact": bool(np.array_equal(got_i, ref_i)),
               "similarities_bit_exact": bool(np.array_equal(got_s, ref_s)),
               "queries_redone_exhaustively": int(nex),
               "fp32_indexflatip_restatement": {
                   "what": "float32 torch.mm (TF32 off) + topk per 131072-key block, merged over blocks and shards",
                   "top5_identical_ordered": same_ordered, "top5_identical_as_sets": same_set,
                   "top1_identical": same_top1}}
    return out


def main():
    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    samples = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = run(Q, K, samples, dev, world, rank)
    if rank == 0:
        print("KNNVERIFY " + json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not (out["indices_bit_exact"] and out["similarities_bit_exact"]):
        sys.exit(1)


if __name__ == "__main__":
    main()
