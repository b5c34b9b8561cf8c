"""Large-batch sweep (BASELINE config 5): fused contrastive loss fwd+bwd for global batch N in a list, d=768, bf16,
image-DNA and image+DNA+text, labels arange (one-hot targets) and randint(0, N/8) (multi-positive).  Prints one
JSON line per case: ms per step (CUDA events), samples/s, algorithmic tensor fraction (6 N^2 d per unordered pair
against MEASURED_PEAKS.json's sustained bf16 rate) and peak device memory -- which must stay O(N d).

    python tools/sweep_batch.py 4096 8192 32768 65536 131072 262144

Under torchrun the batch is row-sharded over the ranks (ClipLoss), as in bench.py.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clibd_b200 as cb  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [32768, 65536, 131072, 262144]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    try:
        peak_tf = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:  # noqa: BLE001
        peak_tf = 1400.0
    d = 768
    for N in sizes:
        n = N // world
        for nmod in (2, 3):
            for kind in ("arange", "randint"):
                gen = torch.Generator().manual_seed(100 + rank)
                feats = [torch.randn(n, d, generator=gen).bfloat16().to(dev) for _ in range(nmod)] + [None] * (3 - nmod)
                if kind == "arange":
                    labels = torch.arange(rank * n, (rank + 1) * n, device=dev)
                else:
                    labels = torch.randint(0, max(1, N // 8), (n,), generator=gen).to(dev)
                if world > 1:
                    mod = cb.ClipLoss(local_loss=False, gather_with_grad=True, rank=rank, world_size=world)
                else:
                    mod = cb.ContrastiveLoss(None, 1 / 0.07)
                scale = 1 / 0.07

                def step():
                    leaves = [None if f is None else f.detach().requires_grad_(True) for f in feats]
                    loss = mod(leaves[0], leaves[1], leaves[2], labels, scale)
                    loss.backward()
                    return loss

                steps = 10 if N <= 32768 else (4 if N <= 65536 else 2)
                torch.cuda.reset_peak_memory_stats(dev)
                base = torch.cuda.memory_allocated(dev)
                # warm-up: the first step of a new shape allocates (and maps) its buffers; launch-bound shapes replay CUDA
                # graphs that are captured at the second sighting of an argument tuple, and the allocator needs a few
                # steps before it hands the same blocks back every step
                for _ in range(8 if N * n <= (1 << 28) else 2):
                    loss = step()
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    loss = step()
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
                pairs = 1 if nmod == 2 else 3
                tf = 6.0 * N * N * d * pairs / (ms * 1e-3) / 1e12 / world
                if rank == 0:
                    print(json.dumps({"N": N, "n_gpus": world, "modalities": nmod, "labels": kind, "ms_per_step": ms,
                                      "samples_per_s": N / (ms * 1e-3), "algorithmic_tflops_per_gpu": tf,
                                      "frac_of_sustained_peak": tf / peak_tf, "loss": float(loss),
                                      "peak_extra_mem_gb": (torch.cuda.max_memory_allocated(dev) - base) / 2 ** 30,
                                      "input_gb_per_gpu": nmod * n * d * 2 / 2 ** 30}), flush=True)
                del feats, labels
                torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
