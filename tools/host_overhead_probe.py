"""Where the HOST time of a sharded (peer-form) loss step goes: cProfile of rank 0 over many small steps.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
        tools/host_overhead_probe.py [n_per_rank] [steps]

Small batches are bound by the time the host needs to enqueue a step (the GPU work of n = 512 rows per rank is ~0.3 ms);
prints wall-clock enqueue time per step (no synchronisation inside the loop) and the 25 most expensive functions."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clibd_b200 as cb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gen = torch.Generator().manual_seed(rank)
    feats = [torch.randn(n, 768, generator=gen).to(torch.bfloat16).to(dev) for _ in range(3)]
    labels = torch.randint(0, 64, (n,), generator=gen).to(dev)
    scale = torch.tensor(1 / 0.07, device=dev)
    mod = cb.ClipLoss(gather_with_grad=True, rank=rank, world_size=world) if world > 1 else cb.ContrastiveLoss(None, 1 / 0.07)

    def step():
        leaves = [f.detach().requires_grad_(True) for f in feats]
        mod(leaves[0], leaves[1], leaves[2], labels, scale).backward()

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    for tag in ("plain", "profiled"):
        prof = cProfile.Profile() if tag == "profiled" else None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if prof:
            prof.enable()
        for _ in range(steps):
            step()
        if prof:
            prof.disable()
        t_host = time.perf_counter() - t0
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"HOSTPROBE {tag} world={world} n={n}: host enqueue {t_host / steps * 1e6:.1f} us/step, "
                  f"GPU {e0.elapsed_time(e1) / steps * 1e3:.1f} us/step", flush=True)
            if prof:
                out = io.StringIO()
                pstats.Stats(prof, stream=out).sort_stats("tottime").print_stats(28)
                print(out.getvalue(), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
