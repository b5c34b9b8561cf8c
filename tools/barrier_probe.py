"""Cost of one barrier across the ranks: this library's flag kernel (clibd_shard_barrier) against torch's
symmetric-memory barrier -- host time to enqueue and GPU time per barrier, back to back and with a small kernel between.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tools/barrier_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clibd_b200 import _peer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    px = _peer.context(None, dev, 1024 * world, 1024, 768, torch.bfloat16, world, rank)
    x = torch.zeros(1 << 16, device=dev)
    for name in ("own", "torch", "own", "torch"):
        os.environ["CLIBD_BARRIER"] = name
        for between in (False, True):
            for _ in range(50):
                px.barrier()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 1000
            t0 = time.perf_counter()
            e0.record()
            for _ in range(reps):
                px.barrier()
                if between:
                    x.add_(1.0)
            t_host = time.perf_counter() - t0
            e1.record()
            torch.cuda.synchronize()
            if rank == 0:
                print(f"BARRIER {name:5s} world={world} kernel_between={between}: host {t_host / reps * 1e6:.1f} us, "
                      f"GPU {e0.elapsed_time(e1) / reps * 1e3:.1f} us per barrier", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
