"""Where does a row-sharded loss step spend its time?  For each exchange form of clibd_b200/loss.py ('local' = two
sweeps per pair, 'nccl' = S once + NCCL reduce-scatter, 'peer' = S once + stores into peer-mapped memory) the full
fwd+bwd step and the forward alone are timed with CUDA events at bench size (max over ranks, mean over steps), and the
library's own event timing of the tensor kernels gives the share that is NOT tensor work (the fixed cost).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P \
        tools/phase_timing.py [N] [steps]
"""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clibd_b200 as cb  # noqa: E402
from clibd_b200 import _lib  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    d = 768
    n = N // world
    gen = torch.Generator().manual_seed(1234 + rank)
    feats = [torch.randn(n, d, generator=gen).bfloat16().to(dev) for _ in range(3)]
    labels = torch.randint(0, N // 8, (n,), generator=gen).to(dev)
    scale = torch.tensor(1 / 0.07, device=dev)
    module = cb.ClipLoss(gather_with_grad=True, rank=rank, world_size=world) if world > 1 else cb.ContrastiveLoss(None, 1 / 0.07)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.clibd_profile_enable(1)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        lib.clibd_profile_enable(0)
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        pm, pn = (ctypes.c_double * 8)(), (ctypes.c_int64 * 8)()
        lib.clibd_profile_read(pm, pn)
        tensor_ms = torch.tensor([(pm[0] + pm[1] + pm[4]) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(tensor_ms, op=dist.ReduceOp.MAX)
        sync()
        return float(ms), float(tensor_ms), [pm[0] / max(pn[0], 1), pm[1] / max(pn[1], 1), pm[4] / max(pn[4], 1)], \
            [int(pn[0]) // steps, int(pn[1]) // steps, int(pn[4]) // steps]

    def step():
        leaves = [f.detach().requires_grad_(True) for f in feats]
        loss = module(leaves[0], leaves[1], leaves[2], labels, scale)
        loss.backward()
        return loss

    def fwd_only():
        with torch.no_grad():
            return module(feats[0], feats[1], feats[2], labels, scale)

    modes = ["local", "nccl", "peer"] if world > 1 else ["single"]
    for mode in modes:
        if world > 1:
            os.environ["CLIBD_SHARD_MODE"] = mode
        try:
            l0 = float(step())
            ms, tms, per, cnt = timed(step)
            fms, ftms, _, _ = timed(fwd_only)
            if rank == 0:
                print("PHASES " + json.dumps({
                    "world": world, "N": N, "mode": mode, "loss": l0, "step_ms": round(ms, 4),
                    "tensor_ms": round(tms, 4), "fixed_ms": round(ms - tms, 4), "fwd_ms": round(fms, 4),
                    "fwd_tensor_ms": round(ftms, 4), "fwd_fixed_ms": round(fms - ftms, 4),
                    "bwd_ms": round(ms - fms, 4), "bwd_fixed_ms": round((ms - fms) - (tms - ftms), 4),
                    "kernel_ms": {"fwd_pair": round(per[0], 4), "bwd_pair": round(per[1], 4), "grad_gemm": round(per[2], 4)},
                    "launches_per_step": {"fwd_pair": cnt[0], "bwd_pair": cnt[1], "grad_gemm": cnt[2]}}), flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"[rank {rank}] mode {mode} FAILED: {ex!r}", flush=True)
            raise
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
