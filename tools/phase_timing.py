"""Where does a row-sharded loss step spend its time?  Replays the phases of clibd_b200/loss.py:_FusedClipLossFn
(collectives and C-ABI calls) at bench size with CUDA events around each phase; max over ranks, mean over steps.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P \
        tools/phase_timing.py
"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clibd_b200 import _lib  # noqa: E402
from clibd_b200.loss import _DT, _coalesced, pair_weights  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    N, d = int(os.environ.get("N", 32768)), 768
    n = N // world
    gen = torch.Generator().manual_seed(1 + rank)
    feats = [torch.randn(n, d, generator=gen).bfloat16().to(dev) for _ in range(3)]
    labels = torch.randint(0, N // 8, (n,), generator=gen).to(dev)
    scale = torch.tensor([1 / 0.07], device=dev)
    weights, _ = pair_weights([True, True, True], None, False)
    path, dtype = _lib.PATH_TC_BF16, torch.bfloat16
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    names = ["gather", "inv_norm", "fwd_stats", "allreduce_stats", "finish", "allreduce_g", "backward", "allreduce_ds"]
    acc = {k: 0.0 for k in names}
    steps, warm = 12, 4
    for it in range(steps + warm):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        k = 0
        ev[k].record(); k += 1
        if world > 1:
            all_labels = torch.empty(N, dtype=torch.int64, device=dev)
            gathered = [torch.empty((N, d), dtype=dtype, device=dev) for _ in feats]
            with _coalesced(None):
                dist.all_gather_into_tensor(all_labels.view(dtype), labels.view(dtype))
                for f, g in zip(feats, gathered):
                    dist.all_gather_into_tensor(g, f)
        else:
            all_labels, gathered = labels, feats
        ev[k].record(); k += 1
        inv = []
        for g in gathered:
            iv = torch.empty(N, dtype=torch.float32, device=dev)
            _lib.check(lib.clibd_row_inv_norm(g.data_ptr(), _DT[dtype], N, d, iv.data_ptr(), stream))
            inv.append(iv)
        ev[k].record(); k += 1
        nbytes = lib.clibd_loss_scratch_bytes(N, n, d, path)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        stats = torch.zeros(6 * N, dtype=torch.float32, device=dev)
        pos = torch.zeros(3, dtype=torch.float64, device=dev)
        xs = _lib.ptr_array3([g.data_ptr() for g in gathered])
        ivs = _lib.ptr_array3([g.data_ptr() for g in inv])
        w = _lib.float_array3(weights)
        _lib.check(lib.clibd_loss_forward_stats(xs, _DT[dtype], ivs, all_labels.data_ptr(), N, d, rank * n, n, 0.0,
                                                scale.data_ptr(), w, path, scratch.data_ptr(), nbytes, stats.data_ptr(),
                                                stats.data_ptr() + 12 * N, pos.data_ptr(), stream))
        ev[k].record(); k += 1
        if world > 1:
            dist.all_reduce(stats)
            dist.all_reduce(pos)
        ev[k].record(); k += 1
        loss = torch.empty((), dtype=torch.float32, device=dev)
        _lib.check(lib.clibd_loss_forward_finish(N, n, d, 0.0, w, path, scratch.data_ptr(), nbytes, stats.data_ptr(),
                                                 stats.data_ptr() + 12 * N, pos.data_ptr(), loss.data_ptr(), stream))
        ev[k].record(); k += 1
        gsum = torch.ones(1, device=dev)
        if world > 1:
            dist.all_reduce(gsum)
        ev[k].record(); k += 1
        dxs = [torch.empty((n, d), dtype=dtype, device=dev) for _ in range(3)]
        dscale = torch.zeros(1, dtype=torch.float64, device=dev)
        outs = _lib.ptr_array3([g.data_ptr() for g in dxs])
        _lib.check(lib.clibd_loss_backward(xs, _DT[dtype], ivs, N, d, rank * n, n, 0.0, w, path, scratch.data_ptr(), nbytes,
                                           1.0, gsum.data_ptr(), outs, dscale.data_ptr(), stream))
        ev[k].record(); k += 1
        if world > 1:
            dist.all_reduce(dscale)
        ev[k].record()
        torch.cuda.synchronize()
        if it >= warm:
            for i, nm in enumerate(names):
                acc[nm] += ev[i].elapsed_time(ev[i + 1]) / steps
    t = torch.tensor([acc[k] for k in names], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("PHASES world", world, "N", N, {k: round(float(v), 3) for k, v in zip(names, t)}, "sum", round(float(t.sum()), 3))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
