"""One rank's exchange-mode loss step on ONE GPU at bench size, for cheap profiling of the W-GPU configuration.

    python tools/sim_rank_step.py [N] [world] [steps]          per-call CUDA-event timing
    ncu --metrics gpu__time_duration.sum ... python tools/sim_rank_step.py 32768 8 2     per-kernel launch list

Rank 0 of a `world`-rank job owns n = N / world rows; its kernels (staging of all N gathered rows, forward over its row
block, row sweep + coefficient strip + gradient GEMM, normalise-backward) do exactly the work they do on a `world`-GPU
box.  The gathered batch is filled in locally and the "peer" slot arrays live on the same device, so only the NVLink
transfer time of the pushes and of the gradient GEMM's epilogue stores is missing (HBM stores instead).

Under torchrun with 2 processes (2 GPUs) the buffers of the simulated ranks 1..world-1 are the OTHER GPU's
peer-mapped memory: every process plays rank 0 of a `world`-rank job and sends (world-1)/world of its pushes and
gradient rows over NVLink -- the per-GPU link load of the real `world`-GPU job, on a 2-GPU box."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clibd_b200 import _lib  # noqa: E402
from clibd_b200.loss import _DT, _column_slots  # noqa: E402
from tools import synth  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    d, rank = 768, 0
    n = N // world
    nproc = int(os.environ.get("WORLD_SIZE", 1))
    me = int(os.environ.get("RANK", 0))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    remote = None
    if nproc > 1:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    dtype = torch.bfloat16
    dt = _DT[dtype]
    path, mode = _lib.PATH_TC_BF16, _lib.MODE_EXCHANGE
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    weights = (1 / 6, 1 / 6, 1 / 6)
    w = _lib.float_array3(weights)
    full = [synth.feature_rows(N, d, m, 0, N).to(dev) for m in range(3)]
    labels = synth.labels_all(N).to(dev)
    # the other ranks' rows are already "gathered"; rank 0 pushes its own block every step
    gx = [f.clone() for f in full]
    ginv = [torch.empty(N, device=dev) for _ in range(3)]
    for m in range(3):
        _lib.check(lib.clibd_row_inv_norm(gx[m].data_ptr(), dt, N, d, ginv[m].data_ptr(), stream))
    glab = labels.clone()
    def other(nbytes, dtype, shape):
        """a buffer that stands for the simulated ranks 1..world-1: local, or the other GPU's memory under torchrun"""
        if nproc == 1:
            return torch.empty(shape, dtype=dtype, device=dev)
        t = symm.empty(shape, dtype=dtype, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        keep.append((t, hdl))
        return hdl.get_buffer((me + 1) % nproc, shape, dtype)

    keep = []
    dummy_x = [other(0, f.dtype, tuple(f.shape)) for f in full]
    dummy_inv = [other(0, torch.float32, (N,)) for _ in range(3)]
    dummy_lab = other(0, torch.int64, (N,))
    stats = torch.zeros(9 * N, device=dev)
    dummy_stats = other(0, torch.float32, (9 * N,))
    colslots = torch.zeros(world * 3 * N, device=dev)
    posslots = torch.zeros(world * 4, dtype=torch.float64, device=dev)
    pos_local = torch.zeros(4, dtype=torch.float64, device=dev)
    pos = torch.zeros(4, dtype=torch.float64, device=dev)
    nbytes = lib.clibd_loss_scratch_bytes(N, n, d, path, mode)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    loss = torch.empty((), device=dev)
    scale_dev = torch.tensor([1 / 0.07], device=dev)
    red = torch.zeros((3, world, n, d), device=dev)
    dummy_red = other(0, torch.float32, (3, world, n, d))  # where the rows owned by the other ranks go
    gslots = torch.ones(world, device=dev)
    go = torch.ones(1, device=dev)
    dx = [torch.empty((n, d), dtype=dtype, device=dev) for _ in range(3)]
    dscale = torch.zeros(1, dtype=torch.float64, device=dev)
    loc = [f[:n].contiguous() for f in full]
    lab_loc = labels[:n].contiguous()
    R = range(world)
    peer_x = _lib.ptr_array([(gx[m] if q == 0 else dummy_x[m]).data_ptr() for q in R for m in range(3)])
    peer_inv = _lib.ptr_array([(ginv[m] if q == 0 else dummy_inv[m]).data_ptr() for q in R for m in range(3)])
    peer_lab = _lib.ptr_array([(glab if q == 0 else dummy_lab).data_ptr() for q in R])
    peer_stats = _lib.ptr_array([(stats if q == 0 else dummy_stats).data_ptr() for q in R])
    peer_col = _lib.ptr_array([colslots.data_ptr() for q in R])
    peer_pos = _lib.ptr_array([posslots.data_ptr() for q in R])
    peer_red = _lib.ptr_array([(red if q == 0 else dummy_red)[p].data_ptr() for q in R for p in range(3)])
    peer_g = _lib.ptr_array([gslots.data_ptr() for q in R])
    xs = _lib.ptr_array3([t.data_ptr() for t in gx])
    ivs = _lib.ptr_array3([t.data_ptr() for t in ginv])
    st = stats.data_ptr()
    first, count = _column_slots(weights, world)
    reduced = _lib.ptr_array3([None if f is None else red[f].data_ptr() for f in first])
    timings = {}

    def call(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(fn())
        e1.record()
        timings.setdefault(name, []).append((e0, e1))

    def step():
        call("push_rows", lambda: lib.clibd_shard_push_rows(_lib.ptr_array3([t.data_ptr() for t in loc]), dt, lab_loc.data_ptr(),
                                                            n, d, rank, world, peer_x, peer_inv, peer_lab, stream))
        call("forward_stats", lambda: lib.clibd_loss_forward_stats(xs, dt, ivs, glab.data_ptr(), N, d, 0, n, 0.0,
                                                                   scale_dev.data_ptr(), w, path, mode, scratch.data_ptr(),
                                                                   nbytes, st, st + 12 * N, st + 24 * N, pos_local.data_ptr(),
                                                                   stream))
        call("push_stats", lambda: lib.clibd_shard_push_stats(st, pos_local.data_ptr(), N, 0, n, rank, world, peer_stats,
                                                              peer_col, peer_pos, stream))
        call("reduce_stats", lambda: lib.clibd_shard_reduce_stats(colslots.data_ptr(), posslots.data_ptr(), N, world, st,
                                                                  pos.data_ptr(), stream))
        call("forward_finish", lambda: lib.clibd_loss_forward_finish(N, n, d, 0.0, w, path, mode, scratch.data_ptr(), nbytes,
                                                                     st, st + 12 * N, pos.data_ptr(), loss.data_ptr(), stream))
        call("push_gradout", lambda: lib.clibd_shard_push_floats(go.data_ptr(), 1, rank, world, peer_g, stream))
        call("backward_sweeps", lambda: lib.clibd_loss_backward_sweeps(xs, dt, ivs, N, d, 0, n, 0.0, w, path,
                                                                       scratch.data_ptr(), nbytes, st + 24 * N, None,
                                                                       peer_red, rank, world, stream))
        call("backward_finish", lambda: lib.clibd_loss_backward_finish(xs, dt, ivs, N, d, 0, n, 0.0, w, path,
                                                                       scratch.data_ptr(), nbytes, reduced,
                                                                       _lib.int_array(count), 1.0, gslots.data_ptr(), world,
                                                                       _lib.ptr_array3([t.data_ptr() for t in dx]),
                                                                       dscale.data_ptr(), stream))

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    timings.clear()
    lib.clibd_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    lib.clibd_profile_enable(0)
    pm, pn = (ctypes.c_double * 8)(), (ctypes.c_int64 * 8)()
    lib.clibd_profile_read(pm, pn)
    per_call = {k: round(sum(a.elapsed_time(b) for a, b in v) / steps, 4) for k, v in timings.items()}
    tensor = {"fwd_pair": round(pm[0] / steps, 4), "bwd_pair": round(pm[1] / steps, 4), "grad_gemm": round(pm[4] / steps, 4)}
    total = e0.elapsed_time(e1) / steps
    print(f"SIMRANK[{me}/{nproc}] " + json.dumps({"N": N, "world": world, "n": n, "step_ms": round(total, 4), "per_call_ms": per_call,
                                   "tensor_ms": tensor, "fixed_ms": round(total - sum(tensor.values()), 4),
                                   "loss": float(loss)}), flush=True)
    if nproc > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
