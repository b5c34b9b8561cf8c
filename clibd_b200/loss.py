"""Drop-in replacements for bioscanclip/model/loss_func.py, backed by the sm_100a CUDA library.

Same names, constructor arguments, forward signatures and error behaviour as the reference:

  * ``construct_label_metrix``   loss_func.py:19-22
  * ``ContrastiveLoss``          loss_func.py:25-69
  * ``gather_features``          loss_func.py:73-106
  * ``ClipLoss``                 loss_func.py:110-201

The N x N logit / target matrices of the reference are never materialised: the fused
kernels (csrc/loss_tc.cu, csrc/loss_simt.cu) produce row/column log-sum-exp statistics in
their epilogue and the label-matched positive term is an O(N d) class sum.  Multi-GPU:
each rank owns its row block; features, inverse norms and labels are all-gathered
(torch.distributed / NCCL), per-rank statistics are all-reduced, and every rank produces
the gradient of ITS rows directly, scaled by sum_r grad_out_r -- the reduce-scatter(SUM)
convention of torch.distributed.nn.all_gather's backward (loss_func.py:97).

There is no CPU or eager-PyTorch fallback: non-CUDA inputs raise.
"""
from __future__ import annotations

import contextlib
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib, _peer

try:  # same optional import dance as loss_func.py:5-11
    import torch.distributed.nn  # noqa: F401
    from torch import distributed as dist

    has_distributed = True
except ImportError:  # pragma: no cover
    dist = None
    has_distributed = False

_MODALITY_INDEX = {"image": 0, "dna": 1, "text": 2}  # loss_func.py:166-173
_PAIR_SLOT = {(0, 1): 0, (0, 2): 1, (1, 2): 2}
_DT = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16, torch.float16: _lib.DT_F16}


def construct_label_metrix(labels):
    """loss_func.py:19-22 (kept for API parity; the fused loss never builds this matrix)."""
    return (labels.unsqueeze(0) == labels.unsqueeze(1)).float()


def pair_weights(present, bind_to=None, no_image_text_loss=False):
    """Weights of the unordered slot pairs (image,dna), (image,text), (dna,text) that make
    sum_p w_p [CE(S_p,T) + CE(S_p^T,T)] equal the reference's mean over its loss list.

    The reference loops over ordered pairs of the None-FILTERED feature list and applies the
    bind_to / no_image_text indices to that filtered list (loss_func.py:159-184); each visit
    appends both directions (loss_func.py:195-198)."""
    slots = [i for i, p in enumerate(present) if p]
    bind_to_idx = _MODALITY_INDEX.get(bind_to) if bind_to is not None else None
    visits = {}
    n_ordered = 0
    for a in range(len(slots)):
        for b in range(len(slots)):
            if bind_to_idx is not None and a != bind_to_idx and b != bind_to_idx:
                continue
            if a == b:
                continue
            if no_image_text_loss and (a == 0 or b == 0) and (a == 2 or b == 2):
                continue
            n_ordered += 1
            key = (min(slots[a], slots[b]), max(slots[a], slots[b]))
            visits[key] = visits.get(key, 0) + 1
    w = [0.0, 0.0, 0.0]
    for key, m in visits.items():
        w[_PAIR_SLOT[key]] = m / (2.0 * n_ordered)
    return w, n_ordered


_EXACT_PATH_MAX_BATCH = 1024


def _select_path(dtype, operands, n_global, d):
    """Which arithmetic the O(N^2 d) kernels use.  An explicit `tensor_core_operands` wins.  By default 16-bit inputs
    take the tensor cores with their own operand type, and float32 inputs take
      * the exact CUDA-core path up to a global batch of 1024 (BASELINE config 1: N = 256, parity 1e-5), and
      * the tensor cores with FLOAT16 operands beyond (gradients within ~5e-5 of the float64 oracle, fp32 accumulation,
        fp32 positive term and statistics): the CUDA-core sweep runs at ~10 TFLOP/s, so at the reference's own training
        batch (500 per GPU x 8 = 4000 float32 embeddings out of its autocast block) the exact path would be slower than
        the reference's cuBLAS calls.  tensor_core_operands="fp32" forces the exact path at any size."""
    if operands is not None:
        return {"fp32": _lib.PATH_SIMT_F32, "bf16": _lib.PATH_TC_BF16, "fp16": _lib.PATH_TC_F16}[operands]
    if dtype == torch.float32:
        return _lib.PATH_SIMT_F32 if n_global <= _EXACT_PATH_MAX_BATCH else _lib.PATH_TC_F16
    return {torch.bfloat16: _lib.PATH_TC_BF16, torch.float16: _lib.PATH_TC_F16}[dtype]


def _stream_ptr(device):
    if device.type != "cuda":  # only reachable with a test double injected through _lib.inject_for_tests
        return ctypes.c_void_p(None)
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _device_ctx(device):
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


def _gather_rows(local, labels, gathered, all_labels, group):
    """All-gather of the labels (their own dtype) and of every present feature set.  On NCCL the feature gathers are
    batched into one grouped launch through torch's coalescing manager when that (private) API is there and can be entered;
    otherwise they run as plain sequential collectives -- same result."""
    dist.all_gather_into_tensor(all_labels, labels, group=group)  # int64, on its own: a grouped call wants one dtype

    def issue():
        for f, g in zip(local, gathered):
            if f is not None:
                dist.all_gather_into_tensor(g, f, group=group)

    manager = None
    try:
        if dist.get_backend(group) == "nccl" and hasattr(dist, "_coalescing_manager"):
            manager = dist._coalescing_manager(group=group)
            manager.__enter__()
    except Exception:  # noqa: BLE001
        manager = None
    if manager is None:
        issue()
        return
    try:
        issue()
    except BaseException as ex:
        manager.__exit__(type(ex), ex, ex.__traceback__)
        raise
    manager.__exit__(None, None, None)


_MAX_EXCHANGE_RANKS = 16  # MAX_PEERS of the library (one NVLink domain); larger jobs use the 'local' form


def _shard_mode(path, d, group, device, world=2):
    """How a row-sharded step exchanges data:
    'peer'  : S once per pair; all-gather, statistics and the reduce-scatter of the column-side gradients are stores
              into peer-mapped symmetric memory over NVLink (csrc/shard_exchange.cu, loss_grad_gemm.cu);
    'nccl'  : the same step with NCCL collectives (all-gather in, all-reduce of statistics, reduce-scatter out);
    'local' : every rank recomputes S for both of its gradients (two sweeps per pair), NCCL in / all-reduce only.
    CLIBD_SHARD_MODE forces one of them; the default is 'peer' where symmetric memory is available."""
    exchange_ok = path != _lib.PATH_SIMT_F32 and (d + 63) // 64 * 64 <= 768 and world <= _MAX_EXCHANGE_RANKS
    want = os.environ.get("CLIBD_SHARD_MODE", "")
    if want == "local" or not exchange_ok:
        return "local"
    if want == "nccl" or device.type != "cuda":
        return "nccl"
    try:
        if dist.get_backend(group) != "nccl":
            return "nccl"
    except Exception:  # noqa: BLE001
        return "nccl"
    if _peer.available() or want == "peer":
        return "peer"
    return "nccl"


def _column_slots(weights, world):
    """(pair whose slot array receives them, slot count) of the column-side gradient partials per modality in the peer
    form.  (image,dna) feeds dna; (image,text) and (dna,text) feed text through ONE gradient GEMM (the library merges
    pairs that share their column modality and weight), so text also receives `world` slots, in the slot array of its
    first weighted pair."""
    first = [None, None, None]
    count = [0, 0, 0]
    for p, b in enumerate((1, 2, 2)):
        if weights[p] != 0.0 and first[b] is None:
            first[b] = p
            count[b] = world
    return first, count


class _FusedClipLossFn(torch.autograd.Function):
    """forward(image, dna, text, labels, scale_tensor|None, scale_value, weights, path, group, world, rank,
    sum_grad_over_ranks) -> 0-d float32 loss"""

    @staticmethod
    def forward(ctx, image, dna, text, labels, scale_tensor, scale_value, weights, path, group, world, rank,
                sum_grads):
        lib = _lib.load()
        feats = [image, dna, text]
        ref = next(f for f in feats if f is not None)
        device, dtype = ref.device, ref.dtype
        n, d = ref.shape
        stream = _stream_ptr(device)
        shard = _shard_mode(path, d, group, device, world) if world > 1 else "single"
        px = entry = None
        if shard == "peer":
            try:
                px = _peer.context(group, device, n * world, n, d, ref.dtype, world, rank)
                entry = px.acquire()
            except Exception as ex:  # noqa: BLE001  (no peer access / no way to pass the handles on this box)
                if os.environ.get("CLIBD_SHARD_MODE", "") == "peer":
                    raise
                _peer.disable(repr(ex))
                shard, px, entry = "nccl", None, None
        mode = _lib.MODE_EXCHANGE if shard in ("peer", "nccl") else _lib.MODE_LOCAL
        with _device_ctx(device):
            local = [None if f is None else f.detach().contiguous() for f in feats]
            labels = labels.detach().to(device=device, dtype=torch.int64).contiguous()

            def inv_norms(t):
                iv = torch.empty(t.shape[0], dtype=torch.float32, device=device)
                _lib.check(lib.clibd_row_inv_norm(t.data_ptr(), _DT[dtype], t.shape[0], d, iv.data_ptr(), stream))
                return iv

            N = n * world
            row0 = rank * n
            w = _lib.float_array3(weights)
            # a tensor scale stays on the device: the library copies it into the scratch and every kernel of the
            # forward and the backward reads it from there (no host read in the training step)
            scale_dev = None
            if scale_tensor is not None:
                scale_dev = scale_tensor.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()
            scale_ptr = None if scale_dev is None else scale_dev.data_ptr()
            nbytes = lib.clibd_loss_scratch_bytes(N, n, d, path, mode)
            if nbytes < 0:
                raise ValueError("clibd_b200: bad loss shape")
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
            loss = torch.empty((), dtype=torch.float32, device=device)
            if shard == "peer":
                # ---- exchanges are stores into peer-mapped memory; a barrier separates writers from readers
                local_ptrs = _lib.ptr_array3([None if f is None else f.data_ptr() for f in local])
                labels_arg = entry.labels_ptr
                if os.environ.get("CLIBD_OVERLAP_PUSH", "0") != "0":
                    # Optional (off by default): the labels (a few KB) go first; their statistics -- hash, class sort,
                    # ranges: ten small launches -- then run on this stream while the feature rows travel on a second one
                    # (NVLink-bound), each side with its own barrier across the ranks.  Measured at 8 GPUs, N = 32768:
                    # 2.694 ms per step with it, 2.681 ms without (profiles/r2n_phase_timing_n8.log) -- the second
                    # barrier and the stream joins cost what the overlap saves, and small batches lose (N = 4096:
                    # 0.74 vs 0.64 ms), so the plain sequence below is the default.
                    main = torch.cuda.current_stream(device)
                    side = px.side_stream()
                    side.wait_stream(main)  # the inputs are ready
                    with torch.cuda.stream(side):
                        _lib.check(lib.clibd_shard_push_rows(local_ptrs, _DT[dtype], None, n, d, rank, world, entry.peer_x,
                                                             entry.peer_inv, entry.peer_labels,
                                                             ctypes.c_void_p(side.cuda_stream)))
                        px.barrier(channel=1)
                    _lib.check(lib.clibd_shard_push_rows(_lib.ptr_array3([None, None, None]), _DT[dtype], labels.data_ptr(),
                                                         n, d, rank, world, entry.peer_x, entry.peer_inv,
                                                         entry.peer_labels, stream))
                    px.barrier(channel=0)
                    _lib.check(lib.clibd_loss_label_stage(entry.labels_ptr, N, n, d, path, mode, scratch.data_ptr(), nbytes,
                                                          stream))
                    main.wait_stream(side)
                    labels_arg = None  # done
                else:
                    _lib.check(lib.clibd_shard_push_rows(local_ptrs, _DT[dtype], labels.data_ptr(), n, d, rank, world,
                                                         entry.peer_x, entry.peer_inv, entry.peer_labels, stream))
                    px.barrier()
                gathered_ptrs = [None if f is None else entry.x_ptr[m] for m, f in enumerate(local)]
                inv_ptrs = [None if f is None else entry.inv_ptr[m] for m, f in enumerate(local)]
                xs, ivs = _lib.ptr_array3(gathered_ptrs), _lib.ptr_array3(inv_ptrs)
                st = entry.stats_ptr
                _lib.check(lib.clibd_loss_forward_stats(xs, _DT[dtype], ivs, labels_arg, N, d, row0, n, scale_value,
                                                        scale_ptr, w, path, mode, scratch.data_ptr(), nbytes, st,
                                                        st + 12 * N, st + 24 * N, entry.pos_local_ptr, stream))
                _lib.check(lib.clibd_shard_push_stats(st, entry.pos_local_ptr, N, row0, n, rank, world, entry.peer_stats,
                                                      entry.peer_colslots, entry.peer_posslots, stream))
                px.barrier()
                _lib.check(lib.clibd_shard_reduce_stats(entry.colslots_ptr, entry.posslots_ptr, N, world, st,
                                                        entry.pos_ptr, stream))
                _lib.check(lib.clibd_loss_forward_finish(N, n, d, scale_value, w, path, mode, scratch.data_ptr(), nbytes,
                                                         st, st + 12 * N, entry.pos_ptr, loss.data_ptr(), stream))
                ctx.keep = (_peer.EntryHolder(px, entry),)
                ctx.x_ptrs, ctx.inv_ptrs, ctx.posrow_ptr = gathered_ptrs, inv_ptrs, st + 24 * N
            else:
                if world > 1:
                    all_labels = torch.empty(N, dtype=torch.int64, device=device)
                    gathered = [None if f is None else torch.empty((N, d), dtype=dtype, device=device) for f in local]
                    # labels + up to three feature sets in one grouped all-gather where NCCL coalescing is available
                    _gather_rows(local, labels, gathered, all_labels, group)
                    # inverse norms of all rows from the gathered features: one 50 MB read per modality instead of
                    # one more latency-bound collective each
                    inv = [None if g is None else inv_norms(g) for g in gathered]
                else:
                    gathered, all_labels = local, labels
                    inv = [None if f is None else inv_norms(f) for f in local]
                # rowsum[3][N] | colsum[3][N] | posrow[3][N] (the last block: exchange mode only)
                stats = torch.zeros(9 * N, dtype=torch.float32, device=device)
                pos = torch.zeros(3, dtype=torch.float64, device=device)
                gathered_ptrs = [None if g is None else g.data_ptr() for g in gathered]
                inv_ptrs = [None if g is None else g.data_ptr() for g in inv]
                xs, ivs = _lib.ptr_array3(gathered_ptrs), _lib.ptr_array3(inv_ptrs)
                st = stats.data_ptr()
                _lib.check(lib.clibd_loss_forward_stats(xs, _DT[dtype], ivs, all_labels.data_ptr(), N, d, row0, n,
                                                        scale_value, scale_ptr, w, path, mode, scratch.data_ptr(), nbytes,
                                                        st, st + 12 * N, st + 24 * N, pos.data_ptr(), stream))
                if world > 1:
                    dist.all_reduce(stats, group=group)  # row sums / posrow: disjoint support, column sums add up
                    dist.all_reduce(pos, group=group)
                _lib.check(lib.clibd_loss_forward_finish(N, n, d, scale_value, w, path, mode, scratch.data_ptr(), nbytes,
                                                         st, st + 12 * N, pos.data_ptr(), loss.data_ptr(), stream))
                ctx.keep = (gathered, inv, stats)
                ctx.x_ptrs, ctx.inv_ptrs, ctx.posrow_ptr = gathered_ptrs, inv_ptrs, st + 24 * N
                if world == 1:
                    # the backward re-reads the inputs: saving them lets autograd's version counters catch an in-place
                    # modification between forward and backward (the kernels read through the pointers kept above)
                    ctx.save_for_backward(*[f for f in feats if f is not None])
        ctx.scratch = scratch
        ctx.meta = (N, n, d, row0, scale_value, tuple(weights), path, group, world, rank, dtype, device, sum_grads, shard)
        ctx.has_scale = scale_tensor is not None
        ctx.scale_dtype = scale_tensor.dtype if scale_tensor is not None else None
        ctx.present = [f is not None for f in feats]
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        N, n, d, row0, scale_value, weights, path, group, world, rank, dtype, device, sum_grads, shard = ctx.meta
        _ = ctx.saved_tensors  # raises if a saved input was modified in place since the forward
        stream = _stream_ptr(device)
        with _device_ctx(device):
            grad_out = grad_out.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()
            # a modality that is present but in no weighted pair (bind_to + no_image_text_loss can leave one out) gets no
            # gradient, as in the reference, where it never enters the graph
            used = (weights[0] != 0 or weights[1] != 0, weights[0] != 0 or weights[2] != 0,
                    weights[1] != 0 or weights[2] != 0)
            dxs = [torch.empty((n, d), dtype=dtype, device=device) if (p and used[i] and ctx.needs_input_grad[i]) else None
                   for i, p in enumerate(ctx.present)]
            dscale = torch.zeros(1, dtype=torch.float64, device=device)
            xs, ivs = _lib.ptr_array3(ctx.x_ptrs), _lib.ptr_array3(ctx.inv_ptrs)
            outs = _lib.ptr_array3([None if g is None else g.data_ptr() for g in dxs])
            w = _lib.float_array3(weights)
            sp, sn = ctx.scratch.data_ptr(), ctx.scratch.numel()
            want_dscale = ctx.has_scale and ctx.needs_input_grad[4]  # else nobody reads it: no exchange over the ranks
            # grad_output stays on the device: the library multiplies it into the feature gradients.  Backward of the
            # differentiable all-gather = reduce-scatter(SUM) over ranks (loss_func.py:97): every rank's loss is the same
            # function, so the local rows receive (sum_r grad_output_r) * dL/dx.
            if shard == "peer":
                holder = ctx.keep[0]
                px, entry = holder.px, holder.entry
                par = px.next_parity()
                if sum_grads:
                    _lib.check(lib.clibd_shard_push_floats(grad_out.data_ptr(), 1, rank, world, px.peer_gslots[par], stream))
                    g_ptr, g_count = px.gslots_ptr[par], world
                else:
                    g_ptr, g_count = grad_out.data_ptr(), 1
                _lib.check(lib.clibd_loss_backward_sweeps(xs, _DT[dtype], ivs, N, d, row0, n, scale_value, w, path, sp, sn,
                                                          ctx.posrow_ptr, None, px.peer_red[par], rank, world, stream))
                px.barrier()
                first, count = _column_slots(weights, world)
                reduced = _lib.ptr_array3([None if f is None else px.red_ptr[par] + f * px.red_pair_bytes for f in first])
                _lib.check(lib.clibd_loss_backward_finish(xs, _DT[dtype], ivs, N, d, row0, n, scale_value, w, path, sp, sn,
                                                          reduced, _lib.int_array(count), 1.0, g_ptr, g_count, outs,
                                                          dscale.data_ptr(), stream))
                if want_dscale:  # dL/ds needs all rows: every rank's share into per-rank slots, summed in rank order
                    _lib.check(lib.clibd_shard_push_floats(dscale.data_ptr(), 2, rank, world, px.peer_dslots[par], stream))
                    px.barrier()
                    dscale = px.dslots[par].sum().reshape(1)
                holder.release()
            elif shard == "nccl":
                gsum = grad_out
                if sum_grads:
                    gsum = grad_out.clone()
                    dist.all_reduce(gsum, group=group)
                first, _ = _column_slots(weights, world)
                part = [None if f is None else torch.empty((N, d), dtype=torch.float32, device=device) for f in first]
                _lib.check(lib.clibd_loss_backward_sweeps(xs, _DT[dtype], ivs, N, d, row0, n, scale_value, w, path, sp, sn,
                                                          ctx.posrow_ptr,
                                                          _lib.ptr_array3([None if t is None else t.data_ptr() for t in part]),
                                                          None, rank, world, stream))
                reduced = [None if t is None else _reduce_scatter_rows(t, n, rank, group) for t in part]
                _lib.check(lib.clibd_loss_backward_finish(xs, _DT[dtype], ivs, N, d, row0, n, scale_value, w, path, sp, sn,
                                                          _lib.ptr_array3([None if t is None else t.data_ptr() for t in reduced]),
                                                          _lib.int_array([0 if t is None else 1 for t in reduced]), 1.0,
                                                          gsum.data_ptr(), 1, outs, dscale.data_ptr(), stream))
                if want_dscale:
                    dist.all_reduce(dscale, group=group)
            else:
                gsum = grad_out
                if world > 1 and sum_grads:
                    gsum = grad_out.clone()
                    dist.all_reduce(gsum, group=group)
                _lib.check(lib.clibd_loss_backward(xs, _DT[dtype], ivs, N, d, row0, n, scale_value, w, path, sp, sn, 1.0,
                                                   gsum.data_ptr(), outs, dscale.data_ptr(), stream))
                if world > 1 and want_dscale:
                    dist.all_reduce(dscale, group=group)  # dL/ds needs all rows
            grads = dxs
            gscale = None
            if want_dscale:
                gscale = (dscale * grad_out).to(ctx.scale_dtype).reshape(())  # float64 x float32 -> float64 product, one kernel
        return grads[0], grads[1], grads[2], None, gscale, None, None, None, None, None, None, None


def _reduce_scatter_rows(part, n, rank, group):
    """reduce-scatter(SUM) of a [N, d] float32 partial over the ranks -> this rank's [n, d] block (the backward of
    torch.distributed.nn.all_gather, loss_func.py:97)."""
    out = torch.empty((n, part.shape[1]), dtype=part.dtype, device=part.device)
    try:
        nccl = dist.get_backend(group) == "nccl"
    except Exception:  # noqa: BLE001
        nccl = False
    if nccl:
        dist.reduce_scatter_tensor(out, part, group=group)
    else:  # gloo (CPU tests): no reduce-scatter
        dist.all_reduce(part, group=group)
        out.copy_(part[rank * n:(rank + 1) * n])
    return out


def _validate_criterion(criterion):
    if criterion is None:
        return
    ok = isinstance(criterion, nn.CrossEntropyLoss) and criterion.reduction == "mean" \
        and getattr(criterion, "label_smoothing", 0.0) == 0.0 and criterion.weight is None
    if not ok:
        raise NotImplementedError(
            "clibd_b200 fuses nn.CrossEntropyLoss() with default arguments (the only criterion the reference "
            "constructs, train_cl.py:260,262); other criteria are not supported")


def _fused_loss(image_features, dna_features, text_features, labels, logit_scale, bind_to, no_image_text_loss,
                operands, group, world, rank, sum_grads):
    feats = [image_features, dna_features, text_features]
    present = [f is not None for f in feats]
    if sum(present) < 2:
        raise ValueError("Too less element for calculating the contrastive loss.")  # loss_func.py:46-47,162-163
    ref = next(f for f in feats if f is not None)
    if not ref.is_cuda and not _lib.test_double_active():
        raise RuntimeError("clibd_b200 runs on CUDA tensors only (there is no CPU fallback)")
    shapes = {tuple(f.shape) for f in feats if f is not None}
    if len(shapes) != 1 or ref.dim() != 2:
        raise ValueError("all feature tensors must share one [batch, dim] shape")
    if ref.shape[0] == 0:
        raise ValueError("empty batch")
    dtypes = {f.dtype for f in feats if f is not None}
    if len(dtypes) != 1 or ref.dtype not in _DT:
        common = torch.float32
        feats = [None if f is None else f.to(common) for f in feats]
    else:
        common = ref.dtype
    weights, n_ordered = pair_weights(present, bind_to, no_image_text_loss)
    if n_ordered == 0:
        raise ZeroDivisionError("float division by zero")  # reference: sum([]) * 1.0 / len([])
    scale_tensor = logit_scale if isinstance(logit_scale, torch.Tensor) else None
    # a tensor scale is handed over as a device pointer (no host read); scale_value is then unused
    scale_value = 0.0 if scale_tensor is not None else float(logit_scale)
    path = _select_path(common, operands, ref.shape[0] * max(1, world), ref.shape[1])
    return _FusedClipLossFn.apply(feats[0], feats[1], feats[2], labels, scale_tensor, scale_value, weights, path,
                                  group, world, rank, sum_grads)


class ContrastiveLoss(nn.Module):
    """Single-process contrastive loss, reference loss_func.py:25-69."""

    def __init__(self, criterion, logit_scale, local_loss=False, gather_with_grad=False, rank=0, world_size=1,
                 use_horovod=False, tensor_core_operands=None):
        super().__init__()
        _validate_criterion(criterion)
        self.criterion = criterion
        self.logit_scale = logit_scale
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self.prev_num_logits = 0
        self.labels = {}
        # None: fp32 inputs -> exact CUDA-core path, bf16/fp16 inputs -> tcgen05 with that operand type;
        # "bf16" / "fp16" force the tensor-core path, "fp32" forces the CUDA-core path.
        self.tensor_core_operands = tensor_core_operands

    def forward(self, image_features, dna_features, text_features, labels, logit_scale):
        scale = logit_scale if logit_scale is not None else self.logit_scale  # loss_func.py:58-63
        return _fused_loss(image_features, dna_features, text_features, labels, scale, None, False,
                           self.tensor_core_operands, None, 1, 0, True)


class _PeerAllGatherFn(torch.autograd.Function):
    """Differentiable all-gather of [n, d] rows as stores into peer-mapped memory (clibd_shard_push_rows); backward =
    reduce-scatter(SUM) of the gathered gradient, like torch.distributed.nn.all_gather's (loss_func.py:97)."""

    @staticmethod
    def forward(ctx, features, group, world, rank):
        lib = _lib.load()
        device, dtype = features.device, features.dtype
        n, d = features.shape
        gb = _peer.gather_buffers(group, device, n * world, d, dtype, world, rank)
        with _device_ctx(device):
            x = features.detach().contiguous()
            _lib.check(lib.clibd_shard_push_rows(_lib.ptr_array3([x.data_ptr(), None, None]), _DT[dtype], None, n, d, rank,
                                                 world, gb.peer_x, gb.peer_inv, gb.peer_labels, _stream_ptr(device)))
            gb.barrier()   # every rank's rows have landed in every buffer
            out = gb.rows.clone()
            gb.barrier()   # every rank has read its buffer: the next call may overwrite it
        ctx.meta = (n, rank, group)
        return out

    @staticmethod
    def backward(ctx, grad):
        n, rank, group = ctx.meta
        return _reduce_scatter_rows(grad.contiguous(), n, rank, group), None, None, None


def _peer_gather_ok(features, world):
    if world <= 1 or world > _MAX_EXCHANGE_RANKS or not isinstance(features, torch.Tensor) or not features.is_cuda \
            or features.dim() != 2:
        return False
    if features.dtype not in _DT or features.shape[0] == 0 or os.environ.get("CLIBD_SHARD_MODE", "") in ("nccl", "local"):
        return False
    try:
        return _peer.available() and dist.is_initialized() and dist.get_backend() == "nccl" \
            and dist.get_world_size() == world
    except Exception:  # noqa: BLE001
        return False


def gather_features(features, local_loss=False, gather_with_grad=False, rank=0, world_size=1, use_horovod=False):
    """loss_func.py:73-106 (public helper of the reference module; the fused ClipLoss gathers inside its own step).
    CUDA [n, d] features under NCCL travel as stores into peer-mapped memory over NVLink (this library's push kernel)
    where the ranks can map each other's buffers; otherwise the reference's collectives are used."""
    assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
    if use_horovod:
        raise NotImplementedError("horovod is not supported (no shipped reference config uses it)")
    if _peer_gather_ok(features, world_size):
        try:
            if gather_with_grad:
                return _PeerAllGatherFn.apply(features, None, world_size, rank)
            n = features.shape[0]
            all_features = _PeerAllGatherFn.apply(features.detach(), None, world_size, rank)
            if not local_loss:  # grads for the local rows when the gathered features carry none
                all_features = torch.cat([all_features[:rank * n], features, all_features[(rank + 1) * n:]], dim=0)
            return all_features
        except RuntimeError as ex:  # mapping the buffers failed on this box: the NCCL form below, from now on
            if os.environ.get("CLIBD_SHARD_MODE", "") == "peer":
                raise
            _peer.disable(repr(ex))
    if gather_with_grad:
        return torch.cat(torch.distributed.nn.all_gather(features), dim=0)
    gathered = [torch.zeros_like(features) for _ in range(world_size)]
    dist.all_gather(gathered, features)
    if not local_loss:
        gathered[rank] = features
    return torch.cat(gathered, dim=0)


class ClipLoss(nn.Module):
    """All-gather contrastive loss, reference loss_func.py:110-201."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False, criterion=None, bind_to=None, no_image_text_loss=False,
                 tensor_core_operands=None, process_group=None):
        super().__init__()
        _validate_criterion(criterion)
        if use_horovod:
            raise NotImplementedError("horovod is not supported (no shipped reference config uses it)")
        if world_size > 1 and local_loss and not gather_with_grad:
            raise NotImplementedError("local_loss=True without gather_with_grad detaches every gathered feature "
                                      "in the reference (loss_func.py:99-104); not supported")
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self.criterion = criterion if criterion is not None else nn.CrossEntropyLoss()
        self.prev_num_logits = 0
        self.labels = {}
        self.bind_to = bind_to
        self.no_image_text_loss = no_image_text_loss
        self.tensor_core_operands = tensor_core_operands
        self.process_group = process_group

    def forward(self, image_features, dna_features, text_features, labels, logit_scale, output_dict=False):
        if self.world_size > 1 and not (has_distributed and dist.is_initialized()):
            raise RuntimeError("ClipLoss(world_size>1) needs an initialised torch.distributed process group")
        # gather_with_grad=True: local grads are summed over every rank's loss (reduce-scatter);
        # gather_with_grad=False, local_loss=False: only this rank's loss reaches the local slot.
        total_loss = _fused_loss(image_features, dna_features, text_features, labels, logit_scale, self.bind_to,
                                 self.no_image_text_loss, self.tensor_core_operands, self.process_group,
                                 self.world_size, self.rank, bool(self.gather_with_grad))
        return {"contrastive_loss": total_loss} if output_dict else total_loss
