"""The row-sharded EXCHANGE step on ONE GPU: W simulated ranks run the C-ABI phases in lockstep.

A sharded job needs W GPUs, the round-end test box has one.  Every phase of the exchange step is an extern "C" call on
caller-owned buffers, so W "ranks" can live on one device: each gets its own gathered buffers, statistics, scratch and
slot arrays, the phases run rank by rank, and the place of a barrier across ranks is simply the end of a loop.  This
covers the push kernels (csrc/shard_exchange.cu), the exchange-mode forward, the row sweep with its local strip, the
gradient GEMM's per-owner epilogue stores (both the reduce-scatter form and the peer-slot form, whose "peer" pointers
here point into the same device) and the split backward -- against the float64 oracle and the single-rank result.
The real multi-GPU run of the same code (NCCL / NVLink) is tools/multigpu_check.py (tests/test_multigpu.py).
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo

pytestmark = pytest.mark.gpu

PAIR_B = (1, 2, 2)


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _sharded_step(feats, labels, scale, world, path, form, weights=(1 / 6, 1 / 6, 1 / 6), grad_outs=None,
                  labels_first=False):
    """feats: list of 3 [N, d] CUDA tensors (or None); returns loss per rank, grads [3][N, d] float32, dscale."""
    from clibd_b200 import _lib
    from clibd_b200.loss import _DT, _column_slots
    lib = _lib.load()
    dev = labels.device
    ref = next(f for f in feats if f is not None)
    N, d = ref.shape
    dtype = ref.dtype
    n = N // world
    assert n * world == N
    dt = _DT[dtype]
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    w = _lib.float_array3(weights)
    grad_outs = grad_outs or [1.0] * world
    mode = _lib.MODE_EXCHANGE
    nbytes = lib.clibd_loss_scratch_bytes(N, n, d, path, mode)
    R = range(world)
    # per-rank "symmetric" buffers
    gx = [[None if f is None else torch.full_like(f, float("nan")) for f in feats] for _ in R]
    ginv = [[None if f is None else torch.full((N,), float("nan"), device=dev) for f in feats] for _ in R]
    glab = [torch.full((N,), -1, dtype=torch.int64, device=dev) for _ in R]
    stats = [torch.full((9 * N,), float("nan"), device=dev) for _ in R]
    colslots = [torch.full((world * 3 * N,), float("nan"), device=dev) for _ in R]
    posslots = [torch.zeros(world * 4, dtype=torch.float64, device=dev) for _ in R]
    pos_local = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in R]
    pos = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in R]
    scratch = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in R]
    loss = [torch.empty((), device=dev) for _ in R]
    gslots = [torch.zeros(world, device=dev) for _ in R]

    def P(t):
        return None if t is None else t.data_ptr()

    # ---- all-gather by pushes (labels_first: the labels travel in their own push, as the overlapped step sends them)
    for r in R:
        loc = [None if f is None else f[r * n:(r + 1) * n].contiguous() for f in feats]
        lab = labels[r * n:(r + 1) * n].contiguous()
        tables = (_lib.ptr_array([P(gx[q][m]) for q in R for m in range(3)]),
                  _lib.ptr_array([P(ginv[q][m]) for q in R for m in range(3)]),
                  _lib.ptr_array([glab[q].data_ptr() for q in R]))
        if labels_first:
            _lib.check(lib.clibd_shard_push_rows(_lib.ptr_array3([None, None, None]), dt, lab.data_ptr(), n, d, r, world,
                                                 *tables, stream))
            _lib.check(lib.clibd_shard_push_rows(_lib.ptr_array3([P(t) for t in loc]), dt, None, n, d, r, world, *tables,
                                                 stream))
        else:
            _lib.check(lib.clibd_shard_push_rows(_lib.ptr_array3([P(t) for t in loc]), dt, lab.data_ptr(), n, d, r, world,
                                                 *tables, stream))
    torch.cuda.synchronize()
    for q in R:  # every rank now holds the whole batch, bit-identical
        assert torch.equal(glab[q], labels)
        for m in range(3):
            if feats[m] is not None:
                assert torch.equal(gx[q][m], feats[m])
                assert torch.equal(ginv[q][m], ginv[0][m]) and bool(torch.isfinite(ginv[q][m]).all())
    # ---- forward statistics of every rank, exchange, finish
    scale_dev = torch.tensor([scale], device=dev)
    for r in R:
        st = stats[r].data_ptr()
        if labels_first:  # label statistics in their own call, then forward_stats with labels == NULL
            _lib.check(lib.clibd_loss_label_stage(glab[r].data_ptr(), N, n, d, path, mode, scratch[r].data_ptr(), nbytes,
                                                  stream))
        _lib.check(lib.clibd_loss_forward_stats(
            _lib.ptr_array3([P(t) for t in gx[r]]), dt, _lib.ptr_array3([P(t) for t in ginv[r]]),
            None if labels_first else glab[r].data_ptr(), N, d,
            r * n, n, 0.0, scale_dev.data_ptr(), w, path, mode, scratch[r].data_ptr(), nbytes, st, st + 12 * N,
            st + 24 * N, pos_local[r].data_ptr(), stream))
    for r in R:
        _lib.check(lib.clibd_shard_push_stats(
            stats[r].data_ptr(), pos_local[r].data_ptr(), N, r * n, n, r, world,
            _lib.ptr_array([stats[q].data_ptr() for q in R]), _lib.ptr_array([colslots[q].data_ptr() for q in R]),
            _lib.ptr_array([posslots[q].data_ptr() for q in R]), stream))
    for r in R:
        st = stats[r].data_ptr()
        _lib.check(lib.clibd_shard_reduce_stats(colslots[r].data_ptr(), posslots[r].data_ptr(), N, world, st,
                                                pos[r].data_ptr(), stream))
        _lib.check(lib.clibd_loss_forward_finish(N, n, d, 0.0, w, path, mode, scratch[r].data_ptr(), nbytes, st,
                                                 st + 12 * N, pos[r].data_ptr(), loss[r].data_ptr(), stream))
    torch.cuda.synchronize()
    used = [p for p in range(3) if weights[p] != 0.0]
    for q in R[1:]:  # the exchanged statistics agree bit for bit on every rank
        for p in used:
            for blk in (0, 3, 6):
                a0 = stats[0][(blk + p) * N:(blk + p + 1) * N]
                assert torch.equal(stats[q][(blk + p) * N:(blk + p + 1) * N], a0)
    # ---- backward: sweeps of every rank, then (barrier) the finish of every rank
    first, count = _column_slots(weights, world)
    go = [torch.tensor([g], device=dev) for g in grad_outs]
    for r in R:
        _lib.check(lib.clibd_shard_push_floats(go[r].data_ptr(), 1, r, world,
                                               _lib.ptr_array([gslots[q].data_ptr() for q in R]), stream))
    if form == "peer":
        red = [torch.full((3, world, n, d), float("nan"), device=dev) for _ in R]
        peer_red = _lib.ptr_array([red[q][p].data_ptr() for q in R for p in range(3)])
    else:
        part = [[None if f is None else torch.full((N, d), float("nan"), device=dev) for f in first] for _ in R]
    for r in R:
        st = stats[r].data_ptr()
        _lib.check(lib.clibd_loss_backward_sweeps(
            _lib.ptr_array3([P(t) for t in gx[r]]), dt, _lib.ptr_array3([P(t) for t in ginv[r]]), N, d, r * n, n, 0.0, w,
            path, scratch[r].data_ptr(), nbytes, st + 24 * N,
            None if form == "peer" else _lib.ptr_array3([P(t) for t in part[r]]),
            peer_red if form == "peer" else None, r, world, stream))
    torch.cuda.synchronize()
    grads = [None if f is None else torch.zeros((N, d), device=dev) for f in feats]
    dscale = torch.zeros(world, dtype=torch.float64, device=dev)
    for r in R:
        if form == "peer":
            reduced = [None if f is None else red[r][f] for f in first]
            slots = count
        else:  # what NCCL's reduce-scatter(SUM) would deliver to rank r
            reduced = [None if f is None else sum(part[q][m][r * n:(r + 1) * n] for q in R).contiguous()
                       for m, f in enumerate(first)]
            slots = [0 if f is None else 1 for f in first]
        dx = [None if f is None else torch.zeros((n, d), dtype=dtype, device=dev) for f in feats]  # no pair: stays 0
        _lib.check(lib.clibd_loss_backward_finish(
            _lib.ptr_array3([P(t) for t in gx[r]]), dt, _lib.ptr_array3([P(t) for t in ginv[r]]), N, d, r * n, n, 0.0, w,
            path, scratch[r].data_ptr(), nbytes, _lib.ptr_array3([P(t) for t in reduced]), _lib.int_array(slots), 1.0,
            gslots[r].data_ptr(), world, _lib.ptr_array3([P(t) for t in dx]), dscale[r:r + 1].data_ptr(), stream))
        for m in range(3):
            if dx[m] is not None:
                grads[m][r * n:(r + 1) * n] = dx[m].float()
    torch.cuda.synchronize()
    return ([float(x) for x in loss], [None if g is None else g.cpu().numpy() for g in grads],
            float(dscale.sum()))


@pytest.mark.parametrize("form", ["reduce_scatter", "peer"])
@pytest.mark.parametrize("N,d,nmod,world,operands,labels_kind", [
    (1024, 768, 3, 4, "bf16", "multi"),   # BASELINE config-2 shape, 4 ranks of 256 rows
    (1536, 768, 3, 2, "fp16", "multi"),   # two ranks of 768 rows (6 row tiles each)
    (768, 200, 2, 3, "bf16", "onehot"),   # three ranks, d not a multiple of 64, image+dna only
    (520, 768, 3, 2, "bf16", "zipf"),     # n = 260 per rank: ragged row tiles, long-tailed classes
])
def test_simulated_ranks_match_oracle(form, N, d, nmod, world, operands, labels_kind):
    from clibd_b200 import _lib
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(21)
    feats = [torch.randn(N, d, generator=gen).bfloat16() for _ in range(nmod)] + [None] * (3 - nmod)
    if labels_kind == "onehot":
        labels = torch.arange(N)
    elif labels_kind == "zipf":
        wz = 1.0 / torch.arange(1, max(2, N // 8) + 1, dtype=torch.float64)
        labels = torch.multinomial(wz / wz.sum(), N, replacement=True, generator=gen)
    else:
        labels = torch.randint(0, max(1, N // 8), (N,), generator=gen)
    scale = 1 / 0.07
    ref = lo.contrastive_loss([None if f is None else f.float().numpy() for f in feats], labels.numpy(), scale)
    weights = (1 / 6, 1 / 6, 1 / 6) if nmod == 3 else (0.5, 0.0, 0.0)
    path = _lib.PATH_TC_BF16 if operands == "bf16" else _lib.PATH_TC_F16
    grad_outs = [1.0 + 0.5 * r for r in range(world)]  # a different upstream gradient on every rank: the SUM scales dx
    gsum = sum(grad_outs)
    losses, grads, ds = _sharded_step([None if f is None else f.to(dev) for f in feats], labels.to(dev), scale, world,
                                      path, form, weights, grad_outs, labels_first=(form == "peer"))
    for l in losses:
        assert abs(l - ref["loss"]) <= 1e-3 * abs(ref["loss"])
        assert l == losses[0]  # bit-identical on every rank
    for i in range(3):
        if ref["grads"][i] is not None:
            assert _rel(grads[i], gsum * ref["grads"][i]) < 1e-3 + 2 ** -8, i
    assert abs(ds - ref["dlogit_scale"]) <= 1e-3 * abs(ref["dlogit_scale"])


def test_simulated_ranks_trained_regime():
    """Aligned modalities (G~ -> 2 T): the lam2 correction must survive the exchange -- lam2 of ALL rows comes from
    the exchanged posrow, the row side subtracts it in the sweep's epilogue and the column side through the weighted
    class sums."""
    from clibd_b200 import _lib
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(21)
    N, d, world, align = 1024, 768, 4, 0.7
    labels = torch.randint(0, N // 4, (N,), generator=gen)
    base = torch.randn(N, d, generator=gen)[labels]
    # fp32 tensors holding fp16-representable values: gradients come back in fp32 (tests/test_loss_gpu.py,
    # test_trained_regime_aligned_modalities: same inputs, same floor of 6e-3 set by the 16-bit S operands)
    feats = [(align * base + (1 - align) * torch.randn(N, d, generator=gen)).half().float() for _ in range(3)]
    scale = 1 / 0.07
    ref = lo.contrastive_loss([f.numpy() for f in feats], labels.numpy(), scale)
    for form in ("reduce_scatter", "peer"):
        losses, grads, ds = _sharded_step([f.to(dev) for f in feats], labels.to(dev), scale, world, _lib.PATH_TC_F16, form)
        assert abs(losses[0] - ref["loss"]) <= 1e-3 * max(abs(ref["loss"]), scale)
        for i in range(3):
            assert _rel(grads[i], world * ref["grads"][i]) < 6e-3, (form, i)


def test_side_stream_overlap_is_bit_identical(monkeypatch):
    """Staging work that runs on the library's side stream next to the tensor kernels (class sums, the operand copies
    of the later pairs; fork / join by events, csrc/loss_api.cu) must not change a single bit: the same sharded step
    with CLIBD_SIDE_STREAM=0 (everything in line) and =1, repeated -- a missing dependency shows up as a mismatch."""
    from clibd_b200 import _lib
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    N, d, world = 2048, 768, 4
    feats = [torch.randn(N, d, generator=gen).bfloat16().to(dev) for _ in range(3)]
    labels = torch.randint(0, N // 8, (N,), generator=gen).to(dev)
    monkeypatch.setenv("CLIBD_SIDE_STREAM", "0")
    ref = _sharded_step(feats, labels, 1 / 0.07, world, _lib.PATH_TC_BF16, "peer")
    for rep in range(3):
        monkeypatch.setenv("CLIBD_SIDE_STREAM", "1")
        got = _sharded_step(feats, labels, 1 / 0.07, world, _lib.PATH_TC_BF16, "peer")
        assert got[0] == ref[0] and got[2] == ref[2]
        for a, b in zip(got[1], ref[1]):
            assert np.array_equal(a, b)


def test_gradient_gemm_on_the_side_stream_is_bit_identical(monkeypatch):
    """The gradient GEMM of the first pair group runs on the side stream next to the following sweeps, each group with
    its own strip buffer (loss_api.cu: shared_s_sweeps); in line (CLIBD_OVERLAP_GEMM=0) it must give the same bits."""
    from clibd_b200 import _lib
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(4)
    N, d, world = 4096, 768, 4
    feats = [torch.randn(N, d, generator=gen).bfloat16().to(dev) for _ in range(3)]
    labels = torch.randint(0, N // 8, (N,), generator=gen).to(dev)
    for form in ("peer", "reduce_scatter"):
        monkeypatch.setenv("CLIBD_OVERLAP_GEMM", "0")
        ref = _sharded_step(feats, labels, 1 / 0.07, world, _lib.PATH_TC_BF16, form)
        monkeypatch.setenv("CLIBD_OVERLAP_GEMM", "1")
        for rep in range(3):
            got = _sharded_step(feats, labels, 1 / 0.07, world, _lib.PATH_TC_BF16, form)
            assert got[0] == ref[0] and got[2] == ref[2]
            for a, b in zip(got[1], ref[1]):
                assert np.array_equal(a, b)


def test_barrier_kernel_single_rank_epochs():
    """clibd_shard_barrier with world = 1 (the rank signals itself): the per-channel epoch counters advance, the flag
    slots follow, channels are independent.  Ranks that wait for each other cannot be simulated with streams of ONE
    GPU -- two streams may share a hardware queue, where the kernel behind a spinning barrier is never dispatched --
    so the multi-rank behaviour is checked on real GPUs (tools/multigpu_check.py runs every exchange through it)."""
    from clibd_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    words = lib.clibd_shard_barrier_bytes() // 8
    assert words == 4 * 16 + 4
    block = torch.zeros(words, dtype=torch.int64, device=dev)
    flags = _lib.ptr_array([block.data_ptr()])
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for _ in range(5):
        _lib.check(lib.clibd_shard_barrier(flags, 0, 1, 0, sp))
    for _ in range(2):
        _lib.check(lib.clibd_shard_barrier(flags, 0, 1, 3, sp))
    torch.cuda.synchronize()
    host = block.cpu()
    assert int(host[4 * 16 + 0]) == 5 and int(host[0]) == 5          # channel 0: counter and the slot rank 0 wrote
    assert int(host[4 * 16 + 3]) == 2 and int(host[3 * 16]) == 2      # channel 3
    assert int(host[4 * 16 + 1]) == 0 and int(host[1]) == 0           # nobody else was touched
    assert lib.clibd_shard_barrier(flags, 1, 1, 0, sp) != 0 and lib.clibd_shard_barrier(flags, 0, 1, 4, sp) != 0


@pytest.mark.parametrize("bind_to,no_it,weights", [
    ("image", False, (0.25, 0.25, 0.0)),   # bind_to="image": (image,dna) and (image,text); text's only column pair is pair 1
    (None, True, (0.25, 0.0, 0.25)),       # no_image_text_loss: (image,dna) and (dna,text); text's column pair is pair 2
    ("text", False, (0.0, 0.25, 0.25)),    # bind_to="text": both weighted pairs share the column modality -> one merged GEMM
])
def test_simulated_ranks_pair_filters(bind_to, no_it, weights):
    """The reference's pair filters (loss_func.py:166-184) change which pairs exist, hence which slot array receives a
    column modality's partial gradients and whether two pairs are merged into one gradient GEMM."""
    from clibd_b200 import _lib
    from clibd_b200.loss import pair_weights
    assert tuple(pair_weights([True, True, True], bind_to, no_it)[0]) == pytest.approx(weights)
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(8)
    N, d, world = 512, 768, 2
    feats = [torch.randn(N, d, generator=gen).bfloat16() for _ in range(3)]
    labels = torch.randint(0, N // 8, (N,), generator=gen)
    scale = 1 / 0.07
    ref = lo.contrastive_loss([f.float().numpy() for f in feats], labels.numpy(), scale, bind_to=bind_to,
                              no_image_text_loss=no_it)
    for form in ("peer", "reduce_scatter"):
        losses, grads, ds = _sharded_step([f.to(dev) for f in feats], labels.to(dev), scale, world, _lib.PATH_TC_F16, form,
                                          weights)
        assert abs(losses[0] - ref["loss"]) <= 1e-3 * abs(ref["loss"])
        for i in range(3):
            assert _rel(grads[i], world * ref["grads"][i]) < 1e-3 + 2 ** -8, (form, i)
        assert abs(ds - ref["dlogit_scale"]) <= 1e-3 * abs(ref["dlogit_scale"])


def test_unequal_pair_weights_are_linear_in_the_pairs(monkeypatch):
    """Pair weights that differ (only reachable through the C ABI: the reference's pair list always weighs its pairs
    equally) keep the two text pairs in separate gradient GEMMs that accumulate into one partial buffer -- three pair
    groups, three strip buffers, the last GEMM joined behind the side-stream ones.  The step is linear in the pair
    weights, so it must equal the sum of the three single-pair steps."""
    from clibd_b200 import _lib
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(21)
    N, d, world = 1024, 768, 4
    feats = [torch.randn(N, d, generator=gen).bfloat16().to(dev) for _ in range(3)]
    labels = torch.randint(0, N // 8, (N,), generator=gen).to(dev)
    scale = 1 / 0.07
    w = (0.2, 0.3, 0.5)
    for ov in ("1", "0"):
        monkeypatch.setenv("CLIBD_OVERLAP_GEMM", ov)
        losses, grads, ds = _sharded_step(feats, labels, scale, world, _lib.PATH_TC_BF16, "reduce_scatter", w)
        loss_sum, ds_sum = 0.0, 0.0
        grad_sum = [np.zeros((N, d), np.float64) for _ in range(3)]
        for p in range(3):
            wp = tuple(w[q] if q == p else 0.0 for q in range(3))
            l1, g1, d1 = _sharded_step(feats, labels, scale, world, _lib.PATH_TC_BF16, "reduce_scatter", wp)
            loss_sum += l1[0]
            ds_sum += d1
            for i in range(3):
                if g1[i] is not None:
                    grad_sum[i] += np.asarray(g1[i], np.float64)
        assert abs(losses[0] - loss_sum) <= 1e-5 * abs(loss_sum)
        assert abs(ds - ds_sum) <= 1e-4 * abs(ds_sum)
        for i in range(3):
            assert _rel(grads[i], grad_sum[i]) < 1e-2, (ov, i)
