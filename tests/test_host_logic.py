"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls), the pair-weight logic matches the reference's pair loop, and the
multi-rank orchestration of ClipLoss reproduces the reference's world-2 golden vectors under gloo
(with a numpy double standing in for the CUDA entry points)."""
import itertools
import os
import re
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import loss_oracle as lo
from tests import _golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from clibd_b200 import _build, _lib
    header = open(os.path.join(ROOT, "include", "clibd_b200.h")).read()
    declared = set(re.findall(r"\b(clibd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()  # builds if needed, dlopens, binds every symbol
    assert os.path.exists(_build.LIB_PATH)
    assert lib.clibd_abi_version() == _lib.ABI_VERSION
    assert lib.clibd_loss_scratch_bytes(4096, 512, 768, 1, 0) > 0
    assert lib.clibd_loss_scratch_bytes(4096, 512, 768, 1, 1) > lib.clibd_loss_scratch_bytes(4096, 512, 768, 1, 0)
    assert lib.clibd_loss_scratch_bytes(0, 0, 768, 1, 0) == -1
    assert lib.clibd_knn_scratch_bytes(1000, 100000, 768, 5, 2) > 0


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "clibd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("present", [p for p in itertools.product([0, 1], repeat=3) if sum(p) >= 2])
@pytest.mark.parametrize("bind_to", [None, "image", "dna", "text"])
@pytest.mark.parametrize("no_it", [False, True])
def test_pair_weights_match_reference_pair_loop(present, bind_to, no_it):
    from clibd_b200.loss import pair_weights
    w, n_ordered = pair_weights(present, bind_to, no_it)
    pairs = lo.ordered_pairs(sum(present), bind_to, no_it)
    assert n_ordered == len(pairs)
    slots = [i for i, p in enumerate(present) if p]
    expect = [0.0, 0.0, 0.0]
    idx = {(0, 1): 0, (0, 2): 1, (1, 2): 2}
    for a, b in pairs:
        key = (min(slots[a], slots[b]), max(slots[a], slots[b]))
        expect[idx[key]] += 1.0 / (2 * len(pairs))
    assert w == pytest.approx(expect)


def test_reference_argument_errors_without_gpu():
    import clibd_b200 as cb
    x = torch.randn(4, 8)
    with pytest.raises(ValueError, match="Too less element"):
        cb.ContrastiveLoss(torch.nn.CrossEntropyLoss(), 1.0)(x, None, None, torch.arange(4), 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cb.ContrastiveLoss(torch.nn.CrossEntropyLoss(), 1.0)(x, x, None, torch.arange(4), 1.0)
    with pytest.raises(NotImplementedError):
        cb.ClipLoss(use_horovod=True)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA device"):
            cb.make_prediction(np.zeros((2, 4)), np.zeros((3, 4)), [{}] * 3)


def test_single_process_orchestration_with_double_matches_golden():
    from clibd_b200 import _lib
    import clibd_b200 as cb
    from tests._fake_lib import FakeLib
    _lib.inject_for_tests(FakeLib())
    try:
        for g in _golden.all_single_process():
            if g.meta.get("inputs_are_bf16_exact"):
                continue
            feats = [None if f is None else torch.from_numpy(f).requires_grad_(True) for f in g.features]
            mod = cb.ClipLoss(gather_with_grad=True, **g.kwargs())
            scale = torch.tensor(g.logit_scale, requires_grad=True)
            loss = mod(feats[0], feats[1], feats[2], torch.from_numpy(g.labels), scale)
            (loss * g.meta.get("grad_mult", 1.0)).backward()
            assert float(loss) == pytest.approx(float(g.outputs["loss"]), rel=2e-6)
            for f, m in zip(feats, _golden.MODS):
                if f is not None and f"grad_{m}" not in g.outputs:
                    assert f.grad is None, m  # present but in no pair: no gradient, as in the reference
                elif f is not None:
                    ref = g.outputs[f"grad_{m}"]
                    assert np.linalg.norm(f.grad.numpy() - ref) / np.linalg.norm(ref) < 2e-5
            if "dlogit_scale" in g.outputs:
                assert float(scale.grad) == pytest.approx(float(g.outputs["dlogit_scale"]), rel=2e-4, abs=1e-7)
    finally:
        _lib.inject_for_tests(None)


def _w2_worker(rank, world, store, ret, operands):
    from clibd_b200 import _lib
    import clibd_b200 as cb
    from tests._fake_lib import FakeLib
    dist.init_process_group("gloo", init_method=f"file://{store}", rank=rank, world_size=world)
    _lib.inject_for_tests(FakeLib())
    g = _golden.load("cliploss_w2_all_n64_d32")
    n = 32
    sl = slice(rank * n, (rank + 1) * n)
    feats = [torch.from_numpy(g.inputs[m][sl].copy()).requires_grad_(True) for m in _golden.MODS]
    scale = torch.tensor(g.logit_scale, requires_grad=True)
    # operands=None: fp32 inputs take the CUDA-core path, whose sharded step is the 'local' form (two sweeps per
    # rank); operands="bf16": a tensor-core path, whose sharded step is the exchange form (here with the collectives
    # of the 'nccl' variant: all-gather, all-reduce of statistics incl. posrow, reduce-scatter of the partials)
    mod = cb.ClipLoss(local_loss=False, gather_with_grad=True, rank=rank, world_size=world,
                      tensor_core_operands=operands)
    loss = mod(feats[0], feats[1], feats[2], torch.from_numpy(g.labels[sl].copy()), scale)
    loss.backward()
    out = {"loss": float(loss), "ds": float(scale.grad)}
    for f, m in zip(feats, _golden.MODS):
        out[m] = f.grad.numpy().tolist()
    # gather_with_grad=False: only this rank's loss reaches the local rows (no sum over ranks)
    feats2 = [torch.from_numpy(g.inputs[m][sl].copy()).requires_grad_(True) for m in _golden.MODS]
    mod2 = cb.ClipLoss(local_loss=False, gather_with_grad=False, rank=rank, world_size=world,
                       tensor_core_operands=operands)
    mod2(feats2[0], feats2[1], feats2[2], torch.from_numpy(g.labels[sl].copy()), g.logit_scale).backward()
    out["image_nograd_gather"] = feats2[0].grad.numpy().tolist()
    # the public gather helper (loss_func.py:73-106) on CPU tensors: the reference's collectives
    for with_grad in (True, False):
        x = (torch.arange(12, dtype=torch.float32).reshape(3, 4) + 100 * rank).requires_grad_(True)
        allf = cb.gather_features(x, local_loss=False, gather_with_grad=with_grad, rank=rank, world_size=world)
        (allf * (rank + 1)).sum().backward()
        out[f"gather_{with_grad}"] = (allf.detach().numpy().tolist(), x.grad.numpy().tolist())
    ret[rank] = out
    dist.destroy_process_group()


@pytest.mark.parametrize("operands", [None, "bf16"], ids=["local_two_sweeps", "exchange_reduce_scatter"])
def test_world2_gloo_orchestration_matches_reference_golden(operands):
    """The reference ran ClipLoss(gather_with_grad=True) under 2 gloo processes (oracle/gen_golden.py);
    our orchestration (all-gather in, all-reduce of statistics, sum-over-ranks gradient scale) must give
    every rank the same full-batch loss and W x the local slice of the full-batch gradient."""
    world = 2
    store = tempfile.mktemp()
    ret = mp.Manager().dict()
    mp.spawn(_w2_worker, args=(world, store, ret, operands), nprocs=world, join=True)
    g = _golden.load("cliploss_w2_all_n64_d32")
    for r in range(world):
        assert ret[r]["loss"] == pytest.approx(float(g.outputs[f"rank{r}_loss"]), rel=2e-6)
        assert ret[r]["ds"] == pytest.approx(float(g.outputs[f"rank{r}_dlogit_scale"]), rel=2e-4)
        for m in _golden.MODS:
            ref = g.outputs[f"rank{r}_grad_{m}"]
            got = np.asarray(ret[r][m], dtype=np.float32)
            assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 2e-5
        half = np.asarray(ret[r]["image_nograd_gather"], dtype=np.float32)
        ref = g.outputs[f"rank{r}_grad_image"] / world
        assert np.linalg.norm(half - ref) / np.linalg.norm(ref) < 2e-5
        if operands is None:
            full = np.concatenate([np.arange(12, dtype=np.float32).reshape(3, 4) + 100 * q for q in range(world)])
            for with_grad, gsum in ((True, sum(q + 1 for q in range(world))), (False, r + 1)):
                vals, grad = ret[r][f"gather_{with_grad}"]
                assert np.array_equal(np.asarray(vals, dtype=np.float32), full)
                assert np.array_equal(np.asarray(grad, dtype=np.float32), np.full((3, 4), gsum, dtype=np.float32))


def test_pipeline_bounds_cover_the_key_range():
    """Block layout of the host-key kNN pipeline (retrieval._pipeline_bounds): contiguous cover of [lo, hi), a small
    first block, every later block at most twice its predecessor (its copy hides behind the previous block's screen),
    few blocks; an explicit block count gives equal blocks."""
    from clibd_b200 import retrieval as R
    for lo, hi in ((0, 1_000_000), (7, 125_007), (0, 3), (0, 65_537), (100, 100 + (1 << 15)), (0, 5_000_000)):
        b = R._pipeline_bounds(lo, hi)
        sizes = [b[i + 1] - b[i] for i in range(len(b) - 1)]
        assert b[0] == lo and b[-1] == hi and all(s > 0 for s in sizes)
        assert len(sizes) <= R._PIPELINE_BLOCKS
        if len(sizes) > 1:
            assert sizes[0] == R._PIPELINE_FIRST_KEYS
            assert all(sizes[i + 1] == 2 * sizes[i] for i in range(len(sizes) - 2))  # all but the last: doubling
            assert sizes[-1] > sizes[-2] or len(sizes) == R._PIPELINE_BLOCKS or sizes[-1] <= 4 * sizes[-2]
    assert R._pipeline_bounds(0, 10, 4) == [0, 2, 5, 7, 10]
    assert R._pipeline_bounds(0, 3, 8) == [0, 1, 2, 3]


def test_shard_mode_selection(monkeypatch):
    """Which exchange form a row-sharded step takes (loss._shard_mode): the peer form only on CUDA under NCCL with
    symmetric memory, for a tensor-core path, d <= 768 and at most 16 ranks; the environment can force a form."""
    from clibd_b200 import _lib, _peer
    from clibd_b200 import loss as L
    cuda = torch.device("cuda", 0)
    monkeypatch.delenv("CLIBD_SHARD_MODE", raising=False)
    monkeypatch.setattr(L.dist, "get_backend", lambda group=None: "nccl")
    monkeypatch.setattr(_peer, "available", lambda: True)
    assert L._shard_mode(_lib.PATH_TC_BF16, 768, None, cuda, 8) == "peer"
    assert L._shard_mode(_lib.PATH_TC_BF16, 768, None, cuda, 32) == "local"       # more ranks than peer slots
    assert L._shard_mode(_lib.PATH_TC_F16, 769, None, cuda, 8) == "local"         # padded d > 768: single-CTA sweep
    assert L._shard_mode(_lib.PATH_SIMT_F32, 768, None, cuda, 8) == "local"       # exact CUDA-core path
    assert L._shard_mode(_lib.PATH_TC_BF16, 768, None, torch.device("cpu"), 2) == "nccl"
    monkeypatch.setattr(_peer, "available", lambda: False)
    assert L._shard_mode(_lib.PATH_TC_BF16, 768, None, cuda, 8) == "nccl"
    monkeypatch.setattr(L.dist, "get_backend", lambda group=None: "gloo")
    monkeypatch.setattr(_peer, "available", lambda: True)
    assert L._shard_mode(_lib.PATH_TC_BF16, 768, None, cuda, 8) == "nccl"
    monkeypatch.setattr(L.dist, "get_backend", lambda group=None: "nccl")
    for forced in ("local", "nccl", "peer"):
        monkeypatch.setenv("CLIBD_SHARD_MODE", forced)
        assert L._shard_mode(_lib.PATH_TC_BF16, 768, None, cuda, 8) == forced
    monkeypatch.setenv("CLIBD_SHARD_MODE", "peer")
    assert L._shard_mode(_lib.PATH_SIMT_F32, 768, None, cuda, 8) == "local"       # no exchange form for this path


def test_peer_disable_is_sticky(monkeypatch):
    """_peer.disable(reason) (called when mapping the ranks' buffers failed) switches the peer form off for the rest
    of the process and warns once."""
    import warnings
    from clibd_b200 import _peer
    monkeypatch.setattr(_peer, "_disabled_reason", None)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        _peer.disable("no fd passing")
        _peer.disable("again")
    assert len(w) == 1 and "no fd passing" in str(w[0].message)
    assert _peer.available() is False and _peer._disabled_reason == "no fd passing"
