"""GPU parity tests of the fused contrastive loss (through the C ABI) against
(a) the golden vectors produced by the reference's own PyTorch code and (b) the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo
from tests import _golden

pytestmark = pytest.mark.gpu

SINGLE = _golden.all_single_process()


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _run(module, feats, labels, scale, grad_mult=1.0, **fw):
    leaves = [None if f is None else f.clone().requires_grad_(True) for f in feats]
    s = scale.clone().requires_grad_(True) if isinstance(scale, torch.Tensor) else scale
    loss = module(leaves[0], leaves[1], leaves[2], labels, s, **fw)
    if isinstance(loss, dict):
        loss = loss["contrastive_loss"]
    (loss * grad_mult).backward()
    torch.cuda.synchronize()
    return (float(loss), [None if (l is None or l.grad is None) else l.grad.float().cpu().numpy() for l in leaves],
            float(s.grad) if isinstance(s, torch.Tensor) else None)


@pytest.mark.parametrize("g", SINGLE, ids=[g.name for g in SINGLE])
def test_fp32_path_matches_reference_golden(g):
    """fp32 CUDA-core path vs the reference's PyTorch output; tolerance 1e-5 (north_star, fp32)."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    feats = [None if f is None else torch.from_numpy(f).to(dev) for f in g.features]
    labels = torch.from_numpy(g.labels).to(dev)
    if g.meta["module"] == "ContrastiveLoss":
        mod = cb.ContrastiveLoss(torch.nn.CrossEntropyLoss(), 1 / 0.07)
    else:
        mod = cb.ClipLoss(gather_with_grad=True, rank=0, world_size=1, **g.kwargs())
    scale = None if g.meta.get("scale_from_ctor") else torch.tensor(g.logit_scale, device=dev)
    loss, grads, ds = _run(mod, feats, labels, scale, grad_mult=g.meta.get("grad_mult", 1.0))
    ref_loss = float(g.outputs["loss"])
    assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)
    for i, m in enumerate(_golden.MODS):
        if f"grad_{m}" in g.outputs:
            assert _rel(grads[i], g.outputs[f"grad_{m}"]) < 1e-5, m
        else:  # absent, or present but in no pair (the reference leaves its .grad None)
            assert grads[i] is None, m
    if "dlogit_scale" in g.outputs:
        ref = float(g.outputs["dlogit_scale"])
        assert abs(ds - ref) <= 1e-5 * abs(ref) + 1e-7
    # ... and against the float64 oracle on the same inputs: the exact path is as close to the true value as the
    # reference's own float32 evaluation (measured: loss <= 7e-8, gradients <= 3e-6, dlogit_scale <= 1.2e-7 on every
    # golden, the reference's outputs 1e-7 / 3e-6 / 3e-7; tools/fp32_error_probe.py, profiles/r2o_fp32_error.log)
    if not g.meta.get("inputs_are_bf16_exact"):
        mult = g.meta.get("grad_mult", 1.0)
        exact = lo.contrastive_loss(g.features, g.labels, g.logit_scale, grad_out=mult, **g.kwargs())
        assert abs(loss - exact["loss"]) <= 1e-6 * abs(exact["loss"])
        for i in range(3):
            if exact["grads"][i] is not None:
                assert _rel(grads[i], exact["grads"][i]) < 1e-5, i
        if ds is not None:
            assert abs(ds - exact["dlogit_scale"]) <= 1e-5 * abs(exact["dlogit_scale"]) + 1e-9


def _synthetic(N, d, nmod, labels_kind, seed, dtype):
    gen = torch.Generator().manual_seed(seed)
    feats = [torch.randn(N, d, generator=gen).to(dtype) for _ in range(nmod)] + [None] * (3 - nmod)
    if labels_kind == "onehot":
        labels = torch.arange(N)
    elif labels_kind == "zipf":
        w = 1.0 / torch.arange(1, max(2, N // 8) + 1, dtype=torch.float64)
        labels = torch.multinomial(w / w.sum(), N, replacement=True, generator=gen)
    else:
        labels = torch.randint(0, max(1, N // 8), (N,), generator=gen)
    return feats, labels


@pytest.fixture(params=["two_sweeps", "shared_s"])
def backward_form(request, monkeypatch):
    """Both backward forms at oracle-sized batches: the two-sweep form (what small batches and sharded jobs run)
    and the single-GPU form that computes S once per pair (normally taken from N = 4096 up)."""
    if request.param == "shared_s":
        monkeypatch.setenv("CLIBD_SHARED_S_MIN_N", "1")
    else:
        monkeypatch.setenv("CLIBD_BWD_TWO_SWEEPS", "1")
    return request.param


@pytest.mark.parametrize("operands", ["bf16", "fp16"])
@pytest.mark.parametrize("N,d,nmod,labels_kind", [
    (1024, 768, 3, "multi"),      # BASELINE config-2 shape at a size the oracle finishes in seconds
    (1000, 200, 2, "onehot"),     # ragged: N not a multiple of 128/256, d not a multiple of 64
    (384, 768, 3, "zipf"),        # long-tailed class sizes
    (130, 64, 2, "multi"),        # one full row tile + 2 rows
])
def test_tensor_core_path_matches_oracle(operands, N, d, nmod, labels_kind, backward_form):
    """tcgen05 path vs the float64 oracle on the same bf16-rounded inputs; tolerance 1e-3
    (north_star: loss and gradients within 1e-3 relative with 16-bit operands)."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    feats, labels = _synthetic(N, d, nmod, labels_kind, seed=11, dtype=torch.bfloat16)
    scale = torch.tensor(1 / 0.07)
    ref = lo.contrastive_loss([None if f is None else f.float().numpy() for f in feats], labels.numpy(), float(scale))
    mod = cb.ContrastiveLoss(torch.nn.CrossEntropyLoss(), 1 / 0.07, tensor_core_operands=operands)
    loss, grads, ds = _run(mod, [None if f is None else f.to(dev) for f in feats], labels.to(dev), scale.to(dev))
    assert abs(loss - ref["loss"]) <= 1e-3 * abs(ref["loss"])
    for i in range(3):
        if ref["grads"][i] is not None:
            # gradients are returned in the input dtype (bf16): allow its rounding on top of 1e-3
            assert _rel(grads[i], ref["grads"][i]) < 1e-3 + 2 ** -8, i
    assert abs(ds - ref["dlogit_scale"]) <= 1e-3 * abs(ref["dlogit_scale"])


@pytest.mark.parametrize("N", [512, 4096])
def test_modality_in_no_pair_gets_no_gradient_tensor_core_path(N):
    """ClipLoss(bind_to="image", no_image_text_loss=True) with all three modalities: only (image, dna) survives the
    reference's filters (loss_func.py:166-184); text is staged nowhere and its leaf keeps grad None.  Both backward
    forms of the tcgen05 path (two sweeps at N = 512, S once per pair at N = 4096) against the oracle."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    feats, labels = _synthetic(N, 768, 3, "multi", seed=31, dtype=torch.bfloat16)
    scale = torch.tensor(1 / 0.07)
    kw = {"bind_to": "image", "no_image_text_loss": True}
    ref = lo.contrastive_loss_streaming([f.float().numpy() for f in feats], labels.numpy(), float(scale), **kw)
    mod = cb.ClipLoss(gather_with_grad=True, rank=0, world_size=1, tensor_core_operands="bf16", **kw)
    loss, grads, ds = _run(mod, [f.to(dev) for f in feats], labels.to(dev), scale.to(dev))
    assert abs(loss - ref["loss"]) <= 1e-3 * abs(ref["loss"])
    assert grads[2] is None and ref["grads"][2] is None
    for i in range(2):
        assert _rel(grads[i], ref["grads"][i]) < 1e-3 + 2 ** -8, i
    assert abs(ds - ref["dlogit_scale"]) <= 1e-3 * abs(ref["dlogit_scale"])


def test_wide_features_take_the_single_cta_sweep():
    """d = 1024 > 768: the CTA-pair kernels hold at most 768 feature columns in TMEM, so the forward runs on the
    pair kernel and the backward on the single-CTA sweep (loss_tc.cu, two 384-column chunks per 128 rows,
    class-sorted column operand like the pair path) -- same tolerance against the oracle."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    feats, labels = _synthetic(300, 1024, 2, "multi", seed=13, dtype=torch.bfloat16)
    scale = torch.tensor(1 / 0.07)
    ref = lo.contrastive_loss([None if f is None else f.float().numpy() for f in feats], labels.numpy(), float(scale))
    mod = cb.ContrastiveLoss(torch.nn.CrossEntropyLoss(), 1 / 0.07, tensor_core_operands="bf16")
    loss, grads, ds = _run(mod, [None if f is None else f.float().to(dev) for f in feats], labels.to(dev), scale.to(dev))
    assert abs(loss - ref["loss"]) <= 1e-3 * abs(ref["loss"])
    for i in range(2):
        assert _rel(grads[i], ref["grads"][i]) < 1e-3, i
    assert abs(ds - ref["dlogit_scale"]) <= 1e-3 * abs(ref["dlogit_scale"])


@pytest.mark.parametrize("operands", ["bf16", "fp16"])
def test_tensor_core_fp32_inputs_gradient_tolerance(operands):
    """fp32 inputs forced through the tensor-core path: gradients come back in fp32, so the
    1e-3 bound is checked without output rounding."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    feats, labels = _synthetic(1024, 768, 3, "multi", seed=12, dtype=torch.float32)
    scale = torch.tensor(1 / 0.07)
    ref = lo.contrastive_loss([f.numpy() for f in feats], labels.numpy(), float(scale))
    mod = cb.ContrastiveLoss(torch.nn.CrossEntropyLoss(), 1 / 0.07, tensor_core_operands=operands)
    loss, grads, ds = _run(mod, [f.to(dev) for f in feats], labels.to(dev), scale.to(dev))
    assert abs(loss - ref["loss"]) <= 1e-3 * abs(ref["loss"])
    for i in range(3):
        assert _rel(grads[i], ref["grads"][i]) < 1e-3, i
    assert abs(ds - ref["dlogit_scale"]) <= 1e-3 * abs(ref["dlogit_scale"])


@pytest.mark.parametrize("operands", ["bf16", "fp16"])
@pytest.mark.parametrize("labels_kind,align", [("onehot", 0.7), ("onehot", 0.5), ("multi", 0.7)])
def test_trained_regime_aligned_modalities(operands, labels_kind, align, backward_form):
    """Late-training regime: the modalities of one specimen are strongly aligned, so the softmax mass sits on the
    positives and dL/dS = c_i p_ij + c_j p_ji - 2 T_ij is a small difference there.  The positives' G~ is formed
    as G~ - lam2 in fp32 inside the tensor-core epilogue BEFORE the 16-bit rounding (class-sorted column
    operand, loss_bwd_pair.cu), so gradients keep their relative tolerance instead of drowning in the rounding
    of a value near 2 (measured before that change: 47 % bf16 / 11 % fp16 relative error at align 0.7)."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    N, d = 768, 768
    gen = torch.Generator().manual_seed(21)
    labels = torch.arange(N) if labels_kind == "onehot" else torch.randint(0, N // 4, (N,), generator=gen)
    centres = torch.randn(N, d, generator=gen)
    base = centres[labels] if labels_kind == "multi" else centres
    dt = torch.bfloat16 if operands == "bf16" else torch.float16
    # fp32 leaves holding 16-bit representable values: gradients come back in fp32 (a 16-bit gradient of this
    # size would underflow in fp16 without GradScaler and hide the kernel's own error)
    feats = [(align * base + (1 - align) * torch.randn(N, d, generator=gen)).to(dt).float() for _ in range(3)]
    scale = torch.tensor(1 / 0.07)
    ref = lo.contrastive_loss([f.float().numpy() for f in feats], labels.numpy(), float(scale))
    mod = cb.ContrastiveLoss(torch.nn.CrossEntropyLoss(), 1 / 0.07, tensor_core_operands=operands)
    loss, grads, ds = _run(mod, [f.to(dev) for f in feats], labels.to(dev), scale.to(dev))
    errs = [_rel(grads[i], ref["grads"][i]) for i in range(3)]
    print("trained regime", operands, labels_kind, align, "loss", loss, ref["loss"], "grad rel", errs,
          "ds", ds, ref["dlogit_scale"])
    # the loss is a difference of terms of size s (LSE - positive logit): tolerance relative to that size
    assert abs(loss - ref["loss"]) <= 1e-3 * max(abs(ref["loss"]), float(scale))
    # Floors set by the 16-bit rounding of the S-GEMM operands alone (tools/emulate_operand_rounding.py, float64
    # emulation: onehot 0.7 -> 2.9e-3 bf16 / 2.5e-4 fp16; multi 0.7 -> 2.7e-2 / 2.8e-3, where e_ij varies by
    # +-30 % inside a class and the class-common part of the gradient cancels)
    tol = {("onehot", 0.7): (6e-3, 1e-3), ("onehot", 0.5): (2e-3, 1e-3), ("multi", 0.7): (4e-2, 6e-3)}[(labels_kind, align)]
    for e in errs:
        assert e < tol[0 if operands == "bf16" else 1]
    assert abs(ds - ref["dlogit_scale"]) <= 2e-2 * abs(ref["dlogit_scale"]) + 1e-6


def test_tensor_core_matches_cuda_core_path_n4096():
    """BASELINE config 2 (N=4096, three modalities, label-matched multi-positives): tcgen05 vs
    the exact fp32 CUDA-core path on the GPU (the oracle would need minutes here)."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    feats, labels = _synthetic(4096, 768, 3, "multi", seed=1, dtype=torch.float32)
    feats = [f.bfloat16().float().to(dev) for f in feats]
    labels = labels.to(dev)
    scale = torch.tensor(1 / 0.07, device=dev)
    exact = _run(cb.ContrastiveLoss(None, 1 / 0.07, tensor_core_operands="fp32"), feats, labels, scale)
    for operands in ("bf16", "fp16"):
        got = _run(cb.ContrastiveLoss(None, 1 / 0.07, tensor_core_operands=operands), feats, labels, scale)
        assert abs(got[0] - exact[0]) <= 1e-3 * abs(exact[0])
        for a, b in zip(got[1], exact[1]):
            assert _rel(a, b) < 1e-3
        assert abs(got[2] - exact[2]) <= 1e-3 * abs(exact[2])


def test_bf16_fed_reference_secondary_target():
    """The reference fed bf16 tensors directly rounds its logits to bf16; we must be at least as
    close to the fp32 reference as it is (secondary target)."""
    import clibd_b200 as cb
    g = _golden.load("loss_imgdna_bf16inputs_n128_d64")
    dev = torch.device("cuda:0")
    feats = [None if f is None else torch.from_numpy(f).to(dev).bfloat16() for f in g.features]
    mod = cb.ContrastiveLoss(torch.nn.CrossEntropyLoss(), 1 / 0.07)
    loss, grads, ds = _run(mod, feats, torch.from_numpy(g.labels).to(dev), torch.tensor(g.logit_scale, device=dev))
    ref = float(g.outputs["loss"])
    assert abs(loss - ref) <= max(1e-3 * abs(ref), abs(float(g.outputs["bf16fed_loss"]) - ref))
    assert _rel(grads[0], g.outputs["grad_image"]) < 1e-3 + 2 ** -8


def test_full_size_properties_n32768():
    """BASELINE north-star size (N=32768, d=768, bf16): properties that need no oracle.
      * x_i . dL/dx_i == 0 for every row (the loss is invariant to the scale of each row);
      * dL/ds matches a central finite difference of the loss;
      * swapping the two modalities leaves the loss unchanged;
      * gradients are linear in the upstream gradient (GradScaler, train_epoch.py:58)."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    N, d = 32768, 768
    gen = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(N, d, generator=gen).bfloat16().to(dev)
    b = (a.float().cpu() * 0.5 + torch.randn(N, d, generator=gen)).bfloat16().to(dev)
    labels = torch.randint(0, N // 8, (N,), generator=gen).to(dev)
    mod = cb.ContrastiveLoss(None, 1 / 0.07)
    s0 = 1 / 0.07
    loss, grads, ds = _run(mod, [a, b, None], labels, torch.tensor(s0, device=dev))
    assert np.isfinite(loss) and all(np.isfinite(g).all() for g in grads[:2])
    for x, g in ((a, grads[0]), (b, grads[1])):
        xr = x.float().cpu().numpy()
        dots = np.abs((xr * g).sum(1))
        scale_ref = np.linalg.norm(xr, axis=1) * np.linalg.norm(g, axis=1)
        assert np.median(dots / scale_ref) < 2e-2  # bf16 output rounding of g limits this
    h = 0.05
    with torch.no_grad():
        lp = float(mod(a, b, None, labels, s0 + h))
        lm = float(mod(a, b, None, labels, s0 - h))
    fd = (lp - lm) / (2 * h)
    assert abs(fd - ds) <= 2e-2 * abs(ds) + 1e-4
    with torch.no_grad():
        assert abs(float(mod(b, a, None, labels, s0)) - loss) <= 1e-5 * abs(loss)
    _, grads2, ds2 = _run(mod, [a, b, None], labels, torch.tensor(s0, device=dev), grad_mult=1024.0)
    assert _rel(grads2[0], grads[0] * 1024.0) < 1e-2
    assert abs(ds2 - 1024.0 * ds) <= 1e-5 * abs(ds2)


@pytest.mark.parametrize("N,nmod,labels_kind", [(20480, 2, "multi"), (6144 + 77, 3, "zipf")])
def test_shared_s_backward_matches_two_sweeps(monkeypatch, N, nmod, labels_kind):
    """Single-GPU backward (S computed once per pair: row sweep + TMA-stored coefficient strip + gradient GEMM,
    loss_api.cu backward_shared_s) against the two-sweep backward every rank of a sharded job runs, on the same
    inputs: identical loss, gradients within the 16-bit operand tolerance.  N = 20480 with a small strip budget
    needs two coefficient strips (18944 + 1536 columns); N = 6221 is ragged in every tile dimension."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    feats, labels = _synthetic(N, 768, nmod, labels_kind, seed=31, dtype=torch.bfloat16)
    feats = [None if f is None else f.float().to(dev) for f in feats]   # fp32 leaves: no output rounding
    labels = labels.to(dev)
    scale = torch.tensor(1 / 0.07, device=dev)
    mod = cb.ContrastiveLoss(None, 1 / 0.07, tensor_core_operands="bf16")
    monkeypatch.setenv("CLIBD_GT_STRIP_MB", "64")
    shared = _run(mod, feats, labels, scale)
    monkeypatch.setenv("CLIBD_BWD_TWO_SWEEPS", "1")
    two = _run(mod, feats, labels, scale)
    assert shared[0] == two[0]
    for i in range(nmod):
        assert _rel(shared[1][i], two[1][i]) < 5e-4, i
    assert abs(shared[2] - two[2]) <= 1e-3 * abs(two[2])


def test_errors_and_edge_cases():
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    x = torch.randn(8, 16, device=dev)
    lab = torch.arange(8, device=dev)
    mod = cb.ContrastiveLoss(torch.nn.CrossEntropyLoss(), 1 / 0.07)
    with pytest.raises(ValueError, match="Too less element"):
        mod(x, None, None, lab, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        mod(x.cpu(), x.cpu(), None, lab.cpu(), 1.0)
    with pytest.raises(ValueError, match="logit_scale"):
        mod(x, x, None, lab, -1.0)
    # a TENSOR scale never leaves the device (no host read per step): an invalid one poisons the loss instead
    assert torch.isnan(mod(x, x, None, lab, torch.tensor(-1.0, device=dev)))
    # no upper limit (the reference never clamps its learnable scale, simple_clip.py:32,61)
    assert torch.isfinite(mod(x, x, None, lab, torch.tensor(100.0, device=dev)))
    assert torch.isfinite(mod(x, x, None, lab, 250.0))
    ok_t = mod(x, x, None, lab, torch.tensor(10.0, device=dev))
    ok_f = mod(x, x, None, lab, 10.0)
    assert float(ok_t) == float(ok_f)
    with pytest.raises(ZeroDivisionError):
        cb.ClipLoss(bind_to="text")(x, x, None, lab, 1.0)
    with pytest.raises(NotImplementedError):
        cb.ContrastiveLoss(torch.nn.MSELoss(), 1.0)
    # zero rows must not produce NaN (train_epoch.py:11 runs under detect_anomaly)
    z = x.clone()
    z[3] = 0
    z.requires_grad_(True)
    loss = mod(z, x, None, lab, torch.tensor(5.0, device=dev))
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(z.grad).all()
    # output_dict switch and all labels equal (c_i = N)
    out = cb.ClipLoss()(x, x.flip(0), None, torch.zeros(8, dtype=torch.int64, device=dev), 2.0, output_dict=True)
    assert set(out) == {"contrastive_loss"} and torch.isfinite(out["contrastive_loss"])


@pytest.mark.parametrize("N", [4096, 32768])
def test_benchmarked_configurations_match_the_float64_oracle(N):
    """BASELINE config 2 (N = 4096) and the benchmarked north-star configuration (N = 32768): three modalities, bf16
    values, labels ~ randint(0, N/8), the backward form bench.py times (S once per pair from N = 4096 up) -- against
    values the float64 streaming oracle produced for the SAME seeded batch (tools/synth.py, oracle/
    gen_golden_fullsize.py; 192 s of host time at N = 32768, so the oracle ran once and its results are stored):
    loss, dL/d(logit_scale) and the gradient rows of three 32-row blocks (first, middle, last) of every modality,
    all within north_star's 1e-3."""
    import clibd_b200 as cb
    from tools import synth
    g = np.load(_golden.os.path.join(_golden.GOLDEN_DIR, f"fullsize_n{N}.npz"))
    dev = torch.device("cuda:0")
    d = int(g["d"])
    # fp32 leaves holding the bf16 values: the gradients come back unrounded
    feats = [synth.feature_rows(N, d, m, 0, N).float().to(dev) for m in range(3)]
    labels = synth.labels_all(N).to(dev)
    scale = torch.tensor(float(g["logit_scale"]), device=dev)
    mod = cb.ContrastiveLoss(None, 1 / 0.07, tensor_core_operands="bf16")
    loss, grads, ds = _run(mod, feats, labels, scale)
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert abs(ds - float(g["dlogit_scale"])) <= 1e-3 * abs(float(g["dlogit_scale"]))
    rows = int(g["rows"])
    for m, name in enumerate(_golden.MODS):
        ref = g[f"grad_{name}"]
        got = np.stack([grads[m][b:b + rows] for b in g["row_starts"]])
        assert _rel(got, ref) < 1e-3, name
        # the whole gradient has the oracle's norm (a wrong block elsewhere would show here)
        assert abs(np.linalg.norm(grads[m].astype(np.float64)) - float(g[f"gradnorm_{name}"])) <= 1e-3 * float(g[f"gradnorm_{name}"])


def test_default_path_for_float32_inputs():
    """float32 embeddings (what the reference's bf16 autocast block hands to the loss: F.normalize returns float32):
    up to a global batch of 1024 the exact CUDA-core path (parity 1e-5, BASELINE config 1), beyond it the tensor
    cores with float16 operands -- loss within 1e-5, gradients within 2e-4 of the float64 oracle; `tensor_core_operands`
    overrides either way (clibd_b200/loss.py:_select_path)."""
    import clibd_b200 as cb
    from clibd_b200 import _lib
    from clibd_b200.loss import _select_path
    assert _select_path(torch.float32, None, 1024, 768) == _lib.PATH_SIMT_F32
    assert _select_path(torch.float32, None, 1025, 768) == _lib.PATH_TC_F16
    assert _select_path(torch.float32, "fp32", 1 << 20, 768) == _lib.PATH_SIMT_F32
    assert _select_path(torch.bfloat16, None, 16, 768) == _lib.PATH_TC_BF16
    dev = torch.device("cuda:0")
    feats, labels = _synthetic(1536, 768, 3, "multi", seed=17, dtype=torch.float32)
    scale = torch.tensor(1 / 0.07)
    ref = lo.contrastive_loss([f.numpy() for f in feats], labels.numpy(), float(scale))
    loss, grads, ds = _run(cb.ContrastiveLoss(None, 1 / 0.07), [f.to(dev) for f in feats], labels.to(dev), scale.to(dev))
    assert abs(loss - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    for i in range(3):
        assert _rel(grads[i], ref["grads"][i]) < 2e-4, i
    assert abs(ds - ref["dlogit_scale"]) <= 2e-4 * abs(ref["dlogit_scale"])


def test_side_stream_overlap_is_bit_identical_single_gpu(monkeypatch):
    """One GPU, S once per pair (N = 4096): the step with the staging work on the side stream equals the in-line step
    bit for bit (loss, gradients, dlogit_scale), run after run."""
    import clibd_b200 as cb
    dev = torch.device("cuda:0")
    feats, labels = _synthetic(4096, 768, 3, "multi", seed=5, dtype=torch.bfloat16)
    feats = [f.to(dev) for f in feats]
    labels = labels.to(dev)
    scale = torch.tensor(1 / 0.07, device=dev)
    mod = cb.ContrastiveLoss(None, 1 / 0.07)
    monkeypatch.setenv("CLIBD_SIDE_STREAM", "0")
    ref = _run(mod, feats, labels, scale)
    monkeypatch.setenv("CLIBD_SIDE_STREAM", "1")
    for rep in range(4):
        got = _run(mod, feats, labels, scale)
        assert got[0] == ref[0] and got[2] == ref[2]
        for a, b in zip(got[1], ref[1]):
            assert np.array_equal(a, b)
