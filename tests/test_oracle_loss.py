"""The numpy loss oracle vs the golden vectors produced by the reference's own PyTorch code."""
import numpy as np
import pytest

from oracle import loss_oracle as lo
from tests import _golden

SINGLE = _golden.all_single_process()


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64))
                 / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-300))


@pytest.mark.parametrize("g", SINGLE, ids=[g.name for g in SINGLE])
def test_dense_oracle_matches_reference(g):
    mult = g.meta.get("grad_mult", 1.0)
    res = lo.contrastive_loss(g.features, g.labels, g.logit_scale, grad_out=mult, **g.kwargs())
    assert abs(res["loss"] - float(g.outputs["loss"])) <= 2e-6 * abs(float(g.outputs["loss"]))
    for i, m in enumerate(_golden.MODS):
        if f"grad_{m}" in g.outputs:
            assert _rel(res["grads"][i], g.outputs[f"grad_{m}"]) < 2e-5, m
        else:
            assert res["grads"][i] is None
    if "dlogit_scale" in g.outputs:
        ref = float(g.outputs["dlogit_scale"])
        assert abs(res["dlogit_scale"] - ref) <= 2e-4 * abs(ref) + 1e-7


@pytest.mark.parametrize("g", SINGLE, ids=[g.name for g in SINGLE])
def test_streaming_oracle_matches_dense(g):
    dense = lo.contrastive_loss(g.features, g.labels, g.logit_scale, **g.kwargs())
    stream = lo.contrastive_loss_streaming(g.features, g.labels, g.logit_scale, block=29, **g.kwargs())
    assert abs(dense["loss"] - stream["loss"]) <= 1e-11 * abs(dense["loss"])
    for a, b in zip(dense["grads"], stream["grads"]):
        assert (a is None) == (b is None)
        if a is not None:
            assert _rel(b, a) < 1e-10
    assert abs(dense["dlogit_scale"] - stream["dlogit_scale"]) <= 1e-9 * abs(dense["dlogit_scale"]) + 1e-13


def test_too_few_modalities_raises():
    x = np.ones((4, 3))
    with pytest.raises(ValueError, match="Too less element"):
        lo.contrastive_loss([x, None, None], np.arange(4), 1.0)


def test_world2_convention_matches_reference():
    g = _golden.load("cliploss_w2_all_n64_d32")
    n = 32
    per_rank_feats = [[g.inputs[m][r * n:(r + 1) * n] for m in _golden.MODS] for r in range(2)]
    per_rank_labels = [g.labels[r * n:(r + 1) * n] for r in range(2)]
    for r in range(2):
        res = lo.clip_loss_rank(per_rank_feats, per_rank_labels, r, g.logit_scale)
        assert abs(res["loss"] - float(g.outputs[f"rank{r}_loss"])) <= 2e-6 * abs(res["loss"])
        for i, m in enumerate(_golden.MODS):
            assert _rel(res["grads"][i], g.outputs[f"rank{r}_grad_{m}"]) < 2e-5
        ref = float(g.outputs[f"rank{r}_dlogit_scale"])
        assert abs(res["dlogit_scale"] - ref) <= 2e-4 * abs(ref)
    # the two ranks see the same full-batch loss (loss_func.py:200 on gathered features)
    assert float(g.outputs["rank0_loss"]) == pytest.approx(float(g.outputs["rank1_loss"]), rel=1e-6)


def test_pair_filter_quirks():
    # bind_to / no_image_text indices refer to the None-filtered list (loss_func.py:159-184)
    assert lo.ordered_pairs(3) == [(0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1)]
    assert lo.ordered_pairs(3, bind_to="dna") == [(0, 1), (1, 0), (1, 2), (2, 1)]
    assert lo.ordered_pairs(3, no_image_text_loss=True) == [(0, 1), (1, 0), (1, 2), (2, 1)]
    assert lo.ordered_pairs(2, bind_to="text") == []  # index 2 does not exist in a 2-entry list
    assert lo.ordered_pairs(2, no_image_text_loss=True) == [(0, 1), (1, 0)]


INFONCE_GOLDENS = ("infonce_b64_d768_t0.07", "infonce_b45_d40_t0.2_views_correlated",
                   "infonce_b16_d96_t0.07_gradscale65536")


@pytest.mark.parametrize("name", INFONCE_GOLDENS)
def test_info_nce_oracle_matches_reference_golden(name):
    """oracle info_nce() vs SimCLR.info_nce_loss + CrossEntropyLoss executed from the reference source."""
    g = _golden.load(name)
    out = lo.info_nce(g.inputs["features"], g.meta["batch_size"], g.meta["n_views"], g.meta["temperature"])
    np.testing.assert_allclose(out["logits"], g.outputs["logits"], rtol=2e-5, atol=2e-5)
    assert out["loss"] == pytest.approx(float(g.outputs["loss"]), rel=2e-6)
    ref = g.outputs["grad"] / g.meta["grad_mult"]
    err = np.linalg.norm(out["grad"] - ref) / np.linalg.norm(ref)
    assert err < 2e-5, err  # the golden is torch float32 autograd


@pytest.mark.parametrize("name", INFONCE_GOLDENS)
def test_info_nce_fused_formulation_matches_oracle(name):
    """The restatement the CUDA path uses (csrc/infonce.cu): fixed-shift exponentials with the diagonal
    dropped, row sums r_i, loss = mean(s + ln r_i - s zhat_i.zhat_p(i)),
    dzhat_k = (s/M) [sum_{j != k} e_kj (1/r_k + 1/r_j) zhat_j - 2 zhat_p(k)]."""
    g = _golden.load(name)
    x = g.inputs["features"].astype(np.float64)
    B, s = g.meta["batch_size"], 1.0 / g.meta["temperature"]
    M = x.shape[0]
    xh, nrm = lo.l2_normalize(x)
    e = np.exp(s * (xh @ xh.T) - s)
    np.fill_diagonal(e, 0.0)
    r = e.sum(axis=1)
    partner = np.roll(xh, -B, axis=0)  # row k -> zhat[(k + B) mod M]
    loss = (s + np.log(r)).mean() - s * (xh * partner).sum() / M
    coef = e * (1.0 / r[:, None] + 1.0 / r[None, :])
    dxh = (s / M) * (coef @ xh - 2.0 * partner)
    grad = (dxh - xh * (xh * dxh).sum(axis=1, keepdims=True)) / nrm
    ref = lo.info_nce(x, B, 2, g.meta["temperature"])
    assert loss == pytest.approx(ref["loss"], rel=1e-12)
    assert np.linalg.norm(grad - ref["grad"]) / np.linalg.norm(ref["grad"]) < 1e-11
