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
