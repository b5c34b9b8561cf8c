"""The kNN / accuracy oracle: C vs numpy, tie-break, and the accuracy functions vs the
reference's own python functions (golden JSON written by oracle/gen_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import knn_oracle as ko
from tests._golden import GOLDEN_DIR


def _data(Q, K, D, seed, dup=0):
    rng = np.random.default_rng(seed)
    keys = rng.standard_normal((K, D))
    q = rng.standard_normal((Q, D))
    if dup:
        src = rng.integers(0, K, dup)
        dst = rng.integers(0, K, dup)
        keys[dst] = keys[src]
    return q, keys


def test_c_matches_numpy_and_breaks_ties_by_lowest_index():
    q, keys = _data(37, 501, 48, seed=0, dup=60)
    q[5] = keys[17]  # a query identical to a (possibly duplicated) key
    q32, k32 = ko.normalize_rows(q), ko.normalize_rows(keys)
    s_c, i_c, s64_c = ko.search(q32, k32, 7, use_c=True)
    s_n, i_n, s64_n = ko.search(q32, k32, 7, use_c=False)
    assert np.array_equal(i_c, i_n)
    assert np.array_equal(s64_c, s64_n)  # same sequential float64 summation, bit for bit
    assert s_c.dtype == np.float32 and i_c.dtype == np.int64
    assert np.all(np.diff(s64_c, axis=1) <= 0)
    full = ko.exact_sims(q32, k32)
    for r in range(q.shape[0]):
        for a in range(6):
            if s64_c[r, a] == s64_c[r, a + 1]:
                assert i_c[r, a] < i_c[r, a + 1]
        assert set(np.flatnonzero(full[r] > s64_c[r, -1])) <= set(i_c[r])


def test_normalize_matches_sklearn():
    from sklearn.preprocessing import normalize
    x = np.random.default_rng(1).standard_normal((50, 33))
    x[3] = 0.0
    assert np.array_equal(ko.normalize_rows(x), normalize(x, norm="l2", axis=1).astype(np.float32))


def test_make_prediction_return_convention():
    q, keys = _data(5, 40, 16, seed=2)
    labels = [{l: f"{l}{i % 7}" for l in ko.LEVELS} for i in range(40)]
    p = ko.make_prediction(q, keys, labels, max_k=3)
    assert isinstance(p, list) and set(p[0]) == set(ko.LEVELS) and len(p[0]["order"]) == 3
    p2, sims = ko.make_prediction(q, keys, labels, with_similarity=True, max_k=3)
    p3, sims3, idx3 = ko.make_prediction(q, keys, labels, with_similarity=True, with_indices=True, max_k=3)
    assert sims.shape == (5, 3) and idx3.shape == (5, 3) and p2 == p3 == p
    assert p[2]["genus"] == [labels[i]["genus"] for i in idx3[2]]


def test_accuracy_matches_reference_functions():
    blob = json.load(open(os.path.join(GOLDEN_DIR, "accuracy_ref_q200_k5.json")))
    pred, gt, k_list = blob["pred_list"], blob["gt_list"], blob["k_list"]
    micro = ko.micro_accuracy_ref_style(pred, gt, k_list)
    macro, per_class = ko.macro_accuracy_ref_style(pred, gt, k_list)
    for k in k_list:
        for l in ko.LEVELS:
            assert micro[k][l] == blob["micro"][str(k)][l]
            assert macro[k][l] == blob["macro"][str(k)][l]
            assert per_class[k][l] == blob["per_class"][str(k)][l]
    # integer-id forms give the same numbers
    vocab = {l: {} for l in ko.LEVELS}
    def ids(d):
        return [vocab[l].setdefault(d[l], len(vocab[l])) for l in ko.LEVELS]
    qid = np.array([ids(g) for g in gt])
    # build a synthetic key table so that idx rows reproduce pred
    key_rows, idx = [], []
    for p in pred:
        row = []
        for j in range(5):
            key_rows.append(ids({l: p[l][j] for l in ko.LEVELS}))
            row.append(len(key_rows) - 1)
        idx.append(row)
    kid, idx = np.array(key_rows), np.array(idx)
    mi = ko.micro_accuracy_ids(idx, kid, qid, k_list)
    ma = ko.macro_accuracy_ids(idx, kid, qid, k_list)
    for a, k in enumerate(k_list):
        for b, l in enumerate(ko.LEVELS):
            assert mi[a, b] == blob["micro"][str(k)][l]
            assert ma[a, b] == blob["macro"][str(k)][l]


def test_derived_feature_types():
    rng = np.random.default_rng(3)
    img, dna, txt = (rng.standard_normal((6, 4)) for _ in range(3))
    out = ko.derived_feature_types(img, dna, txt, for_key_set=True, labels=list(range(6)))
    assert np.array_equal(out["averaged_feature"], (img + dna) / 2)
    assert out["concatenated_feature"].shape == (6, 8)
    assert out["all_key_features"].shape == (18, 4) and out["all_key_features_label"] == list(range(6)) * 3
