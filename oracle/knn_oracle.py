"""CPU oracle (numpy + C) for CLIBD's cosine kNN retrieval and top-k accuracy --
TEST INFRASTRUCTURE ONLY.

Restates /root/reference/bioscanclip/util/util.py:
  * make_prediction            util.py:521-553  -> normalize_rows(), search(), make_prediction()
  * find_closest_match         util.py:759-789  -> same search, dict output
  * top_k_micro_accuracy       util.py:379-395  -> micro_accuracy_ref_style(), micro_accuracy_ids()
  * top_k_macro_accuracy       util.py:555-599  -> macro_accuracy_ref_style(), macro_accuracy_ids()
  * get_features_and_label     util.py:702-742  -> derived_feature_types()

The search itself is third-party: faiss ``IndexFlatIP`` (requirements.txt:22 pins
faiss-gpu==1.7.2, CPU index used at util.py:522-528) after
``sklearn.preprocessing.normalize(norm="l2", axis=1).astype(np.float32)``
(util.py:523-524).  faiss is not in /root/reference nor in this image and the
reference has no tests for it: PARITY UNPINNED for the search.  Published
algorithm restated here: exhaustive inner product of the float32 rows, top-k by
descending similarity.  faiss's tie-break is implementation-defined; BASELINE.json's
north_star fixes LOWEST INDEX WINS, so the order key is (-sim, index).  To make
that order well-defined independently of BLAS summation order, the similarity of a
(query,key) pair is the float64 sum, accumulated sequentially over d = 0..D-1, of the
exact products of the float32 elements; the reported similarity is that value
rounded to float32.  The accuracy functions ARE pinned against the reference's own
python functions (tests/golden/accuracy_*.json, see gen_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

LEVELS = ["order", "family", "genus", "species"]  # util.py:25

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force: bool = False) -> str:
    """Compile oracle/knn_oracle.c -> oracle/_build/libknn_oracle.so (gcc, no fast-math)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "libknn_oracle.so")
    src = os.path.join(_HERE, "knn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
        _LIB.knn_oracle_search.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        _LIB.knn_oracle_search.restype = ctypes.c_int
    return _LIB


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """sklearn.preprocessing.normalize(x, norm="l2", axis=1).astype(np.float32) (util.py:523-524):
    row / ||row||_2 computed in the input dtype (float64 in the reference: embeddings
    arrive via .tolist()), zero rows left as zeros, then cast to float32."""
    x = np.asarray(x)
    work = x.astype(np.float64) if x.dtype != np.float64 else x
    nrm = np.sqrt((work * work).sum(axis=1))
    nrm[nrm == 0.0] = 1.0
    return (work / nrm[:, None]).astype(np.float32)


def exact_sims(q32: np.ndarray, k32: np.ndarray) -> np.ndarray:
    """float64 sims accumulated sequentially over d (numpy; small cases only)."""
    Q, D = q32.shape
    acc = np.zeros((Q, k32.shape[0]), dtype=np.float64)
    q64 = q32.astype(np.float64)
    k64 = k32.astype(np.float64)
    for d in range(D):
        acc += q64[:, d, None] * k64[None, :, d]
    return acc


def search(q32: np.ndarray, k32: np.ndarray, k: int, use_c: bool = True):
    """IndexFlatIP(d).add(keys); search(queries, k) restated (util.py:522,525,528).
    Returns (similarities float32 [Q,k], indices int64 [Q,k], sims64 float64 [Q,k])."""
    q32 = np.ascontiguousarray(q32, dtype=np.float32)
    k32 = np.ascontiguousarray(k32, dtype=np.float32)
    Q, D = q32.shape
    K = k32.shape[0]
    if k > K:
        raise ValueError("k larger than the number of keys")
    if use_c:
        sims = np.empty((Q, k), dtype=np.float64)
        idx = np.empty((Q, k), dtype=np.int64)
        rc = _lib().knn_oracle_search(q32.ctypes.data, Q, k32.ctypes.data, K, D, k,
                                      sims.ctypes.data, idx.ctypes.data)
        if rc != 0:
            raise RuntimeError("knn_oracle_search failed")
    else:
        full = exact_sims(q32, k32)
        ar = np.arange(K)
        idx = np.empty((Q, k), dtype=np.int64)
        sims = np.empty((Q, k), dtype=np.float64)
        for i in range(Q):
            order = np.lexsort((ar, -full[i]))[:k]  # primary: -sim, secondary: index
            idx[i] = order
            sims[i] = full[i, order]
    return sims.astype(np.float32), idx, sims


def make_prediction(query_feature, keys_feature, keys_label, with_similarity=False, with_indices=False,
                    max_k=5, use_c=True):
    """make_prediction (util.py:521-553) with the reference's return convention.
    keys_label: list of {level: label} dicts (util.py:529-541)."""
    keys32 = normalize_rows(keys_feature)
    q32 = normalize_rows(query_feature)
    sims, idx, _ = search(q32, keys32, max_k, use_c=use_c)
    pred_list = []
    for key_indices in idx:
        pred = {level: [keys_label[int(i)][level] for i in key_indices] for level in LEVELS}
        pred_list.append(pred)
    out = [pred_list]
    if with_similarity:
        out.append(sims)
    if with_indices:
        out.append(idx)
    return out[0] if len(out) == 1 else out


def micro_accuracy_ref_style(pred_list, gt_list, k_list):
    """top_k_micro_accuracy (util.py:379-395), dict-of-strings form."""
    total = len(pred_list)
    acc = {}
    for k in k_list:
        acc.setdefault(k, {})
        for level in LEVELS:
            correct = 0
            for pred, gt in zip(pred_list, gt_list):
                if gt[level] in pred[level][:k]:
                    correct += 1
            acc[k][level] = correct * 1.0 / total
    return acc


def macro_accuracy_ref_style(pred_list, gt_list, k_list=None):
    """top_k_macro_accuracy (util.py:555-599): per-GT-class hit rate, unweighted mean over
    the classes present in the query ground truth; also the per-class dict."""
    if k_list is None:
        k_list = [1, 3, 5]
    macro, per_class = {}, {}
    for k in k_list:
        macro[k], per_class[k] = {}, {}
        for level in LEVELS:
            hit, cnt = {}, {}
            for pred, gt in zip(pred_list, gt_list):
                g = gt[level]
                hit.setdefault(g, 0)
                cnt.setdefault(g, 0)
                if g in pred[level][:k]:
                    hit[g] += 1
                cnt[g] += 1
            per_class[k][level] = {g: hit[g] * 1.0 / cnt[g] for g in cnt}
            total = 0  # plain left-to-right sum like util.py:586-597 (builtin sum() compensates in py3.12)
            for g in cnt:
                total = total + hit[g] * 1.0 / cnt[g]
            macro[k][level] = total / len(cnt)
    return macro, per_class


def micro_accuracy_ids(idx, key_ids, query_ids, k_list):
    """Integer-id form: idx [Q,kmax], key_ids [K,4], query_ids [Q,4] -> float64 [len(k_list),4]."""
    pred = key_ids[idx]  # [Q,kmax,4]
    out = np.zeros((len(k_list), 4))
    for a, k in enumerate(k_list):
        hit = (pred[:, :k, :] == query_ids[:, None, :]).any(axis=1)
        out[a] = hit.sum(axis=0) * 1.0 / idx.shape[0]
    return out


def macro_accuracy_ids(idx, key_ids, query_ids, k_list):
    """Integer-id form of top_k_macro_accuracy -> float64 [len(k_list),4]; class order of the
    mean follows first appearance in the query list, like the reference's dict."""
    pred = key_ids[idx]
    out = np.zeros((len(k_list), 4))
    for a, k in enumerate(k_list):
        hit = (pred[:, :k, :] == query_ids[:, None, :]).any(axis=1)
        for l in range(4):
            _, first, inv = np.unique(query_ids[:, l], return_index=True, return_inverse=True)
            cnt = np.bincount(inv)
            h = np.bincount(inv, weights=hit[:, l].astype(np.float64))
            order = np.argsort(first, kind="stable")
            total = 0.0
            for cidx in order:  # sequential sum in first-appearance order (util.py:586-597)
                total = total + h[cidx] * 1.0 / cnt[cidx]
            out[a, l] = total / len(cnt)
    return out


def derived_feature_types(image, dna, text=None, for_key_set=False, labels=None):
    """averaged / concatenated / all_key feature construction (util.py:711-737)."""
    out = {"averaged_feature": None, "concatenated_feature": None,
           "all_key_features": None, "all_key_features_label": None}
    if image is not None and dna is not None:
        out["averaged_feature"] = np.mean([image, dna], axis=0)
        out["concatenated_feature"] = np.concatenate((image, dna), axis=1)
    if for_key_set and image is not None and dna is not None and text is not None:
        out["all_key_features"] = np.concatenate((image, dna, text), axis=0)
        out["all_key_features_label"] = list(labels) * 3 if labels is not None else None
    return out
