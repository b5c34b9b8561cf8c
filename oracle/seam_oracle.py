"""CPU oracle (numpy) for the seams next to the hot path -- TEST INFRASTRUCTURE ONLY.

* ``softmax_mean`` / ``softmax_mean_backward``: the BarcodeBERT head
  ``logits.softmax(dim=-1).mean(dim=1)`` (/root/reference/bioscanclip/model/dna_encoder.py:137) and
  its analytic gradient.
* ``f_normalize_rows``: ``F.normalize(x, dim=-1)`` as applied per batch by
  /root/reference/bioscanclip/epoch/inference_epoch.py:96-101, then widened to float64 like
  ``np.array(list_of_python_floats)`` (:108-119).
* ``derived_feature_types`` lives in knn_oracle.py (util.py:711-737).

Parity PINNED: tests/golden/seam_*.npz were produced by oracle/gen_golden.py executing the reference's own
``CLIBDDNAEncoder.forward`` / ``get_features_and_label`` source (extracted with ast) and torch autograd.
"""
import numpy as np


def softmax_mean(logits: np.ndarray) -> np.ndarray:
    """[n, T, C] -> [n, C], float64 arithmetic."""
    x = logits.astype(np.float64)
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    p = e / e.sum(axis=-1, keepdims=True)
    return p.mean(axis=1)


def softmax_mean_backward(logits: np.ndarray, grad_out: np.ndarray) -> np.ndarray:
    """d/dlogits of sum(softmax_mean(logits) * grad_out): p/T * (g - sum_c g p)."""
    x = logits.astype(np.float64)
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    p = e / e.sum(axis=-1, keepdims=True)
    g = grad_out.astype(np.float64)[:, None, :]
    gbar = (p * g).sum(axis=-1, keepdims=True)
    return p * (g - gbar) / logits.shape[1]


def f_normalize_rows(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """F.normalize(x, p=2, dim=-1) in float32, returned as the float64 array the reference ends up with."""
    x32 = x.astype(np.float32)
    nrm = np.sqrt((x32.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    out = x32 / np.maximum(nrm, np.float32(eps))[:, None]
    return out.astype(np.float64)
