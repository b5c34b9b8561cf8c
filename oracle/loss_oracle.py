"""CPU oracle (numpy) for the CLIBD contrastive loss -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/bioscanclip/model/loss_func.py:
  * construct_label_metrix            loss_func.py:19-22   -> label_matrix()
  * ContrastiveLoss.forward           loss_func.py:41-69   -> contrastive_loss()
  * ClipLoss.forward (pair filters)   loss_func.py:159-200 -> ordered_pairs(), clip_loss_rank()
  * gather_features (gather_with_grad) loss_func.py:96-97  -> clip_loss_rank() (W x local-slice grads)
  * nn.CrossEntropyLoss() with float [N,N] targets (train_cl.py:260,262)
                                                           -> soft_target_ce()

and /root/reference/bioscanclip/util/simclr.py:
  * SimCLR.info_nce_loss + criterion  simclr.py:64-92,119  -> info_nce()

Parity PINNED: tests/test_oracle_loss.py checks every function here against
tests/golden/loss_*.npz, which oracle/gen_golden.py produced by running the
reference's own PyTorch code (imported from /root/reference) on seeded inputs.

Never imported by the product package (clibd_b200).
"""
from __future__ import annotations

import numpy as np

MODALITY_INDEX = {"image": 0, "dna": 1, "text": 2}  # loss_func.py:166-173


def l2_normalize(x: np.ndarray, eps: float = 1e-12):
    """F.normalize(x, p=2, dim=1): x / max(||x||_2, eps) (loss_func.py:55-56,186-187)."""
    nrm = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(x.dtype)
    nrm = np.maximum(nrm, np.asarray(eps, dtype=x.dtype))
    return x / nrm, nrm


def label_matrix(labels: np.ndarray, dtype=np.float32) -> np.ndarray:
    """construct_label_metrix (loss_func.py:19-22): T[i,j] = float(labels[i] == labels[j])."""
    labels = np.asarray(labels)
    return (labels[None, :] == labels[:, None]).astype(dtype)


def soft_target_ce(S: np.ndarray, T: np.ndarray) -> float:
    """nn.CrossEntropyLoss()(S, T) with float class-probability targets (loss_func.py:65-66):
    mean_i( -sum_j T_ij * log_softmax(S_i)_j ).  Rows of T are NOT normalised."""
    m = S.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(S - m).sum(axis=1, keepdims=True))
    return float((-(T * (S - lse)).sum(axis=1)).mean())


def ordered_pairs(n_present: int, bind_to: str | None = None, no_image_text_loss: bool = False):
    """The (idx_a, idx_b) visited by the pair loop over the None-filtered feature list
    (loss_func.py:50-53 for ContrastiveLoss; :176-184 for ClipLoss incl. both filters).
    NOTE the reference applies bind_to / no_image_text indices to the FILTERED list."""
    bind_to_idx = MODALITY_INDEX.get(bind_to) if bind_to is not None else None
    out = []
    for a in range(n_present):
        for b in range(n_present):
            if bind_to_idx is not None and a != bind_to_idx and b != bind_to_idx:
                continue
            if a == b:
                continue
            if no_image_text_loss and (a == 0 or b == 0) and (a == 2 or b == 2):
                continue
            out.append((a, b))
    return out


def contrastive_loss(features, labels, logit_scale, bind_to=None, no_image_text_loss=False,
                     dtype=np.float64, want_grad=True, grad_out=1.0):
    """Dense restatement of ContrastiveLoss.forward / the ClipLoss pair loop on one process.

    features: sequence of 3 entries (image, dna, text), each [N,d] array or None.
    Returns dict(loss, grads=[3 entries or None], dlogit_scale).
    Every ordered pair contributes CE(S_ab,T) and CE(S_ba,T); the result is their mean
    (loss_func.py:50-69)."""
    present = [i for i, f in enumerate(features) if f is not None]
    feats = [np.asarray(features[i], dtype=dtype) for i in present]
    if len(feats) < 2:
        raise ValueError("Too less element for calculating the contrastive loss.")  # loss_func.py:46-47
    N = feats[0].shape[0]
    T = label_matrix(labels, dtype)
    s = dtype(logit_scale)
    pairs = ordered_pairs(len(feats), bind_to, no_image_text_loss)
    n_terms = 2 * len(pairs)
    if n_terms == 0:
        raise ZeroDivisionError("no modality pair passes the filters")  # reference: sum([])/0
    normed = [l2_normalize(f) for f in feats]
    xh = [n[0] for n in normed]
    total = 0.0
    dxh = [np.zeros_like(f) for f in feats]
    ds = 0.0
    w = grad_out / n_terms
    c = T.sum(axis=1, keepdims=True)

    def one_direction(A, B):
        cos = A @ B.T
        S = s * cos
        m = S.max(axis=1, keepdims=True)
        E = np.exp(S - m)
        Z = E.sum(axis=1, keepdims=True)
        lse = m + np.log(Z)
        loss = float((-(T * (S - lse)).sum(axis=1)).mean())
        dS = (c * (E / Z) - T) * (w / N) if want_grad else None
        return loss, dS, cos

    for (a, b) in pairs:
        l_ab, dS_ab, cos_ab = one_direction(xh[a], xh[b])
        l_ba, dS_ba, cos_ba = one_direction(xh[b], xh[a])
        total += l_ab + l_ba
        if want_grad:
            dxh[a] += s * (dS_ab @ xh[b]) + s * (dS_ba.T @ xh[b])
            dxh[b] += s * (dS_ab.T @ xh[a]) + s * (dS_ba @ xh[a])
            ds += float((dS_ab * cos_ab).sum() + (dS_ba * cos_ba).sum())
    out = {"loss": total / n_terms, "grads": [None, None, None], "dlogit_scale": None}
    if want_grad:
        in_a_pair = {k for ab in pairs for k in ab}
        for k, i in enumerate(present):
            if k not in in_a_pair:
                continue  # a modality no pair uses never enters the reference's graph: its leaf keeps grad None
            x_hat, nrm = normed[k]
            dot = (x_hat * dxh[k]).sum(axis=1, keepdims=True)
            out["grads"][i] = (dxh[k] - x_hat * dot) / nrm
        out["dlogit_scale"] = ds
    return out


def contrastive_loss_streaming(features, labels, logit_scale, bind_to=None, no_image_text_loss=False,
                               block=2048, want_grad=True, grad_out=1.0):
    """Row-blocked float64 restatement of the same math for N too large to hold [N,N]
    (the reference needs >= 12 live [N,N] fp32 matrices at N=32k).  Uses
    CE(S,T) = (1/N) sum_i [ c_i*LSE_j(S_ij) - sum_j T_ij S_ij ] with
    the fixed shift exp(S - s), valid because |S_ij| <= s for unit vectors."""
    present = [i for i, f in enumerate(features) if f is not None]
    feats = [np.asarray(features[i], dtype=np.float64) for i in present]
    if len(feats) < 2:
        raise ValueError("Too less element for calculating the contrastive loss.")
    N = feats[0].shape[0]
    labels = np.asarray(labels)
    s = float(logit_scale)
    pairs = ordered_pairs(len(feats), bind_to, no_image_text_loss)
    n_terms = 2 * len(pairs)
    unordered = sorted({(min(a, b), max(a, b)) for a, b in pairs})
    mult = {u: sum(1 for p in pairs if (min(p), max(p)) == u) for u in unordered}  # 2 normally
    normed = [l2_normalize(f) for f in feats]
    xh = [n[0] for n in normed]
    _, inv, cnt = np.unique(labels, return_inverse=True, return_counts=True)
    c = cnt[inv].astype(np.float64)
    total = 0.0
    dxh = [np.zeros_like(f) for f in feats]
    ds = 0.0
    for (a, b) in unordered:
        A, B = xh[a], xh[b]
        rs = np.zeros(N)
        cs = np.zeros(N)
        pos = 0.0
        for i0 in range(0, N, block):
            i1 = min(N, i0 + block)
            S = s * (A[i0:i1] @ B.T)
            E = np.exp(S - s)
            rs[i0:i1] = E.sum(axis=1)
            cs += E.sum(axis=0)
            Tb = labels[i0:i1, None] == labels[None, :]
            pos += float(S[Tb].sum())
        lse_r = s + np.log(rs)
        lse_c = s + np.log(cs)
        # CE(S_ab) + CE(S_ba), each counted mult[u] times in the reference's list
        total += mult[(a, b)] * (float((c * lse_r).sum() + (c * lse_c).sum()) - 2.0 * pos) / N
        if want_grad:
            u = c / rs
            v = c / cs
            kap = mult[(a, b)] * grad_out / (n_terms * N)
            for i0 in range(0, N, block):
                i1 = min(N, i0 + block)
                cos = A[i0:i1] @ B.T
                E = np.exp(s * cos - s)
                G = E * (u[i0:i1, None] + v[None, :])
                G -= 2.0 * (labels[i0:i1, None] == labels[None, :])
                G *= kap
                dxh[a][i0:i1] += s * (G @ B)
                dxh[b] += s * (G.T @ A[i0:i1])
                ds += float((G * cos).sum())
    out = {"loss": total / n_terms, "grads": [None, None, None], "dlogit_scale": None}
    if want_grad:
        in_a_pair = {k for ab in pairs for k in ab}
        for k, i in enumerate(present):
            if k not in in_a_pair:
                continue  # as in contrastive_loss: no pair, no gradient
            x_hat, nrm = normed[k]
            dot = (x_hat * dxh[k]).sum(axis=1, keepdims=True)
            out["grads"][i] = (dxh[k] - x_hat * dot) / nrm
        out["dlogit_scale"] = ds
    return out


def clip_loss_rank(features_per_rank, labels_per_rank, rank, logit_scale, bind_to=None,
                   no_image_text_loss=False, dtype=np.float64, grad_out_per_rank=None):
    """What rank `rank` observes from ClipLoss(gather_with_grad=True) under W processes
    (loss_func.py:138-201 + torch.distributed.nn.all_gather, whose backward is a
    reduce-scatter(SUM)): the loss is the full-batch loss; the gradient w.r.t. the local
    features is sum_r grad_out_r * dL/dx_local (= W x the local slice of the full-batch
    gradient when every rank back-propagates 1.0); d(logit_scale) is grad_out_rank * dL/ds
    (no cross-rank sum: each rank differentiates its own replica of the scalar)."""
    W = len(features_per_rank)
    if grad_out_per_rank is None:
        grad_out_per_rank = [1.0] * W
    full = []
    for m in range(3):
        parts = [features_per_rank[r][m] for r in range(W)]
        full.append(None if parts[0] is None else np.concatenate(parts, axis=0))
    labels = np.concatenate(labels_per_rank, axis=0)
    res = contrastive_loss(full, labels, logit_scale, bind_to, no_image_text_loss, dtype=dtype)
    n = [len(l) for l in labels_per_rank]
    lo = sum(n[:rank])
    hi = lo + n[rank]
    gsum = float(sum(grad_out_per_rank))
    grads = [None if g is None else gsum * g[lo:hi] for g in res["grads"]]
    return {"loss": res["loss"], "grads": grads,
            "dlogit_scale": grad_out_per_rank[rank] * res["dlogit_scale"]}


def row_block_fwd_bwd(features, labels, logit_scale, r0, r1, dtype=np.float32):
    """One bounded SAMPLE of the full-batch step, used as the CPU baseline by bench.py: forward
    statistics and backward products of rows [r0, r1) against all N columns for every unordered
    modality pair (same math as contrastive_loss_streaming, one row block).  The full step costs
    N / (r1 - r0) such blocks.  Returns the block's partial loss terms so the work cannot be elided."""
    feats = [np.asarray(f, dtype=dtype) for f in features if f is not None]
    N = feats[0].shape[0]
    labels = np.asarray(labels)
    s = dtype(logit_scale)
    xh = [l2_normalize(f)[0] for f in feats]
    _, inv, cnt = np.unique(labels, return_inverse=True, return_counts=True)
    c = cnt[inv].astype(dtype)
    acc = 0.0
    for a in range(len(xh)):
        for b in range(a + 1, len(xh)):
            A, B = xh[a][r0:r1], xh[b]
            cos = A @ B.T
            E = np.exp(s * cos - s)
            rs = E.sum(axis=1)
            cs = E.sum(axis=0)  # partial column sums of this block
            Tb = labels[r0:r1, None] == labels[None, :]
            acc += float((c[r0:r1] * np.log(rs)).sum()) - float(cos[Tb].sum())
            G = E * ((c[r0:r1] / rs)[:, None] + (c / np.maximum(cs * (N / (r1 - r0)), 1e-30))[None, :])
            G -= 2.0 * Tb
            dA = G @ B
            dB = G.T @ A
            acc += float(dA[0, 0]) + float(dB[0, 0])
    return acc


def info_nce(features, batch_size, n_views=2, temperature=0.07, dtype=np.float64):
    """SimCLR.info_nce_loss followed by nn.CrossEntropyLoss()(logits, zeros) (simclr.py:64-92, 119), restated
    step by step on the materialised matrices, plus the analytic gradient w.r.t. the un-normalised features.

    Returns {"loss", "grad", "logits"}: logits [M, M-1] are the reference's re-ordered logits (positives
    first, then the negatives in column order), before the division by the temperature is undone."""
    x = np.asarray(features, dtype=dtype)
    M = x.shape[0]
    ids = np.concatenate([np.arange(batch_size) for _ in range(n_views)])           # simclr.py:66-67
    lab = (ids[None, :] == ids[:, None])                                             # :69
    xh, nrm = l2_normalize(x)                                                        # :72
    sim = xh @ xh.T                                                                  # :74
    off = ~np.eye(M, dtype=bool)                                                     # :77
    lab_o = lab[off].reshape(M, M - 1)                                               # :78
    sim_o = sim[off].reshape(M, M - 1)                                               # :79
    pos = sim_o[lab_o].reshape(M, -1)                                                # :83
    neg = sim_o[~lab_o].reshape(M, -1)                                               # :86
    logits = np.concatenate([pos, neg], axis=1) / temperature                        # :88,91
    m = logits.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(logits - m).sum(axis=1, keepdims=True))
    loss = float((lse[:, 0] - logits[:, 0]).mean())                                  # CE with target class 0
    # gradient: dL/dlogits = (softmax - onehot_0) / M, scattered back to the [M, M] similarity matrix
    p = np.exp(logits - lse)
    p[:, 0] -= 1.0
    p /= M
    g_pos, g_neg = p[:, :pos.shape[1]], p[:, pos.shape[1]:]
    g_o = np.zeros((M, M - 1), dtype=dtype)
    g_o[lab_o] = g_pos.reshape(-1)
    g_o[~lab_o] = g_neg.reshape(-1)
    G = np.zeros((M, M), dtype=dtype)
    G[off] = g_o.reshape(-1)
    G /= temperature
    dxh = G @ xh + G.T @ xh                                                          # sim = xh xh^T
    grad = (dxh - xh * (xh * dxh).sum(axis=1, keepdims=True)) / nrm                  # F.normalize backward
    return {"loss": loss, "grad": grad, "logits": logits}
