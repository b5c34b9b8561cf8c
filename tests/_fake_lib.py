"""A numpy test double of the loss entry points of include/clibd_b200.h -- TEST INFRASTRUCTURE.

It lets the host-side orchestration of clibd_b200/loss.py (all-gather of features / inverse
norms / labels, all-reduce of statistics, the sum-over-ranks gradient convention) run on a CPU
box under the gloo backend.  The arithmetic is the oracle's (oracle/loss_oracle.py), split at the
same phase boundaries as the C ABI.  Injected with clibd_b200._lib.inject_for_tests().
"""
import ctypes

import numpy as np

_NP = {0: np.float32}
PAIRS = [(0, 1), (0, 2), (1, 2)]


def _arr(ptr, shape, dtype):
    if isinstance(ptr, ctypes.c_void_p):
        ptr = ptr.value
    n = int(np.prod(shape))
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class FakeLib:
    def __init__(self):
        self.state = {}
        self.err = ""

    def clibd_last_error(self):
        return self.err

    def clibd_row_inv_norm(self, x, dtype, n, d, out, stream):
        xa = _arr(x, (n, d), _NP[dtype])
        nrm = np.sqrt((xa.astype(np.float64) ** 2).sum(1))
        _arr(out, (n,), np.float32)[:] = 1.0 / np.maximum(nrm, 1e-12)
        return 0

    def clibd_loss_scratch_bytes(self, N, n, d, path, mode):
        return 64

    def _inputs(self, xs, ivs, dtype, N, d):
        xh = []
        for m in range(3):
            if not xs[m]:
                xh.append(None)
                continue
            x = _arr(xs[m], (N, d), _NP[dtype]).astype(np.float64)
            iv = _arr(ivs[m], (N,), np.float32).astype(np.float64)
            xh.append(x * iv[:, None])
        return xh

    def clibd_loss_forward_stats(self, xs, dtype, ivs, labels, N, d, row0, n, scale, scale_dev, w, path, mode, scratch,
                                 nbytes, rowsum, colsum, posrow, pos, stream):
        if scale_dev:
            scale = float(_arr(scale_dev, (1,), np.float32)[0])
        xh = self._inputs(xs, ivs, dtype, N, d)
        lab = _arr(labels, (N,), np.int64)
        rs = _arr(rowsum, (3, N), np.float32)
        cs = _arr(colsum, (3, N), np.float32)
        ps = _arr(pos, (3,), np.float64)
        T = lab[row0:row0 + n, None] == lab[None, :]
        for p, (a, b) in enumerate(PAIRS):
            if w[p] == 0.0:
                ps[p] = 0.0
                continue
            cos = xh[a][row0:row0 + n] @ xh[b].T
            E = np.exp(scale * cos - scale)
            rs[p, row0:row0 + n] = E.sum(1)
            cs[p, :] = E.sum(0)
            ps[p] = cos[T].sum()
            if mode == 1 and n < N:  # exchange mode: per-row positive dot products of the local rows
                _arr(posrow, (3, N), np.float32)[p, row0:row0 + n] = (cos * T).sum(1)
        self.state[scratch] = {"labels": lab.copy(), "scale": scale}
        return 0

    def clibd_loss_forward_finish(self, N, n, d, scale, w, path, mode, scratch, nbytes, rowsum, colsum, pos, loss_out,
                                  stream):
        st = self.state[scratch]
        scale = st["scale"]  # the one forward_stats stored
        lab = st["labels"]
        _, inv, cnt = np.unique(lab, return_inverse=True, return_counts=True)
        c = cnt[inv].astype(np.float64)
        rs = _arr(rowsum, (3, N), np.float32).astype(np.float64)
        cs = _arr(colsum, (3, N), np.float32).astype(np.float64)
        ps = _arr(pos, (3,), np.float64)
        total = 0.0
        st["u"], st["v"], st["c"] = {}, {}, c
        for p in range(3):
            if w[p] == 0.0:
                continue
            total += w[p] * ((c * (2 * scale + np.log(rs[p]) + np.log(cs[p]))).sum() - 2 * scale * ps[p])
            st["u"][p], st["v"][p] = c / rs[p], c / cs[p]
        _arr(loss_out, (1,), np.float32)[0] = total / N
        return 0

    def clibd_loss_backward(self, xs, dtype, ivs, N, d, row0, n, scale, w, path, scratch, nbytes, gscale, gscale_dev, dxs,
                            dscale, stream):
        if gscale_dev:
            gscale = gscale * float(_arr(gscale_dev, (1,), np.float32)[0])
        st = self.state[scratch]
        scale = st["scale"]
        xh = self._inputs(xs, ivs, dtype, N, d)
        lab = st["labels"]
        T = (lab[row0:row0 + n, None] == lab[None, :]).astype(np.float64)
        dots = 0.0
        for m in range(3):
            if xh[m] is None:
                continue
            acc = np.zeros((n, d))
            used = False
            for p, (a, b) in enumerate(PAIRS):
                if w[p] == 0.0 or m not in (a, b):
                    continue
                used = True
                other = b if m == a else a
                rc, cc = (st["u"][p], st["v"][p]) if m == a else (st["v"][p], st["u"][p])
                cos = xh[m][row0:row0 + n] @ xh[other].T
                G = np.exp(scale * cos - scale) * (rc[row0:row0 + n, None] + cc[None, :]) - 2 * T
                acc += w[p] * (G @ xh[other])
            if not used:
                continue
            dxh = (scale / N) * acc
            xl = xh[m][row0:row0 + n]
            dot = (xl * dxh).sum(1, keepdims=True)
            dots += dot.sum()
            if dxs[m]:
                iv = _arr(ivs[m], (N,), np.float32).astype(np.float64)[row0:row0 + n, None]
                _arr(dxs[m], (n, d), _NP[dtype])[:] = gscale * (dxh - xl * dot) * iv
        _arr(dscale, (1,), np.float64)[0] = dots / (2 * scale)
        return 0

    # ---- exchange mode (row-sharded, S once per pair): sweeps produce the row-side gradient (kept) and this rank's
    # partial column-side gradient of ALL rows (part[b], reduce-scattered by the caller), finish combines them
    def clibd_loss_backward_sweeps(self, xs, dtype, ivs, N, d, row0, n, scale, w, path, scratch, nbytes, posrow, part,
                                   peer_red, rank, world, stream):
        assert not peer_red, "the test double has no peer memory"
        st = self.state[scratch]
        scale = st["scale"]
        xh = self._inputs(xs, ivs, dtype, N, d)
        lab = st["labels"]
        T = (lab[row0:row0 + n, None] == lab[None, :]).astype(np.float64)
        # the exchanged posrow must be complete (every rank's rows) by now: check it against a direct evaluation
        pr = _arr(posrow, (3, N), np.float32)
        st["row"] = {}
        wrote = set()
        for p, (a, b) in enumerate(PAIRS):
            if w[p] == 0.0:
                continue
            full_T = (lab[:, None] == lab[None, :])
            expect = ((xh[a] @ xh[b].T) * full_T).sum(1)
            assert np.allclose(pr[p], expect, rtol=1e-4, atol=1e-5), "posrow was not exchanged completely"
            xa = xh[a][row0:row0 + n]
            cos = xa @ xh[b].T
            G = np.exp(scale * cos - scale) * (st["u"][p][row0:row0 + n, None] + st["v"][p][None, :]) - 2 * T
            st["row"][a] = st["row"].get(a, 0.0) + w[p] * (G @ xh[b])
            contrib = (w[p] * (G.T @ xa)).astype(np.float32)
            out = _arr(part[b], (N, d), np.float32)
            if b in wrote:
                out += contrib
            else:
                out[:] = contrib
                wrote.add(b)
        return 0

    def clibd_loss_backward_finish(self, xs, dtype, ivs, N, d, row0, n, scale, w, path, scratch, nbytes, reduced,
                                   reduced_slots, gscale, gscale_dev, gcount, dxs, dscale, stream):
        if gscale_dev:
            gscale = gscale * float(_arr(gscale_dev, (gcount,), np.float32).sum())
        st = self.state[scratch]
        scale = st["scale"]
        xh = self._inputs(xs, ivs, dtype, N, d)
        dots = 0.0
        for m in range(3):
            if xh[m] is None:
                continue
            used = any(w[p] != 0.0 and m in PAIRS[p] for p in range(3))
            if not used:
                continue
            acc = np.zeros((n, d)) + st["row"].get(m, 0.0)
            if reduced[m]:
                acc += _arr(reduced[m], (reduced_slots[m], n, d), np.float32).astype(np.float64).sum(0)
            dxh = (scale / N) * acc
            xl = xh[m][row0:row0 + n]
            dot = (xl * dxh).sum(1, keepdims=True)
            dots += dot.sum()
            if dxs[m]:
                iv = _arr(ivs[m], (N,), np.float32).astype(np.float64)[row0:row0 + n, None]
                _arr(dxs[m], (n, d), _NP[dtype])[:] = gscale * (dxh - xl * dot) * iv
        _arr(dscale, (1,), np.float64)[0] = dots / (2 * scale)
        return 0
