"""Stage-by-stage comparison of the tcgen05 path against the fp32 CUDA-core path on the GPU
(development aid; prints relative errors, never asserts)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from clibd_b200 import _lib  # noqa: E402
from clibd_b200.loss import _DT  # noqa: E402


def stats(path, feats, labels, scale, weights):
    lib = _lib.load()
    dev = feats[0].device
    N, d = feats[0].shape
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    inv = []
    for f in feats:
        iv = torch.empty(N, dtype=torch.float32, device=dev)
        _lib.check(lib.clibd_row_inv_norm(f.data_ptr(), _DT[f.dtype], N, d, iv.data_ptr(), stream))
        inv.append(iv)
    nbytes = lib.clibd_loss_scratch_bytes(N, N, d, path)
    scratch = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    st = torch.zeros(6 * N, dtype=torch.float32, device=dev)
    pos = torch.zeros(3, dtype=torch.float64, device=dev)
    xs = _lib.ptr_array3([f.data_ptr() for f in feats] + [None] * (3 - len(feats)))
    ivs = _lib.ptr_array3([f.data_ptr() for f in inv] + [None] * (3 - len(feats)))
    w = _lib.float_array3(weights)
    _lib.check(lib.clibd_loss_forward_stats(xs, _DT[feats[0].dtype], ivs, labels.data_ptr(), N, d, 0, N, scale, w, path,
                                            scratch.data_ptr(), nbytes, st.data_ptr(), st.data_ptr() + 12 * N,
                                            pos.data_ptr(), stream))
    loss = torch.zeros((), dtype=torch.float32, device=dev)
    _lib.check(lib.clibd_loss_forward_finish(N, N, d, scale, w, path, scratch.data_ptr(), nbytes, st.data_ptr(),
                                             st.data_ptr() + 12 * N, pos.data_ptr(), loss.data_ptr(), stream))
    dxs = [torch.zeros(N, d, dtype=f.dtype, device=dev) for f in feats]
    ds = torch.zeros(1, dtype=torch.float64, device=dev)
    outs = _lib.ptr_array3([t.data_ptr() for t in dxs] + [None] * (3 - len(feats)))
    _lib.check(lib.clibd_loss_backward(xs, _DT[feats[0].dtype], ivs, N, d, 0, N, scale, w, path, scratch.data_ptr(),
                                       nbytes, 1.0, None, outs, ds.data_ptr(), stream))
    torch.cuda.synchronize()
    return st[:3 * N].clone(), st[3 * N:].clone(), pos.clone(), float(loss), dxs, float(ds)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    feats = [torch.randn(N, d, device=dev) for _ in range(2)]
    labels = torch.randint(0, max(1, N // 8), (N,), device=dev)
    weights = [0.5, 0.0, 0.0]
    scale = 1 / 0.07
    ref = stats(0, feats, labels, scale, weights)
    # torch check of the fp32 path itself
    a = torch.nn.functional.normalize(feats[0].double(), dim=1)
    b = torch.nn.functional.normalize(feats[1].double(), dim=1)
    e = torch.exp(scale * (a @ b.T) - scale)
    print(f"[fp32 path] rowsum rel {rel(ref[0][:N], e.sum(1)):.2e} colsum rel {rel(ref[1][:N], e.sum(0)):.2e} loss {ref[3]:.6f}")
    for path, name in ((1, "bf16"), (2, "fp16")):
        try:
            got = stats(path, feats, labels, scale, weights)
        except Exception as ex:  # noqa: BLE001
            print(f"[{name}] FAILED: {ex}")
            continue
        print(f"[{name}] rowsum rel {rel(got[0][:N], ref[0][:N]):.2e} colsum rel {rel(got[1][:N], ref[1][:N]):.2e} "
              f"pos {float(got[2][0]):.6f}/{float(ref[2][0]):.6f} loss {got[3]:.6f}/{ref[3]:.6f} "
              f"dx0 rel {rel(got[4][0], ref[4][0]):.2e} dx1 rel {rel(got[4][1], ref[4][1]):.2e} ds {got[5]:.6e}/{ref[5]:.6e}")
        if rel(got[0][:N], ref[0][:N]) > 1e-2:
            print("   rowsum got", got[0][:8].tolist(), "\n   rowsum ref", ref[0][:8].tolist())
            print("   colsum got", got[1][:8].tolist(), "\n   colsum ref", ref[1][:8].tolist())


if __name__ == "__main__":
    main()
