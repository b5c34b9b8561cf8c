"""ctypes binding of include/clibd_b200.h.  There is NO fallback: if the CUDA library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes
import os
import subprocess

from . import _build

_c = ctypes
_P = _c.c_void_p
_I64 = _c.c_int64
_INT = _c.c_int
_F = _c.c_float

ABI_VERSION = 7
DT_F32, DT_BF16, DT_F16, DT_F64 = 0, 1, 2, 3
MODE_LOCAL, MODE_EXCHANGE = 0, 1
PATH_SIMT_F32, PATH_TC_BF16, PATH_TC_F16 = 0, 1, 2

# name -> (restype, argtypes); must list every symbol include/clibd_b200.h declares
SIGNATURES = {
    "clibd_abi_version": (_INT, []),
    "clibd_last_error": (_c.c_char_p, []),
    "clibd_device_supported": (_INT, []),
    "clibd_kernel_launch_count": (_I64, []),
    "clibd_profile_enable": (_INT, [_INT]),
    "clibd_graphs_active": (_INT, [_I64, _I64]),
    "clibd_profile_read": (_INT, [_P, _P]),
    "clibd_row_inv_norm": (_INT, [_P, _INT, _I64, _I64, _P, _P]),
    "clibd_loss_scratch_bytes": (_I64, [_I64, _I64, _I64, _INT, _INT]),
    "clibd_loss_forward_stats": (_INT, [_P, _INT, _P, _P, _I64, _I64, _I64, _I64, _F, _P, _P, _INT, _INT, _P, _I64, _P,
                                        _P, _P, _P, _P]),
    "clibd_loss_label_stage": (_INT, [_P, _I64, _I64, _I64, _INT, _INT, _P, _I64, _P]),
    "clibd_loss_forward_finish": (_INT, [_I64, _I64, _I64, _F, _P, _INT, _INT, _P, _I64, _P, _P, _P, _P, _P]),
    "clibd_loss_backward": (_INT, [_P, _INT, _P, _I64, _I64, _I64, _I64, _F, _P, _INT, _P, _I64, _F, _P, _P, _P, _P]),
    "clibd_loss_backward_sweeps": (_INT, [_P, _INT, _P, _I64, _I64, _I64, _I64, _F, _P, _INT, _P, _I64, _P, _P, _P,
                                          _INT, _INT, _P]),
    "clibd_loss_backward_finish": (_INT, [_P, _INT, _P, _I64, _I64, _I64, _I64, _F, _P, _INT, _P, _I64, _P, _P, _F, _P,
                                          _INT, _P, _P, _P]),
    "clibd_shard_push_rows": (_INT, [_P, _INT, _P, _I64, _I64, _INT, _INT, _P, _P, _P, _P]),
    "clibd_shard_push_stats": (_INT, [_P, _P, _I64, _I64, _I64, _INT, _INT, _P, _P, _P, _P]),
    "clibd_shard_reduce_stats": (_INT, [_P, _P, _I64, _INT, _P, _P, _P]),
    "clibd_shard_push_floats": (_INT, [_P, _I64, _INT, _INT, _P, _P]),
    "clibd_shard_barrier": (_INT, [_P, _INT, _INT, _INT, _P]),
    "clibd_shard_barrier_bytes": (_I64, []),
    "clibd_knn_normalize": (_INT, [_P, _INT, _I64, _I64, _P, _P]),
    "clibd_knn_scratch_bytes": (_I64, [_I64, _I64, _I64, _INT, _INT]),
    "clibd_knn_search": (_INT, [_P, _I64, _P, _I64, _I64, _I64, _INT, _INT, _P, _I64, _P, _P, _P, _P]),
    "clibd_knn_merge": (_INT, [_P, _P, _INT, _I64, _INT, _P, _P, _P, _P]),
    "clibd_topk_accuracy": (_INT, [_P, _I64, _INT, _P, _I64, _P, _P, _INT, _c.c_int32, _P, _P, _P, _P]),
    "clibd_embed_append": (_INT, [_P, _INT, _I64, _I64, _P, _I64, _I64, _I64, _P]),
    "clibd_softmax_mean_scratch_bytes": (_I64, [_I64, _I64, _I64, _INT]),
    "clibd_softmax_mean_forward": (_INT, [_P, _INT, _I64, _I64, _I64, _P, _P, _I64, _P]),
    "clibd_softmax_mean_backward": (_INT, [_P, _P, _INT, _I64, _I64, _I64, _P, _P]),
    "clibd_infonce_forward": (_INT, [_P, _INT, _P, _I64, _I64, _F, _INT, _P, _I64, _P, _P, _P]),
    "clibd_infonce_backward": (_INT, [_P, _INT, _P, _I64, _I64, _F, _INT, _P, _I64, _F, _P, _P, _P]),
}

_lib = None
_test_double = None


def inject_for_tests(double):
    """tests/ only: replace the CUDA library by an object with the same entry points so the host-side
    collective choreography can be exercised under gloo on a CPU box.  Never set by product code."""
    global _test_double
    _test_double = double


def test_double_active() -> bool:
    return _test_double is not None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load the shared library, rebuilding it first when the sources in the tree differ from the ones it was built
    from (content hash, clibd_b200/_build.py)."""
    global _lib
    if _test_double is not None:
        return _test_double
    if _lib is not None:
        return _lib
    path = os.environ.get("CLIBD_B200_LIB")  # development: an instrumented build of the same sources
    if not path:
        # rebuilds when a source is newer than the library (or CLIBD_B200_REBUILD is set); a fresh checkout on a box
        # without nvcc keeps the shipped binary
        try:
            path = _build.build(force=bool(os.environ.get("CLIBD_B200_REBUILD")))
        except (OSError, RuntimeError, subprocess.SubprocessError):
            path = _build.LIB_PATH
            if not os.path.exists(path):
                raise
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.clibd_abi_version() != ABI_VERSION:
        raise RuntimeError("clibd_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().clibd_last_error()
        msg = msg.decode() if isinstance(msg, bytes) else str(msg)
        if rc == 1:
            raise ValueError("clibd_b200: " + msg)
        raise RuntimeError("clibd_b200: " + msg)


def ptr_array3(ptrs):
    """host array of 3 device pointers (None -> NULL)"""
    arr = (_P * 3)()
    for i, p in enumerate(ptrs):
        arr[i] = _P(p) if p else _P(None)
    return arr


def float_array3(vals):
    arr = (_F * 3)()
    for i, v in enumerate(vals):
        arr[i] = float(v)
    return arr


def ptr_array(ptrs):
    """host array of device pointers (None -> NULL)"""
    arr = (_P * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = _P(p) if p else _P(None)
    return arr


def int_array(vals):
    arr = (_INT * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr
