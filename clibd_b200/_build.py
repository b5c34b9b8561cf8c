"""Build the CUDA library IN-TREE: clibd_b200/lib/libclibd_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import fcntl
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libclibd_b200.so")
HASH_PATH = os.path.join(LIB_DIR, "libclibd_b200.sha256")  # hash of the sources the library was built from

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def source_hash() -> str:
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh"))
    deps += glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(deps):
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def _stale() -> bool:
    """The library is stale when it was built from other sources than the ones in the tree (content hash, not
    mtimes: a snapshot copied to another box keeps its binary)."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    return open(HASH_PATH).read().strip() != source_hash()


def build_variant(name: str, defines) -> str:
    """Development builds (e.g. -DCLIBD_BWD_TIMING) next to the product library: lib/libclibd_b200_<name>.so"""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    out = os.path.join(LIB_DIR, f"libclibd_b200_{name}.so")
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-shared", "-o", out, *sources()]
    subprocess.check_call(cmd)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:  # one builder at a time (one process per GPU)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():
            return LIB_PATH
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + out.decode())
    tmp = LIB_PATH + ".tmp"
    link = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(link)
    os.replace(tmp, LIB_PATH)
    with open(HASH_PATH, "w") as f:
        f.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
