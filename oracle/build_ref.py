"""Recipe that puts the UNMODIFIED reference loss module next to the oracle -- TEST / BENCH INFRASTRUCTURE.

    python oracle/build_ref.py            (run in the build container, where /root/reference exists)

/root/reference/bioscanclip/model/loss_func.py needs nothing but torch, so the reference's own ContrastiveLoss /
ClipLoss can be timed on the GPU box's host cores (bench.py --impl reference, cpu_baseline kind "reference").  The file
is copied byte for byte into oracle/_ref/ -- git-ignored, so it never enters this repository's history, but it travels
to the GPU box with the snapshot like the built libraries do.  Nothing under clibd_b200/ may import it."""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/bioscanclip/model/loss_func.py"
DST_DIR = os.path.join(ROOT, "oracle", "_ref")
DST = os.path.join(DST_DIR, "loss_func.py")


def build() -> str:
    """Returns the path of the vendored module ('' when the reference tree is absent, e.g. on the GPU box)."""
    if not os.path.exists(SRC):
        return DST if os.path.exists(DST) else ""
    os.makedirs(DST_DIR, exist_ok=True)
    shutil.copyfile(SRC, DST)
    with open(os.path.join(DST_DIR, "SOURCE.txt"), "w") as f:
        f.write(f"{SRC}\nsha256 {hashlib.sha256(open(SRC, 'rb').read()).hexdigest()}\n")
    return DST


def load():
    """Import oracle/_ref/loss_func.py as a module (None when it was never built)."""
    if not os.path.exists(DST):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("clibd_reference_loss_func", DST)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build()
    print(p if p else "reference tree not found", file=sys.stderr if not p else sys.stdout)
