"""Loader for tests/golden/*.npz (written by oracle/gen_golden.py from the reference)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODS = ("image", "dna", "text")


class Golden:
    def __init__(self, path):
        z = np.load(path, allow_pickle=False)
        self.name = os.path.basename(path)[:-4]
        self.meta = json.loads(str(z["meta"]))
        self.inputs = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
        self.outputs = {k[4:]: z[k] for k in z.files if k.startswith("out_")}

    @property
    def features(self):
        return [self.inputs.get(m) for m in MODS]

    @property
    def labels(self):
        return self.inputs["labels"]

    @property
    def logit_scale(self):
        return float(self.inputs["logit_scale"])

    def kwargs(self):
        return {k: self.meta[k] for k in ("bind_to", "no_image_text_loss") if k in self.meta}


def load(name):
    return Golden(os.path.join(GOLDEN_DIR, name + ".npz"))


def all_single_process():
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        if os.path.basename(p).startswith(("seam_", "infonce_", "fullsize_")):  # other fixture families
            continue
        g = Golden(p)
        if g.meta.get("world", 1) == 1:
            out.append(g)
    return out
