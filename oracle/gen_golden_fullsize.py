"""Golden values of the BENCHMARKED configuration (N = 32768, d = 768, image + DNA + text, bf16 inputs, labels ~
randint(0, N/8), logit_scale 1/0.07) from the float64 streaming oracle -- TEST INFRASTRUCTURE.

    python oracle/gen_golden_fullsize.py [N]        (about 10 minutes on 8 cores at N = 32768)

The reference itself cannot materialise this size (>= 12 [N, N] fp32 matrices); oracle/loss_oracle.py:
contrastive_loss_streaming is the row-blocked restatement of loss_func.py:41-69, pinned to the reference at small
sizes by tests/test_oracle_loss.py.  Stored: the loss, dL/d(logit_scale), and the gradient of 32 rows at the start,
the middle and the end of every modality (tests/golden/fullsize_n<N>.npz); bench.py also compares the loss of its
first step with this value at every GPU count."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_oracle as lo  # noqa: E402
from tools import synth  # noqa: E402

ROWS = 32


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    d = 768
    feats = [synth.feature_rows(N, d, m, 0, N).float().numpy().astype(np.float64) for m in range(3)]
    labels = synth.labels_all(N).numpy()
    t0 = time.time()
    res = lo.contrastive_loss_streaming(feats, labels, 1 / 0.07, block=1024)
    blocks = [0, (N // 2) - ROWS // 2, N - ROWS]
    out = {"N": N, "d": d, "logit_scale": 1 / 0.07, "loss": res["loss"], "dlogit_scale": res["dlogit_scale"],
           "row_starts": np.asarray(blocks), "rows": ROWS, "seconds": time.time() - t0}
    for m, name in enumerate(("image", "dna", "text")):
        g = res["grads"][m]
        out[f"grad_{name}"] = np.stack([g[b:b + ROWS] for b in blocks]).astype(np.float32)
        out[f"gradnorm_{name}"] = float(np.linalg.norm(g))
    path = os.path.join(ROOT, "tests", "golden", f"fullsize_n{N}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "loss", res["loss"], "ds", res["dlogit_scale"], f"{time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
