"""Golden vectors for LARGE logit scales (60, 100), produced by the reference's own code -- TEST INFRASTRUCTURE.

    python oracle/gen_golden_large_scale.py        (build container only: needs /root/reference)

SimpleCLIP.logit_scale is an unclamped learnable parameter (simple_clip.py:32,61); CLIP-style training pushes
logit_scale.exp() from 14.3 towards 100.  The reference's nn.CrossEntropyLoss subtracts the row maximum, so it is exact
at any scale; the fused kernels use one shifted exponential for row sums, column sums and gradient coefficients and
must reproduce these values (csrc/common.cuh: softmax_shift).  Same file format as oracle/gen_golden.py."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden import load_ref_loss, run_case, save  # noqa: E402


def main():
    ref = load_ref_loss()
    crit = ref.ContrastiveLoss(nn.CrossEntropyLoss(), 1 / 0.07)
    for scale_value in (60.0, 100.0):
        # trained-like: the three modalities of a specimen are correlated, multi-positive labels
        torch.manual_seed(int(scale_value))
        N, d = 96, 64
        labels = torch.randint(0, 24, (N,))
        base = torch.randn(N, d)[labels]
        feats = [0.6 * base + 0.4 * torch.randn(N, d) for _ in range(3)]
        scale = torch.tensor(scale_value)
        out = run_case(crit, feats, labels, scale)
        save(f"loss_three_aligned_scale{int(scale_value)}_n96_d64",
             {"image": feats[0].numpy(), "dna": feats[1].numpy(), "text": feats[2].numpy(), "labels": labels.numpy(),
              "logit_scale": np.float32(scale_value)}, out, {"module": "ContrastiveLoss", "present": [1, 1, 1]})
    # untrained: independent modalities at scale 100 (logits spread over +-50: most terms are far below the row maximum)
    torch.manual_seed(7)
    N, d = 128, 64
    A, B = torch.randn(N, d), torch.randn(N, d)
    labels = torch.arange(N)
    out = run_case(crit, [A, B, None], labels, torch.tensor(100.0))
    save("loss_imgdna_independent_scale100_n128_d64",
         {"image": A.numpy(), "dna": B.numpy(), "labels": labels.numpy(), "logit_scale": np.float32(100.0)}, out,
         {"module": "ContrastiveLoss", "present": [1, 1, 0]})


if __name__ == "__main__":
    main()
