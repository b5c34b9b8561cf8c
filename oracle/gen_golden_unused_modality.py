"""Golden for a modality that is present but belongs to no pair: ClipLoss(bind_to="image", no_image_text_loss=True)
with image, DNA and text features -- the reference's pair loop (loss_func.py:166-184) keeps (image, dna) only, text never
enters the graph and its leaf keeps grad None.  Test infrastructure: run in THIS container (imports the reference from
/root/reference), the product never imports it.

    python oracle/gen_golden_unused_modality.py"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gen_golden as G  # noqa: E402


def main():
    ref = G.load_ref_loss()
    dist.init_process_group("gloo", init_method=f"file://{tempfile.mktemp()}", rank=0, world_size=1)
    torch.manual_seed(4)  # the inputs of the other cliploss_w1_* cases
    N, d = 48, 32
    feats = [torch.randn(N, d) for _ in range(3)]
    labels = torch.randint(0, 10, (N,))
    scale = torch.tensor(1 / 0.07)
    kw = {"bind_to": "image", "no_image_text_loss": True}
    crit = ref.ClipLoss(local_loss=False, gather_with_grad=True, rank=0, world_size=1, **kw)
    leaves = [f.clone().requires_grad_(True) for f in feats]
    s = scale.clone().requires_grad_(True)
    loss = crit(leaves[0], leaves[1], leaves[2], labels, s)
    loss.backward()
    assert leaves[2].grad is None, "the reference gave the unused modality a gradient"
    out = {"loss": np.float64(loss.detach().double().item()), "grad_image": leaves[0].grad.numpy(),
           "grad_dna": leaves[1].grad.numpy(), "dlogit_scale": np.float64(s.grad.double().item())}
    inputs = {"labels": labels.numpy(), "logit_scale": np.float32(scale.item()), "image": feats[0].numpy(),
              "dna": feats[1].numpy(), "text": feats[2].numpy()}
    G.save("cliploss_w1_bind_image_no_image_text_unused_text_n48_d32", inputs, out,
           {"module": "ClipLoss", "present": [1, 1, 1], "grad_none": ["text"], **kw})
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
