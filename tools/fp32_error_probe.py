"""How close is the exact CUDA-core path to the float64 oracle on the reference's golden inputs -- next to how close the
reference's own float32 outputs (the goldens) are to the same oracle."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clibd_b200 as cb  # noqa: E402
from oracle import loss_oracle as lo  # noqa: E402
from tests import _golden  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


dev = torch.device("cuda:0")
for g in _golden.all_single_process():
    if g.meta.get("inputs_are_bf16_exact"):
        continue
    mult = g.meta.get("grad_mult", 1.0)
    res = lo.contrastive_loss(g.features, g.labels, g.logit_scale, grad_out=mult, **g.kwargs())
    feats = [None if f is None else torch.from_numpy(f).to(dev).requires_grad_(True) for f in g.features]
    scale = torch.tensor(g.logit_scale, device=dev, requires_grad=True)
    mod = cb.ClipLoss(gather_with_grad=True, rank=0, world_size=1, tensor_core_operands="fp32", **g.kwargs())
    loss = mod(feats[0], feats[1], feats[2], torch.from_numpy(g.labels).to(dev), scale)
    (loss * mult).backward()
    ours = {"loss": abs(float(loss) - res["loss"]) / abs(res["loss"]),
            "grad": max(rel(f.grad.cpu().numpy(), r) for f, r in zip(feats, res["grads"]) if f is not None),
            "ds": abs(float(scale.grad) - res["dlogit_scale"]) / abs(res["dlogit_scale"])}
    ref = {"loss": abs(float(g.outputs["loss"]) - res["loss"]) / abs(res["loss"]),
           "grad": max(rel(g.outputs[f"grad_{m}"], res["grads"][i]) for i, m in enumerate(_golden.MODS)
                       if f"grad_{m}" in g.outputs)}
    if "dlogit_scale" in g.outputs:
        ref["ds"] = abs(float(g.outputs["dlogit_scale"]) - res["dlogit_scale"]) / abs(res["dlogit_scale"])
    print("FP32ERR " + json.dumps({"golden": g.name, "ours_vs_float64": ours, "reference_fp32_vs_float64": ref}))
