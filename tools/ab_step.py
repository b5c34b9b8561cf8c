"""A/B of one environment switch inside ONE process: the 1-GPU loss step (bench shape) is timed in alternating blocks
with the switch set to each value, so that clock / thermal drift between separate runs (+-2.5 % on these boxes) cannot
be mistaken for an effect.

    python tools/ab_step.py CLIBD_SIDE_STREAM 0 1 [N] [rounds] [steps]

Only switches the library reads per call qualify (CLIBD_SIDE_STREAM, CLIBD_SHARD_MODE, CLIBD_BWD_TWO_SWEEPS, ...)."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clibd_b200 as cb  # noqa: E402
from tools import synth  # noqa: E402


def main():
    var, a, b = sys.argv[1], sys.argv[2], sys.argv[3]
    N = int(sys.argv[4]) if len(sys.argv) > 4 else 32768
    rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 6
    steps = int(sys.argv[6]) if len(sys.argv) > 6 else 10
    dev = torch.device("cuda:0")
    feats = [synth.feature_rows(N, 768, m, 0, N).to(dev) for m in range(3)]
    labels = synth.labels_all(N).to(dev)
    mod = cb.ContrastiveLoss(None, 1 / 0.07)
    scale = torch.tensor(1 / 0.07, device=dev)

    def step():
        leaves = [f.detach().requires_grad_(True) for f in feats]
        mod(leaves[0], leaves[1], leaves[2], labels, scale).backward()

    res = {a: [], b: []}
    for r in range(rounds + 1):
        for val in (a, b) if r % 2 == 0 else (b, a):
            os.environ[var] = val
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            if r > 0:  # round 0 warms everything up
                res[val].append(e0.elapsed_time(e1) / steps)
    print("ABSTEP " + json.dumps({"var": var, "N": N, "rounds": rounds, "steps": steps,
                                  **{f"{var}={k}": {"mean_ms": round(statistics.mean(v), 4), "min_ms": round(min(v), 4),
                                                    "all": [round(x, 3) for x in v]} for k, v in res.items()}}))


if __name__ == "__main__":
    main()
