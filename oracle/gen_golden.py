"""Generate tests/golden/* by running the REFERENCE's own code in the build container.

Run once, here (needs /root/reference; the GPU box does not have it):
    python oracle/gen_golden.py

* loss goldens: imports /root/reference/bioscanclip/model/loss_func.py by path (it needs
  only torch) and runs ContrastiveLoss / ClipLoss forward + autograd backward on seeded
  CPU inputs; ClipLoss cases run under a gloo process group (world 1, and world 2 via
  mp.spawn) because loss_func.py:143 all-gathers the labels unconditionally.
* accuracy goldens: /root/reference/bioscanclip/util/util.py cannot be imported (faiss,
  timm, torchtext, ... are absent), so the two pure-python functions
  top_k_micro_accuracy (util.py:379-395) and top_k_macro_accuracy (util.py:555-599) are
  taken from its source text with ast and executed unmodified.

TEST INFRASTRUCTURE ONLY.
"""
import ast
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_ref_loss():
    spec = importlib.util.spec_from_file_location("ref_loss_func", f"{REF}/bioscanclip/model/loss_func.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_accuracy():
    src = open(f"{REF}/bioscanclip/util/util.py").read()
    tree = ast.parse(src)
    ns = {}
    wanted = {"top_k_micro_accuracy", "top_k_macro_accuracy"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    body += [n for n in tree.body if isinstance(n, ast.Assign)
             and any(isinstance(t, ast.Name) and t.id == "LEVELS" for t in n.targets)]
    exec(compile(ast.Module(body=body, type_ignores=[]), "util_extract", "exec"), ns)
    return ns


def run_case(crit, feats, labels, scale, grad_mult=1.0, **fw):
    leaves = [None if f is None else f.clone().requires_grad_(True) for f in feats]
    s = scale.clone().requires_grad_(True) if isinstance(scale, torch.Tensor) else scale
    loss = crit(leaves[0], leaves[1], leaves[2], labels, s, **fw)
    if isinstance(loss, dict):
        loss = loss["contrastive_loss"]
    (loss * grad_mult).backward()
    out = {"loss": np.float64(loss.detach().double().item())}
    for name, leaf in zip(("image", "dna", "text"), leaves):
        if leaf is not None:
            out[f"grad_{name}"] = leaf.grad.float().numpy()
    if isinstance(s, torch.Tensor):
        out["dlogit_scale"] = np.float64(s.grad.double().item())
    return out


def save(name, inputs, outputs, meta):
    path = os.path.join(OUT, name + ".npz")
    payload = {}
    for k, v in inputs.items():
        if v is not None:
            payload["in_" + k] = v
    for k, v in outputs.items():
        payload["out_" + k] = v
    payload["meta"] = np.array(json.dumps(meta))
    np.savez(path, **payload)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in outputs.items()})


def gen_contrastive(ref):
    import torch.nn as nn
    # --- BASELINE configs[0]: image-DNA, N=256, d=768, fp32, one-hot labels
    torch.manual_seed(0)
    N, d = 256, 768
    A, B = torch.randn(N, d), torch.randn(N, d)
    labels = torch.arange(N)
    scale = torch.tensor(1 / 0.07)
    crit = ref.ContrastiveLoss(nn.CrossEntropyLoss(), 1 / 0.07)
    out = run_case(crit, [A, B, None], labels, scale)
    save("loss_cfg1_imgdna_n256_d768_fp32",
         {"image": A.numpy(), "dna": B.numpy(), "labels": labels.numpy(), "logit_scale": np.float32(scale.item())},
         out, {"module": "ContrastiveLoss", "present": [1, 1, 0]})

    # --- three modalities, multi-positive labels, small
    torch.manual_seed(1)
    N, d = 96, 64
    feats = [torch.randn(N, d) for _ in range(3)]
    labels = torch.randint(0, 12, (N,))
    scale = torch.tensor(1 / 0.07)
    out = run_case(crit, feats, labels, scale)
    save("loss_three_multipos_n96_d64_fp32",
         {"image": feats[0].numpy(), "dna": feats[1].numpy(), "text": feats[2].numpy(),
          "labels": labels.numpy(), "logit_scale": np.float32(scale.item())},
         out, {"module": "ContrastiveLoss", "present": [1, 1, 1]})

    # --- GradScaler-style upstream gradient (train_epoch.py:58), python-float scale from ctor
    out = run_case(crit, feats, labels, None, grad_mult=65536.0)
    save("loss_three_multipos_n96_d64_gradscale65536_ctor_scale",
         {"image": feats[0].numpy(), "dna": feats[1].numpy(), "text": feats[2].numpy(),
          "labels": labels.numpy(), "logit_scale": np.float32(1 / 0.07)},
         out, {"module": "ContrastiveLoss", "present": [1, 1, 1], "grad_mult": 65536.0, "scale_from_ctor": True})

    # --- ragged: N not a multiple of anything, d odd, dna+text only, correlated modalities, big scale
    torch.manual_seed(2)
    N, d = 77, 45
    base = torch.randn(N, d)
    feats = [None, base + 0.3 * torch.randn(N, d), base + 0.3 * torch.randn(N, d)]
    labels = torch.randint(0, 9, (N,))
    scale = torch.tensor(30.0)
    out = run_case(crit, feats, labels, scale)
    save("loss_dnatext_ragged_n77_d45_scale30",
         {"dna": feats[1].numpy(), "text": feats[2].numpy(), "labels": labels.numpy(),
          "logit_scale": np.float32(30.0)},
         out, {"module": "ContrastiveLoss", "present": [0, 1, 1]})

    # --- bf16 inputs fed directly to the reference (secondary target)
    torch.manual_seed(3)
    N, d = 128, 64
    feats32 = [torch.randn(N, d).bfloat16() for _ in range(2)]
    labels = torch.randint(0, 32, (N,))
    scale = torch.tensor(1 / 0.07)
    out_bf16 = run_case(crit, [feats32[0], feats32[1], None], labels, scale)
    out_fp32 = run_case(crit, [feats32[0].float(), feats32[1].float(), None], labels, scale)
    merged = {("bf16fed_" + k): v for k, v in out_bf16.items()}
    merged.update(out_fp32)
    save("loss_imgdna_bf16inputs_n128_d64",
         {"image": feats32[0].float().numpy(), "dna": feats32[1].float().numpy(), "labels": labels.numpy(),
          "logit_scale": np.float32(scale.item())},
         merged, {"module": "ContrastiveLoss", "present": [1, 1, 0], "inputs_are_bf16_exact": True})


def gen_cliploss_world1(ref):
    store = tempfile.mktemp()
    dist.init_process_group("gloo", init_method=f"file://{store}", rank=0, world_size=1)
    torch.manual_seed(4)
    N, d = 48, 32
    feats = [torch.randn(N, d) for _ in range(3)]
    labels = torch.randint(0, 10, (N,))
    scale = torch.tensor(1 / 0.07)
    cases = [
        ("all", [1, 1, 1], {}),
        ("bind_image", [1, 1, 1], {"bind_to": "image"}),
        ("bind_dna", [1, 1, 1], {"bind_to": "dna"}),
        ("bind_text", [1, 1, 1], {"bind_to": "text"}),
        ("no_image_text", [1, 1, 1], {"no_image_text_loss": True}),
        ("imgdna_only", [1, 1, 0], {}),
        ("imgtext_bind_dna_quirk", [1, 0, 1], {"bind_to": "dna"}),  # index 1 of the FILTERED list = text
        ("imgtext_no_image_text_quirk", [1, 0, 1], {"no_image_text_loss": True}),  # idx 2 absent -> no filter
    ]
    for name, present, kw in cases:
        crit = ref.ClipLoss(local_loss=False, gather_with_grad=True, rank=0, world_size=1, **kw)
        f = [feats[i] if present[i] else None for i in range(3)]
        out = run_case(crit, f, labels, scale, output_dict=(name == "all"))
        inputs = {"labels": labels.numpy(), "logit_scale": np.float32(scale.item())}
        for i, nm in enumerate(("image", "dna", "text")):
            if present[i]:
                inputs[nm] = feats[i].numpy()
        save(f"cliploss_w1_{name}_n48_d32", inputs, out, {"module": "ClipLoss", "present": present, **kw})
    dist.destroy_process_group()


def _w2_worker(rank, world, store, ret):
    ref = load_ref_loss()
    dist.init_process_group("gloo", init_method=f"file://{store}", rank=rank, world_size=world)
    torch.manual_seed(5)
    n, d = 32, 32
    feats_all = [torch.randn(world * n, d) for _ in range(3)]
    labels_all = torch.randint(0, 12, (world * n,))
    scale = torch.tensor(1 / 0.07)
    crit = ref.ClipLoss(local_loss=False, gather_with_grad=True, rank=rank, world_size=world)
    sl = slice(rank * n, (rank + 1) * n)
    out = run_case(crit, [f[sl] for f in feats_all], labels_all[sl], scale)
    out = {k: (v.tolist() if hasattr(v, "tolist") else float(v)) for k, v in out.items()}
    if rank == 0:
        out["inputs"] = {"image": feats_all[0].tolist(), "dna": feats_all[1].tolist(),
                         "text": feats_all[2].tolist(), "labels": labels_all.tolist()}
    ret[rank] = out
    dist.destroy_process_group()


def gen_cliploss_world2():
    world = 2
    store = tempfile.mktemp()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_w2_worker, args=(world, store, ret), nprocs=world, join=True)
    inp = ret[0]["inputs"]
    inputs = {k: np.asarray(v, dtype=np.float32) for k, v in inp.items() if k != "labels"}
    inputs["labels"] = np.asarray(inp["labels"], dtype=np.int64)
    inputs["logit_scale"] = np.float32(1 / 0.07)
    outputs = {}
    for r in range(world):
        for k, v in ret[r].items():
            if k == "inputs":
                continue
            outputs[f"rank{r}_{k}"] = np.asarray(v, dtype=np.float64 if k in ("loss", "dlogit_scale") else np.float32)
    save("cliploss_w2_all_n64_d32", inputs, outputs,
         {"module": "ClipLoss", "world": 2, "present": [1, 1, 1], "gather_with_grad": True})


def gen_accuracy():
    ns = load_ref_accuracy()
    rng = np.random.default_rng(7)
    levels = ns["LEVELS"]
    Q, kmax = 200, 5
    card = {"order": 4, "family": 9, "genus": 25, "species": 60}
    def lab(level, i):
        return f"{level[:2]}_{i}"
    gt = [{l: lab(l, int(rng.integers(0, card[l]))) for l in levels} for _ in range(Q)]
    pred = []
    for q in range(Q):
        p = {}
        for l in levels:
            row = [lab(l, int(rng.integers(0, card[l]))) for _ in range(kmax)]
            if rng.random() < 0.5:  # plant the right answer somewhere in the top-k
                row[int(rng.integers(0, kmax))] = gt[q][l]
            p[l] = row
        pred.append(p)
    k_list = [1, 3, 5]
    micro = ns["top_k_micro_accuracy"](pred, gt, k_list=k_list)
    macro, per_class = ns["top_k_macro_accuracy"](pred, gt, k_list=k_list)
    blob = {"levels": levels, "k_list": k_list, "pred_list": pred, "gt_list": gt,
            "micro": {str(k): v for k, v in micro.items()},
            "macro": {str(k): v for k, v in macro.items()},
            "per_class": {str(k): v for k, v in per_class.items()}}
    path = os.path.join(OUT, "accuracy_ref_q200_k5.json")
    json.dump(blob, open(path, "w"))
    print("wrote", path)


def load_ref_function(path, func_name, class_name=None, extra_ns=None):
    """Extract one function (or method) from a reference source file with ast and compile it unmodified."""
    tree = ast.parse(open(path).read())
    nodes = tree.body
    if class_name is not None:
        cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
        nodes = cls.body
    fn = next(n for n in nodes if isinstance(n, ast.FunctionDef) and n.name == func_name)
    ns = dict(extra_ns or {})
    exec(compile(ast.Module(body=[fn], type_ignores=[]), os.path.basename(path) + ":" + func_name, "exec"), ns)
    return ns[func_name]


def gen_seam():
    """Goldens of the seams (SURVEY section 8 f): DNA head, embedding hand-off, derived feature types."""
    import types
    from torch import Tensor
    import torch.nn.functional as F

    # --- BarcodeBERT head: execute CLIBDDNAEncoder.forward (dna_encoder.py:131-137) on a stub encoder
    fwd = load_ref_function(f"{REF}/bioscanclip/model/dna_encoder.py", "forward", class_name="CLIBDDNAEncoder",
                            extra_ns={"Tensor": Tensor})
    g = torch.Generator().manual_seed(21)
    for name, (n, t, c), scale in (("seam_softmax_mean_n3_t133_c96", (3, 133, 96), 3.0),
                                   ("seam_softmax_mean_n2_t7_c45", (2, 7, 45), 8.0)):
        logits = (torch.randn(n, t, c, generator=g) * scale).requires_grad_(True)
        stub = types.SimpleNamespace(base_dna_encoder=lambda seq, _l=logits: types.SimpleNamespace(logits=_l))
        out = fwd(stub, None)
        gout = torch.randn(n, c, generator=g)
        out.backward(gout)
        save(name, {"logits": logits.detach().numpy(), "grad_out": gout.numpy()},
             {"out": out.detach().numpy(), "grad_logits": logits.grad.numpy()},
             {"ref": "dna_encoder.py:137 logits.softmax(dim=-1).mean(dim=1) + torch autograd"})

    # --- embedding hand-off: the lines of inference_epoch.py:96-101 and :108-119 applied to two batches
    feats = [torch.randn(5, 768, generator=g) * 3.0, torch.randn(3, 768, generator=g) * 0.01]
    feats[1][1] = 0.0  # a zero row: F.normalize's eps clamp
    lst = []
    for f in feats:
        lst.extend(F.normalize(f, dim=-1).cpu().tolist())
    arr = np.array(lst)
    save("seam_embed_handoff_n8_d768", {"batch0": feats[0].numpy(), "batch1": feats[1].numpy()}, {"features": arr},
         {"ref": "inference_epoch.py:96-101,108-119"})

    # --- derived feature types: execute get_features_and_label (util.py:702-742) on stubbed extraction
    rng = np.random.default_rng(5)
    img, dna, txt = (rng.standard_normal((6, 16)).astype(np.float32).astype(np.float64) for _ in range(3))
    labels = [{"order": f"o{i}", "family": f"f{i}", "genus": f"g{i}", "species": f"s{i}"} for i in range(6)]
    stub_extract = lambda *a, **k: (list(range(6)), img, dna, txt, labels)  # noqa: E731
    gfl = load_ref_function(f"{REF}/bioscanclip/util/util.py", "get_features_and_label",
                            extra_ns={"np": np, "get_feature_and_label": stub_extract})
    d = gfl(None, types.SimpleNamespace(eval=lambda: None), None, for_key_set=True)
    save("seam_derived_feature_types_n6_d16", {"image": img, "dna": dna, "text": txt},
         {"averaged_feature": d["averaged_feature"], "concatenated_feature": d["concatenated_feature"],
          "all_key_features": d["all_key_features"],
          "all_key_features_label_species": np.array([l["species"] for l in d["all_key_features_label"]])},
         {"ref": "util.py:702-742", "keys": sorted(d.keys())})


def gen_info_nce():
    """SimCLR info-NCE goldens: execute SimCLR.info_nce_loss (simclr.py:64-92) from the reference source on a
    stub `self`, then the trainer's criterion (nn.CrossEntropyLoss, simclr.py:62,119) and torch autograd."""
    import types
    import torch.nn.functional as F
    fn = load_ref_function(f"{REF}/bioscanclip/util/simclr.py", "info_nce_loss", class_name="SimCLR",
                           extra_ns={"torch": torch, "F": F})
    g = torch.Generator().manual_seed(33)
    for name, (B, d, tau, gm, corr) in (("infonce_b64_d768_t0.07", (64, 768, 0.07, 1.0, 0.0)),
                                        ("infonce_b45_d40_t0.2_views_correlated", (45, 40, 0.2, 1.0, 0.8)),
                                        ("infonce_b16_d96_t0.07_gradscale65536", (16, 96, 0.07, 65536.0, 0.5))):
        base = torch.randn(B, d, generator=g)
        v1 = corr * base + (1 - corr) * torch.randn(B, d, generator=g)
        v2 = corr * base + (1 - corr) * torch.randn(B, d, generator=g)
        feats = (torch.cat([v1, v2], 0) * 2.5).requires_grad_(True)
        stub = types.SimpleNamespace(
            args=types.SimpleNamespace(model_config=types.SimpleNamespace(batch_size=B, n_views=2, temperature=tau)),
            device="cpu")
        logits, labels = fn(stub, feats)
        loss = torch.nn.CrossEntropyLoss()(logits, labels)
        (loss * gm).backward()
        save(name, {"features": feats.detach().numpy()},
             {"loss": np.float64(loss.item()), "grad": feats.grad.numpy(), "logits": logits.detach().numpy()},
             {"ref": "simclr.py:64-92,119", "batch_size": B, "n_views": 2, "temperature": tau, "grad_mult": gm})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = load_ref_loss()
    gen_contrastive(ref)
    gen_cliploss_world1(ref)
    gen_cliploss_world2()
    gen_accuracy()
    gen_seam()
    gen_info_nce()
