"""BASELINE config 3: the full training step around the loss, with random-init stand-ins of the reference's encoders.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P \
        tools/train_step.py [--per-gpu 500] [--steps 5] [--warmup 2] [--tiny]

The hot path this repo replaces is the loss; this harness is the place where it meets a REAL autograd graph the way
bioscanclip/epoch/train_epoch.py:9-63 drives it:
  * encoders of the reference's shapes, random init (no checkpoints offline): torchvision ViT-B/16 with a 768-d head
    (timm vit_base_patch16_224 + reset_classifier(768), image_encoder.py:92-93), BertForMaskedLM(vocab 1027) with the
    decoder replaced by Linear(768, 768) and `logits.softmax(-1).mean(1)` (dna_encoder.py:121-137), a 4-layer 512-wide
    BertModel + Linear(512, 768) over the mean token (language_encoder.py:78-89); LoRA r = 4 on the query / value
    projections, everything else frozen (simple_clip.py:166-171); outputs L2-normalised (simple_clip.py:45,58,60);
    learnable logit_scale = ln(1/0.07) passed as exp() (simple_clip.py:32,61);
  * DDP(find_unused_parameters=True) (train_cl.py:204), bf16 autocast around the model only, the loss called OUTSIDE
    autocast with keyword arguments, GradScaler, AdamW, torch.autograd.set_detect_anomaly(True) (train_epoch.py:11,
    42-60);
  * the criterion is either the UNMODIFIED reference ClipLoss (oracle/_ref/loss_func.py, NCCL all-gather) or
    clibd_b200.ClipLoss with the same constructor arguments -- one JSON line reports the step time with each and the
    time of the loss's own forward + backward on the step's embeddings (its share of the step).
"""
import argparse
import json
import math
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parallel import DistributedDataParallel as DDP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class LoRALinear(nn.Module):
    """y = W x + (alpha / r) B A x with W frozen (loratorch-style adapter on a q / v projection)."""

    def __init__(self, base: nn.Linear, r=4, alpha=4):
        super().__init__()
        self.base = base
        for p in self.base.parameters():
            p.requires_grad_(False)
        self.a = nn.Parameter(torch.empty(r, base.in_features))
        self.b = nn.Parameter(torch.zeros(base.out_features, r))
        nn.init.kaiming_uniform_(self.a, a=math.sqrt(5))
        self.scaling = alpha / r

    def forward(self, x):
        return self.base(x) + F.linear(F.linear(x, self.a), self.b) * self.scaling


def lora_bert(model, r=4):
    for p in model.parameters():
        p.requires_grad_(False)
    for layer in model.encoder.layer:
        att = layer.attention.self
        att.query = LoRALinear(att.query, r)
        att.value = LoRALinear(att.value, r)
    return model


class LoRAViTAttention(nn.Module):
    """torchvision's encoder block uses nn.MultiheadAttention with a fused in_proj; the LoRA update is added to the
    query and value thirds of that projection (image_encoder.py:13-107 does the same on timm's fused qkv)."""

    def __init__(self, mha: nn.MultiheadAttention, r=4):
        super().__init__()
        self.mha = mha
        for p in self.mha.parameters():
            p.requires_grad_(False)
        e = mha.embed_dim
        self.aq = nn.Parameter(torch.empty(r, e))
        self.bq = nn.Parameter(torch.zeros(e, r))
        self.av = nn.Parameter(torch.empty(r, e))
        self.bv = nn.Parameter(torch.zeros(e, r))
        nn.init.kaiming_uniform_(self.aq, a=math.sqrt(5))
        nn.init.kaiming_uniform_(self.av, a=math.sqrt(5))

    def forward(self, query, key, value, need_weights=False, **kw):
        e = self.mha.embed_dim
        w = self.mha.in_proj_weight
        dw = torch.cat([self.bq @ self.aq, torch.zeros(e, e, device=w.device, dtype=w.dtype), self.bv @ self.av], 0)
        out = F.multi_head_attention_forward(
            query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1), e, self.mha.num_heads, w + dw,
            self.mha.in_proj_bias, None, None, False, 0.0, self.mha.out_proj.weight, self.mha.out_proj.bias,
            training=self.training, need_weights=False)[0]
        return out.transpose(0, 1), None


class StandInCLIP(nn.Module):
    """SimpleCLIP.forward (simple_clip.py:38-61) over stand-in encoders of the reference's shapes."""

    def __init__(self, tiny=False):
        super().__init__()
        from torchvision.models import VisionTransformer
        from transformers import BertConfig, BertForMaskedLM, BertModel
        if tiny:  # CPU smoke test of the harness
            vit = dict(image_size=32, patch_size=16, num_layers=1, num_heads=2, hidden_dim=32, mlp_dim=64, num_classes=768)
            dna = dict(vocab_size=1027, hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64)
            txt = dict(hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=64)
        else:
            vit = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072,
                       num_classes=768)
            dna = dict(vocab_size=1027, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                       intermediate_size=3072)
            txt = dict(hidden_size=512, num_hidden_layers=4, num_attention_heads=8, intermediate_size=2048)
        self.image = VisionTransformer(**vit)
        for p in self.image.parameters():
            p.requires_grad_(False)
        for blk in self.image.encoder.layers:
            blk.self_attention = LoRAViTAttention(blk.self_attention)
        for p in self.image.heads.parameters():
            p.requires_grad_(True)
        self.dna = BertForMaskedLM(BertConfig(**dna))
        lora_bert(self.dna.bert)
        for p in self.dna.cls.parameters():
            p.requires_grad_(False)
        self.dna.cls.predictions.decoder = nn.Linear(dna["hidden_size"], 768)  # dna_encoder.py:121-123
        self.text = lora_bert(BertModel(BertConfig(**txt), add_pooling_layer=False))
        self.text_proj = nn.Linear(txt["hidden_size"], 768)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))  # simple_clip.py:32

    def forward(self, image, dna_tokens, text_ids):
        img = F.normalize(self.image(image), p=2, dim=-1)
        dna = F.normalize(self.dna(dna_tokens).logits.softmax(dim=-1).mean(dim=1), p=2, dim=-1)  # dna_encoder.py:137
        txt = self.text(input_ids=text_ids).last_hidden_state.mean(dim=1)
        txt = F.normalize(self.text_proj(txt), p=2, dim=-1)
        return img, dna, txt, self.logit_scale.exp(), None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=500)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--criteria", default="reference,clibd_b200")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cuda = torch.cuda.is_available()
    dev = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if cuda else "gloo", **({"device_id": dev} if cuda else {}))
    torch.manual_seed(0)  # identical initial weights on every rank (DDP broadcasts rank 0's anyway)
    n = args.per_gpu
    side = 32 if args.tiny else 224
    gen = torch.Generator().manual_seed(100 + rank)
    image = torch.rand(n, 3, side, side, generator=gen).to(dev)
    dna_tokens = torch.randint(0, 1027, (n, 133), generator=gen).to(dev)
    text_ids = torch.randint(0, 30522, (n, 20), generator=gen).to(dev)
    labels = (torch.arange(n) + rank * n).to(dev)  # training label ids (dataset.py:155-165): here one per sample

    from oracle import build_ref
    ref = build_ref.load()
    import clibd_b200 as cb
    kwargs = dict(local_loss=False, gather_with_grad=True, rank=rank, world_size=world, use_horovod=False,
                  criterion=nn.CrossEntropyLoss(), bind_to=None, no_image_text_loss=False)
    criteria = {}
    for name in args.criteria.split(","):
        if name == "reference" and ref is not None:
            criteria[name] = ref.ClipLoss(**kwargs)
        elif name == "clibd_b200" and cuda:
            criteria[name] = cb.ClipLoss(**kwargs)
    out = {"config": "BASELINE config 3: full train step, stand-in encoders", "world": world, "per_gpu": n,
           "global_batch": n * world, "autocast": "bf16", "ddp": True, "detect_anomaly": True, "results": {}}

    def sync():
        if cuda:
            torch.cuda.synchronize()
        dist.barrier()

    for name, criterion in criteria.items():
        torch.manual_seed(0)
        model = StandInCLIP(tiny=args.tiny).to(dev)
        ddp = DDP(model, device_ids=[local] if cuda else None, find_unused_parameters=True)
        opt = torch.optim.AdamW([p for p in ddp.parameters() if p.requires_grad], lr=1e-4)
        scaler = torch.amp.GradScaler("cuda", enabled=cuda)
        torch.autograd.set_detect_anomaly(True)  # train_epoch.py:11
        losses, t_steps = [], []
        feats = None
        for it in range(args.warmup + args.steps):
            sync()
            t0 = time.perf_counter()
            opt.zero_grad()
            with torch.autocast(device_type=dev.type, dtype=torch.bfloat16):
                img, dna, txt, logit_scale, _ = ddp(image, dna_tokens, text_ids)
            loss = criterion(image_features=img, dna_features=dna, text_features=txt, labels=labels,
                             logit_scale=logit_scale)
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            losses.append(loss.item())  # host sync every step, like train_epoch.py:65
            sync()
            if it >= args.warmup:
                t_steps.append(time.perf_counter() - t0)
            feats = [t.detach() for t in (img, dna, txt)]
        torch.autograd.set_detect_anomaly(False)
        # the loss alone on the step's embeddings (fwd + bwd, anomaly mode off: it only adds host-side bookkeeping)
        t_loss = []
        for it in range(3 + 10):
            leaves = [f.clone().requires_grad_(True) for f in feats]
            s = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
            sync()
            t0 = time.perf_counter()
            l2 = criterion(image_features=leaves[0], dna_features=leaves[1], text_features=leaves[2], labels=labels,
                           logit_scale=s)
            l2.backward()
            sync()
            if it >= 3:
                t_loss.append(time.perf_counter() - t0)
        step_ms = 1e3 * sorted(t_steps)[len(t_steps) // 2]
        loss_ms = 1e3 * sorted(t_loss)[len(t_loss) // 2]
        tl = torch.tensor([step_ms, loss_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        out["results"][name] = {"step_ms": float(tl[0]), "loss_fwd_bwd_ms": float(tl[1]),
                                "loss_share_of_step": float(tl[1] / tl[0]),
                                "samples_per_s": n * world / (float(tl[0]) * 1e-3), "losses": losses[:4],
                                "embedding_dtype": str(feats[0].dtype)}
        del ddp, model, opt
        if cuda:
            torch.cuda.empty_cache()
    # ---- the two criteria on the SAME embeddings (the last step's): loss, feature gradients, d/d(logit_scale)
    if "reference" in criteria and "clibd_b200" in criteria:
        got = {}
        for name, criterion in criteria.items():
            leaves = [f.clone().requires_grad_(True) for f in feats]
            s_ = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
            l_ = criterion(image_features=leaves[0], dna_features=leaves[1], text_features=leaves[2], labels=labels,
                           logit_scale=s_)
            (l_ * 65536.0).backward()  # GradScaler's initial scale
            got[name] = (float(l_.detach()), [t.grad.float() for t in leaves], float(s_.grad))
        a, b = got["reference"], got["clibd_b200"]
        out["same_embeddings"] = {
            "loss_rel_diff": abs(a[0] - b[0]) / abs(a[0]),
            "feature_grad_rel_diff": [float((x - y).norm() / x.norm()) for x, y in zip(a[1], b[1])],
            "dlogit_scale_rel_diff": abs(a[2] - b[2]) / abs(a[2])}
    if "reference" in out["results"] and "clibd_b200" in out["results"]:
        a, b = out["results"]["reference"], out["results"]["clibd_b200"]
        out["first_step_loss_rel_diff"] = abs(a["losses"][0] - b["losses"][0]) / abs(a["losses"][0])
        out["loss_speedup"] = a["loss_fwd_bwd_ms"] / b["loss_fwd_bwd_ms"]
        out["step_speedup"] = a["step_ms"] / b["step_ms"]
    if rank == 0:
        print("TRAINSTEP " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
