#!/usr/bin/env python
"""Benchmark of the CLIBD hot path on B200 (see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours      : fused contrastive loss fwd+bwd (image+DNA+text, global batch 32768, d=768, bf16 inputs,
            label-matched multi-positive targets) through the public modules, one process per GPU;
            plus the cosine kNN retrieval (100k queries x 1M keys, d=768, k=5) as a second block.
reference : the reference algorithm's CPU path (oracle port: the reference is Python and cannot
            travel to the GPU box, and cannot materialise N=32768 anyway) on the host cores.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "contrastive fwd+bwd samples/s (B=32k,d=768); kNN queries/s vs 1M keys"
N_GLOBAL = int(os.environ.get("CLIBD_BENCH_BATCH", 32768))
DIM = 768
KNN_Q = int(os.environ.get("CLIBD_BENCH_KNN_Q", 100_000))
KNN_K = int(os.environ.get("CLIBD_BENCH_KNN_K", 1_000_000))
KNN_TOPK = 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-knn", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU leg (oracle port), used by cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_inputs(N, d, seed=1):
    import numpy as np
    rng = np.random.default_rng(seed)
    feats = [rng.standard_normal((N, d), dtype=np.float32) for _ in range(3)]
    labels = rng.integers(0, N // 8, N)
    return feats, labels


def cpu_sample(feats, labels, rows):
    """seconds for one row-block sample of the full step (all threads numpy/BLAS)."""
    from oracle import loss_oracle as lo
    t0 = time.perf_counter()
    lo.row_block_fwd_bwd(feats, labels, 1 / 0.07, 0, rows)
    return time.perf_counter() - t0


def cpu_baseline_block(budget_s=20.0):
    cores = os.cpu_count() or 1
    feats, labels = cpu_inputs(N_GLOBAL, DIM)
    t_probe = cpu_sample(feats, labels, 128)
    rows = int(max(128, min(2048, (budget_s / max(t_probe, 1e-3)) * 128 // 128 * 128)))
    t = cpu_sample(feats, labels, rows)
    return {"value": rows / t, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"rows [0,{rows}) of the N={N_GLOBAL} x 3-modality fwd+bwd step against all columns, "
                      f"numpy fp32 oracle port (oracle/loss_oracle.py:row_block_fwd_bwd), {t:.2f} s; "
                      f"full step = N/rows such blocks; the reference itself needs >= 12 [N,N] fp32 matrices "
                      f"(>= 48 GB) at this N and cannot run"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    feats, labels = cpu_inputs(N_GLOBAL, DIM)
    t_probe = cpu_sample(feats, labels, 128)
    total = args.steps + args.warmup
    rows = int(max(128, min(2048, (90.0 / total / max(t_probe, 1e-3)) * 128 // 128 * 128)))
    for _ in range(args.warmup):
        cpu_sample(feats, labels, rows)
    times = [cpu_sample(feats, labels, rows) for _ in range(args.steps)]
    t = sum(times) / len(times)
    v = rows / t
    sample = (f"each step = rows [0,{rows}) of the N={N_GLOBAL} image+DNA+text fwd+bwd against all columns "
              f"(numpy fp32 oracle port of loss_func.py:41-69), {t:.2f} s per step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3 * (N_GLOBAL / rows),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(world):
    return {"workload": "image+DNA+text contrastive loss fwd+bwd, label-matched multi-positive targets "
                        "(labels ~ randint(0, N/8)), logit_scale 1/0.07",
            "global_batch": N_GLOBAL, "dim": DIM, "modalities": 3, "pairs": 3, "rows_per_gpu": N_GLOBAL // world,
            "parallelism": f"row-block x{world}", "operands": "bf16 (tcgen05, fp32 accumulate)",
            "l2": "per-step working set ~1 GB (operand copies, class sums, gradients) exceeds the 126 MB L2; "
                  "no explicit flush"}


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import clibd_b200 as cb
    from clibd_b200 import _lib
    from clibd_b200 import retrieval as R

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---------------- loss workload
    N, d = N_GLOBAL, DIM
    n = N // world
    gen = torch.Generator().manual_seed(1234 + rank)
    host = [torch.randn(n, d, generator=gen).bfloat16().pin_memory() for _ in range(3)]
    host_labels = torch.randint(0, N // 8, (n,), generator=gen).pin_memory()
    if world > 1:
        module = cb.ClipLoss(local_loss=False, gather_with_grad=True, rank=rank, world_size=world)
    else:
        module = cb.ContrastiveLoss(None, 1 / 0.07)
    scale = torch.tensor(1 / 0.07, device=dev)
    resident = [h.to(dev) for h in host]
    resident_labels = host_labels.to(dev)

    def step_resident():
        leaves = [r.detach().requires_grad_(True) for r in resident]
        loss = module(leaves[0], leaves[1], leaves[2], resident_labels, scale)
        loss.backward()
        return loss

    # End-to-end step: what a training loop with a prefetching loader does.  Every step copies ITS inputs from
    # pinned host memory into one of two device slots on a copy stream (issued one step ahead, so the copy of
    # step k+1 overlaps the kernels of step k) and reads the PREVIOUS step's loss on the host (the last one is
    # read by the synchronize that closes the timed region): K copies and K loss reads per K timed steps.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [{"feats": [torch.empty_like(r) for r in resident], "labels": torch.empty_like(resident_labels),
              "ready": None, "free": None} for _ in range(2)]
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_evt = [None, None]
    pipe = {"k": 0, "last": None}

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            if slot["free"] is not None:
                copy_stream.wait_event(slot["free"])  # the step that used this slot has finished reading it
            for dst, src in zip(slot["feats"], host):
                dst.copy_(src, non_blocking=True)
            slot["labels"].copy_(host_labels, non_blocking=True)
            slot["ready"] = copy_stream.record_event()

    def step_e2e():
        k = pipe["k"]
        cur, nxt = slots[k % 2], slots[(k + 1) % 2]
        if cur["ready"] is None:
            issue_copy(cur)  # very first call: nothing was prefetched yet
        stream = torch.cuda.current_stream(dev)
        stream.wait_event(cur["ready"])
        issue_copy(nxt)
        leaves = [f.detach().requires_grad_(True) for f in cur["feats"]]
        loss = module(leaves[0], leaves[1], leaves[2], cur["labels"], scale)
        loss.backward()
        cur["free"] = stream.record_event()
        loss_host[k % 2].copy_(loss.detach(), non_blocking=True)  # device -> host read of the step's result
        loss_evt[k % 2] = stream.record_event()
        if loss_evt[(k + 1) % 2] is not None:
            loss_evt[(k + 1) % 2].synchronize()
            pipe["last"] = float(loss_host[(k + 1) % 2])
        pipe["k"] = k + 1
        return pipe["last"]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms)

    sampler = ClockSampler(local)
    lib.clibd_profile_enable(0)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    lib.clibd_profile_enable(1)
    launches0 = lib.clibd_kernel_launch_count()
    sampler.start()
    ms_total = timed(step_resident, args.steps, 0)
    launches = lib.clibd_kernel_launch_count() - launches0
    lib.clibd_profile_enable(0)
    prof_ms = (ctypes.c_double * 8)()
    prof_n = (ctypes.c_int64 * 8)()
    lib.clibd_profile_read(prof_ms, prof_n)
    ms_step = ms_total / args.steps
    value = N / (ms_step * 1e-3)

    ms_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 2)) / args.steps
    clocks = sampler.stop()  # sampled (every 100 ms) across both timed regions: the resident and the end-to-end loop
    h2d = sum(h.numel() * h.element_size() for h in host) + host_labels.numel() * 8
    e2e = {"value": N / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "ms_per_step": ms_e2e}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" \
        if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"

    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
    except Exception:  # noqa: BLE001
        pass

    def dram_traffic(key):
        # ncu-measured DRAM bytes per launch; only valid for the configuration it was captured on
        t = traffic.get(key)
        if t and t.get("config") == f"N={N} d={d} n_gpus={world}":
            return t["bytes_per_launch"]
        return None

    def roof(slot, flops_per_launch, name, executed_mult=1.0):
        if prof_n[slot] == 0:
            return None
        avg_ms = prof_ms[slot] / prof_n[slot]
        ach = flops_per_launch / (avg_ms * 1e-3) / 1e12
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf, "traffic": dram_traffic(name), "avg_launch_ms": avg_ms,
                "launches": int(prof_n[slot]),
                "algorithmic_flops_per_launch": flops_per_launch,
                # tensor flops the kernel really executes (the sweep also computes S, which SURVEY 8d does not
                # credit): what the tensor pipe sees
                "executed_flops_per_launch": flops_per_launch * executed_mult,
                "frac_executed": ach * executed_mult / peak_tf, "peak_source": peak_src,
                "share_of_step": prof_ms[slot] / (ms_total if ms_total > 0 else 1)}

    # algorithmic work: forward 2*n*N*d per unordered pair launch; backward 4*n*N*d per unordered
    # pair = 2*n*N*d per ordered-sweep launch (the S recompute is NOT credited)
    roofline = roof(1, 2.0 * n * N * d, "loss_bwd_pair_kernel", executed_mult=2.0)
    roofline_fwd = roof(0, 2.0 * n * N * d, "loss_fwd_pair_kernel")
    # single-GPU backward: the other side's gradient of every pair is a plain GEMM over the stored coefficient
    # strip (2*n*N*d per pair, no S recompute); absent when the rows are sharded (two sweeps per pair instead)
    roofline_grad = roof(4, 2.0 * n * N * d * (prof_n[1] / prof_n[4]) if prof_n[4] else 0.0, "loss_grad_gemm_kernel")
    step_frac = (18.0 * n * N * d) / (ms_step * 1e-3) / 1e12 / peak_tf
    burst_tf = peaks.get("bf16_tflops")

    out = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "roofline_fwd": roofline_fwd, "roofline_grad": roofline_grad, "step_tensor_frac_algorithmic": step_frac,
        # the same 18*n*N*d algorithmic flops of the whole step against the burst cuBLAS figure, for reference
        "step_tensor_frac_algorithmic_vs_burst": (step_frac * peak_tf / burst_tf) if burst_tf else None,
    }

    # ---------------- kNN workload
    if not args.no_knn:
        try:
            out["knn"] = bench_knn(torch, dist, R, lib, dev, rank, world, timed, peak_tf)
        except Exception as ex:  # noqa: BLE001
            out["knn"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_cpu:  # reported baseline: rank 0, single-GPU run only
        try:
            out["cpu_baseline"] = cpu_baseline_block()
        except Exception as ex:  # noqa: BLE001
            out["cpu_baseline"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_knn(torch, dist, R, lib, dev, rank, world, timed, peak_tf):
    Q, K, d, k = KNN_Q, KNN_K, DIM, KNN_TOPK
    per = (K + world - 1) // world
    lo, hi = min(K, rank * per), min(K, (rank + 1) * per)
    gen = torch.Generator(device=dev).manual_seed(77)
    n_species = 50_000
    cent = torch.randn(n_species, d, device=dev, generator=gen) / d ** 0.5
    # every rank draws the same global species assignment, then keeps its shard
    sp_k = torch.randint(0, n_species, (K,), device=dev, generator=gen)
    sp_q = torch.randint(0, n_species, (Q,), device=dev, generator=gen)
    gen2 = torch.Generator(device=dev).manual_seed(1000 + rank)
    keys = cent[sp_k[lo:hi]] + 0.02 * torch.randn(hi - lo, d, device=dev, generator=gen2)
    if hi - lo > 4000:
        keys[2000:3000] = keys[0:1000]  # exact duplicates: the tie-break is exercised
    genq = torch.Generator(device=dev).manual_seed(2000)
    queries = cent[sp_q] + 0.02 * torch.randn(Q, d, device=dev, generator=genq)
    del cent
    q32 = R.normalize_rows(queries, dev)
    k32 = R.normalize_rows(keys, dev)
    state = {}

    def step_resident():
        s64, idx, nex = R.search_normalized(q32, k32, k, key_offset=lo, mode="fp16")
        if world > 1:
            all_s = torch.empty((world,) + tuple(s64.shape), dtype=torch.float64, device=dev)
            all_i = torch.empty((world,) + tuple(idx.shape), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(all_s, s64)
            dist.all_gather_into_tensor(all_i, idx)
            _, _, idx = R.merge_topk(all_s, all_i)
        state["idx"], state["nex"] = idx, nex

    host_q = queries.cpu().pin_memory()
    host_k = keys.cpu().pin_memory()
    del queries, keys

    def step_e2e():
        # host-resident queries and keys: the key set is copied block-wise on a copy stream while the previous
        # block is normalised and screened (what knn_search / make_prediction do for host inputs)
        if world == 1:
            _, idx = R.knn_search(host_q, host_k, k, mode="fp16", device=dev)
            return idx.cpu()
        qd = R.normalize_rows(host_q, dev)
        s64, idx = R._search_host_keys_pipelined(qd, host_k, 0, host_k.shape[0], k, "fp16", dev, index_base=lo)
        if world > 1:
            all_s = torch.empty((world,) + tuple(s64.shape), dtype=torch.float64, device=dev)
            all_i = torch.empty((world,) + tuple(idx.shape), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(all_s, s64)
            dist.all_gather_into_tensor(all_i, idx)
            _, _, idx = R.merge_topk(all_s, all_i)
        return idx.cpu()

    steps, warm = 3, 1
    lib.clibd_profile_enable(1)
    ms = timed(step_resident, steps, warm) / steps
    lib.clibd_profile_enable(0)
    prof_ms = (ctypes.c_double * 8)()
    prof_n = (ctypes.c_int64 * 8)()
    lib.clibd_profile_read(prof_ms, prof_n)
    ms_e2e = timed(step_e2e, 2, 1) / 2
    scr = None
    if prof_n[2]:
        avg = prof_ms[2] / prof_n[2]
        ach = 2.0 * Q * (hi - lo) * d / (avg * 1e-3) / 1e12
        scr = {"kernel": "knn_screen_tc_kernel", "bound": "tensor", "achieved": ach, "peak": peak_tf,
               "unit": "TFLOP/s", "frac": ach / peak_tf, "avg_launch_ms": avg, "launches": int(prof_n[2])}
    return {"metric": "kNN queries/s vs 1M keys", "value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
            "steps": steps, "warmup": warm,
            "config": {"workload": "cosine top-5 retrieval, class-centroid + noise embeddings, 1000 exact duplicate "
                                   "keys per shard", "queries": Q, "keys": K, "dim": d, "k": k,
                       "keys_per_gpu": hi - lo, "operands": "f16 screen + float64 re-rank"},
            "e2e": {"value": Q / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": host_q.numel() * 4 + host_k.numel() * 4, "d2h_bytes_per_step": Q * k * 8},
            "queries_redone_exhaustively": int(state["nex"]), "roofline": scr,
            "rerank_ms": (prof_ms[3] / prof_n[3]) if prof_n[3] else None}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
