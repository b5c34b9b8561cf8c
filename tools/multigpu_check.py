"""Multi-GPU parity check, one process per GPU (launched by tests/test_multigpu.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P \
        tools/multigpu_check.py

Loss: ClipLoss(world_size=W) on row shards (NCCL all-gather / all-reduce) must return the same loss as the
single-process ContrastiveLoss on the concatenated batch, and each rank's gradient must be W x its slice of the
full-batch gradient (the reduce-scatter(SUM) convention of torch.distributed.nn.all_gather's backward,
reference loss_func.py:97).  kNN: keys sharded contiguously, per-rank top-k all-gathered and merged by
(-sim, index) must be bit-identical to the unsharded search.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clibd_b200 as cb  # noqa: E402
from clibd_b200 import retrieval as R  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # every exchange form of the sharded step (clibd_b200/loss.py:_shard_mode): 'local' recomputes S for both gradients
    # on every rank, 'nccl' computes S once and reduce-scatters the column-side partials with NCCL, 'peer' does the
    # all-gather, the statistics exchange and that reduce-scatter as stores into peer-mapped memory over NVLink
    cases = [(192, 96, 3, torch.float32, None, 2e-5),
             (512, 768, 3, torch.bfloat16, None, 6e-3),  # both sides round the gradient to bf16 (2^-9)
             (333, 768, 2, torch.float16, None, 2e-3),
             (1024, 768, 3, torch.float32, "bf16", 1e-3),  # fp32 leaves: no output rounding
             (2048, 768, 3, torch.float32, "fp16", 1e-3)]
    for shard_mode in ("local", "nccl", "peer", "peer+own_barrier"):
        os.environ["CLIBD_SHARD_MODE"] = shard_mode.split("+")[0]
        os.environ["CLIBD_BARRIER"] = "own" if shard_mode.endswith("own_barrier") else "torch"  # clibd_shard_barrier
        for (n, d, nmod, dtype, operands, tol) in cases:
            N = n * world
            gen = torch.Generator().manual_seed(11)
            full = [torch.randn(N, d, generator=gen).to(dtype).to(dev) for _ in range(nmod)] + [None] * (3 - nmod)
            labels = torch.randint(0, max(1, N // 8), (N,), generator=gen).to(dev)
            scale_full = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
            leaves_full = [None if f is None else f.clone().requires_grad_(True) for f in full]
            loss_full = cb.ContrastiveLoss(None, 1 / 0.07, tensor_core_operands=operands)(
                leaves_full[0], leaves_full[1], leaves_full[2], labels, scale_full)
            loss_full.backward()
            sl = slice(rank * n, (rank + 1) * n)
            mod = cb.ClipLoss(local_loss=False, gather_with_grad=True, rank=rank, world_size=world,
                              tensor_core_operands=operands)
            # two steps through the same module: the second one reuses the exchange buffers of the first
            for it in range(2):
                leaves = [None if f is None else f[sl].clone().requires_grad_(True) for f in full]
                scale = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
                loss = mod(leaves[0], leaves[1], leaves[2], labels[sl], scale)
                g_out = 3.0 + rank  # a different upstream gradient per rank: the local rows receive the SUM
                (loss * g_out).backward()
            torch.cuda.synchronize()
            g_sum = sum(3.0 + r for r in range(world))
            e_loss = abs(float(loss.detach()) - float(loss_full.detach())) / abs(float(loss_full.detach()))
            errs = []
            for lf, ll in zip(leaves_full, leaves):
                if lf is None:
                    continue
                ref = lf.grad[sl].float() * g_sum
                errs.append(float((ll.grad.float() - ref).norm() / ref.norm()))
            # d loss / d logit_scale is replicated: every rank holds the full-batch derivative times ITS grad_output
            e_ds = abs(float(scale.grad) - g_out * float(scale_full.grad)) / abs(g_out * float(scale_full.grad))
            good = e_loss < 1e-5 and max(errs) < tol and e_ds < 1e-3
            ok &= good
            print(f"[rank {rank}] {shard_mode:16s} loss n={n} d={d} nmod={nmod} {str(dtype)[6:]} operands={operands}: "
                  f"loss_err={e_loss:.2e} grad_err={max(errs):.2e} dscale_err={e_ds:.2e} {'OK' if good else 'FAIL'}",
                  flush=True)
    os.environ.pop("CLIBD_SHARD_MODE", None)
    os.environ.pop("CLIBD_BARRIER", None)

    # ---- gather_features (loss_func.py:73-106): the peer-memory form against torch's collectives, values and gradients
    # (gather_with_grad: reduce-scatter(SUM) of the gathered gradient; without: only the local rows carry a gradient)
    for dtype, n, d in ((torch.float32, 130, 96), (torch.bfloat16, 257, 768)):
        gen = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(n, d, generator=gen).to(dtype).to(dev)
        wgt = torch.randn(n * world, d, generator=torch.Generator().manual_seed(7)).to(dtype).to(dev) * (1 + rank)
        for with_grad in (True, False):
            res = {}
            for form in ("nccl", "peer"):
                os.environ["CLIBD_SHARD_MODE"] = form
                for it in range(2):  # the second call reuses the receive buffer
                    leaf = x.clone().requires_grad_(True)
                    allf = cb.gather_features(leaf, local_loss=False, gather_with_grad=with_grad, rank=rank,
                                              world_size=world)
                    (allf.float() * wgt.float()).sum().backward()
                res[form] = (allf.detach().clone(), leaf.grad.clone())
            os.environ.pop("CLIBD_SHARD_MODE", None)
            good = bool(torch.equal(res["peer"][0], res["nccl"][0]))
            gerr = float((res["peer"][1].float() - res["nccl"][1].float()).abs().max() / res["nccl"][1].float().abs().max())
            gtol = 0.0 if not with_grad else (1e-6 if dtype == torch.float32 else 2e-2)  # NCCL's own summation order
            good = good and gerr <= gtol and bool(torch.equal(allf[rank * n:(rank + 1) * n], x))
            ok &= good
            print(f"[rank {rank}] gather_features {str(dtype)[6:]} n={n} with_grad={with_grad}: values equal, "
                  f"rel grad diff {gerr:.2e}: {'OK' if good else 'FAIL'}", flush=True)

    # ---- kNN: sharded keys + all-gather + merge == unsharded
    Q, K, d, k = 700, 20011, 768, 5
    gen = torch.Generator().manual_seed(5)
    keys = torch.randn(K, d, generator=gen)
    keys[9000:9500] = keys[100:600]  # exact duplicates straddling the shard boundaries
    keys[K - 300:] = keys[200:500]
    q = keys[torch.randint(0, K, (Q,), generator=gen)] + 0.05 * torch.randn(Q, d, generator=gen)
    q32, k32 = R.normalize_rows(q.to(dev), dev), R.normalize_rows(keys.to(dev), dev)
    s_all, i_all, _ = R.search_normalized(q32, k32, k, mode="fp16")
    per = (K + world - 1) // world
    lo, hi = min(K, rank * per), min(K, (rank + 1) * per)
    s, i, _ = R.search_normalized(q32, k32[lo:hi].contiguous(), k, key_offset=lo, mode="fp16")
    all_s = torch.empty((world,) + tuple(s.shape), dtype=torch.float64, device=dev)
    all_i = torch.empty((world,) + tuple(i.shape), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_s, s)
    dist.all_gather_into_tensor(all_i, i)
    m64, _, mi = R.merge_topk(all_s, all_i)
    good = bool(torch.equal(mi, i_all) and torch.equal(m64, s_all))
    ok &= good
    print(f"[rank {rank}] knn shard merge bit-exact: {'OK' if good else 'FAIL'}", flush=True)

    # ---- the public sharded entry points: knn_search / make_prediction / inference_and_print_result with
    # shard_keys=True (every rank passes the same host arrays, searches its share, merges over NCCL) must return
    # exactly what the unsharded call returns
    import numpy as np
    knp, qnp = keys.numpy(), q.numpy()
    s_one, i_one = R.knn_search(qnp, knp, k, mode="fp16", device=dev)
    s_sh, i_sh = R.knn_search(qnp, knp, k, mode="fp16", device=dev, shard_keys=True)
    good = bool(torch.equal(i_sh, i_one) and torch.equal(s_sh, s_one) and torch.equal(i_one, i_all))
    ok &= good
    print(f"[rank {rank}] knn_search(shard_keys=True) == unsharded: {'OK' if good else 'FAIL'}", flush=True)
    rng = np.random.default_rng(3)
    Kk, Dd = 1501, 64   # not divisible by the world size; 64-d features
    sp = rng.integers(0, 40, Kk)
    cent = rng.standard_normal((40, Dd)).astype(np.float32)
    kf = (cent[sp] + 0.3 * rng.standard_normal((Kk, Dd))).astype(np.float32)
    kf[700:760] = kf[10:70]  # duplicates across the shard boundary
    def lab(ids):
        return [{"order": f"o{i % 3}", "family": f"f{i % 7}", "genus": f"g{i % 20}", "species": f"s{i}"} for i in ids]
    qs, qu = rng.integers(0, 40, 90), rng.integers(0, 40, 70)
    keys_dict = {"label_list": lab(sp), "encoded_image_feature": kf, "encoded_dna_feature": kf[::-1].copy(),
                 "all_key_features": np.concatenate([kf, kf[::-1]], 0), "all_key_features_label": lab(sp) + lab(sp[::-1])}
    seen = {"label_list": lab(qs), "file_name_list": list(range(90)),
            "encoded_image_feature": (cent[qs] + 0.3 * rng.standard_normal((90, Dd))).astype(np.float32),
            "encoded_dna_feature": (cent[qs] + 0.3 * rng.standard_normal((90, Dd))).astype(np.float32)}
    unseen = {"label_list": lab(qu), "file_name_list": list(range(70)),
              "encoded_image_feature": (cent[qu] + 0.5 * rng.standard_normal((70, Dd))).astype(np.float32),
              "encoded_dna_feature": (cent[qu] + 0.5 * rng.standard_normal((70, Dd))).astype(np.float32)}
    one = cb.inference_and_print_result(keys_dict, seen, unseen, args=None, k_list=[1, 3, 5], verbose=False, device=dev)
    sh = cb.inference_and_print_result(keys_dict, seen, unseen, args=None, k_list=[1, 3, 5], verbose=False, device=dev,
                                       shard_keys=True)
    good = one == sh
    ok &= good
    print(f"[rank {rank}] inference_and_print_result(shard_keys=True) == unsharded (acc, per-class, predictions): "
          f"{'OK' if good else 'FAIL'}", flush=True)
    p_one = cb.make_prediction(seen["encoded_image_feature"], kf, keys_dict["label_list"], with_indices=True, max_k=5)
    p_sh = cb.make_prediction(seen["encoded_image_feature"], kf, keys_dict["label_list"], with_indices=True, max_k=5,
                              shard_keys=True)
    good = p_one[0] == p_sh[0] and np.array_equal(p_one[1], p_sh[1])
    ok &= good
    print(f"[rank {rank}] make_prediction(shard_keys=True) == unsharded: {'OK' if good else 'FAIL'}", flush=True)

    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag) != 1:
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
