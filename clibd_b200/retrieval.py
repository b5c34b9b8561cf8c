"""Drop-in replacements for the retrieval / accuracy helpers of bioscanclip/util/util.py,
backed by the sm_100a CUDA library (csrc/knn.cu).

  * ``make_prediction``              util.py:521-553
  * ``find_closest_match``           util.py:759-789
  * ``top_k_micro_accuracy``         util.py:379-395
  * ``top_k_macro_accuracy``         util.py:555-599
  * ``inference_and_print_result``   util.py:601-700

The search replaces sklearn-normalise + faiss ``IndexFlatIP`` (util.py:522-528): exact inner
product top-k with ties broken by LOWEST key index (BASELINE.json north_star).  Keys can be
sharded over the ranks of a torch.distributed group; per-rank top-k lists are all-gathered
and merged with the same (-similarity, index) order, so every rank returns the global result.

No CPU fallback: without a CUDA device (or the built library) these functions raise.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

LEVELS = ["order", "family", "genus", "species"]  # util.py:25
All_TYPE_OF_FEATURES_OF_QUERY = [  # util.py:26-32
    "encoded_image_feature",
    "encoded_dna_feature",
    "encoded_language_feature",
    "averaged_feature",
    "concatenated_feature",
]
All_TYPE_OF_FEATURES_OF_KEY = [  # util.py:33-40
    "encoded_image_feature",
    "encoded_dna_feature",
    "encoded_language_feature",
    "averaged_feature",
    "concatenated_feature",
    "all_key_features",
]

_PATHS = {"exact": _lib.PATH_SIMT_F32, "bf16": _lib.PATH_TC_BF16, "fp16": _lib.PATH_TC_F16}


def _device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("clibd_b200 retrieval needs a CUDA device (there is no CPU fallback)")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def normalize_rows(x, device=None) -> torch.Tensor:
    """sklearn.preprocessing.normalize(x, norm="l2", axis=1).astype(np.float32) on the GPU
    (util.py:523-524): norm and division in float64, result float32 [n, d] on `device`."""
    device = _device(device)
    lib = _lib.load()
    if isinstance(x, torch.Tensor):
        t = x.detach()
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if t.dim() != 2:
        raise ValueError("features must be a 2-D [n, d] array")
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64 if t.dtype in (torch.int64, torch.int32) else torch.float32)
    t = t.to(device, non_blocking=True).contiguous()
    n, d = t.shape
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.clibd_knn_normalize(t.data_ptr(), _lib.DT_F64 if t.dtype == torch.float64 else _lib.DT_F32,
                                           n, d, out.data_ptr(), _stream(device)))
    return out


def search_normalized(q32: torch.Tensor, k32: torch.Tensor, k: int, key_offset: int = 0, mode: str = "fp16"):
    """Exact top-k of already-normalised float32 device tensors.
    Returns (sims64 [Q,k] float64, idx [Q,k] int64, n_exhaustive int32[1]) device tensors."""
    lib = _lib.load()
    device = q32.device
    Q, d = q32.shape
    K = k32.shape[0]
    path = _PATHS[mode]
    nbytes = lib.clibd_knn_scratch_bytes(Q, K, d, k, path)
    if nbytes < 0:
        raise ValueError("clibd_b200: bad kNN shape")
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
    sims64 = torch.empty((Q, k), dtype=torch.float64, device=device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=device)
    nex = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.clibd_knn_search(q32.data_ptr(), Q, k32.data_ptr(), K, key_offset, d, k, path,
                                        scratch.data_ptr(), nbytes, sims64.data_ptr(), idx.data_ptr(),
                                        nex.data_ptr(), _stream(device)))
    return sims64, idx, nex


def merge_topk(sims64_parts: torch.Tensor, idx_parts: torch.Tensor):
    """[parts, Q, k] per-shard results -> global (sims64, sims32, idx) by (-sim, index)."""
    lib = _lib.load()
    parts, Q, k = sims64_parts.shape
    device = sims64_parts.device
    out64 = torch.empty((Q, k), dtype=torch.float64, device=device)
    out32 = torch.empty((Q, k), dtype=torch.float32, device=device)
    oidx = torch.empty((Q, k), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.clibd_knn_merge(sims64_parts.contiguous().data_ptr(), idx_parts.contiguous().data_ptr(), parts,
                                       Q, k, out64.data_ptr(), out32.data_ptr(), oidx.data_ptr(), _stream(device)))
    return out64, out32, oidx


_PIPELINE_MIN_KEYS = 1 << 17   # host-resident key sets at least this large are copied and searched block-wise
_PIPELINE_FIRST_KEYS = 1 << 15  # first block: its copy (100 MB at d = 768, ~2 ms) is the only one nothing hides
_PIPELINE_GROWTH = 2           # a block's copy runs next to the previous block's screen, which takes ~2x as long per key
_PIPELINE_BLOCKS = 6           # every block pays its own re-rank pass (~2.4 ms at 100k queries): keep them few


def _pipeline_bounds(lo: int, hi: int, blocks=None):
    """Block boundaries of the pipelined host-key search.  Default: geometrically growing blocks -- a small first block
    (its copy is exposed), every later one as large as the copy the previous block's compute can hide, the last takes
    the rest.  Equal blocks (tools/knn_blocks.py, 100k x 1M: 2 -> 157 ms, 4 -> 155, 8 -> 161, 16 -> 194) expose a
    quarter of the key copy.  `blocks` given: that many equal blocks."""
    total = hi - lo
    if blocks is not None:
        blocks = max(1, min(blocks, total))
        return [lo + total * b // blocks for b in range(blocks + 1)]
    bounds, size = [lo], _PIPELINE_FIRST_KEYS
    while len(bounds) < _PIPELINE_BLOCKS and hi - bounds[-1] > 2 * size:
        bounds.append(bounds[-1] + size)
        size *= _PIPELINE_GROWTH
    bounds.append(hi)
    return bounds


def normalize_queries_sharded(query_feature, device, world: int, rank: int, process_group=None) -> torch.Tensor:
    """Normalised float32 queries on every rank of a sharded search.  HOST-resident queries are cut into `world`
    row blocks: every rank copies and normalises only its block (1 / world of the host -> device traffic, which is
    what bounds the end-to-end search once the keys are sharded) and the blocks are all-gathered over NVLink."""
    host = _as_host_tensor(query_feature)
    nq = query_feature.shape[0]
    if world == 1 or host is None or nq < world * 1024:
        return normalize_rows(query_feature, device)
    import torch.distributed as dist
    per = (nq + world - 1) // world
    lo, hi = min(nq, rank * per), min(nq, (rank + 1) * per)
    d = host.shape[1]
    mine = torch.zeros((per, d), dtype=torch.float32, device=device)
    if hi > lo:
        mine[:hi - lo] = normalize_rows(host[lo:hi], device)
    out = torch.empty((world * per, d), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(out, mine, group=process_group)
    return out[:nq]


def _as_host_tensor(x):
    """x as a CPU float tensor without copying, or None when it already lives on a device."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return None
        t = x.detach()
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype not in (torch.float32, torch.float64):
        return None
    return t.contiguous()


def _search_block(q32, k32, k, key_offset, mode):
    """top-k of one normalised key block, padded with empty slots to k columns."""
    device = q32.device
    kk = min(k, k32.shape[0])
    s64, idx, _ = search_normalized(q32, k32, kk, key_offset=key_offset, mode=mode)
    if kk < k:
        pad_s = torch.full((q32.shape[0], k - kk), -torch.finfo(torch.float64).max, dtype=torch.float64, device=device)
        pad_i = torch.full((q32.shape[0], k - kk), -1, dtype=torch.int64, device=device)
        s64, idx = torch.cat([s64, pad_s], 1), torch.cat([idx, pad_i], 1)
    return s64, idx


def _search_host_keys_pipelined(q32, host_keys, lo, hi, k, mode, device, blocks=None, index_base=0):
    """Keys [lo, hi) live in host memory: copy them block by block on a copy stream into two staging buffers
    while the previous block is normalised and searched, then merge the per-block lists by (-sim, index) --
    the same order the shard merge uses, so the result equals the one-shot search bit for bit.  Hides the
    host->device copy of the key set (3 GB for 1M x 768 float32) behind the tensor-core screen.
    Returned indices are index_base + the row number inside host_keys."""
    bounds = _pipeline_bounds(lo, hi, blocks)
    blocks = len(bounds) - 1
    longest = max(bounds[b + 1] - bounds[b] for b in range(blocks))
    cur = torch.cuda.current_stream(device)
    copy_stream = torch.cuda.Stream(device=device)
    stage = [torch.empty((longest, host_keys.shape[1]), dtype=host_keys.dtype, device=device) for _ in range(2)]
    ready, free = [None, None], [None, None]
    copy_stream.wait_stream(cur)  # the staging buffers exist before the first copy lands in them

    def issue(b):
        sl = b % 2
        with torch.cuda.stream(copy_stream):
            if free[sl] is not None:
                copy_stream.wait_event(free[sl])
            stage[sl][:bounds[b + 1] - bounds[b]].copy_(host_keys[bounds[b]:bounds[b + 1]], non_blocking=True)
            ready[sl] = copy_stream.record_event()

    issue(0)
    parts_s, parts_i = [], []
    for b in range(blocks):
        if b + 1 < blocks:
            issue(b + 1)
        sl = b % 2
        cur.wait_event(ready[sl])
        k32 = normalize_rows(stage[sl][:bounds[b + 1] - bounds[b]], device)
        free[sl] = cur.record_event()  # the staging buffer has been read
        s64, idx = _search_block(q32, k32, k, index_base + bounds[b], mode)
        parts_s.append(s64)
        parts_i.append(idx)
        del k32
    s64, _, idx = merge_topk(torch.stack(parts_s), torch.stack(parts_i))
    return s64, idx


def _shard_of(nk: int, shard_keys: bool, process_group):
    """(world, rank, lo, hi): the contiguous key range [lo, hi) this rank searches."""
    world, rank = 1, 0
    if shard_keys:
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("shard_keys=True needs an initialised torch.distributed process group")
        world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
    per = (nk + world - 1) // world
    return world, rank, min(nk, rank * per), min(nk, (rank + 1) * per)


def _merge_over_ranks(s64, idx, world, process_group):
    """per-rank top-k lists (global indices) -> the global top-k on every rank: one all-gather + clibd_knn_merge."""
    if world == 1:
        return s64, s64.to(torch.float32), idx
    import torch.distributed as dist
    all_s = torch.empty((world,) + tuple(s64.shape), dtype=torch.float64, device=s64.device)
    all_i = torch.empty((world,) + tuple(idx.shape), dtype=torch.int64, device=s64.device)
    dist.all_gather_into_tensor(all_s, s64.contiguous(), group=process_group)
    dist.all_gather_into_tensor(all_i, idx.contiguous(), group=process_group)
    return merge_topk(all_s, all_i)


def _empty_lists(nq, k, device):
    return (torch.full((nq, k), -torch.finfo(torch.float64).max, dtype=torch.float64, device=device),
            torch.full((nq, k), -1, dtype=torch.int64, device=device))


def knn_search(query_feature, keys_feature, k: int, mode: str = "fp16", device=None, process_group=None,
               shard_keys: bool = False):
    """normalise + search (+ shard merge).  Returns (similarities float32 [Q,k], indices int64 [Q,k])
    as DEVICE tensors, sorted by descending similarity, lowest index first on ties.
    shard_keys=True: every rank of `process_group` passes the same arguments, searches its contiguous share of the
    keys and receives the merged global result (keys sharded over the GPUs, BASELINE.json north_star)."""
    device = _device(device)
    nk = keys_feature.shape[0]
    if k > nk:
        raise ValueError("max_k is larger than the number of keys")
    world, rank, lo, hi = _shard_of(nk, shard_keys, process_group)
    q32 = normalize_queries_sharded(query_feature, device, world, rank, process_group)
    if hi > lo:
        host_keys = _as_host_tensor(keys_feature)
        if host_keys is not None and hi - lo >= _PIPELINE_MIN_KEYS:
            s64, idx = _search_host_keys_pipelined(q32, host_keys, lo, hi, k, mode, device)
        else:
            s64, idx = _search_block(q32, normalize_rows(keys_feature[lo:hi], device), k, lo, mode)
    else:
        s64, idx = _empty_lists(q32.shape[0], k, device)
    _, s32, idx = _merge_over_ranks(s64, idx, world, process_group)
    return s32, idx


def _labels_from_indices(indices: np.ndarray, keys_label):
    pred_list = []
    for key_indices in indices:
        pred = {}
        for level in LEVELS:
            pred[level] = [keys_label[int(i)][level] for i in key_indices]  # KeyError/IndexError propagate
        pred_list.append(pred)
    return pred_list


def make_prediction(query_feature, keys_feature, keys_label, with_similarity=False, with_indices=False, max_k=5,
                    mode="fp16", device=None, process_group=None, shard_keys=False):
    """util.py:521-553.  Returns pred_list, or [pred_list, similarities?, indices?] exactly like the
    reference (similarities float32 [Q,k], indices int64 [Q,k] numpy arrays).  A bad label lookup raises
    instead of the reference's print + sys.exit(1) (util.py:537-540)."""
    s32, idx = knn_search(query_feature, keys_feature, max_k, mode=mode, device=device,
                          process_group=process_group, shard_keys=shard_keys)
    indices = idx.cpu().numpy()
    pred_list = _labels_from_indices(indices, keys_label)
    out = [pred_list]
    if with_similarity:
        out.append(s32.cpu().numpy())
    if with_indices:
        out.append(indices)
    if len(out) == 1:
        return out[0]
    return out


def find_closest_match(query_feature, keys_feature, keys_label, with_similarity=False, with_indices=False, max_k=5,
                       mode="fp16", device=None, process_group=None, shard_keys=False):
    """util.py:759-789: the same search, dictionary output."""
    s32, idx = knn_search(query_feature, keys_feature, max_k, mode=mode, device=device, process_group=process_group,
                          shard_keys=shard_keys)
    indices = idx.cpu().numpy()
    out = {"pred_list": _labels_from_indices(indices, keys_label)}
    if with_similarity:
        out["similarities"] = s32.cpu().numpy()
    if with_indices:
        out["indices"] = indices
    return out


# ----------------------------------------------------------------------------------------------
# accuracy
# ----------------------------------------------------------------------------------------------
class _Vocab:
    """label -> dense id per level, ids issued in first-appearance order."""

    def __init__(self):
        self.maps = [dict() for _ in LEVELS]

    def encode(self, label_dicts) -> np.ndarray:
        out = np.empty((len(label_dicts), 4), dtype=np.int32)
        for li, level in enumerate(LEVELS):
            m = self.maps[li]
            col = out[:, li]
            for r, dct in enumerate(label_dicts):
                lab = dct[level]
                v = m.get(lab)
                if v is None:
                    v = len(m)
                    m[lab] = v
                col[r] = v
        return out

    def max_class(self):
        return max(1, max(len(m) for m in self.maps))

    def names(self, li):
        inv = [None] * len(self.maps[li])
        for lab, v in self.maps[li].items():
            inv[v] = lab
        return inv


def accuracy_counts(idx: torch.Tensor, key_ids: torch.Tensor, query_ids: torch.Tensor, k_list, max_class: int):
    """GPU counts: micro_hits int64 [nk,4]; class_hit / class_cnt int32 [nk,4,max_class]."""
    lib = _lib.load()
    device = idx.device
    Q, kmax = idx.shape
    nk = len(k_list)
    if nk > 4:
        raise ValueError("at most 4 k values per call")
    micro = torch.empty((nk, 4), dtype=torch.int64, device=device)
    chit = torch.empty((nk, 4, max_class), dtype=torch.int32, device=device)
    ccnt = torch.empty((nk, 4, max_class), dtype=torch.int32, device=device)
    # the reference slices pred[level][:k] (util.py:389, 573): a k beyond the number of retrieved neighbours looks at
    # all of them (k_list need not be sorted: max_k is its LAST entry, util.py:607)
    ks = (ctypes.c_int32 * nk)(*[min(int(k), int(kmax)) for k in k_list])
    with torch.cuda.device(device):
        _lib.check(lib.clibd_topk_accuracy(idx.contiguous().data_ptr(), Q, kmax, key_ids.contiguous().data_ptr(),
                                           key_ids.shape[0], query_ids.contiguous().data_ptr(), ks, nk, max_class,
                                           micro.data_ptr(), chit.data_ptr(), ccnt.data_ptr(), _stream(device)))
    return micro, chit, ccnt


def _accuracy_from_counts(micro, chit, ccnt, query_ids_np, k_list, vocab, n_query):
    """Reproduce the reference's float arithmetic from integer counts (util.py:391-393, 586-597)."""
    micro = micro.cpu().numpy()
    chit = chit.cpu().numpy()
    ccnt = ccnt.cpu().numpy()
    micro_acc, macro_acc, per_class = {}, {}, {}
    for a, k in enumerate(k_list):
        micro_acc[k], macro_acc[k], per_class[k] = {}, {}, {}
        for li, level in enumerate(LEVELS):
            micro_acc[k][level] = int(micro[a, li]) * 1.0 / n_query
            names = vocab.names(li)
            _, first = np.unique(query_ids_np[:, li], return_index=True)
            order = query_ids_np[np.sort(first), li]  # classes in first-appearance order, like the reference's dict
            total = 0
            pc = {}
            for cid in order:
                r = int(chit[a, li, cid]) * 1.0 / int(ccnt[a, li, cid])
                total = total + r
                pc[names[cid]] = r
            per_class[k][level] = pc
            macro_acc[k][level] = total / len(order)
    return micro_acc, macro_acc, per_class


def _encode_pred_gt(pred_list, gt_list):
    vocab = _Vocab()
    gt_ids = vocab.encode(gt_list)
    kmax = len(pred_list[0][LEVELS[0]]) if pred_list else 0
    flat = []
    for p in pred_list:
        for j in range(kmax):
            flat.append({level: p[level][j] for level in LEVELS})
    key_ids = vocab.encode(flat) if flat else np.zeros((0, 4), np.int32)
    idx = np.arange(len(pred_list) * kmax, dtype=np.int64).reshape(len(pred_list), kmax)
    return vocab, gt_ids, key_ids, idx


def _accuracy_ref_signature(pred_list, gt_list, k_list):
    device = _device()
    vocab, gt_ids, key_ids, idx = _encode_pred_gt(pred_list, gt_list)
    out = [None, None, None]
    micro_acc, macro_acc, per_class = {}, {}, {}
    for c0 in range(0, len(k_list), 4):
        ks = list(k_list[c0:c0 + 4])
        micro, chit, ccnt = accuracy_counts(torch.from_numpy(idx).to(device), torch.from_numpy(key_ids).to(device),
                                            torch.from_numpy(gt_ids).to(device), ks, vocab.max_class())
        mi, ma, pc = _accuracy_from_counts(micro, chit, ccnt, gt_ids, ks, vocab, len(pred_list))
        micro_acc.update(mi)
        macro_acc.update(ma)
        per_class.update(pc)
    out[0], out[1], out[2] = micro_acc, macro_acc, per_class
    return out


def top_k_micro_accuracy(pred_list, gt_list, k_list=None):
    """util.py:379-395 (same signature; labels are encoded to ids and counted on the GPU)."""
    return _accuracy_ref_signature(pred_list, gt_list, k_list)[0]


def top_k_macro_accuracy(pred_list, gt_list, k_list=None):
    """util.py:555-599 -> (macro_acc_dict, per_class_acc)."""
    if k_list is None:
        k_list = [1, 3, 5]
    _, macro, per_class = _accuracy_ref_signature(pred_list, gt_list, k_list)
    return macro, per_class


def print_micro_and_macro_acc(acc_dict, k_list, args=None):
    """Console table of util.py:397-519.  The reference's CSV / JSON / config dumps under
    args.project_root_path are reporting glue and out of scope."""
    for q_type, by_key in acc_dict.items():
        for k_type, by_split in by_key.items():
            for split in ("seen", "unseen"):
                if split not in by_split:
                    continue
                for kind in ("micro_acc", "macro_acc"):
                    for k in k_list:
                        row = by_split[split][kind][k]
                        cells = " ".join(f"{level}={row[level]:.4f}" for level in LEVELS)
                        print(f"{q_type} -> {k_type} [{split}] {kind}@{k}: {cells}")


def inference_and_print_result(keys_dict, seen_dict, unseen_dict, args=None, small_species_list=None, k_list=None,
                               mode="fp16", device=None, verbose=True, process_group=None, shard_keys=False):
    """util.py:601-700: every (query feature type, key feature type) pair of matching width is searched
    for the seen and unseen query sets; returns (acc_dict, per_class_acc, pred_dict) with the reference's
    schema.  Unlike the reference, the key set is normalised and staged once per key type (not once per
    search) and labels are compared as integer ids on the GPU.
    shard_keys=True (every rank of `process_group` calls this with the same arguments): each rank normalises and
    keeps only ITS contiguous share of every key type, searches it, and the per-rank candidate lists are merged by
    one all-gather + clibd_knn_merge per search -- every rank returns the same dictionaries, bit-identical to the
    unsharded call."""
    device = _device(device)
    acc_dict, per_class_acc = {}, {}
    if k_list is None:
        k_list = [1, 3, 5]
    max_k = k_list[-1]  # util.py:607
    seen_gt_label = seen_dict["label_list"]
    unseen_gt_label = unseen_dict["label_list"]
    keys_label = keys_dict["label_list"]
    try:
        pred_dict = {
            "seen_id": seen_dict["processed_id_list"],
            "seen_gt_label": seen_gt_label,
            "unseen_id": unseen_dict["processed_id_list"],
            "unseen_gt_label": unseen_gt_label,
        }
    except KeyError:
        pred_dict = {
            "seen_id": seen_dict.get("file_name_list", []),
            "seen_gt_label": seen_gt_label,
            "unseen_id": unseen_dict.get("file_name_list", []),
            "unseen_gt_label": unseen_gt_label,
        }
    vocab = _Vocab()
    gt_ids = {"seen": vocab.encode(seen_gt_label), "unseen": vocab.encode(unseen_gt_label)}
    key_id_cache = {}

    def key_ids_for(labels):
        if id(labels) not in key_id_cache:
            key_id_cache[id(labels)] = torch.from_numpy(vocab.encode(labels)).to(device)
        return key_id_cache[id(labels)]

    key_norm_cache = {}
    for query_feature_type in All_TYPE_OF_FEATURES_OF_QUERY:
        if query_feature_type not in seen_dict.keys():
            continue
        acc_dict[query_feature_type] = {}
        per_class_acc[query_feature_type] = {}
        pred_dict[query_feature_type] = {}
        for key_feature_type in All_TYPE_OF_FEATURES_OF_KEY:
            if key_feature_type not in keys_dict.keys():
                continue
            acc_dict[query_feature_type][key_feature_type] = {}
            per_class_acc[query_feature_type][key_feature_type] = {}
            pred_dict[query_feature_type][key_feature_type] = {}
            curr_seen_feature = seen_dict[query_feature_type]
            curr_unseen_feature = unseen_dict[query_feature_type]
            curr_keys_feature = keys_dict[key_feature_type]
            if curr_keys_feature is None:
                continue
            if key_feature_type == "all_key_features":
                keys_label = keys_dict["all_key_features_label"]  # sticks for later key types, like util.py:651-652
            if (curr_seen_feature is None or curr_unseen_feature is None
                    or curr_keys_feature.shape[-1] != curr_seen_feature.shape[-1]
                    or curr_keys_feature.shape[-1] != curr_unseen_feature.shape[-1]):
                continue
            nk = curr_keys_feature.shape[0]
            if max_k > nk:
                raise ValueError("max_k is larger than the number of keys")
            world, _, lo, hi = _shard_of(nk, shard_keys, process_group)
            if key_feature_type not in key_norm_cache:  # this rank's share, normalised and staged once per key type
                key_norm_cache[key_feature_type] = normalize_rows(curr_keys_feature[lo:hi], device) if hi > lo else None
            k32 = key_norm_cache[key_feature_type]
            kid = key_ids_for(keys_label)
            entry = acc_dict[query_feature_type][key_feature_type]
            pc_entry = per_class_acc[query_feature_type][key_feature_type]
            preds = {}
            for split, feats, gts in (("seen", curr_seen_feature, seen_gt_label),
                                      ("unseen", curr_unseen_feature, unseen_gt_label)):
                q32 = normalize_rows(feats, device)
                if k32 is not None:
                    s64, idx = _search_block(q32, k32, max_k, lo, mode)
                else:
                    s64, idx = _empty_lists(q32.shape[0], max_k, device)
                _, _, idx = _merge_over_ranks(s64, idx, world, process_group)
                micro_acc, macro_acc, per_class = {}, {}, {}
                for c0 in range(0, len(k_list), 4):
                    ks = list(k_list[c0:c0 + 4])
                    micro, chit, ccnt = accuracy_counts(idx, kid, torch.from_numpy(gt_ids[split]).to(device), ks,
                                                        vocab.max_class())
                    mi, ma, pc = _accuracy_from_counts(micro, chit, ccnt, gt_ids[split], ks, vocab, len(gts))
                    micro_acc.update(mi)
                    macro_acc.update(ma)
                    per_class.update(pc)
                entry[split] = {"micro_acc": micro_acc, "macro_acc": macro_acc}
                pc_entry[split] = per_class
                preds[split] = _labels_from_indices(idx.cpu().numpy(), keys_label)
            pred_dict[query_feature_type][key_feature_type] = {
                "curr_seen_pred_list": preds["seen"],
                "curr_unseen_pred_list": preds["unseen"],
            }
    if verbose:
        print_micro_and_macro_acc(acc_dict, k_list, args)
    return acc_dict, per_class_acc, pred_dict
