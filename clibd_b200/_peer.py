"""Peer-mapped (symmetric) exchange buffers of the row-sharded loss step.

torch.distributed._symmetric_memory is used as PLUMBING only: it allocates one buffer per rank, maps every rank's
buffer into every process (NVLink / NVSwitch peer access) and provides a barrier across the ranks.  What travels
through those mappings is written by this package's own kernels (csrc/shard_exchange.cu, the gradient GEMM's
epilogue in csrc/loss_grad_gemm.cu).

Layout of one pool ENTRY (everything a forward hands to its backward; identical offsets on every rank):
    x[3]      [N, d] gathered raw features (input dtype)      inv[3]   [N] float32 inverse norms
    labels    [N] int64                                       stats    [9, N] float32 rowsum | colsum | posrow
    colslots  [W, 3, N] float32 column-sum partials per rank  posslots [W, 4] float64, pos_local / pos [4] float64
and of the per-configuration SHARED block (two parities, alternating per backward):
    red       [3 pairs, W, n, d] float32 column-side gradient partials received from every rank
    gslots    [W] float32 grad_output of every rank        dslots  [W] float64 d loss / d logit_scale of every rank
    flags     clibd_shard_barrier's epochs (not duplicated: they only grow)

Why a single entry is enough in a training loop, and when the pool grows: a rank writes into its peers' entry only
at the start of a forward (rows) and after the first barrier of that forward (statistics); its peers read those
regions only before the LAST barrier of the step that used the entry (statistics barrier in a forward-only step, the
gradient barrier of the backward otherwise).  An entry is released -- on the host, in program order -- after that last
barrier has been enqueued, so every later push is ordered behind it on every rank.  A forward whose backward is still
outstanding keeps its entry; a second forward then takes (or allocates, collectively) another one.  `red` / `gslots`
alternate between two copies because two backward passes may follow each other without a forward in between.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

try:
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as _symm
except Exception:  # noqa: BLE001  (no distributed build)
    dist = None
    _symm = None

_contexts = {}


_disabled_reason = None


def available() -> bool:
    return _symm is not None and torch.cuda.is_available() and _disabled_reason is None


def disable(reason: str):
    """Called when mapping the ranks' buffers into each other failed (an environment without peer access / fd passing):
    the sharded step then uses its NCCL form for the rest of the process.  The failure is a property of the box, so
    every rank takes the same branch."""
    global _disabled_reason
    if _disabled_reason is None:
        import warnings
        warnings.warn("clibd_b200: peer-mapped exchange unavailable (" + reason + "); using the NCCL form of the sharded "
                      "loss step")
        _disabled_reason = reason


def _align(v, a=256):
    return (v + a - 1) // a * a


class _Region:
    """One symmetric allocation: local tensor, rendezvous handle, per-rank base pointers."""

    def __init__(self, nbytes, device, group):
        self.tensor = _symm.empty(_align(nbytes), dtype=torch.uint8, device=device)
        self.tensor.zero_()
        self.handle = _symm.rendezvous(self.tensor, group.group_name)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.handle.barrier()  # nobody stores into a peer's buffer before that peer has zeroed it


class Entry:
    def __init__(self, px):
        N, d, W, esize = px.N, px.d, px.world, px.esize
        off = 0
        offs = {}
        for name, nbytes in (("x0", N * d * esize), ("x1", N * d * esize), ("x2", N * d * esize), ("inv0", 4 * N),
                             ("inv1", 4 * N), ("inv2", 4 * N), ("labels", 8 * N), ("stats", 4 * 9 * N),
                             ("colslots", 4 * W * 3 * N), ("posslots", 8 * W * 4), ("pos_local", 32), ("pos", 32)):
            offs[name] = off
            off = _align(off + nbytes)
        self.region = _Region(off, px.device, px.group)
        base = self.region.ptrs
        me = px.rank
        self.busy = False
        self.x_ptr = [base[me] + offs[f"x{m}"] for m in range(3)]
        self.inv_ptr = [base[me] + offs[f"inv{m}"] for m in range(3)]
        self.labels_ptr = base[me] + offs["labels"]
        self.stats_ptr = base[me] + offs["stats"]
        self.colslots_ptr = base[me] + offs["colslots"]
        self.posslots_ptr = base[me] + offs["posslots"]
        self.pos_local_ptr = base[me] + offs["pos_local"]
        self.pos_ptr = base[me] + offs["pos"]
        self.peer_x = _lib.ptr_array([base[q] + offs[f"x{m}"] for q in range(W) for m in range(3)])
        self.peer_inv = _lib.ptr_array([base[q] + offs[f"inv{m}"] for q in range(W) for m in range(3)])
        self.peer_labels = _lib.ptr_array([base[q] + offs["labels"] for q in range(W)])
        self.peer_stats = _lib.ptr_array([base[q] + offs["stats"] for q in range(W)])
        self.peer_colslots = _lib.ptr_array([base[q] + offs["colslots"] for q in range(W)])
        self.peer_posslots = _lib.ptr_array([base[q] + offs["posslots"] for q in range(W)])


class PeerContext:
    """Exchange buffers of one (process group, N, n, d, dtype) configuration."""

    def __init__(self, group, device, N, n, d, dtype, world, rank):
        self.group, self.device = group, device
        self.N, self.n, self.d, self.world, self.rank = N, n, d, world, rank
        self.esize = torch.empty((), dtype=dtype).element_size()
        self.entries = []
        self._side = None
        self._next = 0
        self._parity = 1
        self.red_pair_bytes = 4 * world * n * d  # one pair's slot array [W, n, d]
        red_bytes = _align(3 * self.red_pair_bytes)
        g_off = 2 * red_bytes
        d_off = g_off + 2 * 256                # dslots [2 parities][W] float64: every rank's d loss / d logit_scale
        flag_off = d_off + 2 * _align(8 * world)
        self.shared = _Region(flag_off + _align(_lib.load().clibd_shard_barrier_bytes()), device, group)
        base = self.shared.ptrs
        self.peer_dslots = [_lib.ptr_array([base[q] + d_off + par * _align(8 * world) for q in range(world)])
                            for par in range(2)]
        self.dslots = [self.shared.tensor[d_off + par * _align(8 * world):d_off + par * _align(8 * world) + 8 * world]
                       .view(torch.float64) for par in range(2)]
        self.peer_flags = _lib.ptr_array([base[q] + flag_off for q in range(world)])  # clibd_shard_barrier's blocks
        self.red_ptr = [base[rank] + par * red_bytes for par in range(2)]
        self.gslots_ptr = [base[rank] + g_off + par * 256 for par in range(2)]
        # entry [q * 3 + p]: pair p's slot array inside rank q's memory
        self.peer_red = [_lib.ptr_array([base[q] + par * red_bytes + p * self.red_pair_bytes
                                         for q in range(world) for p in range(3)]) for par in range(2)]
        self.peer_gslots = [_lib.ptr_array([base[q] + g_off + par * 256 for q in range(world)]) for par in range(2)]

    def barrier(self, channel=0):
        """All ranks' earlier work on the current stream (incl. their stores into peer memory) is complete and visible
        before any rank's later work starts.  Barriers issued on two streams at once use different channels.
        Default: torch's symmetric-memory barrier.  CLIBD_BARRIER=own: this library's one-warp kernel over flags in the
        shared region (clibd_shard_barrier, csrc/shard_exchange.cu) -- what a host without torch would call; the two cost
        the same (tools/barrier_probe.py, 2 GPUs: 6.4 vs 6.2 us back to back, 11.5 vs 9.4 us with a kernel in between)."""
        if os.environ.get("CLIBD_BARRIER", "torch") != "own":
            self.shared.handle.barrier(channel=channel)
            return
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.load().clibd_shard_barrier(self.peer_flags, self.rank, self.world, channel, stream))

    def side_stream(self):
        """Second stream of the forward: the push of the feature rows runs there, next to the label statistics."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def acquire(self) -> Entry:
        k = len(self.entries)
        for step in range(k):
            e = self.entries[(self._next + step) % k]
            if not e.busy:
                self._next = (self._next + step + 1) % k
                e.busy = True
                return e
        e = Entry(self)  # collective (rendezvous): every rank reaches this point in the same call (SPMD)
        self.entries.append(e)
        e.busy = True
        return e

    def next_parity(self) -> int:
        self._parity ^= 1
        return self._parity


class GatherBuffers:
    """Receive buffer of the stand-alone differentiable all-gather (loss.gather_features): [N, d] rows + [N] inverse
    norms (clibd_shard_push_rows always writes them).  One copy is enough: the caller reads the gathered rows between
    two barriers, so no rank can push the next call's rows before every rank has finished reading this call's."""

    def __init__(self, group, device, N, d, dtype, world, rank):
        esize = torch.empty((), dtype=dtype).element_size()
        x_bytes = _align(N * d * esize)
        self.region = _Region(x_bytes + 4 * N, device, group)
        base = self.region.ptrs
        self.peer_x = _lib.ptr_array([base[q] if m == 0 else None for q in range(world) for m in range(3)])
        self.peer_inv = _lib.ptr_array([base[q] + x_bytes if m == 0 else None for q in range(world) for m in range(3)])
        self.peer_labels = _lib.ptr_array([None] * world)
        self.rows = self.region.tensor[:N * d * esize].view(dtype).view(N, d)

    def barrier(self):
        self.region.handle.barrier()


_gather_buffers = {}


def gather_buffers(group, device, N, d, dtype, world, rank) -> GatherBuffers:
    group = group if group is not None else dist.group.WORLD
    key = (group.group_name, device.index, N, d, dtype, world, rank)
    gb = _gather_buffers.pop(key, None)
    if gb is None:
        while len(_gather_buffers) >= _MAX_CONTEXTS:
            del _gather_buffers[next(iter(_gather_buffers))]  # least recently used (same sequence on every rank)
        gb = GatherBuffers(group, device, N, d, dtype, world, rank)
    _gather_buffers[key] = gb
    return gb


class EntryHolder:
    """Releases a pool entry when the backward has run, or when autograd drops a forward that will have none."""

    def __init__(self, px, entry):
        self.px, self.entry = px, entry
        self.released = False

    def release(self):
        if not self.released:  # exactly once: the entry may already serve another forward afterwards
            self.released = True
            self.entry.busy = False

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001
            pass


_MAX_CONTEXTS = 4  # distinct (batch, dim, dtype) configurations kept mapped; a training loop uses one or two


def context(group, device, N, n, d, dtype, world, rank) -> PeerContext:
    group = group if group is not None else dist.group.WORLD
    key = (group.group_name, device.index, N, n, d, dtype, world, rank)
    px = _contexts.pop(key, None)
    if px is None:
        # Least recently used configurations whose buffers no forward still holds are dropped first.  Every rank sees
        # the same sequence of shapes (SPMD), so every rank drops and allocates at the same call.
        for old_key in list(_contexts):
            if len(_contexts) < _MAX_CONTEXTS:
                break
            if not any(e.busy for e in _contexts[old_key].entries):
                del _contexts[old_key]
        px = PeerContext(group, device, N, n, d, dtype, world, rank)
    _contexts[key] = px  # (re)inserted last: dict order = recency
    return px


def reset():
    """Drop every cached context (tests; frees the symmetric allocations)."""
    _contexts.clear()
    _gather_buffers.clear()
