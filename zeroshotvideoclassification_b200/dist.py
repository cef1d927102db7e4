"""Data-parallel training across the GPUs of one node: one process per GPU, persistent replicas, gradient
all-reduce (average) overlapped with the backward pass.

Replaces the reference's single-process ``nn.DataParallel`` (main.py:126: per-iteration replicate / scatter /
gather through GPU 0).  Semantics kept: ``--bs`` is per GPU, the loss is a mean over the global
batch (per-rank mean + average of gradients is identical), BatchNorm statistics stay per replica.

The backbone's backward hands each residual block's parameter gradients to ``GradSync.submit`` as soon as they
exist (deepest block first: layer4 holds 23.5 M of the 31.7 M live parameters and finishes first), so NCCL
all-reduces over NVLink run on NCCL's stream while the remaining blocks' dgrad/wgrad kernels execute.  Dead
parameters (network.py:500-517, never reached by forward) have no gradient and are never communicated.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


class GradSync:
    """Asynchronous all-reduce(average) of gradients handed over in backward order.

    Two ways in:
      * ``submit_range(flat, lo, hi)`` -- the backbone's backward writes every gradient into one flat arena
        (ops.GradArena) in the order it is produced and hands over the slice each residual block filled; slices are
        all-reduced IN PLACE as soon as ``bucket_bytes`` are pending, so there is no gather copy, no scaling pass and no
        copy back (with NCCL the average is the collective's own AVG op);
      * ``attach(params)`` -- parameters whose gradients autograd accumulates itself (the MLP head): a
        post-accumulate hook all-reduces ``param.grad`` in place the moment it exists, i.e. before the backbone's
        backward even starts, instead of a second blocking collective after ``loss.backward()``.
    ``finish()`` makes the current stream wait for everything in flight (called by the backbone's backward)."""

    def __init__(self, group=None, bucket_bytes: int = 8 << 20):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_bytes = bucket_bytes
        self._works: List[tuple] = []         # (work, tensor reduced in place, needs_scale)
        self._pending: List[tuple] = []       # dict API: (work, flat, [(name, tensor)])
        self._cur: List[tuple] = []
        self._cur_bytes = 0
        self._range = None                    # (flat, lo, hi) handed over but not yet flushed
        self.bytes_reduced = 0
        self.bytes_per_step = 0           # bytes all-reduced by the most recent backward pass
        self._step_bytes = 0
        self._hooks = []
        self.collectives_per_step = 0
        self._step_collectives = 0
        nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self._op = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM      # gloo (CPU tests) has no AVG: SUM + scale

    def _all_reduce(self, t: torch.Tensor) -> None:
        work = dist.all_reduce(t, op=self._op, group=self.group, async_op=True)
        self._works.append((work, t, self._op == dist.ReduceOp.SUM))
        n = t.numel() * t.element_size()
        self.bytes_reduced += n
        self._step_bytes += n
        self._step_collectives += 1

    # -- flat arena ------------------------------------------------------------------------------------------
    def submit_range(self, flat: torch.Tensor, lo: int, hi: int) -> None:
        """flat[lo:hi] is final on the current stream; consecutive calls extend one pending slice."""
        if self.world == 1 or hi <= lo:
            return
        if self._range is not None and self._range[0] is flat and self._range[2] == lo:
            self._range = (flat, self._range[1], hi)
        else:
            self._flush_range()
            self._range = (flat, lo, hi)
        if (self._range[2] - self._range[1]) * flat.element_size() >= self.bucket_bytes:
            self._flush_range()

    def _flush_range(self) -> None:
        if self._range is not None:
            flat, lo, hi = self._range
            self._range = None
            self._all_reduce(flat[lo:hi])

    # -- parameters accumulated by autograd ------------------------------------------------------------------------
    def attach(self, params) -> None:
        if self.world == 1:
            return
        for p in params:
            if p.requires_grad:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p: torch.Tensor) -> None:
        if p.grad is not None:
            self._all_reduce(p.grad)

    def detach(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    # -- dict API: called with fresh fp32 gradients (kept for callers without an arena) ----------------------------
    def submit(self, grads: Dict[str, torch.Tensor]) -> None:
        if self.world == 1:
            return
        for name, g in grads.items():
            self._cur.append((name, g))
            self._cur_bytes += g.numel() * g.element_size()
        if self._cur_bytes >= self.bucket_bytes:
            self.flush()

    def flush(self) -> None:
        if not self._cur or self.world == 1:
            self._cur, self._cur_bytes = [], 0
            return
        items = self._cur
        self._cur, self._cur_bytes = [], 0
        flat = torch.cat([g.reshape(-1) for _, g in items])
        self._all_reduce(flat)
        self._pending.append((flat, items))

    def finish(self) -> Dict[str, torch.Tensor]:
        """Wait (on the current stream) for every collective in flight; returns the averaged gradients of the dict API
        keyed like the submitted ones (arena slices and attached parameters are averaged in place)."""
        self.flush()
        self._flush_range()
        for work, t, needs_scale in self._works:
            work.wait()
            if needs_scale:
                t.mul_(1.0 / self.world)
        self._works = []
        out: Dict[str, torch.Tensor] = {}
        for flat, items in self._pending:
            off = 0
            for name, g in items:
                n = g.numel()
                out[name] = flat[off:off + n].view(g.shape)
                off += n
        self._pending = []
        if self._step_collectives:        # a second finish() of the same step (nothing in flight) keeps the counters
            self.bytes_per_step, self._step_bytes = self._step_bytes, 0
            self.collectives_per_step, self._step_collectives = self._step_collectives, 0
        return out


_active_sync: Optional[GradSync] = None


def set_grad_sync(sync: Optional[GradSync]) -> None:
    """Install the GradSync the backbone backward reports to (None = single GPU)."""
    global _active_sync
    _active_sync = sync


def active_grad_sync() -> Optional[GradSync]:
    return _active_sync


def init_from_env(backend: Optional[str] = None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def sync_head_grads(params) -> None:
    """All-reduce(average) the (tiny) gradients of parameters outside the backbone Function (the MLP head), blocking.
    Not needed for parameters handed to ``GradSync.attach`` (they are reduced from a hook during backward)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    gs = [p.grad for p in params if p.grad is not None]
    if not gs:
        return
    flat = torch.cat([g.reshape(-1) for g in gs])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / dist.get_world_size())
    off = 0
    for g in gs:
        n = g.numel()
        g.copy_(flat[off:off + n].view(g.shape))
        off += n


def broadcast_module(module: torch.nn.Module, src: int = 0) -> None:
    """Make every replica start from rank ``src``'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


# -- evaluation: rows are independent, so the nearest-class search shards over ranks ----------
def shard_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range [lo, hi) of rank ``rank``; ranges of all ranks tile [0, n) (some may be empty)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def reduce_accuracy_counts(counts: torch.Tensor, group=None) -> Tuple[float, float]:
    """``counts`` = int64 [3] (top-1 hits, top-5 hits, rows) of this rank's rows; SUM over ranks, then the two
    percentages ``compute_accuracy`` (main.py:316-325) returns.  One exchange step of 24 bytes."""
    counts = counts.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    c1, c5, n = (int(v) for v in counts.tolist())
    if n == 0:
        raise ValueError("accuracy of zero samples")
    return 100.0 * c1 / n, 100.0 * c5 / n


def compute_accuracy_sharded(predicted_embed, class_embed, true_embed, group=None) -> Tuple[float, float]:
    """``compute_accuracy`` with the N rows split over the ranks of ``group``: every rank holds the full embedding
    tables (as evaluate() does, main.py:254-256), scores its own row range with the nearest-class kernel and the hit
    counts are all-reduced.  Same result on every rank, equal to the single-GPU call."""
    from . import accuracy
    assert len(predicted_embed) == len(true_embed), "True and predicted labels must have the same number of samples"
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_rows(len(predicted_embed), rank, world)
    return reduce_accuracy_counts(accuracy.count_correct(predicted_embed[lo:hi], class_embed, true_embed[lo:hi]), group)
