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
    """Bucketed asynchronous all-reduce(average) of gradients handed over in backward order."""

    def __init__(self, group=None, bucket_bytes: int = 32 << 20):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_bytes = bucket_bytes
        self._pending: List[tuple] = []       # (work, flat, [(name, tensor)])
        self._cur: List[tuple] = []
        self._cur_bytes = 0
        self.bytes_reduced = 0
        self.bytes_per_step = 0           # bytes all-reduced by the most recent backward pass
        self._step_bytes = 0

    # called from the backward pass with fresh fp32 gradients
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
        self.bytes_reduced += flat.numel() * flat.element_size()
        self._step_bytes += flat.numel() * flat.element_size()
        # SUM + scale keeps gloo (CPU tests) and NCCL on the same code path
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((work, flat, items))

    def finish(self) -> Dict[str, torch.Tensor]:
        """Wait for every bucket; returns the averaged gradients keyed like the submitted ones."""
        self.flush()
        out: Dict[str, torch.Tensor] = {}
        for work, flat, items in self._pending:
            work.wait()
            flat.mul_(1.0 / self.world)
            off = 0
            for name, g in items:
                n = g.numel()
                out[name] = flat[off:off + n].view(g.shape)
                off += n
        self._pending = []
        self.bytes_per_step, self._step_bytes = self._step_bytes, 0
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
    """All-reduce(average) the (tiny) gradients of parameters outside the backbone Function (the MLP head)."""
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
