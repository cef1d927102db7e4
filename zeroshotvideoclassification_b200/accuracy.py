"""GPU replacement for the reference's scipy-based accuracy code (main.py:182-185, main.py:316-325)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _as_cuda(x, device=None) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not x.is_cuda:
        x = x.to(device if device is not None else "cuda")
    return x


def nearest_class(emb, class_embed, k: int = 1) -> torch.Tensor:
    """``cdist(emb, class_embed, 'cosine').argsort(1)[:, :k]`` (argmin for k=1), int64 [N,k] on the GPU."""
    emb = _as_cuda(emb)
    return ops.nearest_class(emb, _as_cuda(class_embed, emb.device), k)


def compute_accuracy(predicted_embed, class_embed, true_embed):
    """main.py:316-325: returns (top-1 %, top-5 %); labels are recovered from the true embeddings."""
    assert len(predicted_embed) == len(true_embed), "True and predicted labels must have the same number of samples"
    pred = _as_cuda(predicted_embed)
    cls = _as_cuda(class_embed, pred.device)
    k = min(5, cls.shape[0])
    y_pred = ops.nearest_class(pred, cls, k)
    y = ops.nearest_class(_as_cuda(true_embed, pred.device), cls, 1)
    top1 = (y_pred[:, :1] == y).float().mean() * 100
    top5 = (y_pred == y).any(dim=1).float().mean() * 100
    return float(top1), float(top5)


def count_correct(predicted_embed, class_embed, true_embed) -> torch.Tensor:
    """int64 [3] on the GPU: (top-1 hits, top-5 hits, rows) of main.py:316-325 for these rows; the summable form of
    ``compute_accuracy`` that ``dist.compute_accuracy_sharded`` all-reduces.  Zero rows give zeros."""
    pred = _as_cuda(predicted_embed)
    cls = _as_cuda(class_embed, pred.device)
    k = min(5, cls.shape[0])
    y_pred = ops.nearest_class(pred, cls, k)
    y = ops.nearest_class(_as_cuda(true_embed, pred.device), cls, 1)
    top1 = (y_pred[:, :1] == y).sum()
    top5 = (y_pred == y).any(dim=1).sum()
    return torch.stack([top1, top5, torch.tensor(pred.shape[0], device=pred.device)]).to(torch.int64)


def class_overlap_mask(class_embedding, other_class_embedding, class_overlap: float) -> torch.Tensor:
    """auxiliary/auxiliary_dataset.py:141-144 (filter_overlapping_classes): keep a training class iff its cosine
    distance to the NEAREST test class exceeds ``class_overlap``:
    ``cdist(class_embedding, ucf_class_embedding, 'cosine').min(1) > class_overlap``.  Same kernel as the nearest-class
    search (fp64, scipy's operation order), so the mask is bit-identical to the reference's.  Returns bool [C] (CUDA)."""
    emb = _as_cuda(class_embedding)
    _, dist = ops.nearest_class(emb, _as_cuda(other_class_embedding, emb.device), 1, return_dist=True)
    return dist[:, 0] > class_overlap
