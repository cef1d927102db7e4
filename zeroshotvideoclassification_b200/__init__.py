"""zeroshotvideoclassification_b200 -- B200 (sm_100a) arithmetic for the R(2+1)D-18 / C3D training hot path
and the zero-shot nearest-class search of damien911224/ZeroShotVideoClassification.

Public surface (mirrors the reference's own names):
  get_network(opt), Model, MLP        -- network.py:24-44, 472-600, 603-618
  r2plus1d_18(), r3d_18()             -- resnet.py:342-362, 293-314
  compute_accuracy(pred, cls, true)   -- main.py:316-325 on the GPU
  nearest_class(emb, cls, k)          -- main.py:183
  class_overlap_mask(cls, other, tau) -- auxiliary/auxiliary_dataset.py:141-144
"""
from .video_models import MLP, Model, default_opt, get_network, r2plus1d_18, r3d_18  # noqa: F401
from .accuracy import class_overlap_mask, compute_accuracy, nearest_class  # noqa: F401
from . import dist  # noqa: F401

__all__ = ["get_network", "Model", "MLP", "r2plus1d_18", "r3d_18", "default_opt", "compute_accuracy", "nearest_class", "class_overlap_mask"]
