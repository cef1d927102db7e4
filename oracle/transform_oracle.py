"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's clip transform (auxiliary/transforms.py:41-56):

    ToFloatTensorInZeroOne   (transforms.py:116-117)   uint8 [T,H,W,C] -> fp32 [C,T,H,W], (x/255 - 1)/2
    Resize(128)              (transforms.py:99-108)    F.interpolate(scale_factor = 128/min(H,W), bilinear,
                                                       align_corners=False)
    CenterCrop / RandomCrop  (transforms.py:76-84, 132-150)
    RandomHorizontalFlip     (transforms.py:189-195)

The random decisions (crop origin, flip) are inputs here, as they are for the CUDA kernel.  Pinned against the
reference's own functions by tests/golden/clip_transform.npz (oracle/make_golden.py:transform_fixture).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def clip_transform(frames_u8: torch.Tensor, crop_ij, flip: bool, resize_short: int = 128, crop: int = 112) -> torch.Tensor:
    """frames_u8: uint8 [T,H,W,3] -> fp32 [3,T,crop,crop]."""
    vid = (frames_u8.permute(3, 0, 1, 2).to(torch.float32) / 255 - 1.0) / 2.0
    scale = float(resize_short) / min(vid.shape[-2:])
    vid = F.interpolate(vid, size=None, scale_factor=scale, mode="bilinear", align_corners=False)
    i, j = int(crop_ij[0]), int(crop_ij[1])
    vid = vid[..., i:i + crop, j:j + crop]
    if flip:
        vid = vid.flip(dims=(-1,))
    return vid.contiguous()


def center_crop_origin(h: int, w: int, crop: int = 112):
    """transforms.py:76-84."""
    return int(round((h - crop) / 2.)), int(round((w - crop) / 2.))
