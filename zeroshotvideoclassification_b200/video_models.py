"""Module trees with the reference's parameter names, shapes, init scheme and construction order, whose
forward/backward run on libzsv_b200.so instead of ATen/cuDNN.

Drop-in contract: ``state_dict()`` keys/shapes/dtypes are those of the reference's
``network.Model`` / ``network.C3D`` / ``resnet.VideoResNet`` (network.py:472-519, network.py:95-141,
resnet.py:190-281), so checkpoints load both ways; ``forward`` takes ``[B, n_clips, 3, T, H, W]`` fp32 CUDA
tensors and returns ``(emb, None)`` (Model, network.py:600) or ``emb`` (C3D, network.py:180); gradients land
in ``param.grad`` as fp32, so ``main.py``'s loss, GradScaler, Adam and evaluation code run unchanged.

The nn.Conv3d / nn.BatchNorm3d / nn.Linear children are parameter containers only: the arithmetic is the
whole-network autograd Function in ``engine.py``.  There is no CPU or eager-PyTorch fallback -- a non-CUDA
input raises.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn

from . import engine


def _factorised_width(cin: int, cout: int) -> int:
    # parameter-matched width of the (2+1)D factorisation, resnet.py:91
    return (cin * cout * 27) // (cin * 9 + 3 * cout)


class SpatioTemporalConv(nn.Sequential):
    """1x3x3 conv -> BN -> ReLU -> 3x1x1 conv (the reference's Conv2Plus1D, resnet.py:37-57).
    Child indices 0..3 are part of the state-dict contract."""

    def __init__(self, cin: int, cout: int, mid: int, stride: int = 1):
        spatial = nn.Conv3d(cin, mid, (1, 3, 3), (1, stride, stride), (0, 1, 1), bias=False)
        norm = nn.BatchNorm3d(mid)
        temporal = nn.Conv3d(mid, cout, (3, 1, 1), (stride, 1, 1), (1, 0, 0), bias=False)
        super().__init__(spatial, norm, nn.ReLU(inplace=True), temporal)


class ResidualUnit(nn.Module):
    """Two SpatioTemporalConv+BN stages with identity or projected shortcut (BasicBlock, resnet.py:79-113)."""

    def __init__(self, cin: int, cout: int, stride: int, shortcut: nn.Module | None):
        super().__init__()
        mid = _factorised_width(cin, cout)
        self.conv1 = nn.Sequential(SpatioTemporalConv(cin, cout, mid, stride), nn.BatchNorm3d(cout),
                                   nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(SpatioTemporalConv(cout, cout, mid), nn.BatchNorm3d(cout))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = shortcut
        self.stride = stride


def _stem() -> nn.Sequential:
    # R2Plus1dStem, resnet.py:176-187
    return nn.Sequential(
        nn.Conv3d(3, 45, (1, 7, 7), (1, 2, 2), (0, 3, 3), bias=False), nn.BatchNorm3d(45), nn.ReLU(inplace=True),
        nn.Conv3d(45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), bias=False), nn.BatchNorm3d(64), nn.ReLU(inplace=True))


class VideoResNet18(nn.Module):
    """R(2+1)D-18 backbone (resnet.r2plus1d_18, resnet.py:342-362).  ``forward`` returns ``(pooled, features)``
    like resnet.py:243-256 (the pooled vector is what the reference computes and then discards)."""

    arch = "r2plus1d_18"

    def __init__(self, num_classes: int = 400):
        super().__init__()
        self.stem = _stem()
        widths = (64, 128, 256, 512)
        cin = 64
        for i, cout in enumerate(widths, start=1):
            stride = 1 if i == 1 else 2
            units = []
            for j in range(2):
                s = stride if j == 0 else 1
                shortcut = None
                if j == 0 and (s != 1 or cin != cout):  # resnet.py:268-273 (built before the block, as there)
                    shortcut = nn.Sequential(nn.Conv3d(cin, cout, 1, (s, s, s), bias=False), nn.BatchNorm3d(cout))
                units.append(ResidualUnit(cin, cout, s, shortcut))
                cin = cout
            setattr(self, f"layer{i}", nn.Sequential(*units))
        self.avgpool = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.fc = nn.Linear(512, num_classes)
        # init scheme of resnet.py:226-236
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor):
        feats = engine.backbone_forward(self, x)           # bf16 NDHWC, autograd-tracked
        f = engine.features_to_ncdhw(feats)                # [B,512,T',H',W'] fp32 view for API parity
        return f.mean(dim=(2, 3, 4)), f


class SimpleResidualUnit(nn.Module):
    """BasicBlock over plain 3x3x3 convolutions (Conv3DSimple, resnet.py:18-34; BasicBlock resnet.py:79-113).
    conv1 = [conv, BN, ReLU], conv2 = [conv, BN]: child indices are part of the state-dict contract."""

    def __init__(self, cin: int, cout: int, stride: int, shortcut: nn.Module | None):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv3d(cin, cout, (3, 3, 3), stride, 1, bias=False), nn.BatchNorm3d(cout),
                                   nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(nn.Conv3d(cout, cout, (3, 3, 3), 1, 1, bias=False), nn.BatchNorm3d(cout))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = shortcut
        self.stride = stride


class VideoResNet18R3D(nn.Module):
    """r3d_18 backbone (resnet.r3d_18, resnet.py:293-314): BasicStem (resnet.py:165-173) + 4 x 2 BasicBlocks of
    3x3x3 convolutions.  Same contract as VideoResNet18."""

    arch = "r3d_18"

    def __init__(self, num_classes: int = 400):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv3d(3, 64, (3, 7, 7), (1, 2, 2), (1, 3, 3), bias=False), nn.BatchNorm3d(64),
                                  nn.ReLU(inplace=True))
        cin = 64
        for i, cout in enumerate((64, 128, 256, 512), start=1):
            stride = 1 if i == 1 else 2
            units = []
            for j in range(2):
                s = stride if j == 0 else 1
                shortcut = None
                if j == 0 and (s != 1 or cin != cout):
                    shortcut = nn.Sequential(nn.Conv3d(cin, cout, 1, (s, s, s), bias=False), nn.BatchNorm3d(cout))
                units.append(SimpleResidualUnit(cin, cout, s, shortcut))
                cin = cout
            setattr(self, f"layer{i}", nn.Sequential(*units))
        self.avgpool = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.fc = nn.Linear(512, num_classes)
        for m in self.modules():                      # init scheme of resnet.py:226-236
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor):
        feats = engine.backbone_forward(self, x)
        f = engine.features_to_ncdhw(feats)
        return f.mean(dim=(2, 3, 4)), f


def r3d_18(pretrained: bool = False, **_) -> VideoResNet18R3D:
    if pretrained:
        raise RuntimeError("pretrained weights need network access; the reference never requests them "
                           "(main.py:42 makes --nopretrained always False)")
    return VideoResNet18R3D()


def r2plus1d_18(pretrained: bool = False, **_) -> VideoResNet18:
    if pretrained:
        raise RuntimeError("pretrained weights need network access; the reference never requests them "
                           "(main.py:42 makes --nopretrained always False)")
    return VideoResNet18()


class MLP(nn.Module):
    """network.MLP (network.py:603-618): Linear layers with ReLU between them."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, last_activate=False):
        super().__init__()
        self.num_layers = num_layers
        self.last_activate = last_activate
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))


class Model(nn.Module):
    """network.Model (network.py:472-600).  The Transformer / embedding members are constructed only so that
    the state dict and the RNG stream match the reference; they never reach the output there either
    (network.py:533-600) and receive no gradient."""

    def __init__(self, network=r2plus1d_18, fixconvs: bool = False, nopretrained: bool = False):
        super().__init__()
        self.model = network(pretrained=nopretrained)
        if fixconvs:
            for p in self.model.parameters():
                p.requires_grad = False
        self.d_model = 256
        self.num_sentences = 1
        self.t_pos_embeds = nn.Embedding(self.num_sentences, 512)
        self.special_tokens = nn.Embedding(1, self.d_model)
        self.feature2input_proj = nn.Linear(512, self.d_model)
        layer = nn.TransformerEncoderLayer(d_model=self.d_model, dim_feedforward=self.d_model * 4, nhead=8,
                                           dropout=0.1, activation="gelu")
        self.encoder = nn.TransformerEncoder(layer, num_layers=6)
        self.output2emb_proj = MLP(512, 512, 300, 2)
        nn.init.normal_(self.t_pos_embeds.weight)
        nn.init.xavier_uniform_(self.special_tokens.weight)

    def forward(self, x: torch.Tensor):
        if x.dtype == torch.bfloat16 and x.dim() == 5 and x.shape[-1] == 8:
            pass                             # clips already transformed on the GPU (ops.clip_transform): [B*nc,T,H,W+8,8]
        else:
            bs, nc = x.shape[:2]
            x = x.reshape(bs * nc, *x.shape[2:])
        feats = engine.backbone_forward(self.model, x)
        lin1, lin2 = self.output2emb_proj.layers
        emb = engine.head_forward(feats, lin1.weight, lin1.bias, lin2.weight, lin2.bias)
        return emb, None


def get_network(opt) -> nn.Module:
    """network.get_network (network.py:24-44): string dispatch on ``opt.network``."""
    name = opt.network
    if "r3d" in name:                  # same test order as network.py:28-36
        return Model(r3d_18, fixconvs=opt.fixconvs, nopretrained=opt.nopretrained)
    if "2plus1d" in name:
        return Model(r2plus1d_18, fixconvs=opt.fixconvs, nopretrained=opt.nopretrained)
    if "c3d" in name:
        from .c3d_model import C3D
        return C3D(fixconvs=opt.fixconvs, nopretrained=opt.nopretrained)
    raise Exception("Network {} not available!".format(name))


def default_opt(network: str = "r2plus1d_18") -> SimpleNamespace:
    return SimpleNamespace(network=network, fixconvs=False, nopretrained=False)
