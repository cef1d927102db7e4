"""Shared helpers for the parity tests (independent of the kernels under test: plain torch ops only)."""
import torch


def cpad(c):
    return (c + 7) & ~7


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


def to_ndhwc(x_ncdhw, device="cuda"):
    """fp32 [N,C,T,H,W] -> bf16 [N,T,H,W,cpad(C)] on `device` using torch only."""
    n, c, t, h, w = x_ncdhw.shape
    out = torch.zeros((n, t, h, w, cpad(c)), dtype=torch.bfloat16)
    out[..., :c] = x_ncdhw.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    return out.to(device)


def from_ndhwc(x, c):
    """bf16 [N,T,H,W,Cp] (any device) -> fp32 CPU [N,C,T,H,W]."""
    return x[..., :c].float().cpu().permute(0, 4, 1, 2, 3).contiguous()


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative error' of north_star (scale of the reference tensor)."""
    a = a.double()
    b = b.double()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def rms_rel_err(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))
