"""CenterOfMass2D -- API mirror of the reference's layers/center_of_mass.py:4-60 on the sm_100a kernels.

The COM trainer's encoder (train_with_env_com_trunc_advantage_ttt.py:42-53) reduces every receiver image of
the history to its centre of mass.  Here that is one pass over the image (helio_com_fwd) and, when a
gradient is requested, one pass back (helio_com_bwd); the reference runs ~10 eager kernels and materialises
two [H,W] coordinate grids and two [B,H,W] products per call.  SURVEY.md section 8f, rank 3.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .functional import _Call, _cf, _ptr, _stream, require_cuda


class _CenterOfMassFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps: float):
        lib = _lib.load()
        B, H, W = x.shape
        coords = torch.empty(B, 2, dtype=torch.float32, device=x.device)
        sums = torch.empty(B, 3, dtype=torch.float32, device=x.device)
        with _Call("com_fwd", x.device):
            rc = lib.helio_com_fwd(_ptr(x), B, H, W, eps, _ptr(coords), _ptr(sums), _stream())
        _lib.check(rc, "helio_com_fwd")
        ctx.save_for_backward(x, sums)
        ctx.eps = eps
        return coords

    @staticmethod
    def backward(ctx, g_coords):
        lib = _lib.load()
        x, sums = ctx.saved_tensors
        B, H, W = x.shape
        g_img = torch.empty_like(x)
        with _Call("com_bwd", x.device):
            rc = lib.helio_com_bwd(_ptr(x), _ptr(sums), _ptr(_cf(g_coords)), B, H, W, ctx.eps, _ptr(g_img), _stream())
        _lib.check(rc, "helio_com_bwd")
        return g_img, None


class CenterOfMass2D(nn.Module):
    """Differentiable centre of mass of (B,H,W) or (B,1,H,W) images -> (B,2) = (x_com, y_com); origin top-left,
    x along columns, y along rows; images without mass give (-1,-1)  (layers/center_of_mass.py:4-60)."""

    def __init__(self, eps: float = 1e-12):
        super().__init__()
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 4:
            if x.size(1) != 1:
                raise ValueError("Expected single-channel images with shape (B, 1, H, W).")
            x = x[:, 0, ...]
        elif x.dim() != 3:
            raise ValueError("Expected input shape (B, H, W) or (B, 1, H, W).")
        require_cuda(x.device, "CenterOfMass2D")
        out = _CenterOfMassFn.apply(_cf(x), float(self.eps))
        return out if x.dtype == torch.float32 else out.to(x.dtype)
