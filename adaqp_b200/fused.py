"""Fused LayerNorm + ReLU between aggregations (host mirror of csrc/norm.cu).

`layer_norm_relu(x, norm)` == `F.relu(norm(x))` for an `nn.LayerNorm` over the last dimension
(AdaQP/model/distGCN.py:81-84, distSAGE.py:93-96), one HBM pass forward and one backward instead of
torch's separate kernels (whose gamma / beta column reduction alone costs 7.6 ms per layer at 2.4 M rows).
The module and its parameters stay `nn.LayerNorm` (state_dict unchanged); dropout stays torch's kernel, so
the dropout mask for a given generator state is unchanged.  CPU tensors, widths that are not a multiple of
4 and `ADAQP_FUSED_NORM=0` use the torch ops."""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch import Tensor, nn
from torch.autograd import Function

from . import _lib

_ENABLED = None
LAUNCHES = {"ln_relu_fwd_kernel": 0, "ln_relu_bwd_kernel": 0}


def enabled() -> bool:
    global _ENABLED
    if _ENABLED is None:
        _ENABLED = os.environ.get("ADAQP_FUSED_NORM", "1") != "0"
    return _ENABLED


def supported(x: Tensor, norm: nn.LayerNorm) -> bool:
    if not (enabled() and isinstance(norm, nn.LayerNorm) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
        return False
    Fdim = x.shape[1]
    if tuple(norm.normalized_shape) != (Fdim,) or norm.weight is None or norm.bias is None:
        return False
    return (x.stride(1) == 1 and Fdim % 4 == 0 and Fdim <= 1024 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0
            and x.shape[0] > 0)


class _LayerNormReLU(Function):
    @staticmethod
    def forward(ctx, x: Tensor, gamma: Tensor, beta: Tensor, eps: float):
        L = _lib.load()
        M, Fdim = x.shape
        y = torch.empty((M, Fdim), dtype=torch.float32, device=x.device)
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty(M, dtype=torch.float32, device=x.device)
        gamma, beta = gamma.contiguous(), beta.contiguous()
        rc = L.adaqp_ln_relu_fwd_f32(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), float(eps), M, Fdim,
                                     y.data_ptr(), y.stride(0), mean.data_ptr(), rstd.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "adaqp_ln_relu_fwd_f32")
        LAUNCHES["ln_relu_fwd_kernel"] += 1
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        L = _lib.load()
        M, Fdim = x.shape
        dy = dy.contiguous()
        dx = torch.empty((M, Fdim), dtype=torch.float32, device=x.device)
        grid = L.adaqp_ln_relu_grid(M)
        partials = torch.empty((grid, 2, Fdim), dtype=torch.float32, device=x.device)
        rc = L.adaqp_ln_relu_bwd_f32(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), mean.data_ptr(), rstd.data_ptr(),
                                     gamma.data_ptr(), beta.data_ptr(), M, Fdim, dx.data_ptr(), dx.stride(0),
                                     partials.data_ptr(), grid, _lib.stream_ptr())
        _lib.check(rc, "adaqp_ln_relu_bwd_f32")
        LAUNCHES["ln_relu_bwd_kernel"] += 1
        sums = partials.sum(0)
        return dx, sums[0], sums[1], None


def layer_norm_relu(x: Tensor, norm: nn.LayerNorm) -> Tensor:
    if supported(x, norm):
        return _LayerNormReLU.apply(x, norm.weight, norm.bias, norm.eps)
    return F.relu(norm(x))
