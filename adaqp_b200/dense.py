"""Dense feature x weight product on the tcgen05 tensor cores.

Host mirror of `adaqp_gemm_tf32x3_f32` (csrc/gemm.cu): `linear(x, weight, bias)` computes
`x @ weight (+ bias)` for `weight` stored [in, out] (DistGCNConv, AdaQP/model/distGCN.py:45) and
`linear_nk(x, weight_nk, bias)` computes `x @ weight_nk.T (+ bias)` for nn.Linear storage [out, in]
(DistSAGEConv, distSAGE.py:51-53).  fp32 in / out; the tensor cores run three tf32 products of
error-compensated operand halves (a = a_hi + a_lo, a_hi = a & 0xFFFFE000), accumulated in fp32.

Backward: dX = dY @ W^T runs through the same kernel; dW = dY^T @ X (reduction over the node
dimension) through the split-K kernel `adaqp_wgrad_tf32x3_f32`, which reads both operands
MN-major straight from their row-major storage; the bias gradient stays a torch reduction.

Shapes the kernel does not take (input row pitch not a multiple of 16 bytes such as F = 602, N > 256, CPU tensors) and
`ADAQP_GEMM=0` use torch.matmul, the reference's own arithmetic; the 47-wide gradient of the last layer is zero-padded
to 48 columns once and runs through the kernels.
"""
from __future__ import annotations

import os

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib

_ENABLED = None
LAUNCHES = {"gemm_tf32x3_kernel": 0, "wgrad_tf32x3_kernel": 0}     # launch counters (bench.py: gpu_launches)


def enabled() -> bool:
    global _ENABLED
    if _ENABLED is None:
        _ENABLED = os.environ.get("ADAQP_GEMM", "1") != "0"
    return _ENABLED


def split_tf32(t: Tensor):
    """t = hi + lo with hi exactly representable in tf32 (low 13 mantissa bits cleared)."""
    hi = (t.contiguous().view(torch.int32) & -8192).view(torch.float32)
    return hi, t - hi


def _pad_cols(t: Tensor) -> Tensor:
    """Row pitch must be a multiple of 4 floats for the TMA descriptor."""
    k = t.shape[1]
    if k % 4 == 0:
        return t.contiguous()
    out = t.new_zeros((t.shape[0], (k + 3) // 4 * 4))
    out[:, :k] = t
    return out


def supported(x: Tensor, n: int, k: int) -> bool:
    if not (enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1):
        return False
    if x.data_ptr() % 16 or x.shape[0] == 0:
        return False
    return bool(_lib.load().adaqp_gemm_tf32x3_supported(x.shape[0], n, k, x.stride(0), (k + 3) // 4 * 4, n))


def gemm_nt(x: Tensor, bt: Tensor, bias: Tensor = None) -> Tensor:
    """x [M, K] @ bt[N, K]^T (+ bias) through the C ABI (no autograd)."""
    M, K = x.shape
    N = bt.shape[0]
    assert bt.shape[1] == K or (bt.shape[1] + 3) // 4 * 4 == K      # x may carry the zero-padded K already
    bt = _pad_cols(bt)
    hi, lo = split_tf32(bt)
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    rc = _lib.load().adaqp_gemm_tf32x3_f32(x.data_ptr(), x.stride(0), hi.data_ptr(), lo.data_ptr(), bt.stride(0),
                                           bias.data_ptr() if bias is not None else None, M, N, K, out.data_ptr(),
                                           out.stride(0), _lib.stream_ptr())
    _lib.check(rc, "adaqp_gemm_tf32x3_f32")
    LAUNCHES["gemm_tf32x3_kernel"] += 1
    return out


def wgrad_supported(dy: Tensor, x: Tensor) -> bool:
    if not (enabled() and dy.is_cuda and x.is_cuda and dy.dtype == x.dtype == torch.float32):
        return False
    if dy.dim() != 2 or x.dim() != 2 or dy.stride(1) != 1 or x.stride(1) != 1 or dy.shape[0] != x.shape[0]:
        return False
    if dy.data_ptr() % 16 or x.data_ptr() % 16 or dy.shape[0] == 0:
        return False
    return bool(_lib.load().adaqp_wgrad_tf32x3_supported(dy.shape[0], dy.shape[1], x.shape[1], dy.stride(0), x.stride(0)))


def gemm_tn(dy: Tensor, x: Tensor) -> Tensor:
    """dy[M, N]^T @ x[M, K] -> [N, K] (the layers' weight gradient), per-CTA partial sums added here."""
    L = _lib.load()
    M, N = dy.shape
    K = x.shape[1]
    grid = L.adaqp_wgrad_tf32x3_grid(M)
    partials = torch.empty((grid, N, K), dtype=torch.float32, device=dy.device)
    rc = L.adaqp_wgrad_tf32x3_f32(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), M, N, K, partials.data_ptr(), grid,
                                  _lib.stream_ptr())
    _lib.check(rc, "adaqp_wgrad_tf32x3_f32")
    LAUNCHES["wgrad_tf32x3_kernel"] += 1
    return partials.sum(0)


class _LinearNK(Function):
    """y = x @ w_nk^T + b with w_nk stored [N, K]."""

    @staticmethod
    def forward(ctx, x: Tensor, w_nk: Tensor, bias):
        ctx.save_for_backward(x, w_nk)
        ctx.has_bias = bias is not None
        return gemm_nt(x, w_nk, bias)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, w_nk = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = db = None
        N = dy.shape[1]
        # a gradient whose row pitch is not a multiple of 16 bytes (the 47-class last layer) is zero-padded once so that
        # TMA can address it; the padded columns meet zero weights / are sliced off the result
        dyp = torch.nn.functional.pad(dy, (0, -N % 4)) if N % 4 and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) else dy
        if ctx.needs_input_grad[0]:
            # dX[M, K] = dY[M, N] @ W_nk[N, K] = dY @ (W_nk^T)[K, N]^T
            wt = w_nk.t().contiguous()
            dx = gemm_nt(dyp, wt) if supported(dyp, wt.shape[0], dyp.shape[1]) else dy @ w_nk
        if ctx.needs_input_grad[1]:
            dw = gemm_tn(dyp, x)[:N] if wgrad_supported(dyp, x) else dy.t() @ x
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dw, db


def linear_nk(x: Tensor, w_nk: Tensor, bias: Tensor = None) -> Tensor:
    if supported(x, w_nk.shape[0], w_nk.shape[1]):
        return _LinearNK.apply(x, w_nk, bias)
    y = x @ w_nk.t()
    return y + bias if bias is not None else y


def linear(x: Tensor, w_kn: Tensor, bias: Tensor = None) -> Tensor:
    """y = x @ w_kn + b with w_kn stored [K, N] (DistGCNConv.weight)."""
    if supported(x, w_kn.shape[1], w_kn.shape[0]):
        return _LinearNK.apply(x, w_kn.t(), bias)
    y = torch.matmul(x, w_kn)
    return y + bias if bias is not None else y
