"""Host side of the tcgen05 dense path (adaqp_b200/dense.py) that needs no GPU: the tf32 operand split is exact,
and on CPU tensors / unsupported shapes `linear` is torch.matmul (the reference's arithmetic, distGCN.py:45)."""
import numpy as np
import torch

from adaqp_b200 import dense


def test_split_is_exact_and_hi_is_tf32():
    torch.manual_seed(0)
    t = torch.randn(257, 100) * torch.logspace(-6, 6, 100)
    t[0, :4] = torch.tensor([0.0, -0.0, float("inf"), 1e-45])
    hi, lo = dense.split_tf32(t)
    assert (hi.view(torch.int32) & 0x1FFF).eq(0).all(), "hi keeps 10 mantissa bits (tf32)"
    fin = torch.isfinite(t)
    assert torch.equal((hi + lo)[fin], t[fin]), "hi + lo == a exactly"
    # lo carries at most the 13 dropped bits: |lo| < 2^-10 |a|
    nz = fin & (t.abs() >= torch.finfo(torch.float32).tiny)          # normal numbers (a subnormal is all "lo")
    assert (lo[nz].abs() <= t[nz].abs() * 2.0 ** -10).all()


def test_cpu_and_unsupported_fall_back_to_matmul():
    x = torch.randn(50, 24, requires_grad=True)
    w = torch.randn(24, 8, requires_grad=True)
    b = torch.randn(8, requires_grad=True)
    assert not dense.supported(x, 8, 24)
    y = dense.linear(x, w, b)
    assert torch.equal(y, torch.matmul(x, w) + b)
    lin = torch.nn.Linear(24, 8, bias=False)
    assert torch.equal(dense.linear_nk(x, lin.weight), x @ lin.weight.t())
    y.sum().backward()
    assert x.grad is not None and w.grad is not None and b.grad is not None


def test_pad_cols_keeps_values():
    t = torch.arange(15.0).reshape(3, 5)
    p = dense._pad_cols(t)
    assert p.shape == (3, 8) and torch.equal(p[:, :5], t) and p[:, 5:].eq(0).all() and p.stride(0) % 4 == 0
