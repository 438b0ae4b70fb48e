"""Host side of the tcgen05 dense path (adaqp_b200/dense.py) that needs no GPU: the tf32 operand split is exact,
and on CPU tensors / unsupported shapes `linear` is torch.matmul (the reference's arithmetic, distGCN.py:45)."""
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


def test_backward_pads_odd_width_gradient(monkeypatch):
    """_LinearNK.backward with a 47-wide gradient: the kernels see a zero-padded [M, 48] gradient, the results come back in
    the original shapes and equal plain autograd.  The two C-ABI calls are replaced by float64 matmuls that enforce the
    kernels' shape rules (row pitch a multiple of 4 floats)."""
    seen = []

    def fake_gemm_nt(x, bt, bias=None):
        bt = dense._pad_cols(bt)
        assert x.shape[1] % 4 == 0 and bt.shape[1] == x.shape[1]
        seen.append(("nt", tuple(x.shape), tuple(bt.shape)))
        y = (x.double() @ bt.double().t()).float()
        return y + bias if bias is not None else y

    def fake_gemm_tn(dy, x):
        assert dy.shape[1] % 4 == 0 and x.shape[1] % 4 == 0
        seen.append(("tn", tuple(dy.shape), tuple(x.shape)))
        return (dy.double().t() @ x.double()).float()

    monkeypatch.setattr(dense, "gemm_nt", fake_gemm_nt)
    monkeypatch.setattr(dense, "gemm_tn", fake_gemm_tn)
    monkeypatch.setattr(dense, "supported", lambda x, n, k: x.shape[1] % 4 == 0)
    monkeypatch.setattr(dense, "wgrad_supported", lambda dy, x: dy.shape[1] % 4 == 0 and x.shape[1] % 4 == 0)
    torch.manual_seed(1)
    x = torch.randn(33, 256, requires_grad=True)
    w = (torch.randn(256, 47) * 0.1).requires_grad_()
    b = torch.zeros(47, requires_grad=True)
    torch.log_softmax(dense.linear(x, w, b), 1)[:, 5].sum().backward()
    x2, w2, b2 = (t.detach().clone().requires_grad_() for t in (x, w, b))
    torch.log_softmax(x2 @ w2 + b2, 1)[:, 5].sum().backward()
    assert ("nt", (33, 48), (256, 48)) in seen and ("tn", (33, 48), (33, 256)) in seen
    for g, g2 in ((x.grad, x2.grad), (w.grad, w2.grad), (b.grad, b2.grad)):
        assert g.shape == g2.shape and torch.allclose(g, g2, rtol=1e-4, atol=1e-6)
