"""tcgen05 3xTF32 GEMM (csrc/gemm.cu) vs a float64 torch reference and vs torch's fp32 matmul.

Tolerance: per product the split drops a_lo*b_lo (<= 2^-22 |ab|) and truncates the lo halves to tf32
(<= 2^-22 |ab| each), so |err| <= ~1e-6 * sum|a||b| per output; asserted at 2e-6 relative to the
row-wise L1 mass (torch's own fp32 SIMT GEMM measures ~1e-7 on the same inputs; a plain 1xTF32 GEMM
~5e-4, which the test also checks we are far below)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dense():
    from adaqp_b200 import build
    build.build()
    from adaqp_b200 import dense as d
    return d


@pytest.mark.parametrize("kb", [32, 16], ids=["K32-sw128-2stages", "K16-sw64-4stages"])
@pytest.mark.parametrize("M,K,N", [(128, 32, 16), (1000, 256, 256), (4099, 100, 256), (333, 256, 47), (257, 200, 256),
                                   (70001, 256, 256), (5, 8, 8), (129, 300, 100)])
def test_matches_float64(dense, M, K, N, kb):
    from adaqp_b200 import _lib
    old = _lib.get_option("gemm_block_k")
    _lib.set_option("gemm_block_k", kb)
    try:
        _matches_float64(dense, M, K, N)
    finally:
        _lib.set_option("gemm_block_k", old)


def _matches_float64(dense, M, K, N):
    torch.manual_seed(M + K + N)
    dev = torch.device("cuda:0")
    x = torch.randn(M, K, device=dev)
    x[::7] *= 100.0
    w = torch.randn(N, K, device=dev) * 0.1
    b = torch.randn(N, device=dev)
    assert dense.supported(x, N, K)
    got = dense.gemm_nt(x, w, b)
    want = (x.double() @ w.double().t() + b.double())
    mass = (x.double().abs() @ w.double().abs().t()) + b.double().abs()
    err = ((got.double() - want).abs() / mass).max().item()
    ref32 = ((x @ w.t() + b).double() - want).abs().div(mass).max().item()
    assert err <= 2e-6, (err, ref32)
    got2 = dense.gemm_nt(x, w)
    assert torch.equal(got2 + b, got) or ((got2 + b) - got).abs().max() <= 1e-5 * got.abs().max()


def test_autograd_and_fallbacks(dense):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = torch.randn(3000, 256, device=dev, requires_grad=True)
    w = (torch.randn(256, 128, device=dev) * 0.1).requires_grad_()
    b = torch.zeros(128, device=dev, requires_grad=True)
    y = dense.linear(x, w, b)
    y.square().sum().backward()
    x2, w2, b2 = (t.detach().double().requires_grad_() for t in (x, w, b))
    (x2 @ w2 + b2).square().sum().backward()
    for g, g2 in ((x.grad, x2.grad), (w.grad, w2.grad), (b.grad, b2.grad)):
        assert ((g.double() - g2).abs().max() / g2.abs().max()).item() < 1e-5
    # the 47-class last layer: forward on the kernel, its 47-wide gradient zero-padded to 48 columns for dX and dW
    before = dict(dense.LAUNCHES)
    w47 = (torch.randn(256, 47, device=dev) * 0.1).requires_grad_()
    b47 = torch.zeros(47, device=dev, requires_grad=True)
    xl = x.detach().clone().requires_grad_()
    torch.log_softmax(dense.linear(xl, w47, b47), 1)[:, 3].sum().backward()
    x4, w4, b4 = (t.detach().double().requires_grad_() for t in (xl, w47, b47))
    torch.log_softmax(x4 @ w4 + b4, 1)[:, 3].sum().backward()
    for g, g2 in ((xl.grad, x4.grad), (w47.grad, w4.grad), (b47.grad, b4.grad)):
        assert g.shape == g2.shape and ((g.double() - g2).abs().max() / g2.abs().max()).item() < 1e-5
    assert dense.LAUNCHES["gemm_tf32x3_kernel"] == before["gemm_tf32x3_kernel"] + 2
    assert dense.LAUNCHES["wgrad_tf32x3_kernel"] == before["wgrad_tf32x3_kernel"] + 1
    # nn.Linear storage
    lin = torch.nn.Linear(256, 64, bias=False).to(dev)
    y = dense.linear_nk(x.detach(), lin.weight)
    assert ((y.double() - x.detach().double() @ lin.weight.double().t()).abs().max() / y.abs().max()).item() < 1e-5
    # shapes the kernel refuses fall back to torch.matmul (F = 602: 8-byte row pitch; strided views)
    x3 = torch.randn(100, 602, device=dev)
    assert not dense.supported(x3, 256, 602)
    w3 = torch.randn(602, 256, device=dev)
    assert torch.equal(dense.linear(x3, w3), torch.matmul(x3, w3))


@pytest.mark.parametrize("M,N,K", [(16, 128, 32), (1000, 256, 256), (4099, 256, 100), (70001, 256, 256), (333, 128, 200),
                                   (50, 64, 256), (7, 256, 8), (3000, 100, 256), (3001, 48, 256)])
def test_weight_gradient_split_k(dense, M, N, K):
    """dW = dY^T X on the tensor cores (MN-major operands, split over the rows) vs float64."""
    torch.manual_seed(M + N + K)
    dev = torch.device("cuda:0")
    dy = torch.randn(M, N, device=dev) * 1e-3
    dy[::5] = 0.0                                 # all-zero gradient rows (non-training nodes)
    x = torch.relu(torch.randn(M, K, device=dev))
    assert dense.wgrad_supported(dy, x)
    got = dense.gemm_tn(dy, x)
    want = dy.double().t() @ x.double()
    mass = dy.double().abs().t() @ x.double().abs()
    err = ((got.double() - want).abs() / (mass + 1e-30)).max().item()
    assert got.shape == (N, K) and err <= 2e-6, err
    again = dense.gemm_tn(dy, x)
    assert torch.equal(got, again), "deterministic for a fixed grid"


def test_wgrad_fallback_shapes(dense):
    dev = torch.device("cuda:0")
    assert not dense.wgrad_supported(torch.randn(100, 47, device=dev), torch.randn(100, 256, device=dev))   # 188-byte pitch
    assert not dense.wgrad_supported(torch.randn(100, 256, device=dev), torch.randn(100, 602, device=dev))  # K > 256
