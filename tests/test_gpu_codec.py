"""GPU parity of the quant_cuda drop-in (C ABI -> sm_100a kernels) against the CPU oracle,
the committed golden vectors and -- when oracle/_ref is present -- the reference's own
kernels run side by side.  Integer path: bit-exact.  Dequantized floats: bit-exact (same
IEEE division and add)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import build as obuild

pytestmark = pytest.mark.gpu

CASES = [(1, 1, 8), (7, 13, 2), (8, 100, 4), (5, 602, 8), (33, 256, 2), (64, 256, 4), (129, 300, 8),
         (9, 200, 1), (1000, 256, 2), (4097, 100, 4), (2, 1024, 2), (3, 7, 4)]


@pytest.fixture(scope="module")
def qc():
    from adaqp_b200 import build, quant
    build.build()
    return quant


def _inputs(N, F, seed, kind="normal"):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((N, F)).astype(np.float32)
    if kind == "relu":
        x = np.maximum(x, 0)
    return x


@pytest.mark.parametrize("N,F,bits", CASES)
def test_pack_unpack_vs_oracle(qc, N, F, bits):
    dev = torch.device("cuda:0")
    x = _inputs(N, F, N * 7 + F + bits)
    xt = torch.from_numpy(x).to(dev)
    rmin, rmax, scale = qc.row_minmax_scale(xt, bits)
    o_min, o_max, o_scale = O.minmax_scale(x, bits)
    np.testing.assert_array_equal(rmin.cpu().numpy().view(np.uint32), o_min.view(np.uint32))
    np.testing.assert_array_equal(rmax.cpu().numpy().view(np.uint32), o_max.view(np.uint32))
    np.testing.assert_array_equal(scale.cpu().numpy().view(np.uint32), o_scale.view(np.uint32))
    # torch's own reductions (what the reference calls) agree as well
    assert torch.equal(rmin, torch.min(xt, dim=1)[0]) and torch.equal(rmax, torch.max(xt, dim=1)[0])
    assert torch.equal(scale, (2 ** bits - 1) / (rmax - rmin))

    torch.cuda.manual_seed(4321 + N)
    gen = torch.cuda.default_generators[0]
    seed, off0 = gen.initial_seed(), gen.get_offset()
    packed = qc.pack_single_precision(xt, rmin, rmax, scale, bits, True)
    assert packed.dtype == torch.int8 and packed.numel() == O.qsize(N, F, bits)
    assert gen.get_offset() - off0 == O.philox_offset_increment(F, bits)
    want = O.pack(x, o_min, o_scale, bits, seed, off0)
    got = packed.cpu().numpy().view(np.uint8)[:want.size]
    np.testing.assert_array_equal(got, want)

    deq = qc.unpack_single_precision(packed, bits, scale, rmin, N, F)
    want_deq = O.unpack(want, bits, o_scale, o_min, N, F)
    np.testing.assert_array_equal(deq.cpu().numpy().view(np.uint32), want_deq.view(np.uint32))


def test_edge_rows_and_errors(qc):
    dev = torch.device("cuda:0")
    x = _inputs(6, 64, 5)
    x[0] = 3.25
    x[1] = 0
    x[2, ::2] = 1e30
    x[3] = np.where(np.arange(64) % 2 == 0, -1.0, 1.0)
    xt = torch.from_numpy(x).to(dev)
    for bits in (2, 4, 8):
        rmin, rmax, scale = qc.row_minmax_scale(xt, bits)
        torch.cuda.manual_seed(5)
        gen = torch.cuda.default_generators[0]
        seed, off = gen.initial_seed(), gen.get_offset()
        p = qc.pack_single_precision(xt, rmin, rmax, scale, bits, True)
        o_min, _, o_scale = O.minmax_scale(x, bits)
        want = O.pack(x, o_min, o_scale, bits, seed, off)
        np.testing.assert_array_equal(p.cpu().numpy().view(np.uint8)[:want.size], want)
        d = qc.unpack_single_precision(p, bits, scale, rmin, 6, 64)
        np.testing.assert_array_equal(d.cpu().numpy().view(np.uint32),
                                      O.unpack(want, bits, o_scale, o_min, 6, 64).view(np.uint32))
    with pytest.raises(RuntimeError):
        qc.pack_single_precision(xt, rmin, rmax, scale, 3, True)
    with pytest.raises(RuntimeError):
        qc.pack_single_precision(xt, rmin, rmax, scale, 4, False)
    with pytest.raises(RuntimeError):
        qc.pack_single_precision(xt.cpu(), rmin, rmax, scale, 4, True)
    with pytest.raises(RuntimeError):
        qc.pack_single_precision(xt.t(), rmin, rmax, scale, 4, True)
    with pytest.raises(RuntimeError):
        qc.unpack_single_precision(p.to(torch.uint8), 8, scale, rmin, 6, 64)
    # empty input
    e = qc.pack_single_precision(torch.zeros(0, 8, device=dev), torch.zeros(0, device=dev),
                                 torch.zeros(0, device=dev), torch.zeros(0, device=dev), 2, True)
    assert e.numel() == 1


def test_nonzero_phase_offsets(qc):
    """Philox offsets that are not multiples of 4 (never produced by torch, allowed by the ABI)."""
    from adaqp_b200 import _lib
    L = _lib.load()
    dev = torch.device("cuda:0")
    x = _inputs(9, 20, 77)
    xt = torch.from_numpy(x).to(dev)
    for bits in (1, 2, 4, 8):
        o_min, _, o_scale = O.minmax_scale(x, bits)
        mn, sc = torch.from_numpy(o_min).to(dev), torch.from_numpy(o_scale).to(dev)
        for off in (1, 2, 3, 7, 2 ** 34 + 5):
            out = torch.zeros(O.packed_nbytes(9, 20, bits), dtype=torch.uint8, device=dev)
            _lib.check(L.adaqp_pack_f32(xt.data_ptr(), mn.data_ptr(), sc.data_ptr(), 9, 20, bits,
                                        2 ** 40 + 17, off, out.data_ptr(), _lib.stream_ptr()))
            np.testing.assert_array_equal(out.cpu().numpy(), O.pack(x, o_min, o_scale, bits, 2 ** 40 + 17, off))


def _golden(prefix):
    d = os.path.join(os.path.dirname(__file__), "golden")
    if not os.path.isdir(d):
        return []
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.startswith(prefix) and f.endswith(".npz"))


@pytest.mark.parametrize("path", _golden("single_") or [None])
def test_kernels_vs_reference_golden(qc, path):
    if path is None:
        pytest.skip("no golden vectors committed yet")
    from adaqp_b200 import _lib
    L = _lib.load()
    g = np.load(path)
    dev = torch.device("cuda:0")
    x, bits = g["x"], int(g["bits"])
    N, F = x.shape
    xt = torch.from_numpy(x).to(dev)
    rmin, rmax, scale = qc.row_minmax_scale(xt, bits)
    np.testing.assert_array_equal(scale.cpu().numpy().view(np.uint32), g["scale"].view(np.uint32))
    out = torch.zeros(len(g["payload"]), dtype=torch.uint8, device=dev)
    _lib.check(L.adaqp_pack_f32(xt.data_ptr(), rmin.data_ptr(), scale.data_ptr(), N, F, bits,
                                int(g["seed"]), int(g["offset"]), out.data_ptr(), _lib.stream_ptr()))
    np.testing.assert_array_equal(out.cpu().numpy(), g["payload"])
    packed = torch.cat([out, torch.zeros(1, dtype=torch.uint8, device=dev)]).view(torch.int8)
    deq = qc.unpack_single_precision(packed, bits, scale, rmin, N, F)
    np.testing.assert_array_equal(deq.cpu().numpy().view(np.uint32), g["deq"].view(np.uint32))


@pytest.mark.skipif(not obuild.ref_available(), reason="oracle/_ref (reference quant_cuda) not built")
@pytest.mark.parametrize("N,F,bits", [(33, 256, 2), (64, 100, 4), (17, 602, 8), (5000, 256, 4), (9, 200, 1)])
def test_side_by_side_with_reference_kernels(qc, N, F, bits):
    """Same inputs, same generator state: reference kernel vs ours, byte for byte."""
    ref = obuild.load_ref()
    dev = torch.device("cuda:0")
    xt = torch.from_numpy(_inputs(N, F, 31 + N, "relu")).to(dev)
    rmin, rmax = torch.min(xt, dim=1)[0], torch.max(xt, dim=1)[0]
    scale = (2 ** bits - 1) / (rmax - rmin)
    gen = torch.cuda.default_generators[0]
    torch.cuda.manual_seed(2024)
    a = ref.pack_single_precision(xt, rmin, rmax, scale, bits, True)
    off_ref = gen.get_offset()
    torch.cuda.manual_seed(2024)
    b = qc.pack_single_precision(xt, rmin, rmax, scale, bits, True)
    assert gen.get_offset() == off_ref
    assert a.shape == b.shape and a.dtype == b.dtype
    assert torch.equal(a[:-1], b[:-1])
    da = ref.unpack_single_precision(a, bits, scale, rmin, N, F)
    db = qc.unpack_single_precision(b, bits, scale, rmin, N, F)
    assert torch.equal(da, db)


def test_dtype_boundary_matches_reference_checks():
    """check.h:22-27 admits float32 and float16 and rejects everything else with 'The type of <name> is not correct!'
    (the fp64 kernel instantiation of quantization_cuda_kernel.cu:81 is unreachable behind that check); mixing dtypes
    fails like data_ptr<scalar_t>() does; a refused call does not consume the generator."""
    from adaqp_b200 import quant
    dev = torch.device("cuda:0")
    x = torch.randn(8, 32, device=dev)
    mn, mx = x.min(1)[0], x.max(1)[0]
    sc = 15.0 / (mx - mn)
    gen = torch.cuda.default_generators[0]
    off = gen.get_offset()
    for bad in (torch.float64, torch.bfloat16, torch.int32):
        with pytest.raises(RuntimeError, match="The type of data is not correct!"):
            quant.pack_single_precision(x.to(bad), mn, mx, sc, 4, True)
    with pytest.raises(RuntimeError, match="The type of scale is not correct!"):
        quant.pack_single_precision(x, mn, mx, sc.double(), 4, True)
    with pytest.raises(RuntimeError, match="expected scalar type"):
        quant.pack_single_precision(x.half(), mn, mx, sc, 4, True)
    q = quant.pack_single_precision(x, mn, mx, sc, 4, True)
    with pytest.raises(RuntimeError, match="expected scalar type"):
        quant.unpack_single_precision(q, 4, sc.half(), mn, 8, 32)
    with pytest.raises(RuntimeError, match="The type of scale is not correct!"):
        quant.unpack_single_precision(q, 4, sc.double(), mn, 8, 32)
    assert gen.get_offset() == off + 64, "only the one successful call consumed the generator (F * 8/bits = 64)"


@pytest.mark.parametrize("N,F,bits", [(7, 13, 2), (8, 100, 4), (5, 602, 8), (33, 256, 2), (9, 200, 1), (1, 1, 8)])
def test_half_codec_vs_oracle_and_reference(qc, N, F, bits):
    """fp16 instantiation (c10::Half arithmetic): our kernels == the C oracle == the reference's own kernels, bit for bit,
    including rows whose half scale overflows to inf and constant rows (scale inf -> NaN -> 0)."""
    from oracle import oracle as O
    from oracle import build as obuild
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(N * 1000 + F + bits)
    x = rng.standard_normal((N, F)).astype(np.float32)
    if N > 3:
        x[0] = 2.5
        x[1] = 1.0
        x[1, ::2] = 1.001
        x[2] *= 1000.0
    xt = torch.from_numpy(x).to(dev).half()
    rmin, rmax = torch.min(xt, dim=1)[0], torch.max(xt, dim=1)[0]
    scale = ((2 ** bits - 1) / (rmax - rmin)).to(torch.float16)
    gen = torch.cuda.default_generators[0]
    torch.cuda.manual_seed(99)
    seed, off = gen.initial_seed(), gen.get_offset()
    packed = qc.pack_single_precision(xt, rmin, rmax, scale, bits, True)
    assert gen.get_offset() - off == O.philox_offset_increment(F, bits)
    payload = O.packed_nbytes(N, F, bits)
    u16 = lambda t: t.view(torch.int16).cpu().numpy().view(np.uint16)
    want = O.pack_f16(u16(xt), u16(rmin), u16(scale), bits, seed, off)
    np.testing.assert_array_equal(packed[:payload].cpu().numpy().view(np.uint8), want)
    deq = qc.unpack_single_precision(packed, bits, scale, rmin, N, F)
    assert deq.dtype == torch.float16 and deq.shape == (N, F)
    want_d = O.unpack_f16(want, bits, u16(scale), u16(rmin), N, F)
    got_d = u16(deq)
    assert ((got_d == want_d.view(np.uint16)) | (np.isnan(want_d) & np.isnan(got_d.view(np.float16)))).all()
    if obuild.ref_available():
        ref = obuild.load_ref()
        torch.cuda.manual_seed(99)
        rp = ref.pack_single_precision(xt, rmin, rmax, scale, bits, True)
        assert torch.equal(rp[:payload], packed[:payload])
        rd = ref.unpack_single_precision(rp, bits, scale, rmin, N, F)
        assert ((u16(rd) == got_d) | (torch.isnan(rd) & torch.isnan(deq)).cpu().numpy()).all()
