"""CPU oracle of the codec: layout, properties, and the golden vectors recorded from the
reference's own quant_cuda kernels on a B200 (tests/golden/, oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as O


def py_pack(data, mn, scale, bits, seed, offset):
    """Pure-Python restatement of quantization_cuda_kernel.cu:34-52 for tiny cases."""
    N, F = data.shape
    wpt = 8 // bits
    groups = (N + wpt - 1) // wpt
    out = np.zeros(groups * F, np.uint8)
    for no in range(groups):
        for d in range(F):
            k = no * F + d
            byte = 0
            for ni in range(wpt):
                n = no * wpt + ni
                if n >= N:
                    break
                u = np.float32(O.curand_uniform(seed, k, offset, ni))
                t = np.float32(data[n, d] - mn[n])
                # fmaf: evaluate exactly in float64 (products of two fp32 are exact there) then round once
                t = np.float32(np.float64(t) * np.float64(scale[n]) + np.float64(u))
                v = max(float(np.float64(t) - 0.5), 0.0)
                q = int(np.rint(np.float32(v)))
                byte |= (q << (ni * bits)) & 0xff
            out[k] = byte
    return out


@pytest.mark.parametrize("bits", [1, 2, 4, 8])
def test_pack_matches_python_loop(bits):
    rng = np.random.RandomState(bits)
    N, F = 11, 7
    x = rng.standard_normal((N, F)).astype(np.float32)
    rmin, rmax, scale = O.minmax_scale(x, bits)
    got = O.pack(x, rmin, scale, bits, seed=42, offset=8)
    want = py_pack(x, rmin, scale, bits, 42, 8)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("N,F", [(1, 1), (5, 3), (64, 100), (31, 256)])
def test_roundtrip_error_bound(bits, N, F):
    rng = np.random.RandomState(N * 1000 + F + bits)
    x = rng.standard_normal((N, F)).astype(np.float32)
    rmin, rmax, scale = O.minmax_scale(x, bits)
    packed = O.pack(x, rmin, scale, bits, seed=7, offset=0)
    assert packed.size == O.packed_nbytes(N, F, bits)
    assert O.qsize(N, F, bits) == packed.size + 1
    y = O.unpack(packed, bits, scale, rmin, N, F)
    step = (rmax - rmin) / (2 ** bits - 1)
    ok = np.isfinite(scale)
    err = np.abs(y - x)[ok]
    assert np.all(err <= step[ok, None] * (1 + 1e-5) + 1e-6)


def test_constant_row_scale_inf():
    x = np.full((2, 8), 1.5, np.float32)
    rmin, rmax, scale = O.minmax_scale(x, 4)
    assert np.all(np.isinf(scale))
    packed = O.pack(x, rmin, scale, 4, 1, 0)
    assert np.all(packed == 0)          # 0 * inf = NaN -> fmax(NaN, 0) = 0
    y = O.unpack(packed, 4, scale, rmin, 2, 8)
    np.testing.assert_array_equal(y, x)  # 0 / inf + min


def test_empty_and_ragged():
    assert O.pack(np.zeros((0, 5), np.float32), np.zeros(0), np.zeros(0), 2, 1, 0).size == 0
    x = np.random.RandomState(0).standard_normal((5, 4)).astype(np.float32)
    rmin, _, scale = O.minmax_scale(x, 2)
    p = O.pack(x, rmin, scale, 2, 3, 0)
    assert p.size == 2 * 4   # ceil(5/4) byte-rows
    # rows past N leave their bit-fields zero
    assert np.all((p[4:] >> 2) == 0)


def test_bf16_rounding():
    vals = np.array([1.0, 1.00390625, 1.005859375, 3.14159, -2.5e-8, np.inf, 65504.0], np.float32)
    h = O.to_bf16(vals)
    back = O.from_bf16(h)
    assert back[0] == 1.0
    assert back[1] == 1.0          # tie -> even
    assert back[2] == np.float32(1.0078125)
    assert np.isinf(back[5])
    assert np.all(np.abs(back[[3, 4, 6]] - vals[[3, 4, 6]]) <= np.abs(vals[[3, 4, 6]]) * 2 ** -8)
    assert O.to_bf16(np.array([np.nan], np.float32))[0] == 0x7FC0


def test_mixed_wire_format_layout():
    rng = np.random.RandomState(3)
    S, F = 13, 6
    rows = rng.standard_normal((S, F)).astype(np.float32)
    assign = np.array([2, 8, 4, 4, 2, 8, 8, 2, 4, 2, 2, 8, 4])
    q, prm, valid, off = O.mixed_quantize(rows, assign, seed=5, offset=100)
    n2, n4, n8 = 5, 4, 4
    sizes = [O.qsize(n2, F, 2), O.qsize(n4, F, 4), O.qsize(n8, F, 8)]
    assert q.size == sum(sizes) and prm.shape == (2, S)
    assert off == 100 + sum(O.philox_offset_increment(F, b) for b in (2, 4, 8))
    assert (~valid).sum() == 3 and not valid[sizes[0] - 1] and not valid[sizes[0] + sizes[1] - 1]
    y = O.mixed_dequantize(q, prm, assign, F)
    # dequantized rows are within one (bf16-perturbed) step of the input
    for b in (2, 4, 8):
        ids = np.nonzero(assign == b)[0]
        rng_ = rows[ids].max(1) - rows[ids].min(1)
        assert np.all(np.abs(y[ids] - rows[ids]) <= (rng_ / (2 ** b - 1))[:, None] * 1.02 + 0.02 * np.abs(rows[ids]).max())


# ---------------------------------------------------------------- golden vectors
def _golden(prefix):
    d = os.path.join(os.path.dirname(__file__), "golden")
    if not os.path.isdir(d):
        return []
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.startswith(prefix) and f.endswith(".npz"))


@pytest.mark.parametrize("path", _golden("single_") or [None])
def test_oracle_vs_reference_golden_single(path):
    if path is None:
        pytest.skip("no golden vectors committed yet")
    g = np.load(path)
    x, bits = g["x"], int(g["bits"])
    N, F = x.shape
    rmin, rmax, scale = O.minmax_scale(x, bits)
    np.testing.assert_array_equal(rmin.view(np.uint32), g["rmin"].view(np.uint32))
    np.testing.assert_array_equal(rmax.view(np.uint32), g["rmax"].view(np.uint32))
    np.testing.assert_array_equal(scale.view(np.uint32), g["scale"].view(np.uint32))
    assert int(g["offset_after"]) - int(g["offset"]) == O.philox_offset_increment(F, bits)
    assert int(g["packed_len"]) == O.qsize(N, F, bits)
    payload = O.pack(x, rmin, scale, bits, int(g["seed"]), int(g["offset"]))
    np.testing.assert_array_equal(payload, g["payload"])           # bit-exact integer path
    deq = O.unpack(payload, bits, scale, rmin, N, F)
    np.testing.assert_array_equal(deq.view(np.uint32), g["deq"].view(np.uint32))
    np.testing.assert_array_equal(O.to_bf16(scale), g["scale_bf16"])
    np.testing.assert_array_equal(O.to_bf16(rmin), g["min_bf16"])
    deq16 = O.unpack(payload, bits, O.from_bf16(g["scale_bf16"]), O.from_bf16(g["min_bf16"]), N, F)
    np.testing.assert_array_equal(deq16.view(np.uint32), g["deq_bf16"].view(np.uint32))


@pytest.mark.parametrize("path", _golden("mixed_") or [None])
def test_oracle_vs_reference_golden_mixed(path):
    if path is None:
        pytest.skip("no golden vectors committed yet")
    g = np.load(path)
    x, assign = g["x"], g["assign"]
    q, prm, valid, off = O.mixed_quantize(x, assign, int(g["seed"]), int(g["offset"]))
    assert off == int(g["offset_after"])
    np.testing.assert_array_equal(valid, g["valid"])
    np.testing.assert_array_equal(q[valid], g["qdata"][valid])
    np.testing.assert_array_equal(prm, g["params"])
    deq = O.mixed_dequantize(q, prm, assign, x.shape[1])
    np.testing.assert_array_equal(deq.view(np.uint32), g["deq"].view(np.uint32))


def _half_goldens():
    import glob
    return sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "half_*.npz")))


@pytest.mark.parametrize("path", _half_goldens() or [None])
def test_half_oracle_vs_reference_golden(path):
    """fp16 instantiation: the C restatement of c10::Half arithmetic reproduces the reference kernels' packed bytes
    and dequantised halves bit for bit (goldens recorded on a B200 by oracle/make_golden.py --half)."""
    if path is None:
        pytest.skip("no fp16 goldens recorded yet")
    from oracle import oracle as O
    g = np.load(path)
    bits, seed, off = int(g["bits"]), int(g["seed"]), int(g["offset"])
    N, F = g["x"].shape
    assert int(g["offset_after"]) - off == O.philox_offset_increment(F, bits)
    packed = O.pack_f16(g["x"], g["rmin"], g["scale"], bits, seed, off)
    np.testing.assert_array_equal(packed, g["payload"])
    deq = O.unpack_f16(g["payload"], bits, g["scale"], g["rmin"], N, F)
    got, want = deq.view(np.uint16), g["deq"]
    same = (got == want) | (np.isnan(deq) & np.isnan(want.view(np.float16)))
    assert same.all()


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("N,F", [(37, 13), (64, 100), (5, 256)])
def test_pack_at_windows_equal_slices_of_the_full_call(bits, N, F):
    """oracle_pack_at restates any window of byte-rows of a pack call with the Philox subsequences that window has in the
    full call -- the property tools/parity_check.py relies on to check full-scale exchanges window by window."""
    rng = np.random.RandomState(bits * 1000 + N + F)
    x = rng.standard_normal((N, F)).astype(np.float32)
    x[3] = 2.5                                               # constant row: scale = inf
    mn, _, scale = O.minmax_scale(x, bits)
    seed, off = 77, 12
    full = O.pack(x, mn, scale, bits, seed, off)
    wpt = 8 // bits
    groups = (N + wpt - 1) // wpt
    for g0, g1 in [(0, 1), (0, groups), (groups - 1, groups), (1, max(2, groups // 2)), (groups // 2, groups)]:
        g1 = min(max(g1, g0 + 1), groups)
        r0, r1 = g0 * wpt, min(g1 * wpt, N)
        got = O.pack_at(x[r0:r1], mn[r0:r1], scale[r0:r1], bits, seed, off, g0)
        assert np.array_equal(got, full[g0 * F:g1 * F]), (g0, g1)
    assert O.pack_at(x[:0], mn[:0], scale[:0], bits, seed, off, 0).size == 0
