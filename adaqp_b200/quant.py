"""`quant_cuda` drop-in: pack_single_precision / unpack_single_precision.

Host-side mirror of the reference's pybind module
(AdaQP/util/quantization/src/quantization.cc:23-51): same names, argument order,
argument checks (check.h:22-27 -> RuntimeError), output shapes/dtypes, current-stream
launch and the same consumption of the default CUDA generator
(quantization_cuda_kernel.cu:66-72: one philox_engine_inputs(F * 8/bits) per call).
The arithmetic runs in libadaqp_b200.so through the C ABI; there is no CPU path.

float32 and float16 instantiations, the two dtypes the reference's boundary check admits
(check.h:22-27; its fp64 kernel instantiation is unreachable behind that check).  The boundary
messages on the hot path are fp32; fp16 follows c10::Half arithmetic (csrc/quant.cu).
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib


def _check_float_tensor(t: Tensor, name: str, ndim: int):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor!")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous!")
    if t.dim() != ndim:
        raise RuntimeError(f"The dimension of {name} is not correct!")
    if t.dtype not in (torch.float32, torch.float16):
        raise RuntimeError(f"The type of {name} is not correct!")


def _same_dtype(ref: Tensor, *others):
    """The reference dispatches on one tensor's dtype and takes data_ptr<scalar_t>() of the others, which raises on a mismatch."""
    for name, t in others:
        if t.dtype != ref.dtype:
            raise RuntimeError(f"expected scalar type {ref.dtype} but found {t.dtype} for {name}")


def philox_engine_inputs(device: torch.device, increment: int):
    """ATen CUDAGeneratorImpl::philox_engine_inputs on the default generator of
    `device`: returns (seed, offset) and advances the offset by `increment`
    rounded up to a multiple of 4."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[idx]
    seed = gen.initial_seed()
    offset = gen.get_offset()
    gen.set_offset(offset + ((int(increment) + 3) // 4) * 4)
    return seed, offset


def pack_single_precision(data: Tensor, min: Tensor, max: Tensor, scale: Tensor, bits: int,
                          stochastic: bool) -> Tensor:
    _check_float_tensor(data, "data", 2)
    _check_float_tensor(min, "min", 1)
    _check_float_tensor(max, "max", 1)
    _check_float_tensor(scale, "scale", 1)
    bits = int(bits)
    if bits <= 0 or 8 % bits != 0:
        raise RuntimeError("Expected 8 % bits == 0 to be true, but got false.")
    _same_dtype(data, ("min", min), ("scale", scale))
    L = _lib.load()
    N, F = data.shape
    with torch.cuda.device(data.device):
        packed = torch.empty((L.adaqp_qsize(N, F, bits),), dtype=torch.int8, device=data.device)
        seed, offset = philox_engine_inputs(data.device, F * (8 // bits))
        if not stochastic:
            raise RuntimeError("Expected stochastic to be true, but got false.")
        fn = L.adaqp_pack_f32 if data.dtype == torch.float32 else L.adaqp_pack_f16
        rc = fn(data.data_ptr(), min.data_ptr(), scale.data_ptr(), N, F, bits,
                seed, offset, packed.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "adaqp_pack")
    return packed


def unpack_single_precision(data: Tensor, bits: int, scale: Tensor, min: Tensor, N: int,
                            group_size: int) -> Tensor:
    if not data.is_cuda:
        raise RuntimeError("data must be a CUDA tensor!")
    if not data.is_contiguous():
        raise RuntimeError("data must be contiguous!")
    if data.dim() != 1:
        raise RuntimeError("The dimension of data is not correct!")
    if data.dtype != torch.int8:
        raise RuntimeError("The type of data is not correct!")
    _check_float_tensor(scale, "scale", 1)
    _check_float_tensor(min, "min", 1)
    bits = int(bits)
    if bits <= 0 or 8 % bits != 0:
        raise RuntimeError("Expected 8 % bits == 0 to be true, but got false.")
    L = _lib.load()
    N = int(N)
    F = int(group_size)
    need = L.adaqp_packed_nbytes(N, F, bits)
    if data.numel() < need:
        raise RuntimeError(f"data holds {data.numel()} bytes, {need} needed for N={N}, F={F}, bits={bits}")
    _same_dtype(scale, ("min", min))
    with torch.cuda.device(data.device):
        out = torch.empty((N, F), dtype=scale.dtype, device=data.device)
        fn = L.adaqp_unpack_f32 if scale.dtype == torch.float32 else L.adaqp_unpack_f16
        rc = fn(data.data_ptr(), scale.data_ptr(), min.data_ptr(), N, F, bits, out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "adaqp_unpack")
    return out


def row_minmax_scale(data: Tensor, bits: int):
    """Fused compute_minmax_params + scale of integer_quantize (op_util.py:20-22,41); fp32 only."""
    _check_float_tensor(data, "data", 2)
    if data.dtype != torch.float32:
        raise RuntimeError("row_minmax_scale: float32 only")
    L = _lib.load()
    N, F = data.shape
    with torch.cuda.device(data.device):
        rmin = torch.empty(N, dtype=torch.float32, device=data.device)
        rmax = torch.empty_like(rmin)
        scale = torch.empty_like(rmin)
        if N:
            rc = L.adaqp_row_minmax_f32(data.data_ptr(), N, F, int(bits), rmin.data_ptr(),
                                        rmax.data_ptr(), scale.data_ptr(), _lib.stream_ptr())
            _lib.check(rc, "adaqp_row_minmax_f32")
    return rmin, rmax, scale
