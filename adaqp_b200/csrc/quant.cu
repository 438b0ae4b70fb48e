// Stand-alone codec kernels: the drop-in replacement of the reference's quant_cuda
// extension (AdaQP/util/quantization/src/quantization_cuda_kernel.cu:34-52,106-122)
// plus the fused row min/max/scale reduction of AdaQP/model/op_util.py:20-22,41.
//
// The reference launches ceil(N/wpt) blocks of F threads (F <= 1024, occupancy set by
// F) and runs Philox two to three times per thread.  Here the packed stream is a flat
// index space k = no*F + d; a thread owns 4 consecutive bytes (one 32-bit store), one
// Philox block per byte, loads coalesced along d.  HBM-bound for bits <= 4; for 8-bit
// the mandated per-byte Philox (parity with curand_init(seed, k, offset)) makes it
// integer-ALU bound.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int kPackThreads = 256;
constexpr int kBytesPerThread = 4;

template <int BITS, bool VEC_STORE>
__global__ void __launch_bounds__(kPackThreads)
pack_flat_kernel(const float *__restrict__ data, const float *__restrict__ mn,
                 const float *__restrict__ scale, int64_t N, int64_t F, int64_t total,
                 uint64_t seed, uint64_t offset, uint8_t *__restrict__ packed) {
    constexpr int WPT = 8 / BITS;
    const int64_t k0 = ((int64_t)blockIdx.x * kPackThreads + threadIdx.x) * kBytesPerThread;
    if (k0 >= total) return;
    int64_t no = k0 / F;
    int64_t d = k0 - no * F;
    uint32_t word = 0;
    const int nb = (total - k0) < kBytesPerThread ? (int)(total - k0) : kBytesPerThread;
#pragma unroll
    for (int j = 0; j < kBytesPerThread; ++j) {
        if (j < nb) {
            float u[WPT];
            byte_noise<WPT>(seed, (uint64_t)(k0 + j), offset, u);
            uint32_t byte = 0;
#pragma unroll
            for (int ni = 0; ni < WPT; ++ni) {
                const int64_t n = no * WPT + ni;
                if (n < N) {
                    const int q = quantize_one(__ldg(data + n * F + d), __ldg(mn + n), __ldg(scale + n), u[ni]);
                    byte |= ((uint32_t)q << (ni * BITS));
                }
            }
            byte &= 0xffu;  // uint8_t local_packed truncation (quantization_cuda_kernel.cu:49)
            if (VEC_STORE) word |= byte << (8 * j);
            else packed[k0 + j] = (uint8_t)byte;
            if (++d == F) { d = 0; ++no; }
        }
    }
    if (VEC_STORE) {
        if (nb == kBytesPerThread) {
            *reinterpret_cast<uint32_t *>(packed + k0) = word;
        } else {
            for (int j = 0; j < nb; ++j) packed[k0 + j] = (uint8_t)(word >> (8 * j));
        }
    }
}

template <int BITS, bool VEC>
__global__ void __launch_bounds__(kPackThreads)
unpack_flat_kernel(const uint8_t *__restrict__ packed, const float *__restrict__ scale,
                   const float *__restrict__ mn, int64_t N, int64_t F, int64_t total,
                   float *__restrict__ out) {
    constexpr int WPT = 8 / BITS;
    constexpr uint32_t MASK = (1u << BITS) - 1u;
    const int64_t k0 = ((int64_t)blockIdx.x * kPackThreads + threadIdx.x) * kBytesPerThread;
    if (k0 >= total) return;
    if (VEC) {
        // F % 4 == 0 and 16-byte aligned pointers: the 4 bytes share one byte-row.
        const int64_t no = k0 / F;
        const int64_t d = k0 - no * F;
        const uint32_t word = *reinterpret_cast<const uint32_t *>(packed + k0);
#pragma unroll
        for (int ni = 0; ni < WPT; ++ni) {
            const int64_t n = no * WPT + ni;
            if (n < N) {
                const float s = __ldg(scale + n), m = __ldg(mn + n);
                const float qz = __fdiv_rn(0.f, s);
                float4 v;
                v.x = __fadd_rn(dequant_div((word >> (ni * BITS)) & MASK, s, qz), m);
                v.y = __fadd_rn(dequant_div((word >> (8 + ni * BITS)) & MASK, s, qz), m);
                v.z = __fadd_rn(dequant_div((word >> (16 + ni * BITS)) & MASK, s, qz), m);
                v.w = __fadd_rn(dequant_div((word >> (24 + ni * BITS)) & MASK, s, qz), m);
                *reinterpret_cast<float4 *>(out + n * F + d) = v;
            }
        }
    } else {
        int64_t no = k0 / F;
        int64_t d = k0 - no * F;
        const int nb = (total - k0) < kBytesPerThread ? (int)(total - k0) : kBytesPerThread;
        for (int j = 0; j < nb; ++j) {
            const uint32_t byte = packed[k0 + j];
#pragma unroll
            for (int ni = 0; ni < WPT; ++ni) {
                const int64_t n = no * WPT + ni;
                if (n < N) {
                    const float q = __fdiv_rn((float)((byte >> (ni * BITS)) & MASK), __ldg(scale + n));
                    out[n * F + d] = __fadd_rn(q, __ldg(mn + n));
                }
            }
            if (++d == F) { d = 0; ++no; }
        }
    }
}

// ---- fp16 instantiation of the reference's dispatch (AT_DISPATCH_FLOATING_TYPES_AND_HALF, .cu:81,138; admitted by
// check.h:22-27).  scalar_t = c10::Half evaluates every Half (op) Half in float and rounds back to half, and
// Half + float in float (c10/util/Half-inl.h):
//   t1 = half(float(x) - float(min)); t2 = half(float(t1) * float(scale)); f = float(t2) + noise;
//   q = float2int_rn(fmax(f - 0.5, 0))     (t2 <= 65504: the fp32 subtraction equals the reference's double one)
// Not on the hot path (every boundary message is fp32): one packed byte per thread, no vector paths.
template <int BITS>
__global__ void __launch_bounds__(kPackThreads)
pack_flat_half_kernel(const __half *__restrict__ data, const __half *__restrict__ mn, const __half *__restrict__ scale,
                      int64_t N, int64_t F, int64_t total, uint64_t seed, uint64_t offset, uint8_t *__restrict__ packed) {
    constexpr int WPT = 8 / BITS;
    const int64_t k = (int64_t)blockIdx.x * kPackThreads + threadIdx.x;
    if (k >= total) return;
    const int64_t no = k / F, d = k - no * F;
    float u[WPT];
    byte_noise<WPT>(seed, (uint64_t)k, offset, u);
    uint32_t byte = 0;
#pragma unroll
    for (int ni = 0; ni < WPT; ++ni) {
        const int64_t n = no * WPT + ni;
        if (n < N) {
            const __half t1 = __float2half_rn(__fsub_rn(__half2float(data[n * F + d]), __half2float(mn[n])));
            const __half t2 = __float2half_rn(__fmul_rn(__half2float(t1), __half2float(scale[n])));
            const float f = __fadd_rn(__half2float(t2), u[ni]);
            const int q = __float2int_rn(fmaxf(__fsub_rn(f, 0.5f), 0.0f));      // inf -> INT_MAX, NaN -> 0 as in the reference
            byte |= ((uint32_t)q << (ni * BITS));
        }
    }
    packed[k] = (uint8_t)byte;
}

template <int BITS>
__global__ void __launch_bounds__(kPackThreads)
unpack_flat_half_kernel(const uint8_t *__restrict__ packed, const __half *__restrict__ scale, const __half *__restrict__ mn,
                        int64_t N, int64_t F, int64_t total, __half *__restrict__ out) {
    constexpr int WPT = 8 / BITS;
    constexpr uint32_t MASK = (1u << BITS) - 1u;
    const int64_t k = (int64_t)blockIdx.x * kPackThreads + threadIdx.x;
    if (k >= total) return;
    const int64_t no = k / F, d = k - no * F;
    const uint32_t byte = packed[k];
#pragma unroll
    for (int ni = 0; ni < WPT; ++ni) {
        const int64_t n = no * WPT + ni;
        if (n < N) {
            const __half q = __float2half_rn((float)((byte >> (ni * BITS)) & MASK));
            const __half dv = __float2half_rn(__fdiv_rn(__half2float(q), __half2float(scale[n])));
            out[n * F + d] = __float2half_rn(__fadd_rn(__half2float(dv), __half2float(mn[n])));
        }
    }
}

// One warp per row; NaN-propagating like torch.min/max(dim=1).
__global__ void __launch_bounds__(256)
row_minmax_kernel(const float *__restrict__ data, int64_t N, int64_t F, float levels,
                  float *__restrict__ rmin, float *__restrict__ rmax, float *__restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = warp; n < N; n += nwarps) {
        const float *row = data + n * F;
        float lo = INFINITY, hi = -INFINITY;
        bool nan = false;
        for (int64_t d = lane; d < F; d += 32) {
            const float v = __ldg(row + d);
            nan |= (v != v);
            lo = fminf(lo, v);
            hi = fmaxf(hi, v);
        }
        lo = warp_min(lo);
        hi = warp_max(hi);
        nan = __any_sync(ADAQP_FULL_MASK, nan);
        if (nan) { lo = __int_as_float(0x7fc00000); hi = lo; }
        if (lane == 0) {
            if (rmin) rmin[n] = lo;
            if (rmax) rmax[n] = hi;
            if (scale) scale[n] = __fmul_rn(__frcp_rn(__fsub_rn(hi, lo)), levels);  // torch: reciprocal(range) * levels
        }
    }
}

inline bool valid_bits(int bits) { return bits == 1 || bits == 2 || bits == 4 || bits == 8; }

}  // namespace

extern "C" {

int64_t adaqp_packed_nbytes(int64_t N, int64_t F, int bits) {
    if (!valid_bits(bits) || N < 0 || F < 0) return -1;
    const int64_t wpt = 8 / bits;
    return ((N + wpt - 1) / wpt) * F;
}

int64_t adaqp_qsize(int64_t N, int64_t F, int bits) {
    if (!valid_bits(bits) || N < 0 || F < 0) return -1;
    const int64_t wpt = 8 / bits;
    const int64_t n_round = N + (wpt - N % wpt) % wpt;
    return ((int64_t)bits * n_round * F + 8) / 8;
}

int adaqp_pack_f32(const float *data, const float *mn, const float *scale, int64_t N, int64_t F,
                   int bits, uint64_t seed, uint64_t offset, uint8_t *packed, void *stream) {
    ADAQP_REQUIRE(valid_bits(bits), ADAQP_EINVAL, "adaqp_pack_f32: 8 %% bits != 0 (bits=%d)", bits);
    ADAQP_REQUIRE(N >= 0 && F >= 0, ADAQP_EINVAL, "adaqp_pack_f32: negative shape");
    const int64_t total = adaqp_packed_nbytes(N, F, bits);
    if (total == 0) return 0;
    ADAQP_REQUIRE(data && mn && scale && packed, ADAQP_EINVAL, "adaqp_pack_f32: null pointer");
    const int64_t threads = (total + kBytesPerThread - 1) / kBytesPerThread;
    const int64_t blocks = (threads + kPackThreads - 1) / kPackThreads;
    ADAQP_REQUIRE(blocks < (1ll << 31), ADAQP_ELIMIT, "adaqp_pack_f32: too many bytes");
    const bool vec = (reinterpret_cast<uintptr_t>(packed) & 3u) == 0;
    cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_PACK(B)                                                                             \
    do {                                                                                           \
        if (vec) pack_flat_kernel<B, true><<<(unsigned)blocks, kPackThreads, 0, s>>>(               \
                data, mn, scale, N, F, total, seed, offset, packed);                               \
        else pack_flat_kernel<B, false><<<(unsigned)blocks, kPackThreads, 0, s>>>(                  \
                data, mn, scale, N, F, total, seed, offset, packed);                               \
    } while (0)
    switch (bits) {
        case 1: LAUNCH_PACK(1); break;
        case 2: LAUNCH_PACK(2); break;
        case 4: LAUNCH_PACK(4); break;
        default: LAUNCH_PACK(8); break;
    }
#undef LAUNCH_PACK
    return adaqp_check_launch("pack_flat_kernel");
}

int adaqp_unpack_f32(const uint8_t *packed, const float *scale, const float *mn, int64_t N,
                     int64_t F, int bits, float *out, void *stream) {
    ADAQP_REQUIRE(valid_bits(bits), ADAQP_EINVAL, "adaqp_unpack_f32: 8 %% bits != 0 (bits=%d)", bits);
    ADAQP_REQUIRE(N >= 0 && F >= 0, ADAQP_EINVAL, "adaqp_unpack_f32: negative shape");
    const int64_t total = adaqp_packed_nbytes(N, F, bits);
    if (total == 0) return 0;
    ADAQP_REQUIRE(packed && scale && mn && out, ADAQP_EINVAL, "adaqp_unpack_f32: null pointer");
    const int64_t threads = (total + kBytesPerThread - 1) / kBytesPerThread;
    const int64_t blocks = (threads + kPackThreads - 1) / kPackThreads;
    ADAQP_REQUIRE(blocks < (1ll << 31), ADAQP_ELIMIT, "adaqp_unpack_f32: too many bytes");
    const bool vec = (F % 4 == 0) && (reinterpret_cast<uintptr_t>(packed) & 3u) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_UNPACK(B)                                                                           \
    do {                                                                                           \
        if (vec) unpack_flat_kernel<B, true><<<(unsigned)blocks, kPackThreads, 0, s>>>(             \
                packed, scale, mn, N, F, total, out);                                              \
        else unpack_flat_kernel<B, false><<<(unsigned)blocks, kPackThreads, 0, s>>>(                \
                packed, scale, mn, N, F, total, out);                                              \
    } while (0)
    switch (bits) {
        case 1: LAUNCH_UNPACK(1); break;
        case 2: LAUNCH_UNPACK(2); break;
        case 4: LAUNCH_UNPACK(4); break;
        default: LAUNCH_UNPACK(8); break;
    }
#undef LAUNCH_UNPACK
    return adaqp_check_launch("unpack_flat_kernel");
}

int adaqp_pack_f16(const void *data, const void *mn, const void *scale, int64_t N, int64_t F, int bits, uint64_t seed,
                   uint64_t offset, uint8_t *packed, void *stream) {
    ADAQP_REQUIRE(valid_bits(bits), ADAQP_EINVAL, "adaqp_pack_f16: 8 %% bits != 0 (bits=%d)", bits);
    ADAQP_REQUIRE(N >= 0 && F >= 0, ADAQP_EINVAL, "adaqp_pack_f16: negative shape");
    const int64_t total = adaqp_packed_nbytes(N, F, bits);
    if (total == 0) return 0;
    ADAQP_REQUIRE(data && mn && scale && packed, ADAQP_EINVAL, "adaqp_pack_f16: null pointer");
    const int64_t blocks = (total + kPackThreads - 1) / kPackThreads;
    ADAQP_REQUIRE(blocks < (1ll << 31), ADAQP_ELIMIT, "adaqp_pack_f16: too many bytes");
    cudaStream_t s = (cudaStream_t)stream;
    const __half *x = (const __half *)data, *m = (const __half *)mn, *sc = (const __half *)scale;
    switch (bits) {
        case 1: pack_flat_half_kernel<1><<<(unsigned)blocks, kPackThreads, 0, s>>>(x, m, sc, N, F, total, seed, offset, packed); break;
        case 2: pack_flat_half_kernel<2><<<(unsigned)blocks, kPackThreads, 0, s>>>(x, m, sc, N, F, total, seed, offset, packed); break;
        case 4: pack_flat_half_kernel<4><<<(unsigned)blocks, kPackThreads, 0, s>>>(x, m, sc, N, F, total, seed, offset, packed); break;
        default: pack_flat_half_kernel<8><<<(unsigned)blocks, kPackThreads, 0, s>>>(x, m, sc, N, F, total, seed, offset, packed); break;
    }
    return adaqp_check_launch("pack_flat_half_kernel");
}

int adaqp_unpack_f16(const uint8_t *packed, const void *scale, const void *mn, int64_t N, int64_t F, int bits, void *out,
                     void *stream) {
    ADAQP_REQUIRE(valid_bits(bits), ADAQP_EINVAL, "adaqp_unpack_f16: 8 %% bits != 0 (bits=%d)", bits);
    ADAQP_REQUIRE(N >= 0 && F >= 0, ADAQP_EINVAL, "adaqp_unpack_f16: negative shape");
    const int64_t total = adaqp_packed_nbytes(N, F, bits);
    if (total == 0) return 0;
    ADAQP_REQUIRE(packed && scale && mn && out, ADAQP_EINVAL, "adaqp_unpack_f16: null pointer");
    const int64_t blocks = (total + kPackThreads - 1) / kPackThreads;
    ADAQP_REQUIRE(blocks < (1ll << 31), ADAQP_ELIMIT, "adaqp_unpack_f16: too many bytes");
    cudaStream_t s = (cudaStream_t)stream;
    const __half *sc = (const __half *)scale, *m = (const __half *)mn;
    __half *o = (__half *)out;
    switch (bits) {
        case 1: unpack_flat_half_kernel<1><<<(unsigned)blocks, kPackThreads, 0, s>>>(packed, sc, m, N, F, total, o); break;
        case 2: unpack_flat_half_kernel<2><<<(unsigned)blocks, kPackThreads, 0, s>>>(packed, sc, m, N, F, total, o); break;
        case 4: unpack_flat_half_kernel<4><<<(unsigned)blocks, kPackThreads, 0, s>>>(packed, sc, m, N, F, total, o); break;
        default: unpack_flat_half_kernel<8><<<(unsigned)blocks, kPackThreads, 0, s>>>(packed, sc, m, N, F, total, o); break;
    }
    return adaqp_check_launch("unpack_flat_half_kernel");
}

int adaqp_row_minmax_f32(const float *data, int64_t N, int64_t F, int bits, float *rmin,
                         float *rmax, float *scale, void *stream) {
    ADAQP_REQUIRE(N >= 0 && F > 0, ADAQP_EINVAL, "adaqp_row_minmax_f32: bad shape");
    ADAQP_REQUIRE(!scale || (bits >= 1 && bits <= 16), ADAQP_EINVAL, "adaqp_row_minmax_f32: bad bits");
    if (N == 0) return 0;
    ADAQP_REQUIRE(data != nullptr, ADAQP_EINVAL, "adaqp_row_minmax_f32: null data");
    const int warps_per_block = 8;
    int64_t blocks = (N + warps_per_block - 1) / warps_per_block;
    const int64_t cap = (int64_t)adaqp_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const float levels = (float)((1 << bits) - 1);
    row_minmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(data, N, F, levels, rmin, rmax, scale);
    return adaqp_check_launch("row_minmax_kernel");
}

}  // extern "C"
