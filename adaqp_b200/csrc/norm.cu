// Fused LayerNorm + ReLU for the hidden layers (forward and backward), sm_100a.
//
// The reference's layers apply dropout -> nn.LayerNorm -> ReLU between aggregations (AdaQP/model/distGCN.py:77-85,
// distSAGE.py:89-97).  On full-graph shapes (10^5..10^6 rows x 256) torch runs that as separate passes whose backward
// is dominated by a column-reduction kernel (GammaBetaBackward: 7.6 ms per layer at 2.4 M rows, LayerNorm + ReLU
// together ~31 ms of a 107 ms epoch, profiles/r02c_launches_bench_n1_gemm.md).  Both directions are HBM-bound
// elementwise + row-reduction work, so here each is ONE pass:
//   forward : y = relu((x - mean) * rstd * gamma + beta), mean / rstd saved            (read 4F, write 4F per row)
//   backward: g = dy * [y_pre > 0] * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat));
//             dgamma += dy' * xhat, dbeta += dy' (dy' = dy * [y_pre > 0])                (read 8F, write 4F per row)
// One warp per row, the row in registers (float4 per lane), biased variance and eps as nn.LayerNorm; the column sums
// of dgamma / dbeta are accumulated per warp in registers over its rows, reduced per CTA in shared memory and
// written as per-CTA partials that the host mirror adds (deterministic for a fixed grid).  Dropout stays torch's own
// kernel so that the mask -- and with it parity with the reference flow from the same generator seed -- is unchanged.
#include "common.cuh"

namespace {

constexpr int kNormWarps = 8;
constexpr int kNormThreads = kNormWarps * 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ADAQP_FULL_MASK, v, o);
    return v;
}

template <int CHUNKS>
__global__ void __launch_bounds__(kNormThreads)
ln_relu_fwd_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ gamma, const float *__restrict__ beta,
                   float eps, int64_t M, int F, float *__restrict__ y, int64_t ldy, float *__restrict__ mean_out,
                   float *__restrict__ rstd_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kNormWarps;
    float g[CHUNKS][4], b[CHUNKS][4];
    bool ok[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = (c * 32 + lane) * 4;
        ok[c] = col < F;
        const float4 gv = ok[c] ? __ldg(reinterpret_cast<const float4 *>(gamma + col)) : make_float4(0, 0, 0, 0);
        const float4 bv = ok[c] ? __ldg(reinterpret_cast<const float4 *>(beta + col)) : make_float4(0, 0, 0, 0);
        g[c][0] = gv.x; g[c][1] = gv.y; g[c][2] = gv.z; g[c][3] = gv.w;
        b[c][0] = bv.x; b[c][1] = bv.y; b[c][2] = bv.z; b[c][3] = bv.w;
    }
    const float inv_f = 1.0f / (float)F;
    for (int64_t row = warp; row < M; row += nwarps) {
        float v[CHUNKS][4];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const float4 t = ok[c] ? ldg_stream_f4(x + row * ldx + (c * 32 + lane) * 4) : make_float4(0, 0, 0, 0);
            v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w;
            s += (t.x + t.y) + (t.z + t.w);
        }
        const float mean = warp_sum(s) * inv_f;
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
            if (ok[c]) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float d = v[c][e] - mean; q = fmaf(d, d, q); }
            }
        const float rstd = rsqrtf(warp_sum(q) * inv_f + eps);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
            if (ok[c]) {
                float4 o;
                o.x = fmaxf(fmaf((v[c][0] - mean) * rstd, g[c][0], b[c][0]), 0.f);
                o.y = fmaxf(fmaf((v[c][1] - mean) * rstd, g[c][1], b[c][1]), 0.f);
                o.z = fmaxf(fmaf((v[c][2] - mean) * rstd, g[c][2], b[c][2]), 0.f);
                o.w = fmaxf(fmaf((v[c][3] - mean) * rstd, g[c][3], b[c][3]), 0.f);
                *reinterpret_cast<float4 *>(y + row * ldy + (c * 32 + lane) * 4) = o;
            }
        if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    }
}

template <int CHUNKS>
__global__ void __launch_bounds__(kNormThreads)
ln_relu_bwd_kernel(const float *__restrict__ dy, int64_t lddy, const float *__restrict__ x, int64_t ldx,
                   const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ gamma,
                   const float *__restrict__ beta, int64_t M, int F, float *__restrict__ dx, int64_t lddx,
                   float *__restrict__ partials /* [grid, 2, F] */) {
    __shared__ float red[kNormWarps][2][32 * 4];      // per-warp column partials of one chunk (reused per chunk)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * kNormWarps + wib;
    const int64_t nwarps = (int64_t)gridDim.x * kNormWarps;
    float g[CHUNKS][4], b[CHUNKS][4], acc_g[CHUNKS][4], acc_b[CHUNKS][4];
    bool ok[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = (c * 32 + lane) * 4;
        ok[c] = col < F;
        const float4 gv = ok[c] ? __ldg(reinterpret_cast<const float4 *>(gamma + col)) : make_float4(0, 0, 0, 0);
        const float4 bv = ok[c] ? __ldg(reinterpret_cast<const float4 *>(beta + col)) : make_float4(0, 0, 0, 0);
        g[c][0] = gv.x; g[c][1] = gv.y; g[c][2] = gv.z; g[c][3] = gv.w;
        b[c][0] = bv.x; b[c][1] = bv.y; b[c][2] = bv.z; b[c][3] = bv.w;
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc_g[c][e] = 0.f; acc_b[c][e] = 0.f; }
    }
    const float inv_f = 1.0f / (float)F;
    // software pipeline: the loads of this warp's NEXT row are issued before the current row's reductions, so that two rows
    // per warp are in flight (at 2 resident CTAs x 8 warps per SM one row per warp leaves HBM half idle: 3.5 TB/s under ncu)
    float4 nx[CHUNKS], nd[CHUNKS];
    float nmu = 0.f, nrs = 0.f;
    auto fetch = [&](int64_t r) {
        nmu = __ldg(mean + r);
        nrs = __ldg(rstd + r);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
            if (ok[c]) {
                nx[c] = ldg_stream_f4(x + r * ldx + (c * 32 + lane) * 4);
                nd[c] = ldg_stream_f4(dy + r * lddy + (c * 32 + lane) * 4);
            }
    };
    if (warp < M) fetch(warp);
    for (int64_t row = warp; row < M; row += nwarps) {
        const float mu = nmu, rs = nrs;
        float4 cx[CHUNKS], cd[CHUNKS];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) { cx[c] = nx[c]; cd[c] = nd[c]; }
        if (row + nwarps < M) fetch(row + nwarps);
        float xh[CHUNKS][4], gg[CHUNKS][4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            if (ok[c]) {
                const float4 xv = cx[c];
                const float4 dv = cd[c];
                const float xe[4] = {xv.x, xv.y, xv.z, xv.w}, de[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float h = (xe[e] - mu) * rs;
                    const float pre = fmaf(h, g[c][e], b[c][e]);
                    const float d = pre > 0.f ? de[e] : 0.f;           // ReLU backward (threshold_backward: x > 0)
                    xh[c][e] = h;
                    acc_g[c][e] = fmaf(d, h, acc_g[c][e]);
                    acc_b[c][e] += d;
                    const float gv = d * g[c][e];
                    gg[c][e] = gv;
                    s1 += gv;
                    s2 = fmaf(gv, h, s2);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) { xh[c][e] = 0.f; gg[c][e] = 0.f; }
            }
        }
        const float m1 = warp_sum(s1) * inv_f, m2 = warp_sum(s2) * inv_f;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
            if (ok[c]) {
                float4 o;
                o.x = rs * (gg[c][0] - m1 - xh[c][0] * m2);
                o.y = rs * (gg[c][1] - m1 - xh[c][1] * m2);
                o.z = rs * (gg[c][2] - m1 - xh[c][2] * m2);
                o.w = rs * (gg[c][3] - m1 - xh[c][3] * m2);
                *reinterpret_cast<float4 *>(dx + row * lddx + (c * 32 + lane) * 4) = o;
            }
    }
    // column sums: warps -> CTA (shared memory, fixed order) -> per-CTA partial
    float *pg = partials + (size_t)blockIdx.x * 2 * F, *pb = pg + F;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { red[wib][0][lane * 4 + e] = acc_g[c][e]; red[wib][1][lane * 4 + e] = acc_b[c][e]; }
        __syncthreads();
        if (wib == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float sg = 0.f, sb = 0.f;
#pragma unroll
                for (int w = 0; w < kNormWarps; ++w) { sg += red[w][0][lane * 4 + e]; sb += red[w][1][lane * 4 + e]; }
                const int col = (c * 32 + lane) * 4 + e;
                if (col < F) { pg[col] = sg; pb[col] = sb; }
            }
        }
        __syncthreads();
    }
}

inline int norm_grid(int64_t M) {
    const int sms = adaqp_sm_count() > 0 ? adaqp_sm_count() : 148;
    int64_t ctas = (M + kNormWarps - 1) / kNormWarps;
    const int64_t cap = (int64_t)sms * 4;          // persistent: ~2 resident CTAs per SM, two waves
    if (ctas > cap) ctas = cap;
    return (int)(ctas < 1 ? 1 : ctas);
}

inline bool norm_ok(int64_t M, int F, int64_t ld) { return M >= 0 && F > 0 && F <= 1024 && (F % 4) == 0 && (ld % 4) == 0 && ld >= F; }

}  // namespace

extern "C" {

int adaqp_ln_relu_grid(int64_t M) { return norm_grid(M); }

int adaqp_ln_relu_fwd_f32(const float *x, int64_t ldx, const float *gamma, const float *beta, float eps, int64_t M, int32_t F,
                          float *y, int64_t ldy, float *mean, float *rstd, void *stream) {
    ADAQP_REQUIRE(norm_ok(M, F, ldx) && norm_ok(M, F, ldy), ADAQP_ELIMIT, "adaqp_ln_relu_fwd_f32: F=%d must be a multiple of 4, <= 1024", F);
    if (M == 0) return 0;
    ADAQP_REQUIRE(x && gamma && beta && y && mean && rstd, ADAQP_EINVAL, "adaqp_ln_relu_fwd_f32: null pointer");
    ADAQP_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0, ADAQP_EALIGN, "adaqp_ln_relu_fwd_f32: 16-byte alignment");
    const int chunks = (F + 127) / 128;
    const int grid = norm_grid(M);
    cudaStream_t s = (cudaStream_t)stream;
#define FWD(C) ln_relu_fwd_kernel<C><<<grid, kNormThreads, 0, s>>>(x, ldx, gamma, beta, eps, M, F, y, ldy, mean, rstd)
    if (chunks <= 1) FWD(1); else if (chunks <= 2) FWD(2); else if (chunks <= 4) FWD(4); else FWD(8);
#undef FWD
    return adaqp_check_launch("ln_relu_fwd_kernel");
}

int adaqp_ln_relu_bwd_f32(const float *dy, int64_t lddy, const float *x, int64_t ldx, const float *mean, const float *rstd,
                          const float *gamma, const float *beta, int64_t M, int32_t F, float *dx, int64_t lddx, float *partials,
                          int32_t grid, void *stream) {
    ADAQP_REQUIRE(norm_ok(M, F, ldx) && norm_ok(M, F, lddy) && norm_ok(M, F, lddx), ADAQP_ELIMIT, "adaqp_ln_relu_bwd_f32: F=%d must be a multiple of 4, <= 1024", F);
    ADAQP_REQUIRE(grid == norm_grid(M), ADAQP_EINVAL, "adaqp_ln_relu_bwd_f32: partials must hold adaqp_ln_relu_grid(M) slices");
    ADAQP_REQUIRE(dy && x && mean && rstd && gamma && beta && dx && partials, ADAQP_EINVAL, "adaqp_ln_relu_bwd_f32: null pointer");
    ADAQP_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0, ADAQP_EALIGN, "adaqp_ln_relu_bwd_f32: 16-byte alignment");
    const int chunks = (F + 127) / 128;
    cudaStream_t s = (cudaStream_t)stream;
#define BWD(C) ln_relu_bwd_kernel<C><<<grid, kNormThreads, 0, s>>>(dy, lddy, x, ldx, mean, rstd, gamma, beta, M, F, dx, lddx, partials)
    if (chunks <= 1) BWD(1); else if (chunks <= 2) BWD(2); else if (chunks <= 4) BWD(4); else BWD(8);
#undef BWD
    return adaqp_check_launch("ln_relu_bwd_kernel");
}

}  // extern "C"
