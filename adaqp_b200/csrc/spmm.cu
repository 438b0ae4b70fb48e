// Normalised CSR SpMM for full-graph GNN aggregation (sm_100a).
//
// Replaces DGL update_all(copy_src, sum|mean) and the surrounding elementwise norm
// kernels / torch.cat copies of AdaQP/model/ops.py:17-67,137-147,169-185:
//   out[v] = post[v] * ( sum_{u in N_in(v)} pre[u] * x[u]  (+ pre[v] x[v]) )  [/ deg(v)]
// Source rows come from two matrices without concatenation: ids < n_split are local
// (inner) rows, ids >= n_split are halo rows written by the exchange kernels.  The
// central / marginal decomposition of the reference (conversion.py:114-172) is a row
// range [row_begin, row_end) of the same CSR: central rows have no halo in-neighbours
// by construction, so no copy buffers are needed.
//
// HBM/L2-bound gather: one warp per destination row, the row's F floats spread across
// lanes as VEC-wide vectors (CHUNKS per lane), neighbour ids fetched 32 at a time and
// broadcast by shuffle, 4 neighbour rows (4*CHUNKS vector loads per lane) in flight.
// fp32 accumulation in CSR order (DGL's order is unspecified: parity is to a stated
// tolerance against a float64 oracle).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kUnroll = 4;

template <int VEC> struct Vec;
template <> struct Vec<4> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec<2> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[2]) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[2]) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    }
};
template <> struct Vec<1> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void store(float *p, const float (&v)[1]) { *p = v[0]; }
};

template <int VEC, int CHUNKS>
__global__ void __launch_bounds__(kThreads)
spmm_csr_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                const float *__restrict__ x0, int64_t ld0, int64_t n_split,
                const float *__restrict__ x1, int64_t ld1,
                const float *__restrict__ pre, const float *__restrict__ post,
                int mean, int add_self, int64_t row_begin, int64_t row_end, int F,
                float *__restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    bool colok[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) colok[c] = ((c * 32 + lane) * VEC) < F;

    for (int64_t row = row_begin + warp; row < row_end; row += nwarps) {
        float acc[CHUNKS][VEC];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[c][e] = 0.f;
        const int64_t b = __ldg(indptr + row), e_ = __ldg(indptr + row + 1);
        for (int64_t j0 = b; j0 < e_; j0 += 32) {
            const int n = (e_ - j0) < 32 ? (int)(e_ - j0) : 32;
            int u = 0;
            float w = 0.f;
            if (lane < n) {
                u = __ldg(indices + j0 + lane);
                w = pre ? __ldg(pre + u) : 1.f;
            }
            for (int k = 0; k < n; k += kUnroll) {
                float v[kUnroll][CHUNKS][VEC];
                float ww[kUnroll];
#pragma unroll
                for (int t = 0; t < kUnroll; ++t) {
                    const int src = (k + t) & 31;
                    const int uu = __shfl_sync(ADAQP_FULL_MASK, u, src);
                    ww[t] = __shfl_sync(ADAQP_FULL_MASK, w, src);
                    const bool live = (k + t) < n;
                    if (!live) ww[t] = 0.f;
                    const float *rp = (uu < n_split) ? (x0 + (int64_t)uu * ld0) : (x1 + ((int64_t)uu - n_split) * ld1);
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) {
                        if (live && colok[c]) {
                            Vec<VEC>::load(rp + (c * 32 + lane) * VEC, v[t][c]);
                        } else {
#pragma unroll
                            for (int e = 0; e < VEC; ++e) v[t][c][e] = 0.f;
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < kUnroll; ++t)
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[c][e] = __fmaf_rn(ww[t], v[t][c][e], acc[c][e]);
            }
        }
        if (add_self) {
            const float ws = pre ? __ldg(pre + row) : 1.f;
            const float *rp = (row < n_split) ? (x0 + row * ld0) : (x1 + (row - n_split) * ld1);
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                if (colok[c]) {
                    float v[VEC];
                    Vec<VEC>::load(rp + (c * 32 + lane) * VEC, v);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[c][e] = __fmaf_rn(ws, v[e], acc[c][e]);
                }
            }
        }
        const float deg = (float)(e_ - b);
        const float ps = post ? __ldg(post + row) : 1.f;
        float *orow = out + (row - row_begin) * ldo;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            if (colok[c]) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    float r = acc[c][e];
                    if (mean && deg > 0.f) r = __fdiv_rn(r, deg);
                    if (post) r = __fmul_rn(r, ps);
                    acc[c][e] = r;
                }
                Vec<VEC>::store(orow + (c * 32 + lane) * VEC, acc[c]);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------
// v2: TMA bulk row gather through a per-warp shared-memory ring.
//
// v1 keeps gathered rows in registers, so bytes in flight per SM are bounded by the
// register file (ncu: 50 % DRAM throughput, 30 % warps active, profiles/r01_*).  Here one
// elected lane issues a `cp.async.bulk` (UBLKCP) per neighbour row into a ring of kStages
// row slots per warp; completion is tracked by one mbarrier per slot, the neighbour's norm
// weight rides along as a 4-byte cp.async, and all lanes consume the slot with LDS.128 +
// FFMA.  Up to kStages KB-sized rows per warp are in flight with no register cost.  Rows
// are handed to warps as contiguous nnz-balanced chunks so the ring stays full across row
// boundaries and hub rows do not serialise a whole CTA.  Each bulk copy carries an L2
// eviction policy: sources close to the destination id (community / partition-order
// locality) are kept (evict_last), far ones stream through (evict_first).
// Requires F % 4 == 0 and 16-byte aligned rows (F = 100, 200, 256, 300 ...); anything else
// (F = 602) uses v1.
constexpr int kStages = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t make_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t make_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// first row r in [lo, hi] with indptr[r] >= target
__device__ __forceinline__ int64_t row_lower_bound(const int64_t *__restrict__ indptr, int64_t lo, int64_t hi, int64_t target) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(indptr + mid) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int CHUNKS>
__global__ void __launch_bounds__(kThreads)
spmm_csr_tma_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                    const float *__restrict__ x0, int64_t ld0, int64_t n_split,
                    const float *__restrict__ x1, int64_t ld1,
                    const float *__restrict__ pre, const float *__restrict__ post,
                    int mean, int add_self, int64_t row_begin, int64_t row_end, int F,
                    float *__restrict__ out, int64_t ldo, int64_t chunk_nnz,
                    int64_t near_window, int use_hints) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int row_bytes = F * 4;
    const int slot_bytes = (row_bytes + 127) & ~127;
    uint8_t *ring = smem_raw + (size_t)wib * kStages * slot_bytes;
    float *wslot = reinterpret_cast<float *>(smem_raw + (size_t)kWarps * kStages * slot_bytes) + wib * kStages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kWarps * kStages * slot_bytes + kWarps * kStages * 4) + wib * kStages;
    const uint32_t ring_u = smem_u32(ring), wslot_u = smem_u32(wslot), bars_u = smem_u32(bars);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bars_u + 8 * s, 2);   // expect_tx arrive + weight arrive
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const uint64_t pol_near = make_policy_evict_last();
    const uint64_t pol_far = make_policy_evict_first();
    const uint64_t pol_norm = make_policy_evict_normal();

    bool colok[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) colok[c] = ((c * 32 + lane) * 4) < F;
    const int64_t nnz_base = __ldg(indptr + row_begin);
    const int64_t nnz_end_all = __ldg(indptr + row_end);
    int64_t n_chunks = (nnz_end_all - nnz_base + chunk_nnz - 1) / chunk_nnz;
    if (n_chunks < 1) n_chunks = 1;   // rows without edges still produce output
    uint32_t issued = 0, consumed = 0;   // ring counters (monotone across chunks)

    const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    for (int64_t chunk = warp; chunk < n_chunks; chunk += nwarps) {
        // rows whose first nnz falls into this chunk's nnz window
        const int64_t lo_nnz = nnz_base + chunk * chunk_nnz;
        int64_t hi_nnz = lo_nnz + chunk_nnz;
        const bool last = (chunk == n_chunks - 1);
        const int64_t r0 = (chunk == 0) ? row_begin : row_lower_bound(indptr, row_begin, row_end, lo_nnz);
        const int64_t r1 = last ? row_end : row_lower_bound(indptr, row_begin, row_end, hi_nnz);
        if (r0 >= r1) continue;
        int64_t pc = __ldg(indptr + r0);
        const int64_t pend = last ? nnz_end_all : __ldg(indptr + r1);
        int64_t pi = pc;
        // index window of the issue side (32 ids per lane-distributed window) + prefetched next window
        int64_t wb = pi;
        int idx_cur = (wb + lane < pend) ? __ldg(indices + wb + lane) : 0;
        int idx_next = (wb + 32 + lane < pend) ? __ldg(indices + wb + 32 + lane) : 0;
        int64_t row = r0;
        int64_t rend = __ldg(indptr + row + 1);
        float acc[CHUNKS][4];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; }

        while (true) {
            // ---- finish every row that is complete (also rows without in-edges)
            while (row < r1 && pc == rend) {
                const int64_t rb = __ldg(indptr + row);
                if (add_self) {
                    const float ws = pre ? __ldg(pre + row) : 1.f;
                    const float *rp = (row < n_split) ? (x0 + row * ld0) : (x1 + (row - n_split) * ld1);
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) if (colok[c]) {
                        const float4 t = __ldg(reinterpret_cast<const float4 *>(rp + (c * 32 + lane) * 4));
                        acc[c][0] = __fmaf_rn(ws, t.x, acc[c][0]); acc[c][1] = __fmaf_rn(ws, t.y, acc[c][1]);
                        acc[c][2] = __fmaf_rn(ws, t.z, acc[c][2]); acc[c][3] = __fmaf_rn(ws, t.w, acc[c][3]);
                    }
                }
                const float deg = (float)(rend - rb);
                const float ps = post ? __ldg(post + row) : 1.f;
                float *orow = out + (row - row_begin) * ldo;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) if (colok[c]) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float r = acc[c][e];
                        if (mean && deg > 0.f) r = __fdiv_rn(r, deg);
                        if (post) r = __fmul_rn(r, ps);
                        acc[c][e] = r;
                    }
                    *reinterpret_cast<float4 *>(orow + (c * 32 + lane) * 4) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                    acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
                }
                ++row;
                if (row < r1) rend = __ldg(indptr + row + 1);
            }
            if (pc >= pend) break;
            // ---- top up the ring: keep kStages row gathers in flight
            while (pi < pend && (uint32_t)(issued - consumed) < (uint32_t)kStages) {
                if (pi >= wb + 32) {            // slide the window; the next one was prefetched 32 ids ago
                    wb += 32;
                    idx_cur = idx_next;
                    idx_next = (wb + 32 + lane < pend) ? __ldg(indices + wb + 32 + lane) : 0;
                }
                const int u = __shfl_sync(ADAQP_FULL_MASK, idx_cur, (int)(pi - wb));
                if (lane == 0) {
                    const uint32_t s = issued % kStages;
                    const uint32_t bar = bars_u + 8 * s;
                    const float *src = (u < n_split) ? (x0 + (int64_t)u * ld0) : (x1 + ((int64_t)u - n_split) * ld1);
                    uint64_t pol = pol_norm;
                    if (use_hints) {
                        const int64_t dist = (int64_t)u - row;
                        pol = (u >= n_split) ? pol_norm : ((dist < near_window && dist > -near_window) ? pol_near : pol_far);
                    }
                    if (pre) {
                        cp_async_4(wslot_u + 4 * s, pre + u);
                        cp_async_mbar_arrive_noinc(bar);
                    } else {
                        mbar_arrive(bar);
                    }
                    mbar_arrive_expect_tx(bar, (uint32_t)row_bytes);
                    bulk_g2s(ring_u + s * slot_bytes, src, (uint32_t)row_bytes, bar, pol);
                }
                ++issued;
                ++pi;
            }
            // ---- consume the oldest slot
            {
                const uint32_t s = consumed % kStages;
                mbar_wait(bars_u + 8 * s, (consumed / kStages) & 1u);
                const float w = pre ? wslot[s] : 1.f;
                const uint8_t *slot = ring + s * slot_bytes;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) if (colok[c]) {
                    const float4 t = *reinterpret_cast<const float4 *>(slot + (c * 32 + lane) * 16);
                    acc[c][0] = __fmaf_rn(w, t.x, acc[c][0]); acc[c][1] = __fmaf_rn(w, t.y, acc[c][1]);
                    acc[c][2] = __fmaf_rn(w, t.z, acc[c][2]); acc[c][3] = __fmaf_rn(w, t.w, acc[c][3]);
                }
                ++consumed;
                ++pc;
                __syncwarp();   // every lane is done with the slot before lane 0 refills it
            }
        }
    }
}

inline bool aligned(const void *p, int vec) { return ((uintptr_t)p & ((uintptr_t)vec * 4 - 1)) == 0; }

}  // namespace

extern "C" {

int adaqp_spmm_csr_f32(const int64_t *indptr, const int32_t *indices, const float *x0, int64_t ld0,
                       int64_t n_split, const float *x1, int64_t ld1, const float *pre,
                       const float *post, int mean, int add_self, int64_t row_begin,
                       int64_t row_end, int32_t F, float *out, int64_t ldo, void *stream) {
    ADAQP_REQUIRE(F > 0 && F <= 1024, ADAQP_ELIMIT, "adaqp_spmm_csr_f32: F=%d outside (0,1024]", F);
    ADAQP_REQUIRE(row_end >= row_begin && row_begin >= 0, ADAQP_EINVAL, "adaqp_spmm_csr_f32: bad row range");
    if (row_end == row_begin) return 0;
    ADAQP_REQUIRE(indptr && indices && x0 && out, ADAQP_EINVAL, "adaqp_spmm_csr_f32: null pointer");
    int vec = 4;
    auto fits = [&](int v) {
        if (F % v || ld0 % v || ldo % v) return false;
        if (!aligned(x0, v) || !aligned(out, v)) return false;
        if (x1 && (!aligned(x1, v) || (ld1 % v))) return false;
        return true;
    };
    while (vec > 1 && !fits(vec)) vec >>= 1;
    const int nchunks = (F + 32 * vec - 1) / (32 * vec);
    const int64_t rows = row_end - row_begin;
    const int sms = adaqp_sm_count() > 0 ? adaqp_sm_count() : 148;
    int64_t grid = (rows + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)sms * 8;
    if (grid > cap) grid = cap;
    cudaStream_t s = (cudaStream_t)stream;
    // v2 (TMA ring) needs 16-byte rows: F % 4 == 0, strides % 4 == 0, 16-byte aligned bases
    static int impl = -1, hints = -1;
    if (impl < 0) { const char *e = getenv("ADAQP_SPMM"); impl = (e && e[0] == '1') ? 1 : ((e && e[0] == '2') ? 2 : 0); }
    if (hints < 0) { const char *e = getenv("ADAQP_SPMM_HINTS"); hints = (e && e[0] == '0') ? 0 : 1; }
    if (impl != 1 && vec == 4 && nchunks <= 8) {
        const int slot_bytes = (F * 4 + 127) & ~127;
        const size_t smem = (size_t)kWarps * kStages * slot_bytes + kWarps * kStages * 4 + kWarps * kStages * 8;
        int ctas_per_sm = (int)((200 * 1024) / (smem + 1024));
        if (ctas_per_sm > 4) ctas_per_sm = 4;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        int64_t g2 = (int64_t)sms * ctas_per_sm;
        const int64_t max_ctas = (rows + kWarps - 1) / kWarps;
        if (g2 > max_ctas) g2 = max_ctas;
        const int64_t chunk_nnz = 2048;       // nnz-balanced work units, many per warp
        const int64_t near_window = 16384;    // |src - dst| below this: keep the row in L2
#define CALL_V2(C)                                                                                   \
    do {                                                                                             \
        static bool attr_set_##C = false;                                                            \
        if (!attr_set_##C) {                                                                         \
            cudaFuncSetAttribute(spmm_csr_tma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
            attr_set_##C = true;                                                                     \
        }                                                                                            \
        spmm_csr_tma_kernel<C><<<(unsigned)g2, kThreads, smem, s>>>(indptr, indices, x0, ld0, n_split, x1, ld1, \
                                                                 pre, post, mean, add_self, row_begin, row_end, \
                                                                 F, out, ldo, chunk_nnz, near_window, hints);   \
    } while (0)
        if (nchunks <= 1) CALL_V2(1);
        else if (nchunks <= 2) CALL_V2(2);
        else if (nchunks <= 3) CALL_V2(3);
        else if (nchunks <= 4) CALL_V2(4);
        else if (nchunks <= 6) CALL_V2(6);
        else CALL_V2(8);
#undef CALL_V2
        return adaqp_check_launch("spmm_csr_tma_kernel");
    }
#define CALL_SPMM(V, C)                                                                           \
    spmm_csr_kernel<V, C><<<(unsigned)grid, kThreads, 0, s>>>(indptr, indices, x0, ld0, n_split, x1, \
                                                             ld1, pre, post, mean, add_self,      \
                                                             row_begin, row_end, F, out, ldo)
    if (vec == 4) {
        if (nchunks <= 1) CALL_SPMM(4, 1);
        else if (nchunks <= 2) CALL_SPMM(4, 2);
        else if (nchunks <= 3) CALL_SPMM(4, 3);
        else if (nchunks <= 4) CALL_SPMM(4, 4);
        else if (nchunks <= 6) CALL_SPMM(4, 6);
        else CALL_SPMM(4, 8);
    } else if (vec == 2) {
        if (nchunks <= 2) CALL_SPMM(2, 2);
        else if (nchunks <= 4) CALL_SPMM(2, 4);
        else if (nchunks <= 6) CALL_SPMM(2, 6);
        else if (nchunks <= 10) CALL_SPMM(2, 10);
        else CALL_SPMM(2, 16);
    } else {
        if (nchunks <= 4) CALL_SPMM(1, 4);
        else if (nchunks <= 8) CALL_SPMM(1, 8);
        else if (nchunks <= 16) CALL_SPMM(1, 16);
        else CALL_SPMM(1, 32);
    }
#undef CALL_SPMM
    return adaqp_check_launch("spmm_csr_kernel");
}

}  // extern "C"
