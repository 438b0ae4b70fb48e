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
#include <cuda.h>      // CUtensorMap (type only: the encoder is fetched with cudaGetDriverEntryPoint)

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kUnroll = 4;

// hint bits (option "spmm_hints"): the output rows and the index stream are touched once per launch,
// the gathered source rows are what should stay in L2
constexpr int kHintStoreStreaming = 1;   // st.global.cs for the output rows (evict-first)
constexpr int kHintIndexStreaming = 2;   // ld.global.cs for `indices`

template <int VEC> struct Vec;
template <> struct Vec<4> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ void store_cs(float *p, const float (&v)[4]) {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
};
template <> struct Vec<2> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[2]) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[2]) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    }
    static __device__ __forceinline__ void store_cs(float *p, const float (&v)[2]) {
        __stcs(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
    }
};
template <> struct Vec<1> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void store(float *p, const float (&v)[1]) { *p = v[0]; }
    static __device__ __forceinline__ void store_cs(float *p, const float (&v)[1]) { __stcs(p, v[0]); }
};

// Row counter of the frontier scheduler: {next_row, finished CTAs}.  The last CTA to finish puts both
// words back to zero, so a launch needs no memset node and one counter pair per (device, stream)
// can never be shared by two launches that are in flight together.
__device__ __forceinline__ void frontier_release(unsigned long long *counter) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long prev = atomicAdd(counter + 1, 1ull);
        if (prev == (unsigned long long)gridDim.x - 1ull) {
            counter[0] = 0ull;
            counter[1] = 0ull;
            __threadfence();
        }
    }
}

template <int VEC, int CHUNKS>
__global__ void __launch_bounds__(kThreads)
spmm_csr_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                const float *__restrict__ x0, int64_t ld0, int64_t n_split,
                const float *__restrict__ x1, int64_t ld1,
                const float *__restrict__ pre, const float *__restrict__ post,
                int mean, int add_self, int64_t row_begin, int64_t row_end, int F,
                float *__restrict__ out, int64_t ldo, unsigned long long *__restrict__ next_row, int rows_per_grab,
                const int64_t *__restrict__ seg_start, const int64_t *__restrict__ seg_end, int accumulate, int hints) {
    const int lane = threadIdx.x & 31;
    bool colok[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) colok[c] = ((c * 32 + lane) * VEC) < F;

    // Frontier scheduling: warps take the next `rows_per_grab` destination rows from a global
    // counter, so the rows in flight are always one contiguous window (~ #warps * rows_per_grab
    // rows) however uneven the degrees are.  With a static row -> warp map the warps drift
    // apart (power-law degrees) and the set of source rows being reused grows far beyond
    // L2: ncu showed a 14 % L2 hit rate on a graph where 60 % of the edges stay inside 4 MB
    // communities (profiles/r01_spmm_frontier.md).
    const int64_t n_rows = row_end - row_begin;
    while (true) {
        unsigned long long grab = 0;
        if (lane == 0) grab = atomicAdd(next_row, (unsigned long long)rows_per_grab);
        grab = __shfl_sync(ADAQP_FULL_MASK, grab, 0);
        if ((int64_t)grab >= n_rows) break;
        const int64_t r_hi = ((int64_t)grab + rows_per_grab < n_rows) ? (int64_t)grab + rows_per_grab : n_rows;
    for (int64_t row = row_begin + (int64_t)grab; row < row_begin + r_hi; ++row) {
        float acc[CHUNKS][VEC];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[c][e] = 0.f;
        // neighbour segment of this launch: the whole row, or [seg_start[row], seg_end[row]) when the
        // caller splits a row into its local-source and halo-source parts (ops.py overlap)
        const int64_t row_b = __ldg(indptr + row), row_e = __ldg(indptr + row + 1);
        const int64_t b = seg_start ? __ldg(seg_start + row) : row_b;
        const int64_t e_ = seg_end ? __ldg(seg_end + row) : row_e;
        for (int64_t j0 = b; j0 < e_; j0 += 32) {
            const int n = (e_ - j0) < 32 ? (int)(e_ - j0) : 32;
            int u = 0;
            float w = 0.f;
            if (lane < n) {
                u = (hints & kHintIndexStreaming) ? __ldcs(indices + j0 + lane) : __ldg(indices + j0 + lane);
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
        const float deg = (float)(row_e - row_b);       // mean divides by the full in-degree
        const float ps = post ? __ldg(post + row) : 1.f;
        float *orow = out + (row - row_begin) * ldo;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            if (colok[c]) {
                float prev[VEC];
                if (accumulate) Vec<VEC>::load(orow + (c * 32 + lane) * VEC, prev);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    float r = acc[c][e];
                    if (mean && deg > 0.f) r = __fdiv_rn(r, deg);
                    if (post) r = __fmul_rn(r, ps);
                    if (accumulate) r = __fadd_rn(prev[e], r);
                    acc[c][e] = r;
                }
                if (hints & kHintStoreStreaming) Vec<VEC>::store_cs(orow + (c * 32 + lane) * VEC, acc[c]);
                else Vec<VEC>::store(orow + (c * 32 + lane) * VEC, acc[c]);
            }
        }
    }
    }
    frontier_release(next_row);
}

// ---------------------------------------------------------------------------------------
// v2: asynchronous row gather through a lane-private shared-memory ring (cp.async / LDGSTS).
//
// v1 stages gathered rows in registers: bytes in flight per SM are bounded by the register
// file and the load -> wait -> consume phases do not overlap (ncu, profiles/r01_*: 50 % DRAM
// throughput, 30 % warps active).  Here every lane copies its own 16-byte column chunks of
// each neighbour row with cp.async into a ring of kStages row slots per warp and consumes
// the oldest slot (LDS.128 + FFMA) while the next kStages-1 rows are still in flight:
// a continuous software pipeline with no register cost per in-flight row and no cross-lane
// synchronisation (a lane only ever reads the bytes it copied; completion is tracked by
// per-thread cp.async groups).  Neighbour ids and their norm weights travel in
// lane-distributed register windows that are prefetched one and two windows ahead.  Rows
// are handed to warps as contiguous nnz-balanced chunks, so the pipeline stays full across
// row boundaries and a hub row costs one warp its own length, not a whole CTA's.
//
// Measured alternative (dropped): one `cp.async.bulk` (TMA, UBLKCP) per neighbour row with
// an mbarrier per slot ran 1.6x-2.2x SLOWER than v1 -- the per-SM bulk-copy engine retires
// roughly one 0.4-1 KB request per ~60 cycles, far below what a row gather needs
// (profiles/r01_spmm_variants.md).
// Requires F % 4 == 0 and 16-byte aligned rows (F = 100, 200, 256, 300 ...); anything else
// (F = 602) uses v1.
constexpr int kStages = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
spmm_csr_ring_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                     const float *__restrict__ x0, int64_t ld0, int64_t n_split,
                     const float *__restrict__ x1, int64_t ld1,
                     const float *__restrict__ pre, const float *__restrict__ post,
                     int mean, int add_self, int64_t row_begin, int64_t row_end, int F,
                     float *__restrict__ out, int64_t ldo, int64_t chunk_nnz) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    // ring[warp][stage][chunk][lane] of 16 bytes: a lane's chunk is contiguous with its
    // neighbours' (conflict-free LDS.128 / LDGSTS.128)
    const uint32_t ring_u = smem_u32(smem_raw) + (uint32_t)wib * kStages * CHUNKS * 512u + (uint32_t)lane * 16u;
    const float4 *ring = reinterpret_cast<const float4 *>(smem_raw) + (size_t)wib * kStages * CHUNKS * 32 + lane;

    bool colok[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) colok[c] = ((c * 32 + lane) * 4) < F;
    const int64_t nnz_base = __ldg(indptr + row_begin);
    const int64_t nnz_end_all = __ldg(indptr + row_end);
    int64_t n_chunks = (nnz_end_all - nnz_base + chunk_nnz - 1) / chunk_nnz;
    if (n_chunks < 1) n_chunks = 1;   // rows without edges still produce output

    const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    for (int64_t chunk = warp; chunk < n_chunks; chunk += nwarps) {
        // rows whose first nnz falls into this chunk's nnz window
        const int64_t lo_nnz = nnz_base + chunk * chunk_nnz;
        const bool last = (chunk == n_chunks - 1);
        const int64_t r0 = (chunk == 0) ? row_begin : row_lower_bound(indptr, row_begin, row_end, lo_nnz);
        const int64_t r1 = last ? row_end : row_lower_bound(indptr, row_begin, row_end, lo_nnz + chunk_nnz);
        if (r0 >= r1) continue;
        int64_t pc = __ldg(indptr + r0);                       // consume cursor
        const int64_t pend = last ? nnz_end_all : __ldg(indptr + r1);
        int64_t pi = pc;                                       // issue cursor
        // lane-distributed windows of 32 neighbour ids / weights: [cur | nxt | nx2 (ids only)]
        int64_t wb = pc;
        int idx_cur = (wb + lane < pend) ? __ldg(indices + wb + lane) : 0;
        int idx_nxt = (wb + 32 + lane < pend) ? __ldg(indices + wb + 32 + lane) : 0;
        int idx_nx2 = (wb + 64 + lane < pend) ? __ldg(indices + wb + 64 + lane) : 0;
        float w_cur = (pre && wb + lane < pend) ? __ldg(pre + idx_cur) : 1.f;
        float w_nxt = (pre && wb + 32 + lane < pend) ? __ldg(pre + idx_nxt) : 1.f;
        float w_prev = 1.f;   // the consume side lags the issue side by < kStages <= 32 ids
        int64_t row = r0;
        int64_t rend = __ldg(indptr + row + 1);
        float acc[CHUNKS][4];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; }
        uint32_t si = 0, sc = 0;                               // ring positions

        auto issue_one = [&]() {
            if (pi < pend) {
                if (pi >= wb + 32) {                           // slide the issue windows
                    wb += 32;
                    idx_cur = idx_nxt; idx_nxt = idx_nx2;
                    idx_nx2 = (wb + 64 + lane < pend) ? __ldg(indices + wb + 64 + lane) : 0;
                    w_prev = w_cur;
                    w_cur = w_nxt;
                    w_nxt = (pre && wb + 32 + lane < pend) ? __ldg(pre + idx_nxt) : 1.f;
                }
                const int u = __shfl_sync(ADAQP_FULL_MASK, idx_cur, (int)(pi - wb));
                const float *src = (u < n_split) ? (x0 + (int64_t)u * ld0) : (x1 + ((int64_t)u - n_split) * ld1);
                const uint32_t dst = ring_u + (si % kStages) * (CHUNKS * 512u);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                    if (colok[c]) cp_async_16(dst + c * 512u, src + (c * 32 + lane) * 4);
                ++pi;
                ++si;
            }
            cp_async_commit();                                 // empty groups keep the queue depth constant
        };

#pragma unroll 1
        for (int k = 0; k < kStages - 1; ++k) issue_one();     // prologue: fill the pipeline

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
            issue_one();                                       // one more row in flight ...
            cp_async_wait<kStages - 1>();                      // ... and the oldest has landed (for this lane)
            const float wa = __shfl_sync(ADAQP_FULL_MASK, w_cur, (int)((pc - wb) & 31));
            const float wp = __shfl_sync(ADAQP_FULL_MASK, w_prev, (int)((pc - wb + 32) & 31));
            const float w = (pc >= wb) ? wa : wp;
            const float4 *slot = ring + (size_t)(sc % kStages) * (CHUNKS * 32);
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) if (colok[c]) {
                const float4 t = slot[c * 32];
                acc[c][0] = __fmaf_rn(w, t.x, acc[c][0]); acc[c][1] = __fmaf_rn(w, t.y, acc[c][1]);
                acc[c][2] = __fmaf_rn(w, t.z, acc[c][2]); acc[c][3] = __fmaf_rn(w, t.w, acc[c][3]);
            }
            ++sc;
            ++pc;
        }
        cp_async_wait<0>();
    }
}

// ---------------------------------------------------------------------------------------
// v3 / v4: TMA row gather into a per-warp shared-memory ring (opt-in, option spmm_impl = 3 / 4).
//
// v3 uses Blackwell's row-gather TMA, `cp.async.bulk.tensor.2d ... tile::gather4` (SASS UTMALDG):
// ONE request fetches FOUR neighbour rows (4 x F floats) of a 2-D tensor map over the source
// matrix, selected by four row coordinates, and completes on an mbarrier.  v4 issues one
// `cp.async.bulk` (UBLKCP) per neighbour row into the same ring -- the round-1 design that was
// request-rate bound (profiles/r01_spmm_variants.md), kept so that the comparison stays
// reproducible from the tree.  Scheduling is the frontier scheme of v1 (rows from a global
// counter, so the source rows being reused stay inside L2); the ring pipelines across the rows
// of one grab.  A group never mixes local and halo sources (two tensor maps): the columns of a
// row are sorted, so every 32-id window splits into a local prefix and a halo suffix.
// Needs 16-byte rows and F <= 256 (one TMA box); anything else runs v1.
constexpr int kRingWarps = 16;
constexpr int kRingThreads = kRingWarps * 32;

struct __align__(16) RingMeta {
    float w[4];     // pre-norm weights of the slot's rows
    int cnt;        // valid rows in the slot (0..4)
    int last;       // 1: the destination row is complete after this slot
    int pad[2];
};

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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

template <int CHUNKS, bool GATHER4>
__global__ void __launch_bounds__(kRingThreads, 1)
spmm_csr_tma_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                    const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                    const float *__restrict__ x0, int64_t ld0, int64_t n_split,
                    const float *__restrict__ x1, int64_t ld1,
                    const float *__restrict__ pre, const float *__restrict__ post,
                    int mean, int add_self, int64_t row_begin, int64_t row_end, int F,
                    float *__restrict__ out, int64_t ldo, unsigned long long *__restrict__ next_row, int rows_per_grab,
                    const int64_t *__restrict__ seg_start, const int64_t *__restrict__ seg_end, int accumulate,
                    int stages) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t rowbytes = (uint32_t)F * 4u;
    const uint32_t slot_bytes = (4u * rowbytes + 127u) & ~127u;
    uint8_t *ring = smem_raw + (size_t)wib * stages * slot_bytes;
    uint8_t *aux = smem_raw + (size_t)kRingWarps * stages * slot_bytes;
    RingMeta *metas = reinterpret_cast<RingMeta *>(aux) + wib * stages;
    const uint32_t bars_u = smem_u32(aux + (size_t)kRingWarps * stages * sizeof(RingMeta)) + (uint32_t)(wib * stages) * 8u;
    const uint32_t ring_u = smem_u32(ring);
    if (lane == 0)
        for (int st = 0; st < stages; ++st) mbar_init(bars_u + 8u * st, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    bool colok[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) colok[c] = ((c * 32 + lane) * 4) < F;
    const int64_t n_rows = row_end - row_begin;
    int si = 0, sc = 0;                 // issue / consume slot
    uint32_t pc = 0;                    // parity of the consume side's current lap
    int inflight = 0;

    while (true) {
        unsigned long long grab = 0;
        if (lane == 0) grab = atomicAdd(next_row, (unsigned long long)rows_per_grab);
        grab = __shfl_sync(ADAQP_FULL_MASK, grab, 0);
        if ((int64_t)grab >= n_rows) break;
        const int64_t r_hi = ((int64_t)grab + rows_per_grab < n_rows) ? (int64_t)grab + rows_per_grab : n_rows;
        int64_t rc = row_begin + (int64_t)grab;          // row the consume side is accumulating
        float acc[CHUNKS][4];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;

        auto consume_one = [&]() {
            mbar_wait(bars_u + 8u * sc, pc);
            const RingMeta m = metas[sc];
            const float4 *slot = reinterpret_cast<const float4 *>(ring + (size_t)sc * slot_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < m.cnt) {
                    const float w = m.w[k];
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) {
                        if (colok[c]) {
                            const float4 t = slot[k * (F >> 2) + c * 32 + lane];
                            acc[c][0] = __fmaf_rn(w, t.x, acc[c][0]); acc[c][1] = __fmaf_rn(w, t.y, acc[c][1]);
                            acc[c][2] = __fmaf_rn(w, t.z, acc[c][2]); acc[c][3] = __fmaf_rn(w, t.w, acc[c][3]);
                        }
                    }
                }
            }
            if (++sc == stages) { sc = 0; pc ^= 1u; }
            --inflight;
            if (m.last) {               // destination row rc is complete
                const int64_t row = rc;
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
                const float deg = (float)(__ldg(indptr + row + 1) - __ldg(indptr + row));
                const float ps = post ? __ldg(post + row) : 1.f;
                float *orow = out + (row - row_begin) * ldo;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) if (colok[c]) {
                    float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (accumulate) prev = *reinterpret_cast<const float4 *>(orow + (c * 32 + lane) * 4);
                    const float pv[4] = {prev.x, prev.y, prev.z, prev.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float r = acc[c][e];
                        if (mean && deg > 0.f) r = __fdiv_rn(r, deg);
                        if (post) r = __fmul_rn(r, ps);
                        if (accumulate) r = __fadd_rn(pv[e], r);
                        acc[c][e] = r;
                    }
                    *reinterpret_cast<float4 *>(orow + (c * 32 + lane) * 4) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                    acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
                }
                ++rc;
            }
            __syncwarp();               // every lane is done with the slot before it is refilled
        };

        for (int64_t row = row_begin + (int64_t)grab; row < row_begin + r_hi; ++row) {
            const int64_t b = seg_start ? __ldg(seg_start + row) : __ldg(indptr + row);
            const int64_t e_ = seg_end ? __ldg(seg_end + row) : __ldg(indptr + row + 1);
            if (b >= e_) {              // no neighbours in this launch's segment: an empty, final slot
                while (inflight >= stages) consume_one();
                if (lane == 0) {
                    metas[si].cnt = 0;
                    metas[si].last = 1;
                    mbar_arrive(bars_u + 8u * si);
                }
                __syncwarp();
                if (++si == stages) si = 0;
                ++inflight;
                continue;
            }
            for (int64_t j0 = b; j0 < e_; j0 += 32) {
                const int n = (e_ - j0) < 32 ? (int)(e_ - j0) : 32;
                int u = 0;
                float w = 0.f;
                if (lane < n) {
                    u = __ldg(indices + j0 + lane);
                    w = pre ? __ldg(pre + u) : 1.f;
                }
                const int nl = __popc(__ballot_sync(ADAQP_FULL_MASK, lane < n && u < n_split));   // sorted: local prefix
                const int gl = (nl + 3) >> 2;
                const int ng = gl + ((n - nl + 3) >> 2);
                for (int g = 0; g < ng; ++g) {
                    while (inflight >= stages) consume_one();
                    const bool halo = g >= gl;
                    const int base = halo ? nl + 4 * (g - gl) : 4 * g;
                    const int lim = halo ? n : nl;
                    const int cnt = (lim - base) < 4 ? (lim - base) : 4;
                    const int srcl = (base + (lane & 3)) & 31;
                    int uv = __shfl_sync(ADAQP_FULL_MASK, u, srcl);
                    const float wv = __shfl_sync(ADAQP_FULL_MASK, w, srcl);
                    const int ufirst = __shfl_sync(ADAQP_FULL_MASK, u, base & 31);
                    if ((lane & 3) >= cnt) uv = ufirst;          // pad with the group's first row (its FMA is skipped)
                    const int r1 = __shfl_sync(ADAQP_FULL_MASK, uv, 1), r2 = __shfl_sync(ADAQP_FULL_MASK, uv, 2),
                              r3 = __shfl_sync(ADAQP_FULL_MASK, uv, 3);
                    if (lane < 4) metas[si].w[lane] = wv;
                    if (lane == 0) {
                        metas[si].cnt = cnt;
                        metas[si].last = (j0 + 32 >= e_ && g == ng - 1) ? 1 : 0;
                        const uint32_t bar = bars_u + 8u * si;
                        const uint32_t dst = ring_u + (uint32_t)si * slot_bytes;
                        const int off = halo ? (int)n_split : 0;
                        if (GATHER4) {
                            mbar_arrive_expect_tx(bar, 4u * rowbytes);
                            tma_gather4(dst, halo ? &map1 : &map0, 0, uv - off, r1 - off, r2 - off, r3 - off, bar);
                        } else {
                            mbar_arrive_expect_tx(bar, (uint32_t)cnt * rowbytes);
                            const float *sb = halo ? x1 : x0;
                            const int64_t ld = halo ? ld1 : ld0;
                            const int rr[4] = {uv - off, r1 - off, r2 - off, r3 - off};
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < cnt) bulk_g2s(dst + (uint32_t)k * rowbytes, sb + (int64_t)rr[k] * ld, rowbytes, bar);
                        }
                    }
                    __syncwarp();
                    if (++si == stages) si = 0;
                    ++inflight;
                }
            }
        }
        while (inflight > 0) consume_one();
    }
    frontier_release(next_row);
}

inline bool aligned(const void *p, int vec) { return ((uintptr_t)p & ((uintptr_t)vec * 4 - 1)) == 0; }

}  // namespace

namespace {

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is not linked: the encoder comes from the driver the process already runs on
TensorMapEncodeTiledFn tensor_map_encoder() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return q == cudaDriverEntryPointSuccess ? (TensorMapEncodeTiledFn)fn : nullptr;
}

// 2-D fp32 map {F, rows} with box {F, 1}: gather4 fetches the box at four row coordinates
int make_row_map(CUtensorMap *m, const float *base, int64_t rows, int F, int64_t ld) {
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    ADAQP_REQUIRE(enc != nullptr, ADAQP_EINVAL, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t dims[2] = {(cuuint64_t)F, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)F, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ADAQP_REQUIRE(r == CUDA_SUCCESS, ADAQP_EINVAL, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// One {next_row, finished} pair per (device, stream): launches on one stream are serialised and the
// kernel zeroes the pair when it finishes, so the pair is never shared by two live launches.
unsigned long long *frontier_counter(int dev, cudaStream_t s) {
    static std::mutex mu;
    static std::unordered_map<uint64_t, unsigned long long *> pairs;
    std::lock_guard<std::mutex> lock(mu);
    const uint64_t key = ((uint64_t)(uintptr_t)s << 6) ^ (uint64_t)dev;
    auto it = pairs.find(key);
    if (it != pairs.end()) return it->second;
    unsigned long long *p = nullptr;
    if (cudaMalloc(&p, 2 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    if (cudaMemset(p, 0, 2 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    pairs.emplace(key, p);
    return p;
}

}  // namespace

extern "C" {

int adaqp_spmm_csr_seg_f32(const int64_t *indptr, const int64_t *seg_start, const int64_t *seg_end,
                           const int32_t *indices, const float *x0, int64_t ld0,
                           int64_t n_split, const float *x1, int64_t ld1, const float *pre,
                           const float *post, int mean, int add_self, int accumulate, int64_t row_begin,
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
    const AdaqpOptions &opt = adaqp_options();
    cudaStream_t s = (cudaStream_t)stream;
    const int impl = opt.spmm_impl;
    // v2 (cp.async ring) needs 16-byte rows: F % 4 == 0, strides % 4 == 0, 16-byte aligned bases
    if (impl == 2 && vec == 4 && nchunks <= 8 && !seg_start && !seg_end && !accumulate) {
        auto launch = [&](auto kernel, int C) {
            const size_t smem = (size_t)kWarps * kStages * C * 512;
            int ctas_per_sm = (int)((200 * 1024) / (smem + 1024));
            if (ctas_per_sm > 4) ctas_per_sm = 4;
            if (ctas_per_sm < 1) ctas_per_sm = 1;
            int64_t g2 = (int64_t)sms * ctas_per_sm;
            const int64_t max_ctas = (rows + kWarps - 1) / kWarps;
            if (g2 > max_ctas) g2 = max_ctas;
            const int64_t chunk_nnz = 2048;   // nnz-balanced work units, many per warp
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            kernel<<<(unsigned)g2, kThreads, smem, s>>>(indptr, indices, x0, ld0, n_split, x1, ld1, pre, post, mean,
                                                        add_self, row_begin, row_end, F, out, ldo, chunk_nnz);
        };
        if (nchunks <= 1) launch(spmm_csr_ring_kernel<1>, 1);
        else if (nchunks <= 2) launch(spmm_csr_ring_kernel<2>, 2);
        else if (nchunks <= 3) launch(spmm_csr_ring_kernel<3>, 3);
        else if (nchunks <= 4) launch(spmm_csr_ring_kernel<4>, 4);
        else if (nchunks <= 6) launch(spmm_csr_ring_kernel<6>, 6);
        else launch(spmm_csr_ring_kernel<8>, 8);
        return adaqp_check_launch("spmm_csr_ring_kernel");
    }
    int dev = 0;
    ADAQP_CUDA(cudaGetDevice(&dev));
    unsigned long long *counter = frontier_counter(dev, s);
    ADAQP_REQUIRE(counter != nullptr, ADAQP_EINVAL, "adaqp_spmm_csr_seg_f32: row counter allocation failed");
    // v3 / v4 (TMA ring): one TMA box per row -> F <= 256, 16-byte rows
    if ((impl == 3 || impl == 4) && vec == 4 && F <= 256 && nchunks <= 2) {
        CUtensorMap map0, map1;
        memset(&map0, 0, sizeof(map0));
        memset(&map1, 0, sizeof(map1));
        if (impl == 3) {
            // x0 holds the ids below n_split; the halo matrix's row count is not part of the ABI, the
            // map only bounds-checks coordinates, so give it the largest extent a row coordinate can have
            const int64_t rows0 = n_split > 0 ? n_split : 1;
            int rc = make_row_map(&map0, x0, rows0, F, ld0);
            if (rc) return rc;
            if (x1) { rc = make_row_map(&map1, x1, (int64_t)1 << 31, F, ld1); if (rc) return rc; }
        }
        const uint32_t slot_bytes = (16u * (uint32_t)F + 127u) & ~127u;
        int stages = (int)((200u * 1024u) / ((size_t)kRingWarps * (slot_bytes + sizeof(RingMeta) + 8)));
        if (stages > 8) stages = 8;
        ADAQP_REQUIRE(stages >= 2, ADAQP_ELIMIT, "adaqp_spmm_csr_seg_f32: ring does not fit shared memory");
        const size_t smem = (size_t)kRingWarps * stages * (slot_bytes + sizeof(RingMeta) + 8);
        int64_t grid = (rows + kRingWarps - 1) / kRingWarps;
        if (grid > sms) grid = sms;
        const int grab = opt.spmm_rows_per_grab > 0 ? opt.spmm_rows_per_grab : 4;
        auto launch = [&](auto kernel) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kernel<<<(unsigned)grid, kRingThreads, smem, s>>>(map0, map1, indptr, indices, x0, ld0, n_split, x1, ld1, pre, post,
                                                             mean, add_self, row_begin, row_end, F, out, ldo, counter, grab,
                                                             seg_start, seg_end, accumulate, stages);
        };
        if (impl == 3) { if (nchunks <= 1) launch(spmm_csr_tma_kernel<1, true>); else launch(spmm_csr_tma_kernel<2, true>); }
        else { if (nchunks <= 1) launch(spmm_csr_tma_kernel<1, false>); else launch(spmm_csr_tma_kernel<2, false>); }
        return adaqp_check_launch("spmm_csr_tma_kernel");
    }
    int64_t grid = (rows + kWarps - 1) / kWarps;
    // measured on B200 (profiles/r01_spmm_frontier.md): 1 row per grab for wide rows, 2 for F <= 128
    const int grab_now = opt.spmm_rows_per_grab > 0 ? opt.spmm_rows_per_grab : (F > 128 ? 1 : 2);
    { const int64_t cap2 = (int64_t)sms * (opt.spmm_ctas_per_sm > 0 ? opt.spmm_ctas_per_sm : 8); if (grid > cap2) grid = cap2; }
    const int hints = opt.spmm_hints;
#define CALL_SPMM(V, C)                                                                           \
    spmm_csr_kernel<V, C><<<(unsigned)grid, kThreads, 0, s>>>(indptr, indices, x0, ld0, n_split, x1, \
                                                             ld1, pre, post, mean, add_self,      \
                                                             row_begin, row_end, F, out, ldo, counter, grab_now, seg_start, seg_end, accumulate, hints)
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


int adaqp_spmm_csr_f32(const int64_t *indptr, const int32_t *indices, const float *x0, int64_t ld0,
                       int64_t n_split, const float *x1, int64_t ld1, const float *pre,
                       const float *post, int mean, int add_self, int64_t row_begin,
                       int64_t row_end, int32_t F, float *out, int64_t ldo, void *stream) {
    return adaqp_spmm_csr_seg_f32(indptr, nullptr, nullptr, indices, x0, ld0, n_split, x1, ld1, pre, post, mean,
                                  add_self, 0, row_begin, row_end, F, out, ldo, stream);
}

}  // extern "C"
