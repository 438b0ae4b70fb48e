// Dense feature x weight GEMM on the 5th-generation tensor cores (sm_100a): C = A . Bt^T (+ bias), fp32 in / out.
//
// Replaces the `torch.matmul(rst, self.weight)` of AdaQP/model/distGCN.py:45 and the two Linear layers of
// distSAGE.py:51-53 (the one true contraction on the path) for the tall-skinny shapes of full-graph
// training: A [M, K] with M = number of inner nodes (10^5 .. 10^6) and K, N <= 256.
//
// fp32 results from tf32 tensor cores by error-compensated splitting ("3xTF32"): every operand is written
// a = a_hi + a_lo with a_hi = a & 0xFFFFE000 (exactly representable in tf32, so the hardware conversion is the
// identity) and a_lo = a - a_hi (exact in fp32), and the product is accumulated in fp32 TMEM as
//     a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo          (dropped: a_lo.b_lo <= 2^-22 |a.b|)
// i.e. three tcgen05.mma.kind::tf32 per K step.  Relative error per product <= 3 * 2^-22 (fp32 FMA: 2^-24).
//
// Structure (one CTA per SM, persistent over 128-row tiles of A; whole N and K per tile):
//   warp 0      TMA producer: per 32-wide K block one box of A [128 x 32] and of Bt_hi / Bt_lo [N x 32]
//               (SWIZZLE_128B, K-major) into a 2-stage ring, mbarrier complete_tx
//   warps 8-11  operand split: A_raw -> A_hi (in place) and A_lo (second buffer, same swizzled offsets), then
//               fence.proxy.async + mbarrier arrive
//   warp 1      one elected thread issues 12 tcgen05.mma (M128 x N x K8, 3 products x 4 K steps) per K block into a
//               double-buffered fp32 accumulator in TMEM (2 x 256 columns); tcgen05.commit frees the smem stage and,
//               after the last K block, hands the accumulator to the epilogue
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> (+ bias) -> 128-byte row segments to global
// Bt_hi / Bt_lo (the small weight matrix, transposed and split) are prepared by the host mirror once per call.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int kGemmThreads = 384;       // 12 warps
constexpr int kBlockM = 128;
constexpr int kBoxCols = 32;            // fp32 elements of one 128-byte swizzle row (TMA box inner extent, default)
constexpr int kMaxN = 256;
constexpr uint32_t kTmemCols = 512;     // two accumulators of up to 256 fp32 columns

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must end in a trap (the process dies, the GPU stays usable), never in a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major swizzled operand tiles (umma_desc_k below): rows of 128 (64) bytes, 8-row groups SBO = 1024 (512) bytes apart,
// LBO unused (1), descriptor version 1 (sm_100), layout type SWIZZLE_128B = 2 / SWIZZLE_64B = 4
// (cute/arch/mma_sm100_desc.hpp: SmemDescriptor).
// Instruction descriptor (UMMA::InstrDescriptor): D fp32, A/B tf32, K-major both, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 2u << 7;                    // a_format = TF32
    d |= 2u << 10;                   // b_format = TF32
    d |= (uint32_t)(n >> 3) << 17;   // n_dim
    d |= (uint32_t)(kBlockM >> 4) << 24;   // m_dim
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// KB = 32: 128-byte rows (SWIZZLE_128B), 96 KB stages, 2 of them; KB = 16: 64-byte rows (SWIZZLE_64B), 48 KB stages,
// 4 of them -- twice as many TMA -> split -> MMA chains in flight for the same shared memory.
template <int KB>
struct GemmCfg {
    static constexpr int kStages = KB == 32 ? 2 : 4;
    static constexpr uint32_t kATile = kBlockM * KB * 4;
    static constexpr uint32_t kBTile = kMaxN * KB * 4;
    static constexpr uint32_t kStageBytes = 2 * kATile + 2 * kBTile;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 + 256;
    static constexpr uint32_t kSbo = 8 * KB * 4;                 // bytes between 8-row groups
    static constexpr uint64_t kLayout = KB == 32 ? 2 : 4;        // UMMA LayoutType: SWIZZLE_128B / SWIZZLE_64B
};

template <int KB>
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(GemmCfg<KB>::kSbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= GemmCfg<KB>::kLayout << 61;
    return d;
}

template <int KB>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                   const __grid_constant__ CUtensorMap map_blo, const float *__restrict__ bias,
                   float *__restrict__ C, int64_t ldc, int M, int N, int K) {
    constexpr int kStages = GemmCfg<KB>::kStages;
    constexpr int kBlockK = KB;
    constexpr uint32_t kATile = GemmCfg<KB>::kATile, kBTile = GemmCfg<KB>::kBTile, kStageBytes = GemmCfg<KB>::kStageBytes;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_addr(smem_raw) + 1023u) & ~1023u;          // swizzled tiles: 1024-byte aligned
    uint8_t *gen = smem_raw + (base - smem_addr(smem_raw));
    const uint32_t bars = base + kStages * kStageBytes;
    // barriers: full[s] (TMA landed), split[s] (A_hi / A_lo written), empty[s] (MMAs done with the stage),
    //           acc_full[a], acc_empty[a]
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_split = [&](int s) { return bars + 8u * (kStages + s); };
    auto bar_empty = [&](int s) { return bars + 8u * (2 * kStages + s); };
    auto bar_accf = [&](int a) { return bars + 8u * (3 * kStages + a); };
    auto bar_acce = [&](int a) { return bars + 8u * (3 * kStages + 2 + a); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + kStages * kStageBytes + 8 * (3 * kStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mma = ((N + 15) / 16) * 16;                 // M = 128 needs N % 16 == 0
    const int k_blocks = (K + kBlockK - 1) / kBlockK;
    const int tiles = (M + kBlockM - 1) / kBlockM;
    const uint32_t stage_tx = kATile + 2u * (uint32_t)n_mma * kBlockK * 4u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_split(s), 4); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_accf(a), 1); mbar_init(bar_acce(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {     // TMEM allocation: one warp, address published through shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(bar_empty(s), ph ^ 1u);
                    const uint32_t st = base + s * kStageBytes;
                    mbar_expect_tx(bar_full(s), stage_tx);
                    tma_load_2d(st, &map_a, kb * kBlockK, t * kBlockM, bar_full(s));
                    tma_load_2d(st + 2 * kATile, &map_bhi, kb * kBlockK, 0, bar_full(s));
                    tma_load_2d(st + 2 * kATile + kBTile, &map_blo, kb * kBlockK, 0, bar_full(s));
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(n_mma);
            int s = 0; uint32_t ph = 0;
            int a = 0; uint32_t aph = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(bar_acce(a), aph ^ 1u);              // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)a * kMaxN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(bar_full(s), ph);
                    mbar_wait(bar_split(s), ph);
                    tcgen05_fence_after();
                    const uint32_t st = base + s * kStageBytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / 8; ++k) {
                        const uint64_t ahi = umma_desc_k<KB>(st + k * 32);
                        const uint64_t alo = umma_desc_k<KB>(st + kATile + k * 32);
                        const uint64_t bhi = umma_desc_k<KB>(st + 2 * kATile + k * 32);
                        const uint64_t blo = umma_desc_k<KB>(st + 2 * kATile + kBTile + k * 32);
                        umma_tf32(tmem_d, alo, bhi, idesc, (kb | k) != 0);     // small terms first
                        umma_tf32(tmem_d, ahi, blo, idesc, 1);
                        umma_tf32(tmem_d, ahi, bhi, idesc, 1);
                    }
                    tcgen05_commit(bar_empty(s));                              // smem stage free once these MMAs retire
                    if (kb == k_blocks - 1) tcgen05_commit(bar_accf(a));       // accumulator complete
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
                if (++a == 2) { a = 0; aph ^= 1u; }
            }
        }
    } else if (warp >= 8) {
        // ===== operand split: A_raw -> A_hi (in place), A_lo =====
        const int tid = threadIdx.x - 256;                       // 0..127
        int s = 0; uint32_t ph = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(bar_full(s), ph);
                float4 *hi = reinterpret_cast<float4 *>(gen + s * kStageBytes);
                float4 *lo = reinterpret_cast<float4 *>(gen + s * kStageBytes + kATile);
#pragma unroll
                for (int i = 0; i < (int)(kATile / 16 / 128); ++i) {
                    const int idx = i * 128 + tid;
                    float4 v = hi[idx];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    hi[idx] = h;
                    lo[idx] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_split(s));
                if (++s == kStages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> (+ bias) -> global =====
        const int q = warp & 3;                                   // TMEM lane quarter this warp may touch
        int a = 0; uint32_t aph = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            mbar_wait(bar_accf(a), aph);
            tcgen05_fence_after();
            const int row = t * kBlockM + q * 32 + lane;
            float *crow = C + (int64_t)row * ldc;
            for (int c0 = 0; c0 < n_mma; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * kMaxN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < M) {
                    if (c0 + 32 <= N && (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 o;
                            o.x = __uint_as_float(r[j]); o.y = __uint_as_float(r[j + 1]);
                            o.z = __uint_as_float(r[j + 2]); o.w = __uint_as_float(r[j + 3]);
                            if (bias) {
                                const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c0 + j));
                                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                            }
                            *reinterpret_cast<float4 *>(crow + c0 + j) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (c0 + j < N) {
                                float o = __uint_as_float(r[j]);
                                if (bias) o += __ldg(bias + c0 + j);
                                crow[c0 + j] = o;
                            }
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce(a));
            if (++a == 2) { a = 0; aph ^= 1u; }
        }
    }
    // teardown
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// Weight gradient: dW[N, K] = dY[M, N]^T . X[M, K], the reduction running over the M node rows (split-K).
//
// Both operands have the reduction index as their SLOW memory dimension (row-major [rows, features]), i.e. they are
// MN-major for the tensor core.  For 32-bit MN-major operands the only swizzled shared-memory layout tcgen05 accepts is
// SWIZZLE_128B_BASE32B (32-byte chunks permuted inside 128-byte feature runs, atoms of 4 rows = 512 bytes; CUTLASS:
// "for mn-major tf32 operands, SW128_32B is the only available smem layout"), which is exactly what a TMA box
// {32 features, 16 rows} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B lands: per 32-feature chunk one box, chunks
// LBO = 2048 bytes apart, 4-row groups SBO = 512 bytes apart; one tcgen05.mma (K = 8 rows) reads two atoms per chunk.  Every CTA reduces a contiguous range of 16-row blocks into TWO fp32 accumulators in
// TMEM (output features 0-127 and 128-255: all 512 columns) with the same 3xTF32 splitting as above -- here both
// operand tiles are split by 8 warps -- and writes its partial [N, K] to `partials[cta]`; the host mirror sums the
// partials (deterministic for a fixed grid).
constexpr int kWgThreads = 512;               // 16 warps
constexpr int kWgRows = 16;                   // rows (MMA K) per stage = 2 MMA K steps
constexpr int kWgStages = 3;
constexpr uint32_t kWgBox = kWgRows * 128;    // bytes of one {32 features x 16 rows} box
constexpr uint32_t kWgOperand = 8 * kWgBox;   // up to 256 features = 8 boxes = 16 KB
constexpr uint32_t kWgStageBytes = 4 * kWgOperand;   // dY_hi | dY_lo | X_hi | X_lo = 64 KB
constexpr uint32_t kWgSmemBytes = kWgStages * kWgStageBytes + 1024 + 256;

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(kWgBox >> 4) << 16;              // leading byte offset: next 32-feature chunk
    d |= (uint64_t)(512 >> 4) << 32;                 // stride byte offset: next 4-row group
    d |= (uint64_t)1 << 46;                          // version
    d |= (uint64_t)1 << 61;                          // SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                    float *__restrict__ partials, int M, int N, int K) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_addr(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_addr(smem_raw));
    const uint32_t bars = base + kWgStages * kWgStageBytes;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_split = [&](int s) { return bars + 8u * (kWgStages + s); };
    auto bar_empty = [&](int s) { return bars + 8u * (2 * kWgStages + s); };
    const uint32_t bar_acc = bars + 8u * (3 * kWgStages);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + kWgStages * kWgStageBytes + 8 * (3 * kWgStages + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (N + 31) / 32, k_chunks = (K + 31) / 32;      // 32-feature boxes of dY / X
    const int m_tiles = (N + 127) / 128;                               // accumulators (MMA M = 128 output features)
    const int n_mma = ((K + 15) / 16) * 16;                            // MMA N = input features
    const int blocks = (M + kWgRows - 1) / kWgRows;
    const int per = (blocks + gridDim.x - 1) / gridDim.x;
    const int b0 = blockIdx.x * per;
    const int b1 = (b0 + per < blocks) ? b0 + per : blocks;
    const uint32_t stage_tx = (uint32_t)(n_chunks + k_chunks) * kWgBox;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_split(s), 8); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int b = b0; b < b1; ++b) {
                mbar_wait(bar_empty(s), ph ^ 1u);
                const uint32_t st = base + s * kWgStageBytes;
                mbar_expect_tx(bar_full(s), stage_tx);
                for (int c = 0; c < n_chunks; ++c) tma_load_2d(st + c * kWgBox, &map_dy, c * 32, b * kWgRows, bar_full(s));
                for (int c = 0; c < k_chunks; ++c) tma_load_2d(st + 2 * kWgOperand + c * kWgBox, &map_x, c * 32, b * kWgRows, bar_full(s));
                if (++s == kWgStages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t idesc = umma_idesc_tf32(n_mma) | (1u << 15) | (1u << 16);     // A and B MN-major
            int s = 0; uint32_t ph = 0;
            for (int b = b0; b < b1; ++b) {
                mbar_wait(bar_full(s), ph);
                mbar_wait(bar_split(s), ph);
                tcgen05_fence_after();
                const uint32_t st = base + s * kWgStageBytes;
#pragma unroll
                for (int k = 0; k < kWgRows / 8; ++k) {
                    const uint64_t xhi = umma_desc_mn_sw128(st + 2 * kWgOperand + k * 1024);
                    const uint64_t xlo = umma_desc_mn_sw128(st + 3 * kWgOperand + k * 1024);
                    for (int mt = 0; mt < m_tiles; ++mt) {
                        const uint64_t yhi = umma_desc_mn_sw128(st + mt * 4 * kWgBox + k * 1024);
                        const uint64_t ylo = umma_desc_mn_sw128(st + kWgOperand + mt * 4 * kWgBox + k * 1024);
                        const uint32_t d = tmem_base + (uint32_t)mt * kMaxN;
                        umma_tf32(d, ylo, xhi, idesc, (b != b0) || (k != 0));
                        umma_tf32(d, yhi, xlo, idesc, 1);
                        umma_tf32(d, yhi, xhi, idesc, 1);
                    }
                }
                tcgen05_commit(bar_empty(s));
                if (++s == kWgStages) { s = 0; ph ^= 1u; }
            }
            tcgen05_commit(bar_acc);
        }
    } else if (warp >= 8) {
        // split both operand tiles: hi in place, lo into the second buffer (same swizzled offsets)
        const int tid = threadIdx.x - 256;                       // 0..255
        int s = 0; uint32_t ph = 0;
        const int y_vec = n_chunks * (int)(kWgBox / 16), x_vec = k_chunks * (int)(kWgBox / 16);
        for (int b = b0; b < b1; ++b) {
            mbar_wait(bar_full(s), ph);
            uint8_t *st = gen + s * kWgStageBytes;
            for (int part = 0; part < 2; ++part) {
                float4 *hi = reinterpret_cast<float4 *>(st + part * 2 * kWgOperand);
                float4 *lo = reinterpret_cast<float4 *>(st + part * 2 * kWgOperand + kWgOperand);
                const int nvec = part == 0 ? y_vec : x_vec;
                for (int idx = tid; idx < nvec; idx += 256) {
                    float4 v = hi[idx];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    hi[idx] = h;
                    lo[idx] = l;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_split(s));
            if (++s == kWgStages) { s = 0; ph ^= 1u; }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        float *out = partials + (size_t)blockIdx.x * N * K;
        if (b1 > b0) {
            mbar_wait(bar_acc, 0);
            tcgen05_fence_after();
        }
        for (int mt = 0; mt < m_tiles; ++mt) {
            const int row = mt * 128 + q * 32 + lane;            // output feature n
            for (int c0 = 0; c0 < n_mma; c0 += 32) {
                uint32_t r[32];
                if (b1 > b0) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * kMaxN + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;      // a CTA without rows contributes zeros
                }
                if (row < N) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < K) out[(size_t)row * K + c0 + j] = __uint_as_float(r[j]);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encoder() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return q == cudaDriverEntryPointSuccess ? (EncodeTiledFn)fn : nullptr;
}

// 2-D fp32 map {cols, rows} (cols contiguous), box {32, box_rows}, SWIZZLE_128B, out-of-bounds elements read as zero
int make_tile_map(CUtensorMap *m, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                  CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, int box_cols = kBoxCols) {
    // box = {32 columns (one 128-byte swizzle row), box_rows}
    EncodeTiledFn enc = encoder();
    ADAQP_REQUIRE(enc != nullptr, ADAQP_EINVAL, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ADAQP_REQUIRE(r == CUDA_SUCCESS, ADAQP_EINVAL, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

}  // namespace

extern "C" {

int adaqp_gemm_tf32x3_supported(int64_t M, int32_t N, int32_t K, int64_t lda, int64_t ldb, int64_t ldc) {
    if (M <= 0 || N <= 0 || K <= 0 || N > kMaxN || K > 4096) return 0;
    if ((lda & 3) || (ldb & 3) || lda < K || ldb < K || ldc < N) return 0;   // TMA: 16-byte row pitch
    return 1;
}

int adaqp_gemm_tf32x3_f32(const float *A, int64_t lda, const float *Bt_hi, const float *Bt_lo, int64_t ldb,
                          const float *bias, int64_t M, int32_t N, int32_t K, float *C, int64_t ldc, void *stream) {
    ADAQP_REQUIRE(adaqp_gemm_tf32x3_supported(M, N, K, lda, ldb, ldc), ADAQP_ELIMIT,
                  "adaqp_gemm_tf32x3_f32: unsupported shape M=%lld N=%d K=%d lda=%lld ldb=%lld", (long long)M, N, K,
                  (long long)lda, (long long)ldb);
    ADAQP_REQUIRE(A && Bt_hi && Bt_lo && C, ADAQP_EINVAL, "adaqp_gemm_tf32x3_f32: null pointer");
    ADAQP_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)Bt_hi & 15) == 0 && ((uintptr_t)Bt_lo & 15) == 0, ADAQP_EALIGN,
                  "adaqp_gemm_tf32x3_f32: operands must be 16-byte aligned");
    ADAQP_REQUIRE(M < (1ll << 31), ADAQP_ELIMIT, "adaqp_gemm_tf32x3_f32: M too large");
    const int n_mma = ((N + 15) / 16) * 16;
    const int kb = adaqp_options().gemm_block_k == 32 ? 32 : 16;
    const CUtensorMapSwizzle sw = kb == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap ma, mh, ml;
    int rc = make_tile_map(&ma, A, M, K, lda, kBlockM, sw, kb);
    if (rc) return rc;
    rc = make_tile_map(&mh, Bt_hi, N, K, ldb, n_mma, sw, kb);
    if (rc) return rc;
    rc = make_tile_map(&ml, Bt_lo, N, K, ldb, n_mma, sw, kb);
    if (rc) return rc;
    const int sms = adaqp_sm_count() > 0 ? adaqp_sm_count() : 148;
    const int64_t tiles = (M + kBlockM - 1) / kBlockM;
    const int grid = (int)(tiles < sms ? tiles : sms);
    if (kb == 32) {
        ADAQP_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<32>::kSmemBytes));
        gemm_tf32x3_kernel<32><<<grid, kGemmThreads, GemmCfg<32>::kSmemBytes, (cudaStream_t)stream>>>(ma, mh, ml, bias, C, ldc, (int)M, N, K);
    } else {
        ADAQP_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<16>::kSmemBytes));
        gemm_tf32x3_kernel<16><<<grid, kGemmThreads, GemmCfg<16>::kSmemBytes, (cudaStream_t)stream>>>(ma, mh, ml, bias, C, ldc, (int)M, N, K);
    }
    return adaqp_check_launch("gemm_tf32x3_kernel");
}

int adaqp_wgrad_tf32x3_supported(int64_t M, int32_t N, int32_t K, int64_t ldy, int64_t ldx) {
    if (M <= 0 || N <= 0 || K <= 0 || N > kMaxN || K > kMaxN || M >= (1ll << 31)) return 0;
    if ((ldy & 3) || (ldx & 3) || ldy < N || ldx < K) return 0;
    return 1;
}

int adaqp_wgrad_tf32x3_grid(int64_t M) {
    const int sms = adaqp_sm_count() > 0 ? adaqp_sm_count() : 148;
    const int64_t blocks = (M + kWgRows - 1) / kWgRows;
    return (int)(blocks < sms ? (blocks < 1 ? 1 : blocks) : sms);
}

int adaqp_wgrad_tf32x3_f32(const float *dY, int64_t ldy, const float *X, int64_t ldx, int64_t M, int32_t N, int32_t K,
                           float *partials, int32_t grid, void *stream) {
    ADAQP_REQUIRE(adaqp_wgrad_tf32x3_supported(M, N, K, ldy, ldx), ADAQP_ELIMIT,
                  "adaqp_wgrad_tf32x3_f32: unsupported shape M=%lld N=%d K=%d", (long long)M, N, K);
    ADAQP_REQUIRE(dY && X && partials, ADAQP_EINVAL, "adaqp_wgrad_tf32x3_f32: null pointer");
    ADAQP_REQUIRE(((uintptr_t)dY & 15) == 0 && ((uintptr_t)X & 15) == 0, ADAQP_EALIGN, "adaqp_wgrad_tf32x3_f32: 16-byte alignment");
    ADAQP_REQUIRE(grid == adaqp_wgrad_tf32x3_grid(M), ADAQP_EINVAL, "adaqp_wgrad_tf32x3_f32: partials must hold adaqp_wgrad_tf32x3_grid(M) slices");
    CUtensorMap my, mx;
    int rc = make_tile_map(&my, dY, M, N, ldy, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    rc = make_tile_map(&mx, X, M, K, ldx, kWgRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    ADAQP_CUDA(cudaFuncSetAttribute(wgrad_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmemBytes));
    wgrad_tf32x3_kernel<<<grid, kWgThreads, kWgSmemBytes, (cudaStream_t)stream>>>(my, mx, partials, (int)M, N, K);
    return adaqp_check_launch("wgrad_tf32x3_kernel");
}

}  // extern "C"
