// Fused boundary-message exchange kernels (sm_100a).
//
// Sender: gather halo rows -> per-row min/max -> stochastic quantize at the row's
// assigned bit-width -> bit-pack in the reference wire layout -> store packed bytes and
// bf16 (scale, min) straight into the destination GPU's slab over NVLink/NVSwitch P2P ->
// publish a per-(key, peer) sequence flag.  Receiver: acquire flag -> unpack ->
// dequantize with the bf16 parameters -> scatter into the halo matrix -> ack.
// No NCCL, no host staging, one launch per side per layer key.
//
// Replaces AdaQP/model/op_util.py:137-236 (msg_all2all_GLOO and everything below it),
// AdaQP/communicator/comm.py:166-222 and the per-(peer, bit) launches of
// quantization_cuda_kernel.cu.  The bytes landing in the receive slab are exactly the
// reference's wire format; parity is checked against the oracle and the
// reference-built quant_cuda in tests/.
//
// Work decomposition: one warp per "byte-row" (8/bits consecutive rows of one
// (peer, bit) segment that share packed bytes).  Rows live in registers (one pass over
// HBM), the F packed bytes are staged in shared memory at the same 16-byte phase as
// their destination so the peer stores are 16-byte vectors regardless of the odd
// segment offsets of the reference layout.
#include "common.cuh"

static_assert(sizeof(adaqp_send_item) == 64, "adaqp_send_item layout");
static_assert(sizeof(adaqp_recv_item) == 32, "adaqp_recv_item layout");
static_assert(sizeof(adaqp_fp_item) == 16, "adaqp_fp_item layout");

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ void report(uint32_t *status, uint32_t code, uint32_t slot) {
    if (status && atomicCAS(status, 0u, code) == 0u) status[1] = slot;
}

// ------------------------------------------------------------------ row loads
template <int VEC> struct RowVec;
template <> struct RowVec<4> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
        const float4 t = ldg_stream_f4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct RowVec<2> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[2]) {
        const float2 t = ldg_stream_f2(p); v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[2]) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    }
};
template <> struct RowVec<1> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[1]) { v[0] = ldg_stream_f1(p); }
    static __device__ __forceinline__ void store(float *p, const float (&v)[1]) { *p = v[0]; }
};

// Copy nbytes from shared memory `src` to global `dst` where (src & 15) == (dst & 15):
// byte head to the next 16-byte boundary, 16-byte body, byte tail.  Warp-cooperative.
__device__ __forceinline__ void warp_copy_out(uint8_t *dst, const uint8_t *src, int nbytes, int lane) {
    int head = (int)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
    if (head > nbytes) head = nbytes;
    if (lane < head) dst[lane] = src[lane];
    const int body = (nbytes - head) >> 4;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
    for (int i = lane; i < body; i += 32) d4[i] = s4[i];
    const int done = head + (body << 4);
    const int tail = nbytes - done;
    if (lane < tail) dst[done + lane] = src[done + lane];
}

// ------------------------------------------------------------------ sender
// FULL: every row of the byte-row exists and F == 32 * VEC * CHUNKS, so no row / column
// predicates are needed (the common case: F = 256, all but the last byte-row of a segment).
template <int BITS, int VEC, int CHUNKS, bool FULL>
__device__ __forceinline__ void send_item(const adaqp_send_item &it, const adaqp_send_chan &ch,
                                          const float *__restrict__ x, int64_t ld, int F,
                                          float *__restrict__ trace, float trace_coef,
                                          const PhiloxKeys &keys, uint64_t base_offset,
                                          uint8_t *stage, int lane) {
    constexpr int WPT = 8 / BITS;
    float v[WPT][CHUNKS][VEC];
    float lo[WPT], hi[WPT];
    bool nan[WPT];
    const int nrows = it.nrows;
    // (1) one pass over HBM: all loads issued before any use
#pragma unroll
    for (int r = 0; r < WPT; ++r) {
        const int row = it.src_row[r];
        lo[r] = INFINITY; hi[r] = -INFINITY; nan[r] = false;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int col = (c * 32 + lane) * VEC;
            if (FULL || (r < nrows && col < F)) {
                RowVec<VEC>::load(x + (int64_t)row * ld + col, v[r][c]);
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[r][c][e] = 0.f;
            }
        }
    }
    // (2) per-row min / max (torch.min/max semantics: NaN propagates)
#pragma unroll
    for (int r = 0; r < WPT; ++r) {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int col = (c * 32 + lane) * VEC;
            if (FULL || col < F) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const float t = v[r][c][e];
                    nan[r] |= (t != t);
                    lo[r] = fminf(lo[r], t);
                    hi[r] = fmaxf(hi[r], t);
                }
            }
        }
        lo[r] = warp_min(lo[r]);
        hi[r] = warp_max(hi[r]);
        if (__any_sync(ADAQP_FULL_MASK, nan[r])) { lo[r] = __int_as_float(0x7fc00000); hi[r] = lo[r]; }
    }
    // (3) scale = reciprocal(max - min) * (2^b - 1) as torch evaluates (2**b-1)/(rmax-rmin);
    //     wire params are bf16(scale), bf16(min)
    float scale[WPT];
    constexpr float levels = (float)((1 << BITS) - 1);
#pragma unroll
    for (int r = 0; r < WPT; ++r) {
        const float range = __fsub_rn(hi[r], lo[r]);
        scale[r] = __fmul_rn(__frcp_rn(range), levels);  // Tensor.__rtruediv__: reciprocal(range) * levels
        if (lane == r && (FULL || r < nrows)) {
            ch.params[it.param_pos + r] = f32_to_bf16_bits(scale[r]);
            ch.params[ch.S + it.param_pos + r] = f32_to_bf16_bits(lo[r]);
            if (trace) {  // trace_input: (dim / 6) * (rmax - rmin) ** 2, op_util.py:95-97
                const int pos = it.send_pos[r];
                trace[pos] = __fadd_rn(trace[pos], __fmul_rn(trace_coef, __fmul_rn(range, range)));
            }
        }
    }
    // (4) quantize + pack; byte k = group*F + col is its own Philox subsequence
    uint8_t *dst = ch.qdata + it.dst_off;
    const int phase16 = (int)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const uint64_t offset = base_offset + it.rel_offset;
    const uint64_t kbase = (uint64_t)it.group * (uint64_t)F;
    // offset % 4 == 0 (enforced by the launcher; torch's Philox offsets always are), so the
    // WPT <= 4 draws of a byte are the four words of ONE block at counter (offset/4, k).
    const uint64_t blk = offset >> 2;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = (c * 32 + lane) * VEC;
        if (FULL || col0 < F) {
            uint4 ctr[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const uint64_t k = kbase + (uint64_t)(col0 + e);
                ctr[e] = make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)k, (uint32_t)(k >> 32));
            }
            philox4x32_10_xN<VEC>(ctr, keys);          // VEC interleaved chains
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const uint32_t w[4] = {ctr[e].x, ctr[e].y, ctr[e].z, ctr[e].w};
                uint32_t byte = 0;
#pragma unroll
                for (int r = 0; r < WPT; ++r) {
                    if (FULL || r < nrows) {
                        const int q = quantize_one(v[r][c][e], lo[r], scale[r], uniform_from_u32(w[r]));
                        byte |= ((uint32_t)q << (r * BITS));
                    }
                }
                stage[phase16 + col0 + e] = (uint8_t)byte;
            }
        }
    }
    __syncwarp();
    // (5) 16-byte peer stores
    warp_copy_out(dst, stage + phase16, F, lane);
    __syncwarp();
}

template <int VEC, int CHUNKS>
__global__ void __launch_bounds__(kThreads, (VEC * CHUNKS <= 8) ? 3 : 1)
send_quant_kernel(const float *__restrict__ x, int64_t ld, int F,
                  const adaqp_send_item *__restrict__ items, int64_t n_items,
                  const adaqp_send_chan *__restrict__ chans, int n_chans,
                  float *__restrict__ trace, const __grid_constant__ PhiloxKeys keys, uint64_t seed,
                  uint64_t base_offset, uint32_t seq, uint32_t *work, uint32_t *status, uint64_t timeout_ns) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int stage_stride = ((F + 16 + 15) >> 4) << 4;
    uint8_t *stage = smem + (size_t)wib * stage_stride;
    // channel table in shared memory: a 30-cycle LDS instead of a dependent global load per item
    adaqp_send_chan *s_chans = reinterpret_cast<adaqp_send_chan *>(smem + (size_t)kWarps * stage_stride);
    for (int t = threadIdx.x; t < n_chans * (int)(sizeof(adaqp_send_chan) / 8); t += kThreads)
        reinterpret_cast<uint64_t *>(s_chans)[t] = reinterpret_cast<const uint64_t *>(chans)[t];
    __syncthreads();
    const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const float trace_coef = (float)((double)F / 6.0);
    int acked = -1;  // channel whose slab is known to be free for this seq
    for (int64_t i = warp; i < n_items; i += nwarps) {
        adaqp_send_item it;
        {   // 64-byte item through four uniform 16-byte loads (broadcast by L1)
            const uint4 *p = reinterpret_cast<const uint4 *>(items + i);
            uint4 *q = reinterpret_cast<uint4 *>(&it);
            q[0] = __ldg(p); q[1] = __ldg(p + 1); q[2] = __ldg(p + 2); q[3] = __ldg(p + 3);
        }
        const adaqp_send_chan ch = s_chans[it.chan];
        if (it.chan != acked) {
            // the peer must have consumed the previous payload of this key (seq - 1)
            bool ok = true;
            if (lane == 0) ok = spin_wait_ge(ch.ack, seq - 1u, timeout_ns);
            ok = __shfl_sync(ADAQP_FULL_MASK, ok, 0);
            if (!ok && lane == 0) report(status, ADAQP_ST_ACK_TIMEOUT, (uint32_t)it.chan);
            acked = it.chan;
        }
        const bool full = (F == 32 * VEC * CHUNKS) && (it.nrows * it.bits == 8);
#define ADAQP_SEND(B)                                                                                          \
    do {                                                                                                       \
        if (full) send_item<B, VEC, CHUNKS, true>(it, ch, x, ld, F, trace, trace_coef, keys, base_offset, stage, lane);  \
        else send_item<B, VEC, CHUNKS, false>(it, ch, x, ld, F, trace, trace_coef, keys, base_offset, stage, lane);      \
    } while (0)
        switch (it.bits) {
            case 2: ADAQP_SEND(2); break;
            case 4: ADAQP_SEND(4); break;
            default: ADAQP_SEND(8); break;
        }
#undef ADAQP_SEND
    }
    // publish: every thread's peer stores are fenced, the last CTA raises the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(work, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            for (int c = 0; c < n_chans; ++c) st_release_sys(chans[c].flag, seq);
            atomicExch(work, 0u);
        }
    }
}

// ------------------------------------------------------------------ receiver
template <int BITS, int VEC, int CHUNKS>
__device__ __forceinline__ void recv_item(const adaqp_recv_item &it, const adaqp_recv_chan &ch,
                                          float *__restrict__ halo, int64_t ld, int F, int lane) {
    constexpr int WPT = 8 / BITS;
    constexpr uint32_t MASK = (1u << BITS) - 1u;
    const uint8_t *src = ch.qdata + it.src_off;
    const int nrows = it.nrows;
    // 2/4-bit: a byte-row has only WPT * 2^BITS (= 16 / 32) distinct outputs val/scale + min.
    // Lane (r * 2^BITS + val) computes that one value with the reference's IEEE division and
    // add; every element is then a single shuffle lookup instead of a division.
    constexpr int NV = 1 << BITS;
    float lut = 0.f;
    float scale[WPT], mn[WPT], qz[WPT];
    if (BITS <= 4) {
        const int lr = lane / NV, lv = lane % NV;
        if (lr < nrows) {
            const float sc = bf16_bits_to_f32(__ldcg(ch.params + it.param_pos + lr));
            const float m = bf16_bits_to_f32(__ldcg(ch.params + ch.S + it.param_pos + lr));
            lut = __fadd_rn(__fdiv_rn((float)lv, sc), m);
        }
    } else {
#pragma unroll
        for (int r = 0; r < WPT; ++r) {
            if (r < nrows) {
                scale[r] = bf16_bits_to_f32(__ldcg(ch.params + it.param_pos + r));
                mn[r] = bf16_bits_to_f32(__ldcg(ch.params + ch.S + it.param_pos + r));
            } else {
                scale[r] = 1.f; mn[r] = 0.f;
            }
            qz[r] = __fdiv_rn(0.f, scale[r]);
        }
    }
    uint32_t bytes[CHUNKS][VEC];
    if (VEC == 4 && (reinterpret_cast<uintptr_t>(src) & 3u) == 0) {  // warp-uniform
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int col = (c * 32 + lane) * VEC;
            const uint32_t w = (col < F) ? __ldcg(reinterpret_cast<const uint32_t *>(src + col)) : 0u;
#pragma unroll
            for (int e = 0; e < VEC; ++e) bytes[c][e] = (w >> (8 * e)) & 0xffu;
        }
    } else {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int col = (c * 32 + lane) * VEC + e;
                bytes[c][e] = (col < F) ? (uint32_t)__ldcg(src + col) : 0u;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < WPT; ++r) {
        if (r < nrows) {
            float *orow = halo + (int64_t)it.dst_row[r] * ld;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const int col = (c * 32 + lane) * VEC;
                {
                    float o[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const uint32_t ival = (bytes[c][e] >> (r * BITS)) & MASK;
                        if (BITS <= 4) o[e] = __shfl_sync(ADAQP_FULL_MASK, lut, r * NV + (int)ival);
                        else o[e] = __fadd_rn(dequant_div(ival, scale[r], qz[r]), mn[r]);  // IEEE div then add, as unpack
                    }
                    if (col < F) RowVec<VEC>::store(orow + col, o);   // lanes past F still take part in the shuffles
                }
            }
        }
    }
}

template <int VEC, int CHUNKS>
__global__ void __launch_bounds__(kThreads)
recv_quant_kernel(float *__restrict__ halo, int64_t ld, int F,
                  const adaqp_recv_item *__restrict__ items, int64_t n_items,
                  const adaqp_recv_chan *__restrict__ chans, int n_chans, uint32_t seq,
                  uint32_t *work, uint32_t *status, uint64_t timeout_ns) {
    extern __shared__ __align__(16) uint8_t smem[];
    adaqp_recv_chan *s_chans = reinterpret_cast<adaqp_recv_chan *>(smem);
    for (int t = threadIdx.x; t < n_chans * (int)(sizeof(adaqp_recv_chan) / 8); t += kThreads)
        reinterpret_cast<uint64_t *>(s_chans)[t] = reinterpret_cast<const uint64_t *>(chans)[t];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    int ready = -1;
    for (int64_t i = warp; i < n_items; i += nwarps) {
        adaqp_recv_item it;
        {
            const uint4 *p = reinterpret_cast<const uint4 *>(items + i);
            uint4 *q = reinterpret_cast<uint4 *>(&it);
            q[0] = __ldg(p); q[1] = __ldg(p + 1);
        }
        const adaqp_recv_chan ch = s_chans[it.chan];
        if (it.chan != ready) {
            bool ok = true;
            if (lane == 0) ok = spin_wait_ge(ch.flag, seq, timeout_ns);
            ok = __shfl_sync(ADAQP_FULL_MASK, ok, 0);
            __syncwarp();  // order the other lanes' slab reads after lane 0's acquire
            if (!ok && lane == 0) report(status, ADAQP_ST_FLAG_TIMEOUT, (uint32_t)it.chan);
            ready = it.chan;
        }
        switch (it.bits) {
            case 2: recv_item<2, VEC, CHUNKS>(it, ch, halo, ld, F, lane); break;
            case 4: recv_item<4, VEC, CHUNKS>(it, ch, halo, ld, F, lane); break;
            default: recv_item<8, VEC, CHUNKS>(it, ch, halo, ld, F, lane); break;
        }
    }
    // all slab reads done -> tell the senders the regions may be overwritten
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(work, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            for (int c = 0; c < n_chans; ++c) st_release_sys(chans[c].ack, seq);
            atomicExch(work, 0u);
        }
    }
}

// ------------------------------------------------------------------ fp32 exchange
template <int VEC>
__global__ void __launch_bounds__(kThreads)
send_fp32_kernel(const float *__restrict__ x, int64_t ld, int F,
                 const adaqp_fp_item *__restrict__ items, int64_t n_items,
                 const adaqp_send_chan *__restrict__ chans, int n_chans, int64_t dst_ld,
                 uint32_t seq, uint32_t *work, uint32_t *status, uint64_t timeout_ns) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    int acked = -1;
    for (int64_t i = warp; i < n_items; i += nwarps) {
        const adaqp_fp_item it = items[i];
        const adaqp_send_chan ch = chans[it.chan];
        if (it.chan != acked) {
            bool ok = true;
            if (lane == 0) ok = spin_wait_ge(ch.ack, seq - 1u, timeout_ns);
            ok = __shfl_sync(ADAQP_FULL_MASK, ok, 0);
            if (!ok && lane == 0) report(status, ADAQP_ST_ACK_TIMEOUT, (uint32_t)it.chan);
            acked = it.chan;
        }
        const float *src = x + (int64_t)it.src_row * ld;
        float *dst = ch.fp_rows + it.dst_row * dst_ld;
        const int nvec = F / VEC;
        // 4 independent vector loads in flight per lane
        for (int c0 = lane; c0 < nvec; c0 += 128) {
            float v[4][VEC];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j * 32;
                if (c < nvec) RowVec<VEC>::load(src + c * VEC, v[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j * 32;
                if (c < nvec) RowVec<VEC>::store(dst + c * VEC, v[j]);
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(work, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            for (int c = 0; c < n_chans; ++c) st_release_sys(chans[c].flag, seq);
            atomicExch(work, 0u);
        }
    }
}

__global__ void wait_flags_kernel(const uint32_t *const *flags, int n, uint32_t seq,
                                  uint32_t *status, uint64_t timeout_ns) {
    const int i = threadIdx.x;
    if (i < n) {
        if (!spin_wait_ge(flags[i], seq, timeout_ns)) report(status, ADAQP_ST_FLAG_TIMEOUT, (uint32_t)i);
    }
}

__global__ void post_acks_kernel(uint32_t *const *acks, int n, uint32_t seq) {
    const int i = threadIdx.x;
    if (i < n) {
        __threadfence_system();
        st_release_sys(acks[i], seq);
    }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(const float *__restrict__ x, int64_t ld, const int64_t *__restrict__ idx,
                   int64_t n, int F, float *__restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const int nvec = F / VEC;
    for (int64_t i = warp; i < n; i += nwarps) {
        const float *src = x + idx[i] * ld;
        float *dst = out + i * ldo;
        for (int c = lane; c < nvec; c += 32) {
            float v[VEC];
            RowVec<VEC>::load(src + c * VEC, v);
            RowVec<VEC>::store(dst + c * VEC, v);
        }
    }
}

// ------------------------------------------------------------------ dispatch
inline int pick_vec(int F, int64_t ld, const void *p0, const void *p1 = nullptr, int64_t ld1 = 0) {
    auto ok = [&](int vec) {
        const uintptr_t m = (uintptr_t)vec * 4 - 1;
        if (F % vec) return false;
        if (ld % vec) return false;
        if ((uintptr_t)p0 & m) return false;
        if (p1 && (((uintptr_t)p1 & m) || (ld1 % vec))) return false;
        return true;
    };
    if (ok(4)) return 4;
    if (ok(2)) return 2;
    return 1;
}

// One wave of resident CTAs (persistent-style): ask the runtime how many CTAs of this kernel fit per SM,
// so that the grid is not the requested cap rounded into a second, partly filled wave.
// `total_cap` > 0 (options exch_send_ctas / exch_recv_ctas) bounds the grid to a small persistent
// set of CTAs, so that the SMs left over go to the aggregation the exchange overlaps with.
template <class K>
inline int resident_grid(K kernel, size_t smem, int64_t n_items, int cap_ctas_per_sm, int total_cap) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (per_sm > cap_ctas_per_sm) per_sm = cap_ctas_per_sm;
    const int sms = adaqp_sm_count() > 0 ? adaqp_sm_count() : 148;
    int64_t ctas = (n_items + kWarps - 1) / kWarps;
    int64_t cap = (int64_t)sms * per_sm;
    if (total_cap > 0 && cap > total_cap) cap = total_cap;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

inline int grid_for(int64_t n_items, int max_ctas_per_sm) {
    const int sms = adaqp_sm_count() > 0 ? adaqp_sm_count() : 148;
    int64_t ctas = (n_items + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)sms * max_ctas_per_sm;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

}  // namespace

// CHUNKS ladders per vector width: F <= 32 * VEC * CHUNKS, F <= 1024 like the reference,
// whose pack kernel launches F threads per block.
extern "C" {

int adaqp_send_quant(const float *x, int64_t ld, int32_t F, const adaqp_send_item *items,
                     int64_t n_items, const adaqp_send_chan *chans, int32_t n_chans,
                     float *trace, uint64_t seed, uint64_t base_offset, uint32_t seq,
                     uint32_t *work, uint32_t *status, uint64_t timeout_ns, void *stream) {
    ADAQP_REQUIRE(F > 0 && F <= 1024, ADAQP_ELIMIT, "adaqp_send_quant: F=%d outside (0,1024]", F);
    ADAQP_REQUIRE(n_items >= 0 && n_chans >= 0, ADAQP_EINVAL, "adaqp_send_quant: negative count");
    ADAQP_REQUIRE(work != nullptr, ADAQP_EINVAL, "adaqp_send_quant: null work");
    if (n_chans == 0) return 0;
    ADAQP_REQUIRE(chans && (n_items == 0 || (x && items)), ADAQP_EINVAL, "adaqp_send_quant: null pointer");
    const int vec = pick_vec(F, ld, x);
    const int nchunks = (F + 32 * vec - 1) / (32 * vec);
    const int stage_stride = ((F + 16 + 15) >> 4) << 4;
    ADAQP_REQUIRE((base_offset & 3u) == 0, ADAQP_EINVAL, "adaqp_send_quant: base_offset must be a multiple of 4 (torch Philox offsets are)");
    ADAQP_REQUIRE(n_chans <= 64, ADAQP_ELIMIT, "adaqp_send_quant: more than 64 channels");
    const size_t smem = (size_t)kWarps * stage_stride + (size_t)n_chans * sizeof(adaqp_send_chan);
    const PhiloxKeys keys = make_philox_keys(seed);
    cudaStream_t s = (cudaStream_t)stream;
#define CALL_SEND(V, C)                                                                          \
    send_quant_kernel<V, C><<<resident_grid(send_quant_kernel<V, C>, smem, n_items, 8, adaqp_options().exch_send_ctas), kThreads, smem, s>>>(  \
        x, ld, F, items, n_items, chans, n_chans, trace, keys, seed, base_offset, seq, work, status, timeout_ns)
    if (vec == 4) {
        if (nchunks <= 1) CALL_SEND(4, 1);
        else if (nchunks <= 2) CALL_SEND(4, 2);
        else if (nchunks <= 3) CALL_SEND(4, 3);
        else if (nchunks <= 4) CALL_SEND(4, 4);
        else if (nchunks <= 6) CALL_SEND(4, 6);
        else CALL_SEND(4, 8);
    } else if (vec == 2) {
        if (nchunks <= 2) CALL_SEND(2, 2);
        else if (nchunks <= 4) CALL_SEND(2, 4);
        else if (nchunks <= 6) CALL_SEND(2, 6);
        else if (nchunks <= 8) CALL_SEND(2, 8);
        else if (nchunks <= 10) CALL_SEND(2, 10);
        else if (nchunks <= 12) CALL_SEND(2, 12);
        else CALL_SEND(2, 16);
    } else {
        if (nchunks <= 4) CALL_SEND(1, 4);
        else if (nchunks <= 8) CALL_SEND(1, 8);
        else if (nchunks <= 16) CALL_SEND(1, 16);
        else CALL_SEND(1, 32);
    }
#undef CALL_SEND
    return adaqp_check_launch("send_quant_kernel");
}

int adaqp_recv_quant(float *halo, int64_t ld, int32_t F, const adaqp_recv_item *items,
                     int64_t n_items, const adaqp_recv_chan *chans, int32_t n_chans, uint32_t seq,
                     uint32_t *work, uint32_t *status, uint64_t timeout_ns, void *stream) {
    ADAQP_REQUIRE(F > 0 && F <= 1024, ADAQP_ELIMIT, "adaqp_recv_quant: F=%d outside (0,1024]", F);
    ADAQP_REQUIRE(n_items >= 0 && n_chans >= 0, ADAQP_EINVAL, "adaqp_recv_quant: negative count");
    ADAQP_REQUIRE(work != nullptr, ADAQP_EINVAL, "adaqp_recv_quant: null work");
    if (n_chans == 0) return 0;
    ADAQP_REQUIRE(chans && (n_items == 0 || (halo && items)), ADAQP_EINVAL, "adaqp_recv_quant: null pointer");
    const int vec = pick_vec(F, ld, halo);
    const int nchunks = (F + 32 * vec - 1) / (32 * vec);
    ADAQP_REQUIRE(n_chans <= 64, ADAQP_ELIMIT, "adaqp_recv_quant: more than 64 channels");
    cudaStream_t s = (cudaStream_t)stream;
#define CALL_RECV(V, C)                                                                          \
    recv_quant_kernel<V, C><<<resident_grid(recv_quant_kernel<V, C>, (size_t)n_chans * sizeof(adaqp_recv_chan), n_items, 8, adaqp_options().exch_recv_ctas), kThreads, (size_t)n_chans * sizeof(adaqp_recv_chan), s>>>(halo, ld, F, items, n_items, chans, n_chans, \
                                                      seq, work, status, timeout_ns)
    if (vec == 4) {
        if (nchunks <= 1) CALL_RECV(4, 1);
        else if (nchunks <= 2) CALL_RECV(4, 2);
        else if (nchunks <= 3) CALL_RECV(4, 3);
        else if (nchunks <= 4) CALL_RECV(4, 4);
        else if (nchunks <= 6) CALL_RECV(4, 6);
        else CALL_RECV(4, 8);
    } else if (vec == 2) {
        if (nchunks <= 2) CALL_RECV(2, 2);
        else if (nchunks <= 4) CALL_RECV(2, 4);
        else if (nchunks <= 6) CALL_RECV(2, 6);
        else if (nchunks <= 8) CALL_RECV(2, 8);
        else if (nchunks <= 10) CALL_RECV(2, 10);
        else if (nchunks <= 12) CALL_RECV(2, 12);
        else CALL_RECV(2, 16);
    } else {
        if (nchunks <= 4) CALL_RECV(1, 4);
        else if (nchunks <= 8) CALL_RECV(1, 8);
        else if (nchunks <= 16) CALL_RECV(1, 16);
        else CALL_RECV(1, 32);
    }
#undef CALL_RECV
    return adaqp_check_launch("recv_quant_kernel");
}

int adaqp_send_fp32(const float *x, int64_t ld, int32_t F, const adaqp_fp_item *items,
                    int64_t n_items, const adaqp_send_chan *chans, int32_t n_chans, int64_t dst_ld,
                    uint32_t seq, uint32_t *work, uint32_t *status, uint64_t timeout_ns,
                    void *stream) {
    ADAQP_REQUIRE(F > 0, ADAQP_EINVAL, "adaqp_send_fp32: F=%d", F);
    ADAQP_REQUIRE(n_items >= 0 && n_chans >= 0, ADAQP_EINVAL, "adaqp_send_fp32: negative count");
    ADAQP_REQUIRE(work != nullptr, ADAQP_EINVAL, "adaqp_send_fp32: null work");
    if (n_chans == 0) return 0;
    ADAQP_REQUIRE(chans && (n_items == 0 || (x && items)), ADAQP_EINVAL, "adaqp_send_fp32: null pointer");
    // destination rows live in slabs allocated 256-byte aligned: alignment follows dst_ld
    int vec = pick_vec(F, ld, x);
    while (vec > 1 && (dst_ld % vec)) vec >>= 1;
    int grid = grid_for(n_items, 8);
    if (adaqp_options().exch_send_ctas > 0 && grid > adaqp_options().exch_send_ctas) grid = adaqp_options().exch_send_ctas;
    cudaStream_t s = (cudaStream_t)stream;
    if (vec == 4)
        send_fp32_kernel<4><<<grid, kThreads, 0, s>>>(x, ld, F, items, n_items, chans, n_chans, dst_ld, seq, work, status, timeout_ns);
    else if (vec == 2)
        send_fp32_kernel<2><<<grid, kThreads, 0, s>>>(x, ld, F, items, n_items, chans, n_chans, dst_ld, seq, work, status, timeout_ns);
    else
        send_fp32_kernel<1><<<grid, kThreads, 0, s>>>(x, ld, F, items, n_items, chans, n_chans, dst_ld, seq, work, status, timeout_ns);
    return adaqp_check_launch("send_fp32_kernel");
}

int adaqp_wait_flags(const uint32_t *const *flags, int32_t n, uint32_t seq, uint32_t *status,
                     uint64_t timeout_ns, void *stream) {
    ADAQP_REQUIRE(n >= 0 && n <= 1024, ADAQP_ELIMIT, "adaqp_wait_flags: n=%d", n);
    if (n == 0) return 0;
    ADAQP_REQUIRE(flags != nullptr, ADAQP_EINVAL, "adaqp_wait_flags: null flags");
    wait_flags_kernel<<<1, ((n + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(flags, n, seq, status, timeout_ns);
    return adaqp_check_launch("wait_flags_kernel");
}

int adaqp_post_acks(uint32_t *const *acks, int32_t n, uint32_t seq, void *stream) {
    ADAQP_REQUIRE(n >= 0 && n <= 1024, ADAQP_ELIMIT, "adaqp_post_acks: n=%d", n);
    if (n == 0) return 0;
    ADAQP_REQUIRE(acks != nullptr, ADAQP_EINVAL, "adaqp_post_acks: null acks");
    post_acks_kernel<<<1, ((n + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(acks, n, seq);
    return adaqp_check_launch("post_acks_kernel");
}

int adaqp_gather_rows_f32(const float *x, int64_t ld, const int64_t *idx, int64_t n, int32_t F,
                          float *out, int64_t ldo, void *stream) {
    ADAQP_REQUIRE(F > 0 && n >= 0, ADAQP_EINVAL, "adaqp_gather_rows_f32: bad shape");
    if (n == 0) return 0;
    ADAQP_REQUIRE(x && idx && out, ADAQP_EINVAL, "adaqp_gather_rows_f32: null pointer");
    const int vec = pick_vec(F, ld, x, out, ldo);
    const int grid = grid_for(n, 8);
    cudaStream_t s = (cudaStream_t)stream;
    if (vec == 4) gather_rows_kernel<4><<<grid, kThreads, 0, s>>>(x, ld, idx, n, F, out, ldo);
    else if (vec == 2) gather_rows_kernel<2><<<grid, kThreads, 0, s>>>(x, ld, idx, n, F, out, ldo);
    else gather_rows_kernel<1><<<grid, kThreads, 0, s>>>(x, ld, idx, n, F, out, ldo);
    return adaqp_check_launch("gather_rows_kernel");
}

}  // extern "C"
