/*
 * adaqp_b200.h -- C ABI of libadaqp_b200.so (sm_100a).
 *
 * Drop-in boundary for AdaQP's per-layer boundary-message exchange + local
 * aggregation path.  The reference has no C/FFI plugin interface: its native
 * boundary is the pybind11 torch extension `quant_cuda` plus Python methods on
 * Communicator / CommBuffer.  Every entry point below names the
 * reference interface it replaces (paths relative to the reference tree) and
 * takes plain pointers, sizes and a cudaStream_t passed as void*; no torch
 * types.  The Python side (adaqp_b200/_lib.py, ctypes) is the binding a
 * maintainer would add -- see INTEGRATION.md.
 *
 * Conventions: all pointers are DEVICE pointers unless the name says host;
 * every function returns 0 on success, a positive cudaError_t on a CUDA
 * failure, or a negative ADAQP_E* code on bad arguments.  Launches are
 * asynchronous on `stream` (NULL = legacy default stream).
 * adaqp_last_error() returns a thread-local message for the last failure.
 */
#ifndef ADAQP_B200_H
#define ADAQP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADAQP_ABI_VERSION 2

#define ADAQP_EINVAL (-1)   /* bad argument (bits not in {1,2,4,8}, negative size ...) */
#define ADAQP_EALIGN (-2)   /* pointer alignment requirement violated */
#define ADAQP_ELIMIT (-3)   /* size outside what the kernels support */

/* status words written by the exchange kernels into `status` (device uint32[4]):
 * status[0] != 0 -> a flag/ack wait timed out (value = ADAQP_ST_*), status[1] = slot */
#define ADAQP_ST_OK 0u
#define ADAQP_ST_FLAG_TIMEOUT 1u
#define ADAQP_ST_ACK_TIMEOUT 2u

/* ---------------------------------------------------------------- runtime */
int adaqp_abi_version(void);
const char *adaqp_last_error(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int adaqp_sm_count(void);
/* Process-wide tunables (the library never reads the environment; adaqp_b200/_lib.py maps ADAQP_*
 * variables onto these once at load).  Names: "spmm_impl" (1 register gather [default], 2 cp.async
 * ring, 3 TMA tile::gather4 ring, 4 TMA bulk-per-row ring), "spmm_rows_per_grab" (0 = default),
 * "spmm_ctas_per_sm", "spmm_hints" (bit 0 streaming output stores, bit 1 streaming index loads),
 * "gemm_block_k" (32 or 16: K block / swizzle width of the dense GEMM),
 * "exch_send_ctas" / "exch_recv_ctas" (0 = one resident wave over all SMs, n = at most n CTAs, so
 * that the exchange kernels leave the remaining SMs to the overlapped aggregation,
 * SURVEY.md 7 step 7 -- the reference instead serialises them, ops.py:119-130). */
int adaqp_set_option(const char *name, int64_t value);
int adaqp_get_option(const char *name, int64_t *value);

/* ------------------------------------------------------------ single codec
 * Replaces quant_cuda.pack_single_precision / unpack_single_precision
 * (AdaQP/util/quantization/src/quantization.cc:23-45,
 *  quantization_cuda_kernel.cu:34-52,106-122), fp32 instantiation.
 *
 * pack: packed[k = no*F + d] |= q(n = no*wpt + ni, d) << (ni*bits),
 *   q = rn(max(fma(data[n,d] - min[n], scale[n], U) - 0.5, 0)),
 *   U = curand_uniform of Philox4_32_10(seed, subsequence = k, offset), draw ni.
 * Writes exactly adaqp_packed_nbytes(N,F,bits) bytes.  The reference returns a
 * tensor one byte longer (adaqp_qsize) whose last byte is never written; the
 * host mirror allocates that size.  (seed, offset) are the values ATen's
 * philox_engine_inputs(F*8/bits) would hand the reference kernel; advancing
 * the generator is the host mirror's job (adaqp_b200/quant.py).
 */
int64_t adaqp_packed_nbytes(int64_t N, int64_t F, int bits);
int64_t adaqp_qsize(int64_t N, int64_t F, int bits);
int adaqp_pack_f32(const float *data, const float *min, const float *scale,
                   int64_t N, int64_t F, int bits, uint64_t seed, uint64_t offset,
                   uint8_t *packed, void *stream);
/* out[n,d] = float((packed[no*F+d] >> ni*bits) & mask) / scale[n] + min[n] */
int adaqp_unpack_f32(const uint8_t *packed, const float *scale, const float *min,
                     int64_t N, int64_t F, int bits, float *out, void *stream);

/* fp16 instantiation (the reference dispatches AT_DISPATCH_FLOATING_TYPES_AND_HALF, quantization_cuda_kernel.cu:81,138,
 * and check.h:22-27 admits float16): same layout and generator protocol, c10::Half arithmetic -- every Half (op) Half is
 * evaluated in float and rounded back to half.  data / min / scale / out are IEEE half arrays.  Not used by the hot
 * path (boundary messages are fp32); provided for API parity. */
int adaqp_pack_f16(const void *data, const void *min, const void *scale, int64_t N, int64_t F, int bits,
                   uint64_t seed, uint64_t offset, uint8_t *packed, void *stream);
int adaqp_unpack_f16(const uint8_t *packed, const void *scale, const void *min, int64_t N, int64_t F, int bits,
                     void *out, void *stream);

/* Row min / max / scale = (2^bits-1)/(max-min) in one pass
 * (AdaQP/model/op_util.py:20-22,41).  Any of rmin/rmax/scale may be NULL. */
int adaqp_row_minmax_f32(const float *data, int64_t N, int64_t F, int bits,
                         float *rmin, float *rmax, float *scale, void *stream);

/* ------------------------------------------------------------- P2P slabs
 * Replace the pinned-host staging buffers + gloo isend/irecv of
 * AdaQP/communicator/buffer.py:154-248 and comm.py:166-222: each rank owns one
 * device slab holding its receive regions, flags and acks; peers map it with
 * CUDA IPC and write into it directly over NVLink.  Handles travel through the
 * (gloo) control plane as 64 opaque bytes.
 */
#define ADAQP_IPC_HANDLE_BYTES 64
int adaqp_slab_alloc(void **ptr, size_t bytes);           /* cudaMalloc + zero fill */
int adaqp_slab_free(void *ptr);
int adaqp_ipc_export(void *ptr, unsigned char handle[ADAQP_IPC_HANDLE_BYTES]);
int adaqp_ipc_open(const unsigned char handle[ADAQP_IPC_HANDLE_BYTES], void **ptr);
int adaqp_ipc_close(void *ptr);
/* 1 if the current device can map `peer_device` memory */
int adaqp_can_access_peer(int peer_device);
/* cudaDeviceEnablePeerAccess(peer_device) for the current device (idempotent): lets ONE process
 * that drives several GPUs address their slabs directly (tools/bench_exchange.py --devices, the
 * single-process harness ncu can profile; multi-process runs map slabs with the IPC calls above). */
int adaqp_enable_peer_access(int peer_device);

/* --------------------------------------------------------- fused exchange
 * One channel = (this rank, one peer) for one layer key.
 */
typedef struct adaqp_send_chan {
    uint8_t *qdata;        /* peer-mapped: region receiving this rank's packed bytes   */
    uint16_t *params;      /* peer-mapped: bf16 [2, S] (row 0 scale, row 1 min)        */
    float *fp_rows;        /* peer-mapped: fp32 rows (fp32 exchange), else NULL        */
    uint32_t *flag;        /* peer-mapped: written with `seq` when the data is visible */
    const uint32_t *ack;   /* local: peer writes seq here after consuming              */
    int64_t S;             /* rows on this channel (params row stride)                 */
} adaqp_send_chan;

typedef struct adaqp_recv_chan {
    const uint8_t *qdata;  /* local region the peer writes                              */
    const uint16_t *params;
    const uint32_t *flag;  /* local flag the peer sets                                  */
    uint32_t *ack;         /* peer-mapped ack word                                      */
    int64_t S;
} adaqp_recv_chan;

/* One work item = one byte-row of a segment: 8/bits consecutive rows of one
 * (peer, bit-width) segment that share packed bytes. 64 bytes. */
typedef struct adaqp_send_item {
    int32_t src_row[4];    /* rows of the local message matrix (-1 = past the end)      */
    int32_t send_pos[4];   /* position in send_messages (total_send_idx order), tracing */
    int64_t dst_off;       /* byte offset of the item's F bytes inside chan.qdata       */
    int32_t param_pos;     /* column of the item's first row in chan.params             */
    int32_t group;         /* byte-row index inside its segment (Philox subsequence/F)  */
    uint32_t rel_offset;   /* Philox offset of the segment's pack call minus the base   */
    int16_t chan;          /* index into the channel table                              */
    int8_t bits;           /* 2, 4 or 8 (BITS_SET, buffer.py:20)                        */
    int8_t nrows;          /* valid rows (1..8/bits)                                    */
    int32_t _pad;
} adaqp_send_item;

typedef struct adaqp_recv_item {
    int32_t dst_row[4];    /* rows of the halo matrix to write (-1 = none)              */
    int64_t src_off;       /* byte offset of the item's F bytes inside chan.qdata       */
    int32_t param_pos;
    int16_t chan;
    int8_t bits;
    int8_t nrows;
} adaqp_recv_item;

/* fp32 exchange: one item per row sent. 16 bytes. */
typedef struct adaqp_fp_item {
    int32_t src_row;       /* row of the local message matrix                           */
    int32_t chan;
    int64_t dst_row;       /* row inside chan.fp_rows                                   */
} adaqp_fp_item;

/* Fused gather -> row min/max -> stochastic quantize -> bit-pack -> store into the
 * peers' slabs (+ bf16 params) -> publish flag.  Replaces, per layer key,
 * local_messages[total_send_idx] (ops.py:134,164), mixed_msg_quantization
 * (op_util.py:189-209), the D2H staging + isend half of qt_msg_exchange
 * (comm.py:193-222) and the tracing reductions of trace_input (op_util.py:91-99).
 * Byte-for-byte the reference wire format.
 *   x[n_rows, F] fp32 row-major (ld = row stride in floats);
 *   trace: optional [S_total] fp32 accumulator, trace[pos] += (F/6)(max-min)^2;
 *   (seed, base_offset): generator state the first pack call would have seen;
 *   seq: per-key exchange sequence number (>=1), written to every chan.flag;
 *   work: device uint32[2] scratch, zero on first use (kernel leaves it zero);
 *   status: device uint32[4], see ADAQP_ST_*; timeout_ns bounds every spin. */
int adaqp_send_quant(const float *x, int64_t ld, int32_t F,
                     const adaqp_send_item *items, int64_t n_items,
                     const adaqp_send_chan *chans, int32_t n_chans,
                     float *trace, uint64_t seed, uint64_t base_offset, uint32_t seq,
                     uint32_t *work, uint32_t *status, uint64_t timeout_ns, void *stream);

/* Wait flags -> unpack -> dequantize with the bf16 params -> scatter into the halo
 * matrix -> ack.  Replaces the irecv/H2D half of qt_msg_exchange and
 * mixed_msg_dequantization (op_util.py:211-236).  halo[n_remote, F], ld in floats. */
int adaqp_recv_quant(float *halo, int64_t ld, int32_t F,
                     const adaqp_recv_item *items, int64_t n_items,
                     const adaqp_recv_chan *chans, int32_t n_chans,
                     uint32_t seq, uint32_t *work, uint32_t *status,
                     uint64_t timeout_ns, void *stream);

/* fp32 exchange (Vanilla / AdaQP-p training and every eval forward): gather rows and
 * store them straight into the peers' halo rows, then publish flags.  Replaces
 * fp_msg_exchange (comm.py:166-191) + the scatter of fp_msg_transfer_process
 * (op_util.py:168-170): the sender already knows each row's halo position. */
int adaqp_send_fp32(const float *x, int64_t ld, int32_t F,
                    const adaqp_fp_item *items, int64_t n_items,
                    const adaqp_send_chan *chans, int32_t n_chans, int64_t dst_ld,
                    uint32_t seq, uint32_t *work, uint32_t *status,
                    uint64_t timeout_ns, void *stream);

/* Spin until every flags[i] >= seq (i < n).  Receiver half of the fp32 exchange. */
int adaqp_wait_flags(const uint32_t *const *flags, int32_t n, uint32_t seq,
                     uint32_t *status, uint64_t timeout_ns, void *stream);
/* Store seq into every acks[i]; launched after the consumer of a halo buffer. */
int adaqp_post_acks(uint32_t *const *acks, int32_t n, uint32_t seq, void *stream);

/* ------------------------------------------------------------ aggregation
 * Normalised CSR SpMM replacing DGL update_all(copy_src, sum|mean) plus the two
 * elementwise norm multiplies and the torch.cat of local and halo rows
 * (AdaQP/model/ops.py:17-67,137-147,169-185):
 *   out[v - row_begin] = post[v] * ( sum_{j in [indptr[v], indptr[v+1])} pre[u_j] x[u_j]
 *                                    (+ pre[v] x[v] if add_self) )      v in [row_begin,row_end)
 *   x[u] = u < n_split ? x0[u*ld0 ..] : x1[(u-n_split)*ld1 ..]   (local rows | halo rows)
 *   mean != 0: divide the sum by (indptr[v+1]-indptr[v]) (0-degree rows give 0).
 * pre (per source id, may be NULL) and post (per destination id, may be NULL) are fp32.
 * indices are int32 source ids; indptr is int64.  F <= 1024 (as the reference's codec). */
int adaqp_spmm_csr_f32(const int64_t *indptr, const int32_t *indices,
                       const float *x0, int64_t ld0, int64_t n_split,
                       const float *x1, int64_t ld1,
                       const float *pre, const float *post, int mean, int add_self,
                       int64_t row_begin, int64_t row_end, int32_t F,
                       float *out, int64_t ldo, void *stream);

/* Same aggregation over a per-row neighbour SEGMENT [seg_start[v], seg_end[v]) of the CSR row
 * (NULL = the row's own bounds), optionally accumulating into `out` (out += ...).  With the
 * columns of a row sorted, [indptr[v], split[v]) are the local sources and [split[v],
 * indptr[v+1]) the halo sources: the local part of the marginal rows can then run while the
 * exchange is still in flight and only the halo part waits for it (a finer overlap than the
 * reference's central / marginal split, ops.py:156-193).  `mean` still divides by the full
 * in-degree indptr[v+1] - indptr[v]; add_self belongs to exactly one of the two calls. */
int adaqp_spmm_csr_seg_f32(const int64_t *indptr, const int64_t *seg_start, const int64_t *seg_end,
                           const int32_t *indices, const float *x0, int64_t ld0, int64_t n_split,
                           const float *x1, int64_t ld1, const float *pre, const float *post,
                           int mean, int add_self, int accumulate, int64_t row_begin,
                           int64_t row_end, int32_t F, float *out, int64_t ldo, void *stream);

/* --------------------------------------------------------------- dense GEMM
 * C[M, N] = A[M, K] . Bt[N, K]^T (+ bias[N]) in fp32 on the tcgen05 tensor cores by 3xTF32 error-compensated
 * splitting (csrc/gemm.cu): replaces torch.matmul(rst, self.weight) (AdaQP/model/distGCN.py:45) and the Linear
 * layers of distSAGE.py:51-53 for tall-skinny shapes (M = inner nodes, N <= 256).  Bt_hi / Bt_lo are the
 * transposed weight split as b_hi = b & 0xFFFFE000, b_lo = b - b_hi (host mirror: adaqp_b200/dense.py).
 * lda / ldb multiples of 4 floats, 16-byte aligned bases; adaqp_gemm_tf32x3_supported tells whether a shape
 * qualifies (otherwise the caller keeps torch.matmul). */
int adaqp_gemm_tf32x3_supported(int64_t M, int32_t N, int32_t K, int64_t lda, int64_t ldb, int64_t ldc);
int adaqp_gemm_tf32x3_f32(const float *A, int64_t lda, const float *Bt_hi, const float *Bt_lo, int64_t ldb,
                          const float *bias, int64_t M, int32_t N, int32_t K, float *C, int64_t ldc, void *stream);

/* Weight gradient dW[N, K] = dY[M, N]^T . X[M, K] (the `X^T . dY` of the layers' backward), split over the M node rows:
 * CTA c writes its partial sum to partials[c] ([grid, N, K] fp32, grid = adaqp_wgrad_tf32x3_grid(M)); the caller adds the
 * slices.  Same 3xTF32 arithmetic; both operands are read MN-major straight from their row-major storage. */
int adaqp_wgrad_tf32x3_supported(int64_t M, int32_t N, int32_t K, int64_t ldy, int64_t ldx);
int adaqp_wgrad_tf32x3_grid(int64_t M);
int adaqp_wgrad_tf32x3_f32(const float *dY, int64_t ldy, const float *X, int64_t ldx, int64_t M, int32_t N, int32_t K,
                           float *partials, int32_t grid, void *stream);

/* ------------------------------------------------------ LayerNorm + ReLU
 * y = relu(LayerNorm(x) * gamma + beta) over the last dimension (biased variance, eps as nn.LayerNorm), forward and
 * backward in one pass each: the `self.norms[i](feats)` + `F.relu` between aggregations (AdaQP/model/distGCN.py:81-84,
 * distSAGE.py:93-96).  F % 4 == 0, F <= 1024, 16-byte aligned rows.  The backward writes per-CTA partial column sums
 * partials[grid][2][F] (dgamma, dbeta; grid = adaqp_ln_relu_grid(M)) that the caller adds.  Dropout stays torch's. */
int adaqp_ln_relu_grid(int64_t M);
int adaqp_ln_relu_fwd_f32(const float *x, int64_t ldx, const float *gamma, const float *beta, float eps, int64_t M,
                          int32_t F, float *y, int64_t ldy, float *mean, float *rstd, void *stream);
int adaqp_ln_relu_bwd_f32(const float *dy, int64_t lddy, const float *x, int64_t ldx, const float *mean,
                          const float *rstd, const float *gamma, const float *beta, int64_t M, int32_t F, float *dx,
                          int64_t lddx, float *partials, int32_t grid, void *stream);

/* Row gather out[i] = x[idx[i]] (copy-buffer fills of ops.py:159-164; API parity only). */
int adaqp_gather_rows_f32(const float *x, int64_t ld, const int64_t *idx, int64_t n,
                          int32_t F, float *out, int64_t ldo, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ADAQP_B200_H */
