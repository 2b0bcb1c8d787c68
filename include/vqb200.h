/*
 * vqb200.h -- C ABI of the B200-native engine for the vq hot path.
 *
 * This is the drop-in boundary: exactly the entry points a host-language shim
 * (Rust `extern "C"` + build.rs, PyO3, ctypes, C++) binds to replace the
 * reference's CPU implementation of
 *     k-means/LBG codebook training + nearest-centroid encoding behind
 *     ProductQuantizer and TSVQ, and the BQ/SQ element-wise codecs.
 * Plain pointers and sizes only; no C++/torch types.  Every entry point cites the
 * reference interface it replaces (paths relative to the vq repository, commit
 * d54c906).  The conventions follow the reference's only native boundary,
 * hsdlib (src/core/hsdlib_ffi.rs:10-35,68-83; external/hsdlib/include/hsdlib.h:32-38):
 *   - every function returns an int status, 0 == success;
 *   - results go through out-pointers, the caller owns every buffer;
 *   - strings returned by the library are static or owned by the context.
 *
 * Pointer mode.  Every data pointer may be a HOST pointer (pageable or pinned) or a
 * DEVICE pointer of the context's GPU; the library classifies it with
 * cudaPointerGetAttributes.  Calls whose data pointers are all device pointers are
 * enqueued on the context stream and return without synchronising (use
 * vqb_ctx_synchronize); calls that touch host memory are complete on return.  Host-pointer
 * batch calls are pipelined in row chunks over three streams through grow-only device staging
 * buffers owned by the context (no allocation per call after the first); calls on one context
 * are serialised by its mutex, so several threads may share a context.
 *
 * There is NO CPU fallback: every function fails with VQB_ERR_UNSUPPORTED_DEVICE
 * when no sm_100 GPU is present.
 */
#ifndef VQB200_H
#define VQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes: hsdlib's (hsdlib.h:32-38) plus the two VqError kinds that
 *      the shim must be able to distinguish (src/core/error.rs:5-28) ---------- */
#define VQB_SUCCESS                 0
#define VQB_ERR_NULL_PTR           (-1)   /* HSD_ERR_NULL_PTR                               */
#define VQB_ERR_EMPTY_INPUT        (-2)   /* VqError::EmptyInput                            */
#define VQB_ERR_INVALID_INPUT      (-3)   /* HSD_ERR_INVALID_INPUT / VqError::InvalidParameter */
#define VQB_ERR_UNSUPPORTED_DEVICE (-4)   /* HSD_ERR_CPU_NOT_SUPPORTED: no sm_100 device     */
#define VQB_ERR_DIM_MISMATCH       (-5)   /* VqError::DimensionMismatch                     */
#define VQB_FAILURE                (-99)  /* HSD_FAILURE: CUDA error, see vqb_last_error     */

/* ---- Distance, enum order of src/core/distance.rs:8-17 ---------------------- */
#define VQB_SQUARED_EUCLIDEAN 0
#define VQB_EUCLIDEAN         1
#define VQB_MANHATTAN         2
#define VQB_COSINE            3
/* EXTENSION, not part of the reference (its Distance has exactly the four kinds above, src/core/distance.rs:8-17):
 * Chebyshev distance max_i |a_i - b_i| (components whose difference is NaN are skipped).  Accepted by
 * vqb_distance_batch and by vqb_pq_create / vqb_pq_encode (CUDA-core assignment kernel); training is squared-L2 whatever
 * the metric, as in the reference (src/pq.rs:121-132).  There is nothing in the reference to be parity-checked against:
 * the tests compare with oracle/vq_oracle.c's restatement of this definition only. */
#define VQB_CHEBYSHEV         5

typedef struct vqb_ctx vqb_ctx;   /* one GPU, its streams and scratch memory          */
typedef struct vqb_pq vqb_pq;     /* immutable trained ProductQuantizer (pq.rs:39-45) */
typedef struct vqb_tsvq vqb_tsvq; /* immutable trained TSVQ tree (tsvq.rs:13-17,159-163) */

/* ======================= context / memory ==================================== */

/* Creates the engine on CUDA device `device`.  Fails with
 * VQB_ERR_UNSUPPORTED_DEVICE unless the device is compute capability 10.x. */
int vqb_ctx_create(int device, vqb_ctx** out);
int vqb_ctx_destroy(vqb_ctx* ctx);
int vqb_ctx_synchronize(vqb_ctx* ctx);
/* The cudaStream_t all work of this context is enqueued on (for event timing). */
void* vqb_ctx_stream(vqb_ctx* ctx);
/* Adopt a caller-owned cudaStream_t (e.g. the host framework's current stream). */
int vqb_ctx_set_stream(vqb_ctx* ctx, void* cuda_stream);
/* Message of the last failure on this context (owned by the context). */
const char* vqb_last_error(vqb_ctx* ctx);
/* Replaces hsd_get_backend() / get_simd_backend() (src/core/hsdlib_ffi.rs:144-155). */
const char* vqb_backend_name(void);
/* Number of kernels this context has launched so far (bench.py `gpu_launches`). */
uint64_t vqb_ctx_launch_count(vqb_ctx* ctx);

/* Device / pinned-host memory for hosts that do not link CUDA themselves. */
int vqb_malloc(vqb_ctx* ctx, size_t bytes, void** dptr);
int vqb_free(vqb_ctx* ctx, void* dptr);
int vqb_host_alloc(vqb_ctx* ctx, size_t bytes, void** hptr);   /* pinned */
int vqb_host_free(vqb_ctx* ctx, void* hptr);
int vqb_memcpy(vqb_ctx* ctx, void* dst, const void* src, size_t bytes); /* any direction, blocking */

/* ======================= multi-GPU: one process per GPU ======================
 * PQ training shards the rows over the GPUs of a box and exchanges ONE fused buffer
 * [sums | count_lo | count_hi] per k-means iteration (src/core/vector.rs:432-447 over row shards); encode, BQ, SQ and TSVQ
 * encode shard rows with no collective.  The library owns the NCCL communicator (libnccl.so.2 is opened at run time, so
 * a single-GPU host needs no NCCL): rank 0 calls vqb_comm_unique_id, ships the VQB_COMM_ID_BYTES to the other ranks by
 * any means (MPI, a file, torch.distributed), every rank calls vqb_comm_init_rank (collective), then trains with
 * vqb_train_opts.flags = VQB_TRAIN_USE_COMM, row_offset and n_global. */
#define VQB_COMM_ID_BYTES 128
int vqb_comm_unique_id(void* id_out /* VQB_COMM_ID_BYTES */);
int vqb_comm_init_rank(vqb_ctx* ctx, const void* id, int rank, int world);
int vqb_comm_destroy(vqb_ctx* ctx);
int vqb_comm_info(vqb_ctx* ctx, int* rank, int* world);
/* In-place float sum of a device buffer over the communicator's ranks, enqueued on the context stream. */
int vqb_comm_allreduce(vqb_ctx* ctx, float* buf, size_t count);

/* ======================= Distance ============================================ */

/* Distance::compute (src/core/distance.rs:48-65) for `rows` independent pairs:
 * out[r] = metric(a[r*n .. r*n+n), b[r*n .. r*n+n)).  Bit-exact with the `simd`
 * build on an AVX-512 host (hsdlib euclidean.c:131-163, manhattan.c:132-163,
 * cosine.c:163-198 + 28-63, incl. the scalar fallbacks of distance.rs:75-83,93-95,106-119). */
int vqb_distance_batch(vqb_ctx* ctx, int metric, const float* a, const float* b,
                       size_t rows, size_t n, float* out);

/* ======================= BinaryQuantizer / ScalarQuantizer =================== */

/* BinaryQuantizer::quantize (src/bq.rs:94-105): out[i] = x[i] >= thr ? high : low. */
int vqb_bq_quantize(vqb_ctx* ctx, const float* x, size_t n, float threshold,
                    uint8_t low, uint8_t high, uint8_t* out);
/* BinaryQuantizer::dequantize (src/bq.rs:107-118): out[i] = c[i] >= high ? high : low as f32. */
int vqb_bq_dequantize(vqb_ctx* ctx, const uint8_t* codes, size_t n, uint8_t low, uint8_t high,
                      float* out);
/* ScalarQuantizer::quantize (src/sq.rs:123-144); `step` as computed by
 * ScalarQuantizer::new (src/sq.rs:94): (max - min) / (levels - 1) as f32. */
int vqb_sq_quantize(vqb_ctx* ctx, const float* x, size_t n, float min, float max, float step,
                    uint32_t levels, uint8_t* out);
/* ScalarQuantizer::dequantize (src/sq.rs:146-151): out[i] = min + codes[i] as f32 * step. */
int vqb_sq_dequantize(vqb_ctx* ctx, const uint8_t* codes, size_t n, float min, float step,
                      float* out);
/* ProductQuantizer::dequantize / TSVQ::dequantize (src/pq.rs:201-209, src/tsvq.rs:257-265):
 * f16 (IEEE binary16 bit patterns) -> f32. */
int vqb_f16_dequantize(vqb_ctx* ctx, const uint16_t* q, size_t n, float* out);

/* ======================= ProductQuantizer ==================================== */

/* Stands in for `data.choose(&mut rng)` (src/core/vector.rs:450): returns the (global)
 * row that re-seeds the next empty cluster of `subspace`.  Called on the host, in
 * ascending cluster order within an iteration, exactly as the reference consumes its
 * per-subspace StdRng (seed + subspace, src/pq.rs:130). */
typedef uint64_t (*vqb_reseed_fn)(void* user, uint32_t subspace);

/* In-place sum over all ranks of `count` floats at device pointer `buf`, ordered after
 * prior work on `cuda_stream`.  Supplied by a multi-GPU host (one process per GPU; e.g.
 * ncclAllReduce or torch.distributed.all_reduce); NULL for a single GPU. */
typedef int (*vqb_allreduce_fn)(void* user, float* buf, size_t count, void* cuda_stream);

#define VQB_UPDATE_ORDERED 0 /* per-cluster sums in ascending row order == vector.rs:368-384, bit-exact */
#define VQB_UPDATE_FAST    1 /* fixed-shape partial sums (one pass over X, no second copy): deterministic, not the reference's order */
#define VQB_ASSIGN_AUTO    0
#define VQB_ASSIGN_EXACT   1 /* CUDA-core kernel evaluating the reference's formula for every centroid */
#define VQB_ASSIGN_TENSOR  2 /* tcgen05 GEMM-form scores + exact re-check of the candidates: sub_dim 8, 16, 24 or 32, k <= 256,
                                 squared L2 / L2 / cosine (and training), 16-byte aligned rows; VQB_ERR_INVALID_INPUT otherwise.
                                 AUTO takes it from 1024 rows on, and a warp-per-pair kernel for up to 64 rows */

#define VQB_TRAIN_USE_COMM 1u /* flags: rows are sharded over the ranks of the context's communicator (vqb_comm_init_rank);
                                 the per-iteration exchange is one ncclAllReduce issued by the library on the context stream */

typedef struct vqb_train_opts {
    uint32_t struct_size;        /* sizeof(vqb_train_opts) */
    uint32_t update_mode;        /* VQB_UPDATE_*  */
    uint32_t assign_mode;        /* VQB_ASSIGN_*  */
    uint32_t flags;              /* VQB_TRAIN_*  */
    vqb_reseed_fn reseed;        /* may be NULL: empty clusters then keep their centroid */
    void* reseed_user;
    vqb_allreduce_fn allreduce;  /* NULL: single GPU, or the context's own communicator with VQB_TRAIN_USE_COMM */
    void* allreduce_user;
    uint64_t row_offset;         /* global id of this rank's first row (0 on a single GPU) */
    uint64_t n_global;           /* total rows over all ranks (0: == n) */
    float* iter_ms;              /* diagnostics, may be NULL: host array of max_iters floats; iter_ms[t] receives the
                                    device time of iteration t in ms (CUDA events on the context stream, the
                                    all-reduce included), entries of iterations that did not run are left alone */
} vqb_train_opts;

/* ProductQuantizer::new (src/pq.rs:83-141) == m x lbg_quantize (src/core/vector.rs:390-461).
 *   x          n x dim row-major f32 (this rank's rows)
 *   init_idx   m*k global row ids; stands in for choose_multiple (vector.rs:413)
 *   codebooks  out, m*k*(dim/m) f32, host or device
 *   iters_run  out (host, may be NULL), loop bodies executed per subspace
 * Validation order and error kinds follow pq.rs:91-117 and vector.rs:396-410
 * (empty -> dim<m -> dim%m -> k==0 -> n<k).  Training always uses squared L2
 * (vector.rs:352-363), whatever metric is later used to encode. */
int vqb_pq_train(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k,
                 size_t max_iters, const uint64_t* init_idx, const vqb_train_opts* opts,
                 float* codebooks, uint32_t* iters_run);

/* One assignment step of training, exposed for teacher-forced parity tests:
 * codes[s*n + i] = argmin_j ||x_i^(s) - c_j^(s)||^2 (vector.rs:352-363), u32. */
int vqb_pq_assign_train(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k,
                        const float* codebooks, uint32_t assign_mode, uint32_t* codes_out);

/* One full iteration (assign + update + epsilon test) from given centroids, in place.
 * changed_out[m] (host), counts_out[m*k] (host, may be NULL). */
int vqb_pq_train_step(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k,
                      float* codebooks_inout, const vqb_train_opts* opts,
                      uint32_t* changed_out, uint32_t* counts_out);

/* Wraps trained codebooks (m*k*sub_dim f32, host or device) for encoding. */
int vqb_pq_create(vqb_ctx* ctx, const float* codebooks, size_t m, size_t k, size_t sub_dim,
                  int metric, vqb_pq** out);
int vqb_pq_destroy(vqb_pq* pq);
int vqb_pq_codebooks(vqb_pq* pq, float* out /* m*k*sub_dim, host or device */);

/* ProductQuantizer::quantize (src/pq.rs:167-199) over a batch of n vectors.
 *   codes_out  n*m code indices of `code_bytes` (1, 2 or 4) bytes each, or NULL
 *   recon_out  n*dim f16 bit patterns == the reference's Vec<f16> output, or NULL
 * Distances follow the `simd` build on an AVX-512 host (see vqb_distance_batch); the
 * arg-min is the reference's: strict '<', lowest index wins, NaN never wins. */
int vqb_pq_encode(vqb_pq* pq, const float* x, size_t n, uint32_t assign_mode,
                  void* codes_out, uint32_t code_bytes, uint16_t* recon_out);
/* Reconstruction from codes (batch form of pq.rs:193-195 + 201-209): out n*dim f32. */
int vqb_pq_decode(vqb_pq* pq, const void* codes, uint32_t code_bytes, size_t n, float* out);

/* Diagnostics for the tensor-core assignment kernel (evidence for its error margin):
 * runs one training-metric (cosine == 0) or cosine-encode (cosine != 0) assignment pass and returns
 * the raw tcgen05 scores of subspace `sub` (n*256 f32, may be NULL), the number of (row, subspace)
 * pairs that fell inside the margin and were re-scanned exactly (may be NULL), and the codes
 * [m][n] u32 (may be NULL).  Not used by any quantizer call. */
int vqb_debug_tc_scores(vqb_ctx* ctx, int cosine, const float* x, size_t n, size_t dim, size_t m, size_t k,
                        const float* codebooks, int sub, float* scores_out, uint64_t* rescans_out,
                        uint32_t* codes_out);

/* Diagnostics: SM-clock stamps of one CTA's per-unit hand-offs during a cosine-encode pass of the tensor-core kernel
 * (ts_out[units][8], host): 0 MMA issuer saw the accumulator free, 1 MMAs issued, 2 scan saw the accumulator full,
 * 3 scan released it, 4 scan published its result, 5 resolve saw it, 6 resolve done, 7 splitter published x_lo.
 * Evidence for DESIGN.md's account of what bounds the kernel; not used by any quantizer call. */
int vqb_debug_tc_timeline(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k,
                          const float* codebooks, uint64_t* ts_out, int units);

/* Diagnostics: selects a timing variant of the tensor-core kernel (warp-role order; every variant produces the same
 * codes).  Process-wide; not used by any quantizer call. */
int vqb_debug_tc_variant(int variant);

/* ======================= TSVQ ================================================ */

/* TSVQ::new (src/tsvq.rs:195-223) == TSVQNode::build (src/tsvq.rs:31-115), level-synchronous.
 * Nodes are numbered breadth-first, root = 0. */
int vqb_tsvq_train(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t max_depth,
                   int metric, vqb_tsvq** out);
/* Wraps an existing tree (arrays as produced by vqb_tsvq_export). */
int vqb_tsvq_create(vqb_ctx* ctx, const float* centroids, const int32_t* left, const int32_t* right,
                    size_t n_nodes, size_t dim, int metric, vqb_tsvq** out);
int vqb_tsvq_destroy(vqb_tsvq* t);
int vqb_tsvq_num_nodes(vqb_tsvq* t, size_t* n_nodes, size_t* dim);
/* Any out pointer may be NULL.  centroids n_nodes*dim f32; left/right child ids or -1;
 * split_dim / median / count: the split taken at each internal node (-1 / NaN for leaves). */
int vqb_tsvq_export(vqb_tsvq* t, float* centroids, int32_t* left, int32_t* right,
                    int32_t* split_dim, float* median, uint64_t* count);
/* TSVQ::quantize (src/tsvq.rs:239-255) + find_leaf (:117-132) over a batch.
 *   leaf_out  n node ids (u32) or NULL;  recon_out n*dim f16 bit patterns or NULL. */
int vqb_tsvq_encode(vqb_tsvq* t, const float* x, size_t n, uint32_t* leaf_out, uint16_t* recon_out);

#ifdef __cplusplus
}
#endif
#endif /* VQB200_H */
