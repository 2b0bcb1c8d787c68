/*
 * vq_oracle.h -- CPU restatement of the CogitatorTech/vq hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker
 * or as the timed CPU baseline.  The product (vq_b200/) never links it.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the vq repository root, commit d54c906; hsdlib submodule 48d928b).
 *
 * Parity status: pinned against the reference's own known-answer tests
 * (tests/test_oracle_golden.py) and against hsdlib compiled verbatim from
 * /root/reference (oracle/_ref/libhsd_ref.so).  The `rand 0.9` stream that
 * picks initial/reseed rows is NOT in the reference tree (un-vendored crate),
 * so PQ training parity is conditional on an explicit index stream
 * ("parity unpinned at the rand boundary").
 */
#ifndef VQ_ORACLE_H
#define VQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Distance metric ids == enum order of src/core/distance.rs:8-17 */
enum { VQO_SQEUCLIDEAN = 0, VQO_EUCLIDEAN = 1, VQO_MANHATTAN = 2, VQO_COSINE = 3,
       VQO_CHEBYSHEV = 5 /* EXTENSION: not in the reference (distance.rs:8-17); restated definition only, parity n/a */ };

/* Which build of the reference the distance follows.
 *   SCALAR : `simd` feature off  -> Rust loops, src/core/distance.rs:75-83,93-95,106-119
 *   AVX512 : `simd` on, hsdlib resolved to its AVX-512F kernels (euclidean.c:131-163 ...)
 *   AVX2   : `simd` on, hsdlib resolved to its AVX2+FMA kernels (euclidean.c:97-129 ...)
 *   HSDLIB : `simd` on, calls the real hsdlib through pointers given to vqo_set_hsdlib()
 */
enum { VQO_SEM_SCALAR = 0, VQO_SEM_AVX512 = 1, VQO_SEM_AVX2 = 2, VQO_SEM_HSDLIB = 3 };

typedef int (*vqo_hsd_fn)(const float*, const float*, size_t, float*);
void vqo_set_hsdlib(vqo_hsd_fn sqeuclid, vqo_hsd_fn manhattan, vqo_hsd_fn cosine);

/* reseed source: returns a row index in [0,n) for the next empty cluster of `subspace`
 * (stands in for data.choose(&mut rng), src/core/vector.rs:450) */
typedef uint64_t (*vqo_reseed_fn)(void* user, uint32_t subspace);

float vqo_distance2(const float* a, const float* b, size_t n);                 /* vector.rs:135-143 */
float vqo_distance(int metric, int sem, const float* a, const float* b, size_t n); /* distance.rs:48-120 */
uint16_t vqo_f32_to_f16(float x);                                              /* half::f16::from_f32 */
float vqo_f16_to_f32(uint16_t h);

/* hsdlib restatements (return hsdlib status code, 0 == success) */
int vqo_hsd_sqeuclid(int sem, const float* a, const float* b, size_t n, float* out);
int vqo_hsd_manhattan(int sem, const float* a, const float* b, size_t n, float* out);
int vqo_hsd_cosine(int sem, const float* a, const float* b, size_t n, float* out);

/* One LBG iteration on one subspace (vector.rs:415-457 body).
 * x: n rows, row stride `ld` floats, subspace columns [col0, col0+d).
 * centroids: k x d, updated in place.  assign_out (n, may be NULL) receives the argmin.
 * empties_out (k, may be NULL) receives the ids of empty clusters in ascending order,
 * *n_empty their number; empty clusters are NOT reseeded here.
 * returns `changed` (vector.rs:438-447). */
int vqo_lbg_step(const float* x, size_t n, size_t ld, size_t col0, size_t d,
                 float* centroids, size_t k, uint32_t* assign_out,
                 uint32_t* empties_out, uint32_t* n_empty, int threads);

/* lbg_quantize for all m subspaces (pq.rs:121-132 + vector.rs:390-461).
 * init_idx: m*k row indices (stands in for choose_multiple, vector.rs:413).
 * iters_run (m) receives the number of loop bodies executed per subspace. */
int vqo_pq_train(const float* x, size_t n, size_t dim, size_t m, size_t k, size_t max_iters,
                 const uint64_t* init_idx, vqo_reseed_fn reseed, void* user,
                 float* codebooks_out, uint32_t* iters_run, int threads);

/* ProductQuantizer::quantize over a batch (pq.rs:167-199).
 * codes_out (n*m u32, may be NULL), recon_out (n*dim f16 bits, may be NULL). */
int vqo_pq_encode(const float* codebooks, size_t m, size_t k, size_t sub_dim, int metric, int sem,
                  const float* x, size_t n, uint32_t* codes_out, uint16_t* recon_out, int threads);

/* f16 -> f32 (pq.rs:201-209, tsvq.rs:257-265) */
void vqo_dequantize_f16(const uint16_t* q, size_t n, float* out);

/* BinaryQuantizer (bq.rs:94-118), ScalarQuantizer (sq.rs:94,123-151) */
void vqo_bq_quantize(const float* x, size_t n, float thr, uint8_t low, uint8_t high, uint8_t* out);
void vqo_bq_dequantize(const uint8_t* c, size_t n, uint8_t low, uint8_t high, float* out);
float vqo_sq_step(float mn, float mx, size_t levels);
void vqo_sq_quantize(const float* x, size_t n, float mn, float mx, float step, size_t levels, uint8_t* out);
void vqo_sq_dequantize(const uint8_t* c, size_t n, float mn, float step, float* out);

/* TSVQ (tsvq.rs:31-132).  Nodes are numbered breadth-first; arrays are sized for
 * max_nodes = 2^(max_depth+1)-1.  left/right hold child node ids or -1.
 * Returns the node count (>0) or a negative error. */
int64_t vqo_tsvq_build(const float* x, size_t n, size_t dim, size_t max_depth,
                       float* centroids_out, int32_t* left_out, int32_t* right_out,
                       int32_t* split_dim_out, float* median_out, uint64_t* count_out,
                       size_t max_nodes);
int vqo_tsvq_encode(const float* centroids, const int32_t* left, const int32_t* right, size_t dim,
                    int metric, int sem, const float* x, size_t n,
                    uint32_t* leaf_out, uint16_t* recon_out, int threads);

/* src/bin/common.rs:61-78 reconstruction MSE: sum (x - f16->f32(recon))^2 / (n*dim), f64 accum */
double vqo_recon_mse(const float* x, const uint16_t* recon, size_t count);

#ifdef __cplusplus
}
#endif
#endif
