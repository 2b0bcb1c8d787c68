/*
 * vq_oracle.c -- CPU restatement of the CogitatorTech/vq hot path (plain C11).
 *
 * TEST INFRASTRUCTURE ONLY (see vq_oracle.h).  Compile with
 *   gcc -O2 -std=c11 -ffp-contract=off -fopenmp -fPIC -shared
 * -ffp-contract=off matters: Rust never contracts a*b+c into an FMA, and the
 * places where hsdlib DOES use FMA are written with explicit fmaf() below.
 *
 * Paths in comments are relative to the vq repository (commit d54c906).
 */
#include "vq_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* f16 <-> f32: half::f16::from_f32 / to_f32 (half 2.4.1, Cargo.toml:42).     */
/* IEEE round-to-nearest-even, overflow -> inf, subnormals kept, NaN quieted. */
/* ------------------------------------------------------------------------- */
uint16_t vqo_f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t man = x & 0x007FFFFFu;
    int32_t exp = (int32_t)((x >> 23) & 0xFF);
    if (exp == 0xFF) { /* inf / nan */
        if (man == 0) return (uint16_t)(sign | 0x7C00u);
        return (uint16_t)(sign | 0x7C00u | 0x0200u | (man >> 13)); /* quiet NaN, keep payload top bits */
    }
    int32_t e = exp - 127 + 15;
    if (e >= 0x1F) return (uint16_t)(sign | 0x7C00u); /* overflow */
    if (e <= 0) {                                      /* subnormal half or zero */
        if (e < -10) return (uint16_t)sign;            /* too small: rounds to 0 */
        man |= 0x00800000u;                            /* implicit 1 */
        uint32_t shift = (uint32_t)(14 - e);           /* 14..24 */
        uint32_t half_man = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1u);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half_man & 1u))) half_man++;
        return (uint16_t)(sign | half_man);
    }
    uint32_t half = sign | ((uint32_t)e << 10) | (man >> 13);
    uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++; /* may carry into exp: correct */
    return (uint16_t)half;
}

float vqo_f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t x;
    if (exp == 0) {
        if (man == 0) {
            x = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3FFu;
            x = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 0x1F) {
        x = sign | 0x7F800000u | (man << 13);
    } else {
        x = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}

/* ------------------------------------------------------------------------- */
/* Vector::distance2, src/core/vector.rs:135-143: sequential fold, no FMA.    */
/* ------------------------------------------------------------------------- */
float vqo_distance2(const float* a, const float* b, size_t n) {
    float acc = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float diff = a[i] - b[i];
        acc = acc + diff * diff;
    }
    return acc;
}

/* ------------------------------------------------------------------------- */
/* hsdlib restatements.  HSD_ALLOW_FP_CHECKS is 1 in the vq build             */
/* (build.rs never defines HSDLIB_NO_CHECKS; hsdlib.h:4-8).                   */
/* ------------------------------------------------------------------------- */
#define HSD_OK 0
#define HSD_INVALID (-3)

static int bad(float v) { return isnan(v) || isinf(v); }

/* _mm512_reduce_add_ps as GCC/clang expand it: 16->8->4->2->1 halving tree. */
static float reduce16(const float* l) {
    float a8[8], a4[4], a2[2];
    for (int i = 0; i < 8; ++i) a8[i] = l[i] + l[i + 8];
    for (int i = 0; i < 4; ++i) a4[i] = a8[i] + a8[i + 4];
    for (int i = 0; i < 2; ++i) a2[i] = a4[i] + a4[i + 2];
    return a2[0] + a2[1];
}
/* hsd_internal_hsum_avx_f32, hsdlib.h:240-245: (lo128+hi128), hadd, hadd. */
static float reduce8(const float* l) {
    float h[4];
    for (int i = 0; i < 4; ++i) h[i] = l[i] + l[i + 4];
    float p0 = h[0] + h[1], p1 = h[2] + h[3];
    return p0 + p1;
}

static size_t lanes_of(int sem) { return sem == VQO_SEM_AVX2 ? 8 : 16; }

/* euclidean.c:35-57 (scalar), :97-129 (AVX2), :131-163 (AVX-512F) */
int vqo_hsd_sqeuclid(int sem, const float* a, const float* b, size_t n, float* out) {
    if (n == 0) { *out = 0.0f; return HSD_OK; } /* euclidean.c:251-254 */
    size_t W = (sem == VQO_SEM_SCALAR) ? 0 : lanes_of(sem);
    size_t i = 0;
    float sum = 0.0f;
    if (W) {
        float acc[16] = {0};
        for (; i + W <= n; i += W)
            for (size_t l = 0; l < W; ++l) {
                float d = a[i + l] - b[i + l];
                acc[l] = fmaf(d, d, acc[l]);
            }
        sum = (W == 16) ? reduce16(acc) : reduce8(acc);
    }
    for (; i < n; ++i) {
        if (bad(a[i]) || bad(b[i])) { *out = NAN; return HSD_INVALID; }
        float d = a[i] - b[i];
        sum += d * d;
    }
    *out = sum;
    return bad(sum) ? HSD_INVALID : HSD_OK;
}

/* manhattan.c:39-60 (scalar), :97-130 (AVX2), :132-163 (AVX-512F) */
int vqo_hsd_manhattan(int sem, const float* a, const float* b, size_t n, float* out) {
    if (n == 0) { *out = 0.0f; return HSD_OK; }
    size_t W = (sem == VQO_SEM_SCALAR) ? 0 : lanes_of(sem);
    size_t i = 0;
    float sum = 0.0f;
    if (W) {
        float acc[16] = {0};
        for (; i + W <= n; i += W)
            for (size_t l = 0; l < W; ++l) acc[l] = acc[l] + fabsf(a[i + l] - b[i + l]);
        sum = (W == 16) ? reduce16(acc) : reduce8(acc);
    }
    for (; i < n; ++i) {
        if (bad(a[i]) || bad(b[i])) { *out = NAN; return HSD_INVALID; }
        sum += fabsf(a[i] - b[i]);
    }
    *out = sum;
    return bad(sum) ? HSD_INVALID : HSD_OK;
}

/* cosine.c:28-63 */
static int cosine_from_sums(float dot, float na, float nb, float* out) {
    if (bad(dot) || bad(na) || bad(nb)) { *out = NAN; return HSD_INVALID; }
    int az = na < FLT_MIN, bz = nb < FLT_MIN;
    float sim;
    if (az && bz) sim = 1.0f;
    else if (az || bz) sim = 0.0f;
    else {
        float denom = sqrtf(na) * sqrtf(nb);
        if (denom < FLT_MIN) sim = 0.0f;
        else {
            sim = dot / denom;
            if (sim > 1.0f) sim = 1.0f;
            if (sim < -1.0f) sim = -1.0f;
        }
    }
    if (bad(sim)) { *out = NAN; return HSD_INVALID; }
    *out = sim;
    return HSD_OK;
}

/* cosine.c:65-81 (scalar), :126-161 (AVX2), :163-198 (AVX-512F); returns SIMILARITY */
int vqo_hsd_cosine(int sem, const float* a, const float* b, size_t n, float* out) {
    if (n == 0) { *out = 1.0f; return HSD_OK; } /* cosine.c:288-291 */
    size_t W = (sem == VQO_SEM_SCALAR) ? 0 : lanes_of(sem);
    size_t i = 0;
    float dot = 0.0f, na = 0.0f, nb = 0.0f;
    if (W) {
        float d[16] = {0}, x[16] = {0}, y[16] = {0};
        for (; i + W <= n; i += W)
            for (size_t l = 0; l < W; ++l) {
                d[l] = fmaf(a[i + l], b[i + l], d[l]);
                x[l] = fmaf(a[i + l], a[i + l], x[l]);
                y[l] = fmaf(b[i + l], b[i + l], y[l]);
            }
        if (W == 16) { dot = reduce16(d); na = reduce16(x); nb = reduce16(y); }
        else { dot = reduce8(d); na = reduce8(x); nb = reduce8(y); }
    }
    for (; i < n; ++i) {
        if (bad(a[i]) || bad(b[i])) { *out = NAN; return HSD_INVALID; }
        dot += a[i] * b[i];
        na += a[i] * a[i];
        nb += b[i] * b[i];
    }
    return cosine_from_sums(dot, na, nb, out);
}

static vqo_hsd_fn g_hsd_sq = NULL, g_hsd_l1 = NULL, g_hsd_cos = NULL;
void vqo_set_hsdlib(vqo_hsd_fn sq, vqo_hsd_fn l1, vqo_hsd_fn cs) {
    g_hsd_sq = sq; g_hsd_l1 = l1; g_hsd_cos = cs;
}

/* ------------------------------------------------------------------------- */
/* Distance::compute, src/core/distance.rs:48-120                             */
/* ------------------------------------------------------------------------- */
static float rust_sq(const float* a, const float* b, size_t n) { /* distance.rs:75-83 */
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) { float d = a[i] - b[i]; s = s + d * d; }
    return s;
}
static float rust_l1(const float* a, const float* b, size_t n) { /* distance.rs:93-95 */
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) s = s + fabsf(a[i] - b[i]);
    return s;
}
static float rust_cos(const float* a, const float* b, size_t n) { /* distance.rs:106-119 */
    float dot = 0.0f, na = 0.0f, nb = 0.0f;
    for (size_t i = 0; i < n; ++i) dot = dot + a[i] * b[i];
    for (size_t i = 0; i < n; ++i) na = na + a[i] * a[i];
    for (size_t i = 0; i < n; ++i) nb = nb + b[i] * b[i];
    na = sqrtf(na); nb = sqrtf(nb);
    if (na < 1e-10f || nb < 1e-10f) return 1.0f;
    float v = 1.0f - (dot / (na * nb));
    /* f32::clamp(0,1): NaN stays NaN */
    if (v < 0.0f) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    return v;
}

static float sq_dispatch(int sem, const float* a, const float* b, size_t n) { /* distance.rs:68-83 */
    float r;
    if (sem == VQO_SEM_HSDLIB && g_hsd_sq) { if (g_hsd_sq(a, b, n, &r) == 0) return r; }
    else if (sem != VQO_SEM_SCALAR) { if (vqo_hsd_sqeuclid(sem, a, b, n, &r) == 0) return r; }
    return rust_sq(a, b, n);
}

float vqo_distance(int metric, int sem, const float* a, const float* b, size_t n) {
    float r;
    switch (metric) {
    case VQO_SQEUCLIDEAN: return sq_dispatch(sem, a, b, n);
    case VQO_EUCLIDEAN:   return sqrtf(sq_dispatch(sem, a, b, n)); /* distance.rs:58 */
    case VQO_CHEBYSHEV: {   /* EXTENSION, no reference counterpart: max_i |a_i - b_i|, NaN differences skipped */
        float mx = 0.0f;
        for (size_t i = 0; i < n; ++i) { float v = fabsf(a[i] - b[i]); if (v > mx) mx = v; }
        return mx;
    }
    case VQO_MANHATTAN:                                              /* distance.rs:86-95 */
        if (sem == VQO_SEM_HSDLIB && g_hsd_l1) { if (g_hsd_l1(a, b, n, &r) == 0) return r; }
        else if (sem != VQO_SEM_SCALAR) { if (vqo_hsd_manhattan(sem, a, b, n, &r) == 0) return r; }
        return rust_l1(a, b, n);
    default:                                                         /* distance.rs:98-119 */
        if (sem == VQO_SEM_HSDLIB && g_hsd_cos) { if (g_hsd_cos(a, b, n, &r) == 0) return 1.0f - r; }
        else if (sem != VQO_SEM_SCALAR) { if (vqo_hsd_cosine(sem, a, b, n, &r) == 0) return 1.0f - r; }
        return rust_cos(a, b, n);
    }
}

/* ------------------------------------------------------------------------- */
/* LBG / k-means, src/core/vector.rs:350-461                                  */
/* ------------------------------------------------------------------------- */
static uint32_t nearest_centroid(const float* v, const float* c, size_t k, size_t d) { /* :352-363 */
    uint32_t best = 0;
    float best_dist = vqo_distance2(v, c, d);
    for (size_t j = 1; j < k; ++j) {
        float dist = vqo_distance2(v, c + j * d, d);
        if (dist < best_dist) { best_dist = dist; best = (uint32_t)j; }
    }
    return best;
}

int vqo_lbg_step(const float* x, size_t n, size_t ld, size_t col0, size_t d,
                 float* cent, size_t k, uint32_t* assign_out,
                 uint32_t* empties_out, uint32_t* n_empty, int threads) {
    uint32_t* assign = assign_out ? assign_out : (uint32_t*)malloc(n * sizeof(uint32_t));
    float* sums = (float*)calloc(k * d, sizeof(float));
    uint64_t* cnt = (uint64_t*)calloc(k, sizeof(uint64_t));
    (void)threads;
    /* vector.rs:417-429: pure map over points (rayon par_iter with `parallel`) */
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        assign[i] = nearest_centroid(x + (size_t)i * ld + col0, cent, k, d);
    /* vector.rs:432-435 + :368-384: members pushed in ascending id, summed in that order */
    for (size_t i = 0; i < n; ++i) {
        const float* v = x + i * ld + col0;
        float* s = sums + (size_t)assign[i] * d;
        for (size_t t = 0; t < d; ++t) s[t] += v[t];
        cnt[assign[i]]++;
    }
    int changed = 0;
    uint32_t ne = 0;
    for (size_t j = 0; j < k; ++j) { /* vector.rs:440-453 */
        if (cnt[j]) {
            float nn = (float)cnt[j]; /* `indices.len() as f32` */
            for (size_t t = 0; t < d; ++t) {
                float nv = sums[j * d + t] / nn;
                if (!(fabsf(nv - cent[j * d + t]) < 1e-6f)) changed = 1; /* approx_eq :232-240 */
                cent[j * d + t] = nv;
            }
        } else {
            if (empties_out) empties_out[ne] = (uint32_t)j;
            ne++;
        }
    }
    if (n_empty) *n_empty = ne;
    free(sums); free(cnt);
    if (!assign_out) free(assign);
    return changed;
}

int vqo_pq_train(const float* x, size_t n, size_t dim, size_t m, size_t k, size_t max_iters,
                 const uint64_t* init_idx, vqo_reseed_fn reseed, void* user,
                 float* cb, uint32_t* iters_run, int threads) {
    if (n == 0) return -2;              /* EmptyInput, pq.rs:91-93 */
    if (m == 0 || dim < m || dim % m) return -3; /* pq.rs:106-117 */
    if (k == 0 || n < k) return -3;     /* vector.rs:399-410 */
    size_t d = dim / m;
    uint32_t* empt = (uint32_t*)malloc(k * sizeof(uint32_t));
    for (size_t s = 0; s < m; ++s) { /* pq.rs:121-132: subspaces trained serially */
        float* c = cb + s * k * d;
        for (size_t j = 0; j < k; ++j) { /* vector.rs:413: sampled rows become the centroids */
            uint64_t r = init_idx[s * k + j];
            if (r >= n) { free(empt); return -3; }
            memcpy(c + j * d, x + r * dim + s * d, d * sizeof(float));
        }
        uint32_t it = 0;
        for (; it < max_iters;) { /* vector.rs:415 */
            uint32_t ne = 0;
            int changed = vqo_lbg_step(x, n, dim, s * d, d, c, k, NULL, empt, &ne, threads);
            ++it;
            for (uint32_t e = 0; e < ne; ++e) { /* vector.rs:448-452, ascending j */
                uint64_t r = reseed ? reseed(user, (uint32_t)s) : 0;
                if (r >= n) r = r % n;
                memcpy(c + (size_t)empt[e] * d, x + r * dim + s * d, d * sizeof(float));
            }
            if (!changed) break; /* vector.rs:455-457 */
        }
        if (iters_run) iters_run[s] = it;
    }
    free(empt);
    return 0;
}

/* ProductQuantizer::quantize, src/pq.rs:167-199 */
int vqo_pq_encode(const float* cb, size_t m, size_t k, size_t d, int metric, int sem,
                  const float* x, size_t n, uint32_t* codes, uint16_t* recon, int threads) {
    size_t dim = m * d;
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
        size_t i = (size_t)ii;
        for (size_t s = 0; s < m; ++s) {
            const float* v = x + i * dim + s * d;
            const float* c = cb + s * k * d;
            size_t best = 0;
            float best_dist = vqo_distance(metric, sem, v, c, d);
            for (size_t j = 1; j < k; ++j) {
                float dist = vqo_distance(metric, sem, v, c + j * d, d);
                if (dist < best_dist) { best_dist = dist; best = j; }
            }
            if (codes) codes[i * m + s] = (uint32_t)best;
            if (recon)
                for (size_t t = 0; t < d; ++t) recon[i * dim + s * d + t] = vqo_f32_to_f16(c[best * d + t]);
        }
    }
    return 0;
}

void vqo_dequantize_f16(const uint16_t* q, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) out[i] = vqo_f16_to_f32(q[i]);
}

/* ------------------------------------------------------------------------- */
/* BinaryQuantizer src/bq.rs:94-118, ScalarQuantizer src/sq.rs:94,123-151     */
/* ------------------------------------------------------------------------- */
void vqo_bq_quantize(const float* x, size_t n, float thr, uint8_t low, uint8_t high, uint8_t* out) {
    for (size_t i = 0; i < n; ++i) out[i] = (x[i] >= thr) ? high : low;
}
void vqo_bq_dequantize(const uint8_t* c, size_t n, uint8_t low, uint8_t high, float* out) {
    for (size_t i = 0; i < n; ++i) out[i] = (c[i] >= high) ? (float)high : (float)low;
}
float vqo_sq_step(float mn, float mx, size_t levels) { return (mx - mn) / (float)(levels - 1); }

void vqo_sq_quantize(const float* x, size_t n, float mn, float mx, float step, size_t levels, uint8_t* out) {
    for (size_t i = 0; i < n; ++i) {
        float v = x[i];
        /* f32::clamp: NaN passes through */
        float c = v;
        if (c < mn) c = mn;
        if (c > mx) c = mx;
        float q = roundf((c - mn) / step); /* f32::round = half away from zero */
        size_t idx;
        if (isnan(q) || q <= 0.0f) idx = 0;             /* `as usize` saturates; NaN -> 0 */
        else if (q >= 1.8446744e19f) idx = SIZE_MAX;
        else idx = (size_t)q;
        if (idx > levels - 1) idx = levels - 1;
        out[i] = (uint8_t)idx;
    }
}
void vqo_sq_dequantize(const uint8_t* c, size_t n, float mn, float step, float* out) {
    for (size_t i = 0; i < n; ++i) {
        float p = (float)c[i] * step; /* two roundings: mul then add (sq.rs:149) */
        out[i] = mn + p;
    }
}

/* ------------------------------------------------------------------------- */
/* TSVQ, src/tsvq.rs:31-132 (breadth-first instead of recursive; nodes are     */
/* independent so the result is the same tree).                                */
/* ------------------------------------------------------------------------- */
static int total_cmp_f32(const void* pa, const void* pb) { /* f32::total_cmp, tsvq.rs:75 */
    int32_t a, b;
    memcpy(&a, pa, 4); memcpy(&b, pb, 4);
    a ^= (int32_t)(((uint32_t)(a >> 31)) >> 1);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return (a > b) - (a < b);
}

int64_t vqo_tsvq_build(const float* x, size_t n, size_t dim, size_t max_depth,
                       float* cent, int32_t* left, int32_t* right,
                       int32_t* split_dim_out, float* median_out, uint64_t* count_out,
                       size_t max_nodes) {
    if (n == 0) return -2;
    /* per-node member lists (row ids, parent order preserved) kept in two ping-pong arrays */
    uint32_t* ids = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint32_t* ids2 = (uint32_t*)malloc(n * sizeof(uint32_t));
    float* vals = (float*)malloc(n * sizeof(float));
    float* var = (float*)malloc(dim * sizeof(float));
    /* queue entries: node id, start, len, depth-left */
    size_t qcap = max_nodes + 1;
    size_t* q_start = (size_t*)malloc(qcap * sizeof(size_t));
    size_t* q_len = (size_t*)malloc(qcap * sizeof(size_t));
    size_t* q_depth = (size_t*)malloc(qcap * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) ids[i] = (uint32_t)i;
    size_t n_nodes = 1, head = 0;
    q_start[0] = 0; q_len[0] = n; q_depth[0] = max_depth;
    int64_t rc = 0;
    /* Each level permutes `ids` in place segment by segment (stable), so one array suffices;
       ids2 is scratch for the partition. */
    while (head < n_nodes) {
        size_t node = head++;
        size_t st = q_start[node], len = q_len[node], depth = q_depth[node];
        float* c = cent + node * dim;
        /* mean_vector, vector.rs:332-348: sequential sums in member order, then / n */
        for (size_t t = 0; t < dim; ++t) c[t] = 0.0f;
        for (size_t i = 0; i < len; ++i) {
            const float* v = x + (size_t)ids[st + i] * dim;
            for (size_t t = 0; t < dim; ++t) c[t] = c[t] + v[t];
        }
        float nn = (float)len;
        for (size_t t = 0; t < dim; ++t) c[t] = c[t] / nn;
        left[node] = -1; right[node] = -1;
        if (split_dim_out) split_dim_out[node] = -1;
        if (median_out) median_out[node] = NAN;
        if (count_out) count_out[node] = len;
        if (depth == 0 || len <= 1) continue; /* tsvq.rs:38-44 */
        /* tsvq.rs:47-57: per-dimension sum of squared deviations, sequential over members */
        for (size_t t = 0; t < dim; ++t) var[t] = 0.0f;
        for (size_t i = 0; i < len; ++i) {
            const float* v = x + (size_t)ids[st + i] * dim;
            for (size_t t = 0; t < dim; ++t) { float df = v[t] - c[t]; var[t] = var[t] + df * df; }
        }
        /* tsvq.rs:59-66: max_by over non-NaN, LAST maximal element wins; none -> 0 */
        size_t sd = 0; int have = 0; float best = 0.0f;
        for (size_t t = 0; t < dim; ++t) {
            if (isnan(var[t])) continue;
            if (!have || !(var[t] < best)) { best = var[t]; sd = t; have = 1; }
        }
        /* tsvq.rs:68-81 */
        size_t nv = 0;
        for (size_t i = 0; i < len; ++i) {
            float v = x[(size_t)ids[st + i] * dim + sd];
            if (!isnan(v)) vals[nv++] = v;
        }
        if (nv == 0) { rc = -3; break; } /* reference would index out of bounds (panic) */
        qsort(vals, nv, sizeof(float), total_cmp_f32);
        float median = (nv % 2 == 0) ? (vals[nv / 2 - 1] + vals[nv / 2]) / 2.0f : vals[nv / 2];
        if (split_dim_out) split_dim_out[node] = (int32_t)sd;
        if (median_out) median_out[node] = median;
        /* tsvq.rs:84-85: stable partition, `<= median` left (NaN goes right) */
        size_t nl = 0, nr = 0;
        for (size_t i = 0; i < len; ++i) {
            uint32_t id = ids[st + i];
            if (x[(size_t)id * dim + sd] <= median) ids[st + nl++] = id; /* nl <= i: safe in place */
            else ids2[nr++] = id;
        }
        memcpy(ids + st + nl, ids2, nr * sizeof(uint32_t));
        /* tsvq.rs:88-108: child only if non-empty and strictly smaller than the parent */
        if (nl > 0 && nl < len) {
            if (n_nodes >= max_nodes) { rc = -99; break; }
            left[node] = (int32_t)n_nodes;
            q_start[n_nodes] = st; q_len[n_nodes] = nl; q_depth[n_nodes] = depth - 1; n_nodes++;
        }
        if (nr > 0 && nr < len) {
            if (n_nodes >= max_nodes) { rc = -99; break; }
            right[node] = (int32_t)n_nodes;
            q_start[n_nodes] = st + nl; q_len[n_nodes] = nr; q_depth[n_nodes] = depth - 1; n_nodes++;
        }
    }
    free(ids); free(ids2); free(vals); free(var); free(q_start); free(q_len); free(q_depth);
    return rc < 0 ? rc : (int64_t)n_nodes;
}

/* TSVQNode::find_leaf + TSVQ::quantize, tsvq.rs:117-132, :239-255 */
int vqo_tsvq_encode(const float* cent, const int32_t* left, const int32_t* right, size_t dim,
                    int metric, int sem, const float* x, size_t n,
                    uint32_t* leaf_out, uint16_t* recon, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
        size_t i = (size_t)ii;
        const float* v = x + i * dim;
        int32_t node = 0;
        for (;;) {
            int32_t l = left[node], r = right[node];
            if (l >= 0 && r >= 0) {
                float dl = vqo_distance(metric, sem, v, cent + (size_t)l * dim, dim);
                float dr = vqo_distance(metric, sem, v, cent + (size_t)r * dim, dim);
                node = (dl <= dr) ? l : r;
            } else if (l >= 0) node = l;
            else if (r >= 0) node = r;
            else break;
        }
        if (leaf_out) leaf_out[i] = (uint32_t)node;
        if (recon)
            for (size_t t = 0; t < dim; ++t) recon[i * dim + t] = vqo_f32_to_f16(cent[(size_t)node * dim + t]);
    }
    return 0;
}

/* quality metric of src/bin/common.rs:61-78 (f64 accumulation here: it is a report, not the path) */
double vqo_recon_mse(const float* x, const uint16_t* recon, size_t count) {
    double s = 0.0;
    for (size_t i = 0; i < count; ++i) {
        double d = (double)x[i] - (double)vqo_f16_to_f32(recon[i]);
        s += d * d;
    }
    return count ? s / (double)count : 0.0;
}
