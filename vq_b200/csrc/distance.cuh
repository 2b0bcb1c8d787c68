// distance.cuh -- device functions that reproduce, operation for operation, the f32
// arithmetic of the reference's distance code, so that exact re-checks and the
// CUDA-core kernels are bit-identical with the CPU results:
//   * training distance  Vector::distance2, src/core/vector.rs:135-143 (sequential, no FMA)
//   * encode distance    Distance::compute, src/core/distance.rs:48-120, `simd` build on an
//     AVX-512F host: hsdlib euclidean.c:131-163, manhattan.c:132-163, cosine.c:163-198,28-63
//     (16 FMA lanes + _mm512_reduce_add_ps tree + sequential no-FMA tail), including the
//     Rust scalar fallbacks taken when hsdlib reports HSD_ERR_INVALID_INPUT.
// All arithmetic uses the __f*_rn intrinsics so nvcc can never contract mul+add into FMA.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>

#define VQB_DEV __device__ __forceinline__

VQB_DEV bool vqb_bad(float v) { return isnan(v) || isinf(v); }

// Generic element accessors let the same code run on registers, smem or global memory.
struct PtrAcc {
    const float* p;
    VQB_DEV float operator()(int i) const { return p[i]; }
};

// vector.rs:135-143
template <int N = 0, typename A, typename B>
VQB_DEV float dist2_seq(const A& a, const B& b, int n) {
    if (N > 0) n = N;  // compile-time length: loops below unroll fully
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < n; ++i) {
        float d = __fsub_rn(a(i), b(i));
        acc = __fadd_rn(acc, __fmul_rn(d, d));
    }
    return acc;
}

// _mm512_reduce_add_ps: halving tree 16 -> 8 -> 4 -> 2 -> 1
VQB_DEV float reduce16(const float* l) {
    float a8[8], a4[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a8[i] = __fadd_rn(l[i], l[i + 8]);
#pragma unroll
    for (int i = 0; i < 4; ++i) a4[i] = __fadd_rn(a8[i], a8[i + 4]);
    float b0 = __fadd_rn(a4[0], a4[2]), b1 = __fadd_rn(a4[1], a4[3]);
    return __fadd_rn(b0, b1);
}

// hsd_dist_sqeuclidean_f32, AVX-512F kernel.  ok=false <=> HSD_ERR_INVALID_INPUT.
template <int N = 0, typename A, typename B>
VQB_DEV float hsd_sqeuclid(const A& a, const B& b, int n, bool& ok) {
    if (N > 0) n = N;  // compile-time length: loops below unroll fully
    ok = true;
    if (n == 0) return 0.0f;
    int i = 0;
    float sum = 0.0f;
    if (n >= 16) {
        float acc[16];
#pragma unroll
        for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
#pragma unroll
        for (; i + 16 <= n; i += 16) {
#pragma unroll
            for (int l = 0; l < 16; ++l) {
                float d = __fsub_rn(a(i + l), b(i + l));
                acc[l] = __fmaf_rn(d, d, acc[l]);
            }
        }
        sum = reduce16(acc);
    }
#pragma unroll
    for (; i < n; ++i) {
        float x = a(i), y = b(i);
        if (vqb_bad(x) || vqb_bad(y)) { ok = false; return nanf(""); }
        float d = __fsub_rn(x, y);
        sum = __fadd_rn(sum, __fmul_rn(d, d));
    }
    if (vqb_bad(sum)) ok = false;
    return sum;
}

template <int N = 0, typename A, typename B>
VQB_DEV float hsd_manhattan(const A& a, const B& b, int n, bool& ok) {
    if (N > 0) n = N;  // compile-time length: loops below unroll fully
    ok = true;
    if (n == 0) return 0.0f;
    int i = 0;
    float sum = 0.0f;
    if (n >= 16) {
        float acc[16];
#pragma unroll
        for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
#pragma unroll
        for (; i + 16 <= n; i += 16) {
#pragma unroll
            for (int l = 0; l < 16; ++l) acc[l] = __fadd_rn(acc[l], fabsf(__fsub_rn(a(i + l), b(i + l))));
        }
        sum = reduce16(acc);
    }
#pragma unroll
    for (; i < n; ++i) {
        float x = a(i), y = b(i);
        if (vqb_bad(x) || vqb_bad(y)) { ok = false; return nanf(""); }
        sum = __fadd_rn(sum, fabsf(__fsub_rn(x, y)));
    }
    if (vqb_bad(sum)) ok = false;
    return sum;
}

// The three running sums of hsd_sim_cosine_f32 (cosine.c:163-198).  tail_ok=false when the
// scalar tail met a NaN/Inf element (cosine.c:187-192).
template <int N = 0, typename A, typename B>
VQB_DEV void hsd_cosine_sums(const A& a, const B& b, int n, float& dot, float& na, float& nb, bool& tail_ok) {
    if (N > 0) n = N;  // compile-time length: loops below unroll fully
    tail_ok = true;
    int i = 0;
    dot = na = nb = 0.0f;
    if (n >= 16) {
        float d[16], x[16], y[16];
#pragma unroll
        for (int l = 0; l < 16; ++l) d[l] = x[l] = y[l] = 0.0f;
#pragma unroll
        for (; i + 16 <= n; i += 16) {
#pragma unroll
            for (int l = 0; l < 16; ++l) {
                float u = a(i + l), v = b(i + l);
                d[l] = __fmaf_rn(u, v, d[l]);
                x[l] = __fmaf_rn(u, u, x[l]);
                y[l] = __fmaf_rn(v, v, y[l]);
            }
        }
        dot = reduce16(d); na = reduce16(x); nb = reduce16(y);
    }
#pragma unroll
    for (; i < n; ++i) {
        float u = a(i), v = b(i);
        if (vqb_bad(u) || vqb_bad(v)) { tail_ok = false; return; }
        dot = __fadd_rn(dot, __fmul_rn(u, v));
        na = __fadd_rn(na, __fmul_rn(u, u));
        nb = __fadd_rn(nb, __fmul_rn(v, v));
    }
}

// One-sided version: the squared norm exactly as hsd_cosine_sums accumulates it (it depends on
// one operand only), so it can be hoisted out of the per-centroid loop.
template <int N = 0, typename A>
VQB_DEV float hsd_cosine_norm(const A& a, int n, bool& tail_ok) {
    if (N > 0) n = N;  // compile-time length: loops below unroll fully
    tail_ok = true;
    int i = 0;
    float na = 0.0f;
    if (n >= 16) {
        float x[16];
#pragma unroll
        for (int l = 0; l < 16; ++l) x[l] = 0.0f;
#pragma unroll
        for (; i + 16 <= n; i += 16) {
#pragma unroll
            for (int l = 0; l < 16; ++l) { float u = a(i + l); x[l] = __fmaf_rn(u, u, x[l]); }
        }
        na = reduce16(x);
    }
#pragma unroll
    for (; i < n; ++i) {
        float u = a(i);
        if (vqb_bad(u)) tail_ok = false;  // keep summing: caller only needs the flag
        na = __fadd_rn(na, __fmul_rn(u, u));
    }
    return na;
}
template <int N = 0, typename A, typename B>
VQB_DEV float hsd_cosine_dot(const A& a, const B& b, int n) {
    if (N > 0) n = N;  // compile-time length: loops below unroll fully
    int i = 0;
    float dot = 0.0f;
    if (n >= 16) {
        float d[16];
#pragma unroll
        for (int l = 0; l < 16; ++l) d[l] = 0.0f;
#pragma unroll
        for (; i + 16 <= n; i += 16) {
#pragma unroll
            for (int l = 0; l < 16; ++l) d[l] = __fmaf_rn(a(i + l), b(i + l), d[l]);
        }
        dot = reduce16(d);
    }
#pragma unroll
    for (; i < n; ++i) dot = __fadd_rn(dot, __fmul_rn(a(i), b(i)));
    return dot;
}

// calculate_cosine_similarity_from_sums, cosine.c:28-63.  sa/sb = sqrtf(na)/sqrtf(nb).
VQB_DEV float hsd_cosine_from_sums(float dot, float na, float nb, float sa, float sb, bool& ok) {
    ok = true;
    if (vqb_bad(dot) || vqb_bad(na) || vqb_bad(nb)) { ok = false; return nanf(""); }
    bool az = na < FLT_MIN, bz = nb < FLT_MIN;
    float sim;
    if (az && bz) sim = 1.0f;
    else if (az || bz) sim = 0.0f;
    else {
        float denom = __fmul_rn(sa, sb);
        if (denom < FLT_MIN) sim = 0.0f;
        else {
            sim = __fdiv_rn(dot, denom);
            if (sim > 1.0f) sim = 1.0f;
            if (sim < -1.0f) sim = -1.0f;
        }
    }
    if (vqb_bad(sim)) { ok = false; return nanf(""); }
    return sim;
}

// Rust scalar fallbacks (distance.rs:75-83, 93-95, 106-119)
template <int N = 0, typename A, typename B>
VQB_DEV float rust_l1(const A& a, const B& b, int n) {
    if (N > 0) n = N;  // compile-time length: loops below unroll fully
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < n; ++i) s = __fadd_rn(s, fabsf(__fsub_rn(a(i), b(i))));
    return s;
}
template <int N = 0, typename A, typename B>
VQB_DEV float rust_cos(const A& a, const B& b, int n) {
    if (N > 0) n = N;  // compile-time length: loops below unroll fully
    float dot = 0.0f, na = 0.0f, nb = 0.0f;
#pragma unroll
    for (int i = 0; i < n; ++i) dot = __fadd_rn(dot, __fmul_rn(a(i), b(i)));
#pragma unroll
    for (int i = 0; i < n; ++i) na = __fadd_rn(na, __fmul_rn(a(i), a(i)));
#pragma unroll
    for (int i = 0; i < n; ++i) nb = __fadd_rn(nb, __fmul_rn(b(i), b(i)));
    na = __fsqrt_rn(na); nb = __fsqrt_rn(nb);
    if (na < 1e-10f || nb < 1e-10f) return 1.0f;
    float v = __fsub_rn(1.0f, __fdiv_rn(dot, __fmul_rn(na, nb)));
    if (v < 0.0f) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    return v;
}

// Distance::compute for one pair (distance.rs:48-65), metric = VQB_* id.
template <int N = 0, typename A, typename B>
VQB_DEV float vq_distance(int metric, const A& a, const B& b, int n) {
    if (N > 0) n = N;  // compile-time length: loops below unroll fully
    bool ok;
    if (metric == 0 || metric == 1) {
        float s = hsd_sqeuclid<N>(a, b, n, ok);
        if (!ok) s = dist2_seq<N>(a, b, n);  // distance.rs:75-83 has the same order as distance2
        return metric == 1 ? __fsqrt_rn(s) : s;
    } else if (metric == 5) {
        // EXTENSION (VQB_CHEBYSHEV, no reference counterpart): max_i |a_i - b_i|, NaN differences skipped
        float mx = 0.0f;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float v = fabsf(__fsub_rn(a(i), b(i)));
            if (v > mx) mx = v;
        }
        return mx;
    } else if (metric == 2) {
        // n < 16: hsdlib's kernel is its scalar tail only (manhattan.c:132-163), i.e. the same sequential sum as
        // the Rust fallback it defers to on NaN/Inf -- one formula for every input, no per-element checks
        if (N > 0 && N < 16) return rust_l1<N>(a, b, n);
        float s = hsd_manhattan<N>(a, b, n, ok);
        return ok ? s : rust_l1<N>(a, b, n);
    } else {
        if (n == 0) return 0.0f;  // hsd: similarity 1 for n == 0 (cosine.c:288-291)
        float dot, na, nb;
        bool tail_ok;
        hsd_cosine_sums<N>(a, b, n, dot, na, nb, tail_ok);
        if (tail_ok) {
            float sim = hsd_cosine_from_sums(dot, na, nb, __fsqrt_rn(na), __fsqrt_rn(nb), ok);
            if (ok) return __fsub_rn(1.0f, sim);
        }
        return rust_cos<N>(a, b, n);
    }
}
