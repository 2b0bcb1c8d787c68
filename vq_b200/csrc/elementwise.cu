// elementwise.cu -- the HBM-bound codecs of the path:
//   BinaryQuantizer  (src/bq.rs:94-118)      f32 -> u8, u8 -> f32        5 B/element
//   ScalarQuantizer  (src/sq.rs:123-151)     f32 -> u8, u8 -> f32        5 B/element
//   PQ/TSVQ dequantize (src/pq.rs:201-209, src/tsvq.rs:257-265)  f16 -> f32   6 B/element
// and the batched Distance::compute used by the host mirror of `Distance`.
//
// Layout: flat arrays, no reuse -> no shared memory.  Every warp-level load/store instruction
// touches one contiguous 512 B (f32) / 128 B (u8) span; four independent 16 B loads are in
// flight per thread before the first use; streaming cache hints (.cs) keep the one-shot data
// from evicting anything useful.  Grid = multiple of the SM count, grid-stride over 4096-element
// block tiles.
#include "common.cuh"
#include "distance.cuh"

namespace {

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;
constexpr size_t EW_TILE = (size_t)EW_THREADS * EW_UNROLL * 4;  // elements per block iteration

struct BqQ {
    float thr; uint8_t lo, hi;
    __device__ __forceinline__ uint8_t operator()(float x) const { return x >= thr ? hi : lo; }  // bq.rs:98
};
struct SqQ {
    float mn, mx, step; uint32_t top;  // top = levels - 1
    float rinv;                        // RN(1 / step), or 0 when the fast division below is not applicable
    // Correctly rounded a / step without the generic division sequence: q0 = RN(a * rinv) is within 2 ulp,
    // one residual correction makes it faithful, a second one (Markstein: y = RN(1/b), q faithful,
    // r = a - b*q exact => RN(q + r*y) = RN(a/b)) makes it the IEEE quotient.  The residuals are exact only
    // away from underflow/overflow, so tiny, huge and non-finite numerators take __fdiv_rn.
    __device__ __forceinline__ float div_step(float a) const {
        if (rinv != 0.0f && ((a >= 1e-18f && a <= 1e18f) || a == 0.0f)) {
            float q = __fmul_rn(a, rinv);
            float r = __fmaf_rn(-q, step, a);
            q = __fmaf_rn(r, rinv, q);
            r = __fmaf_rn(-q, step, a);
            return __fmaf_rn(r, rinv, q);
        }
        return __fdiv_rn(a, step);
    }
    __device__ __forceinline__ uint8_t operator()(float x) const {
        float c = x;                    // f32::clamp: NaN falls through both tests (sq.rs:124)
        if (c < mn) c = mn;
        if (c > mx) c = mx;
        float q = roundf(div_step(__fsub_rn(c, mn)));  // f32::round: half away from zero
        uint32_t idx = __float2uint_rz(q);                    // `as usize`: saturating, NaN -> 0
        return (uint8_t)min(idx, top);                        // sq.rs:126
    }
};
struct BqD {
    uint8_t lo, hi;
    __device__ __forceinline__ float operator()(uint8_t c) const { return c >= hi ? (float)hi : (float)lo; }  // bq.rs:111
};
struct SqD {
    float mn, step;
    __device__ __forceinline__ float operator()(uint8_t c) const {
        return __fadd_rn(mn, __fmul_rn((float)c, step));  // sq.rs:149: two roundings, never an FMA
    }
};

template <typename Op>
__global__ void __launch_bounds__(EW_THREADS) k_f32_to_u8(const float* __restrict__ x, uint8_t* __restrict__ out,
                                                          size_t n, Op op) {
    const size_t n_tiles = n / EW_TILE;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    uchar4* o4 = reinterpret_cast<uchar4*>(out);
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        size_t base = t * (EW_TILE / 4) + threadIdx.x;
        float4 v[EW_UNROLL];
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) v[u] = __ldcs(x4 + base + (size_t)u * EW_THREADS);
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) {
            uchar4 r = make_uchar4(op(v[u].x), op(v[u].y), op(v[u].z), op(v[u].w));
            __stcs(o4 + base + (size_t)u * EW_THREADS, r);
        }
    }
    // ragged tail (< one tile): scalar, spread over the whole grid
    size_t tail0 = n_tiles * EW_TILE;
    for (size_t i = tail0 + (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS)
        out[i] = op(x[i]);
}

template <typename Op>
__global__ void __launch_bounds__(EW_THREADS) k_f32_to_u8_unaligned(const float* __restrict__ x,
                                                                    uint8_t* __restrict__ out, size_t n, Op op) {
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS)
        out[i] = op(x[i]);
}

template <typename Op>
__global__ void __launch_bounds__(EW_THREADS) k_u8_to_f32(const uint8_t* __restrict__ c, float* __restrict__ out,
                                                          size_t n, Op op) {
    const size_t n_tiles = n / EW_TILE;
    const uchar4* c4 = reinterpret_cast<const uchar4*>(c);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        size_t base = t * (EW_TILE / 4) + threadIdx.x;
        uchar4 v[EW_UNROLL];
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) v[u] = __ldcs(c4 + base + (size_t)u * EW_THREADS);
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u)
            __stcs(o4 + base + (size_t)u * EW_THREADS, make_float4(op(v[u].x), op(v[u].y), op(v[u].z), op(v[u].w)));
    }
    size_t tail0 = n_tiles * EW_TILE;
    for (size_t i = tail0 + (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS)
        out[i] = op(c[i]);
}

template <typename Op>
__global__ void __launch_bounds__(EW_THREADS) k_u8_to_f32_unaligned(const uint8_t* __restrict__ c,
                                                                    float* __restrict__ out, size_t n, Op op) {
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS)
        out[i] = op(c[i]);
}

__global__ void __launch_bounds__(EW_THREADS) k_f16_to_f32(const __half* __restrict__ q, float* __restrict__ out,
                                                           size_t n, int aligned) {
    size_t n_tiles = aligned ? n / EW_TILE : 0;
    const uint2* q4 = reinterpret_cast<const uint2*>(q);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        size_t base = t * (EW_TILE / 4) + threadIdx.x;
        uint2 v[EW_UNROLL];
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) v[u] = __ldcs(q4 + base + (size_t)u * EW_THREADS);
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) {
            float2 a = __half22float2(*reinterpret_cast<__half2*>(&v[u].x));
            float2 b = __half22float2(*reinterpret_cast<__half2*>(&v[u].y));
            __stcs(o4 + base + (size_t)u * EW_THREADS, make_float4(a.x, a.y, b.x, b.y));
        }
    }
    size_t tail0 = n_tiles * EW_TILE;
    for (size_t i = tail0 + (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS)
        out[i] = __half2float(q[i]);
}

__global__ void k_distance_batch(int metric, const float* __restrict__ a, const float* __restrict__ b, size_t rows,
                                 int n, float* __restrict__ out) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    PtrAcc pa{a + r * (size_t)n}, pb{b + r * (size_t)n};
    out[r] = vq_distance(metric, pa, pb, n);
}

inline unsigned ew_grid(vqb_ctx* ctx, size_t n) {
    size_t want = (n + EW_TILE - 1) / EW_TILE;
    size_t cap = (size_t)ctx->sm_count * 8;  // 8 resident 256-thread CTAs per SM
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename Op>
int run_f32_to_u8(vqb_ctx* ctx, const float* x, size_t n, uint8_t* out, Op op) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if (n == 0) return VQB_SUCCESS;  // empty input is legal (tests/integration_tests.rs:296-310)
    if (!x || !out) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null data pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    ChunkIo io; io.in = x; io.in_unit = 4; io.out0 = out; io.out0_unit = 1;
    // host buffers: 64 MB chunks, copies overlapped with the kernels (BASELINE config 2 does not fit any other way)
    return vqb_chunk_pipeline(ctx, n, io, [&](const void* din, void* d0, void*, size_t units, size_t) -> int {
        const float* dx = static_cast<const float*>(din);
        uint8_t* dout = static_cast<uint8_t*>(d0);
        if (aligned_to(dx, 16) && aligned_to(dout, 4))
            k_f32_to_u8<<<ew_grid(ctx, units), EW_THREADS, 0, ctx->stream>>>(dx, dout, units, op);
        else
            k_f32_to_u8_unaligned<<<ew_grid(ctx, units), EW_THREADS, 0, ctx->stream>>>(dx, dout, units, op);
        VQB_LAUNCHED(ctx);
        return VQB_SUCCESS;
    });
}

template <typename Op>
int run_u8_to_f32(vqb_ctx* ctx, const uint8_t* c, size_t n, float* out, Op op) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if (n == 0) return VQB_SUCCESS;
    if (!c || !out) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null data pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    ChunkIo io; io.in = c; io.in_unit = 1; io.out0 = out; io.out0_unit = 4;
    return vqb_chunk_pipeline(ctx, n, io, [&](const void* din, void* d0, void*, size_t units, size_t) -> int {
        const uint8_t* dc = static_cast<const uint8_t*>(din);
        float* dout = static_cast<float*>(d0);
        if (aligned_to(dc, 4) && aligned_to(dout, 16))
            k_u8_to_f32<<<ew_grid(ctx, units), EW_THREADS, 0, ctx->stream>>>(dc, dout, units, op);
        else
            k_u8_to_f32_unaligned<<<ew_grid(ctx, units), EW_THREADS, 0, ctx->stream>>>(dc, dout, units, op);
        VQB_LAUNCHED(ctx);
        return VQB_SUCCESS;
    });
}

}  // namespace

extern "C" {

int vqb_bq_quantize(vqb_ctx* ctx, const float* x, size_t n, float threshold, uint8_t low, uint8_t high,
                    uint8_t* out) {
    return run_f32_to_u8(ctx, x, n, out, BqQ{threshold, low, high});
}

int vqb_bq_dequantize(vqb_ctx* ctx, const uint8_t* codes, size_t n, uint8_t low, uint8_t high, float* out) {
    return run_u8_to_f32(ctx, codes, n, out, BqD{low, high});
}

int vqb_sq_quantize(vqb_ctx* ctx, const float* x, size_t n, float mn, float mx, float step, uint32_t levels,
                    uint8_t* out) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if (levels < 2 || levels > 256) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "levels must be in [2,256]");
    // fast exact division only for steps whose reciprocal and residual products stay in the normal range
    const float rinv = (step >= 1e-18f && step <= 1e18f) ? 1.0f / step : 0.0f;
    return run_f32_to_u8(ctx, x, n, out, SqQ{mn, mx, step, levels - 1, rinv});
}

int vqb_sq_dequantize(vqb_ctx* ctx, const uint8_t* codes, size_t n, float mn, float step, float* out) {
    return run_u8_to_f32(ctx, codes, n, out, SqD{mn, step});
}

int vqb_f16_dequantize(vqb_ctx* ctx, const uint16_t* q, size_t n, float* out) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if (n == 0) return VQB_SUCCESS;
    if (!q || !out) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null data pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    ChunkIo io; io.in = q; io.in_unit = 2; io.out0 = out; io.out0_unit = 4;
    return vqb_chunk_pipeline(ctx, n, io, [&](const void* din, void* d0, void*, size_t units, size_t) -> int {
        int al = aligned_to(din, 8) && aligned_to(d0, 16);
        k_f16_to_f32<<<ew_grid(ctx, units), EW_THREADS, 0, ctx->stream>>>(static_cast<const __half*>(din), static_cast<float*>(d0), units, al);
        VQB_LAUNCHED(ctx);
        return VQB_SUCCESS;
    });
}

int vqb_distance_batch(vqb_ctx* ctx, int metric, const float* a, const float* b, size_t rows, size_t n,
                       float* out) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if ((metric < 0 || metric > 3) && metric != VQB_CHEBYSHEV) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "unknown metric %d", metric);
    if (rows == 0) return VQB_SUCCESS;
    if (!out || (n && (!a || !b))) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null data pointer");
    if (n > (size_t)INT32_MAX) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "vector too long");
    std::lock_guard<std::mutex> lk(ctx->mu);
    InputView ia, ib; OutputView ov;
    VQB_TRY(ia.bind(ctx, a, rows * n * sizeof(float)));
    VQB_TRY(ib.bind(ctx, b, rows * n * sizeof(float)));
    VQB_TRY(ov.bind(ctx, out, rows * sizeof(float)));
    k_distance_batch<<<cdiv(rows, 128), 128, 0, ctx->stream>>>(metric, static_cast<const float*>(ia.dev),
                                                               static_cast<const float*>(ib.dev), rows, (int)n,
                                                               static_cast<float*>(ov.dev));
    VQB_LAUNCHED(ctx);
    VQB_TRY(ov.finish(ctx));
    if (ia.was_host || ib.was_host || ov.host) VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}

}  // extern "C"
