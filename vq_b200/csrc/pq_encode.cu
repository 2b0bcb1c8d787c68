// pq_encode.cu -- the immutable ProductQuantizer handle and its batch encode / decode.
//
// Reference: ProductQuantizer::quantize / dequantize (src/pq.rs:167-209).  The reference encodes
// one vector per call and returns the reconstructed centroid values as Vec<f16>; the batch form
// here returns the same f16 values (recon_out) and/or the compact code indices (codes_out).
//
// Host-pointer calls are pipelined in row chunks over three streams (H2D | kernels | D2H) with
// double-buffered device staging, so the PCIe copies overlap the kernels; device-pointer calls
// are a single asynchronous launch on the context stream.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

struct vqb_pq {
    vqb_ctx* ctx = nullptr;
    size_t m = 0, k = 0, d = 0;
    int metric = 0;
    DevBuf cb;       // [m][k][d] f32
    DevBuf tc_prep;  // prepared tensor-core operand images (pq_tc.cu), built once at creation
    bool tc_ready = false;
};

namespace {

// decode: out[row][s*d + t] = f16->f32(f16(codebook[s][code][t]))  (pq.rs:193-195 then :201-209)
__global__ void k_pq_decode(const void* __restrict__ codes, uint32_t code_bytes, size_t n, int m, int k, int d,
                            const float* __restrict__ cb, float* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = n * (size_t)m * d;
    if (t >= total) return;
    int comp = (int)(t % d);
    size_t rs = t / d;  // row * m + s
    int s = (int)(rs % m);
    uint32_t c = code_bytes == 1 ? static_cast<const uint8_t*>(codes)[rs]
               : code_bytes == 2 ? static_cast<const uint16_t*>(codes)[rs]
                                 : static_cast<const uint32_t*>(codes)[rs];
    if (c >= (uint32_t)k) c = (uint32_t)k - 1;  // defensive: never read outside the codebook
    out[t] = __half2float(__float2half_rn(cb[((size_t)s * k + c) * d + comp]));
}

// decode, vector form for sub_dim % 4 == 0: thread = one float4 of the output, so a warp's store covers 512
// contiguous bytes; the codebook (m*k*d*4 bytes, L1/L2 resident) is read through the read-only path
template <int CB>
__global__ void __launch_bounds__(256) k_pq_decode_v4(const void* __restrict__ codes, size_t n, int m, int k, int d,
                                                      const float* __restrict__ cb, float* __restrict__ out) {
    const int q4 = d >> 2;                       // float4s per (row, subspace)
    const size_t total = n * (size_t)m * q4;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t rs = t / q4;
        const int q = (int)(t - rs * q4);
        const int s = (int)(rs % m);
        uint32_t c = CB == 1 ? static_cast<const uint8_t*>(codes)[rs]
                   : CB == 2 ? static_cast<const uint16_t*>(codes)[rs]
                             : static_cast<const uint32_t*>(codes)[rs];
        if (c >= (uint32_t)k) c = (uint32_t)k - 1;  // defensive: never read outside the codebook
        float4 v = __ldg(reinterpret_cast<const float4*>(cb + ((size_t)s * k + c) * d) + q);
        v.x = __half2float(__float2half_rn(v.x)); v.y = __half2float(__float2half_rn(v.y));
        v.z = __half2float(__float2half_rn(v.z)); v.w = __half2float(__float2half_rn(v.w));
        __stcs(reinterpret_cast<float4*>(out) + t, v);
    }
}

// non-owning views of the staging buffers / events cached in the context
struct Buf {
    void* p = nullptr;
    template <typename T> T* as() const { return static_cast<T*>(p); }
};
struct Ev { cudaEvent_t e = nullptr; };

int encode_device(vqb_pq* pq, const float* x, size_t n, uint32_t assign_mode, void* codes, uint32_t code_bytes,
                  __half* recon, cudaStream_t stream_override = nullptr) {
    vqb_ctx* ctx = pq->ctx;
    const size_t dim = pq->m * pq->d;
    const bool tc_ok = pq->tc_ready && vqb_tc_supported(pq->metric, x, n, dim, pq->m, pq->k, pq->d);
    if (assign_mode == VQB_ASSIGN_TENSOR && !tc_ok)
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT,
                        "tensor-core assignment needs sub_dim 8, 16, 24 or 32, k <= 256, a contraction metric and 16-byte aligned rows");
    const bool use_tc = tc_ok && (assign_mode == VQB_ASSIGN_TENSOR || (assign_mode == VQB_ASSIGN_AUTO && n >= VQB_TC_MIN_ROWS));
    cudaStream_t saved = ctx->stream;
    if (stream_override) ctx->stream = stream_override;
    int rc;
    if (use_tc)
        rc = vqb_tc_assign_launch(ctx, pq->metric, x, n, dim, pq->m, pq->k, pq->tc_prep.p, nullptr, codes, code_bytes,
                                  /*stride_row=*/pq->m, /*stride_sub=*/1, recon);
    else if (vqb_l1_tiles_supported(pq->metric, pq->d, pq->k) && n >= 4096 && assign_mode != VQB_ASSIGN_EXACT)
        rc = vqb_assign_l1_tiles_launch(ctx, x, n, dim, pq->m, pq->k, pq->cb.as<float>(), codes, code_bytes,
                                        /*stride_row=*/pq->m, /*stride_sub=*/1, recon);
    else
        rc = vqb_pq_assign_exact_launch(ctx, pq->metric, x, n, pq->m * pq->d, pq->m, pq->k, pq->d,
                                        pq->cb.as<float>(), nullptr, (int)pq->m, codes, code_bytes,
                                        /*stride_row=*/pq->m, /*stride_sub=*/1, recon);
    ctx->stream = saved;
    return rc;
}

}  // namespace

extern "C" {

int vqb_pq_create(vqb_ctx* ctx, const float* codebooks, size_t m, size_t k, size_t sub_dim, int metric,
                  vqb_pq** out) {
    if (!ctx || !out) return VQB_ERR_NULL_PTR;
    *out = nullptr;
    if (!codebooks) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null codebooks");
    if (m == 0 || k == 0 || sub_dim == 0) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "m, k, sub_dim must be > 0");
    if ((metric < 0 || metric > 3) && metric != VQB_CHEBYSHEV) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "unknown metric %d", metric);
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    vqb_pq* p = new vqb_pq();
    p->ctx = ctx; p->m = m; p->k = k; p->d = sub_dim; p->metric = metric;
    size_t bytes = m * k * sub_dim * sizeof(float);
    cudaError_t e = p->cb.alloc(bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->cb.p, codebooks, bytes, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess && metric != VQB_MANHATTAN && metric != VQB_CHEBYSHEV && k <= 256 &&
        (sub_dim == 8 || sub_dim == 16 || sub_dim == 24 || sub_dim == 32)) {
        e = p->tc_prep.alloc(vqb_tc_prep_bytes(m, sub_dim));
        if (e == cudaSuccess && vqb_tc_prepare(ctx, metric, p->cb.as<float>(), m, k, sub_dim, p->tc_prep.p) == VQB_SUCCESS)
            p->tc_ready = true;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        delete p;
        return vqb_fail(ctx, VQB_FAILURE, "codebook upload failed: %s", cudaGetErrorString(e));
    }
    *out = p;
    return VQB_SUCCESS;
}

int vqb_pq_destroy(vqb_pq* pq) {
    if (!pq) return VQB_ERR_NULL_PTR;
    cudaStreamSynchronize(pq->ctx->stream);
    delete pq;
    return VQB_SUCCESS;
}

int vqb_pq_codebooks(vqb_pq* pq, float* out) {
    if (!pq || !out) return VQB_ERR_NULL_PTR;
    vqb_ctx* ctx = pq->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaMemcpyAsync(out, pq->cb.p, pq->cb.bytes, cudaMemcpyDefault, ctx->stream));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}

int vqb_pq_encode(vqb_pq* pq, const float* x, size_t n, uint32_t assign_mode, void* codes_out, uint32_t code_bytes,
                  uint16_t* recon_out) {
    if (!pq) return VQB_ERR_NULL_PTR;
    vqb_ctx* ctx = pq->ctx;
    if (n == 0) return VQB_SUCCESS;
    if (!x) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null input");
    if (!codes_out && !recon_out) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "no output requested");
    if (codes_out) {
        if (code_bytes != 1 && code_bytes != 2 && code_bytes != 4)
            return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "code_bytes must be 1, 2 or 4");
        if ((code_bytes == 1 && pq->k > 256) || (code_bytes == 2 && pq->k > 65536))
            return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "k = %zu does not fit %u-byte codes", pq->k, code_bytes);
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t dim = pq->m * pq->d;
    const bool x_dev = vqb_is_device_ptr(x);
    const bool c_dev = !codes_out || vqb_is_device_ptr(codes_out);
    const bool r_dev = !recon_out || vqb_is_device_ptr(recon_out);
    (void)x_dev; (void)c_dev; (void)r_dev;
    // host buffers: row chunks over three streams (common.cuh); device buffers: one asynchronous launch
    ChunkIo io;
    io.in = x; io.in_unit = dim * sizeof(float);
    io.out0 = codes_out; io.out0_unit = pq->m * code_bytes;
    io.out1 = recon_out; io.out1_unit = dim * sizeof(uint16_t);
    return vqb_chunk_pipeline(ctx, n, io, [&](const void* din, void* d0, void* d1, size_t rows, size_t) -> int {
        return encode_device(pq, static_cast<const float*>(din), rows, assign_mode, d0, code_bytes, static_cast<__half*>(d1));
    });
}

int vqb_pq_decode(vqb_pq* pq, const void* codes, uint32_t code_bytes, size_t n, float* out) {
    if (!pq) return VQB_ERR_NULL_PTR;
    vqb_ctx* ctx = pq->ctx;
    if (n == 0) return VQB_SUCCESS;
    if (!codes || !out) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null pointer");
    if (code_bytes != 1 && code_bytes != 2 && code_bytes != 4)
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "code_bytes must be 1, 2 or 4");
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t dim = pq->m * pq->d;
    ChunkIo io;
    io.in = codes; io.in_unit = pq->m * code_bytes;
    io.out0 = out; io.out0_unit = dim * sizeof(float);
    return vqb_chunk_pipeline(ctx, n, io, [&](const void* din, void* d0, void*, size_t rows, size_t) -> int {
        float* o = static_cast<float*>(d0);
        if (pq->d % 4 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            const unsigned grid = (unsigned)std::min<size_t>(cdiv(rows * pq->m * (pq->d / 4), 256), (size_t)ctx->sm_count * 64);
            if (code_bytes == 1) k_pq_decode_v4<1><<<grid, 256, 0, ctx->stream>>>(din, rows, (int)pq->m, (int)pq->k, (int)pq->d, pq->cb.as<float>(), o);
            else if (code_bytes == 2) k_pq_decode_v4<2><<<grid, 256, 0, ctx->stream>>>(din, rows, (int)pq->m, (int)pq->k, (int)pq->d, pq->cb.as<float>(), o);
            else k_pq_decode_v4<4><<<grid, 256, 0, ctx->stream>>>(din, rows, (int)pq->m, (int)pq->k, (int)pq->d, pq->cb.as<float>(), o);
        } else {
            k_pq_decode<<<cdiv(rows * dim, 256), 256, 0, ctx->stream>>>(din, code_bytes, rows, (int)pq->m, (int)pq->k, (int)pq->d,
                                                                       pq->cb.as<float>(), o);
        }
        VQB_LAUNCHED(ctx);
        return VQB_SUCCESS;
    });
}

}  // extern "C"
