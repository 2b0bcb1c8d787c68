// common.cuh -- context, error plumbing and pointer classification shared by the
// kernels of libvqb200.  sm_100a only; there is no CPU path anywhere in this library.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vqb200.h"

struct vqb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;   // all kernels
    bool own_stream = true;
    cudaStream_t copy_in = nullptr;  // H2D staging for host-pointer calls
    cudaStream_t copy_out = nullptr; // D2H staging
    std::string last_error;
    uint64_t launches = 0;
    std::mutex mu;                   // handles are immutable, the context is not: serialise calls
    // pinned mailbox for small per-iteration read-backs
    void* mailbox = nullptr;
    size_t mailbox_bytes = 0;
    // grow-only device staging for host-pointer calls (double-buffered x / codes / f16 chunks) and the
    // events that order the three streams: kept across calls so a call costs no cudaMalloc / cudaFree
    void* stage[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t stage_bytes[6] = {0, 0, 0, 0, 0, 0};
    cudaEvent_t stage_ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // grow-only device workspace of the training / tree-building calls (one slab, bump-allocated per call)
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // multi-GPU: NCCL communicator owned by the context (comm.cu; libnccl is loaded at run time), or null
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 1;
};

// Bump allocator over the context's grow-only workspace slab.  Usage: size pass (base == nullptr) to learn the
// total, vqb_ws_reserve, then the same sequence of take() calls with the real base.  Callers hold ctx->mu and all
// previous users of the slab are complete (every call that uses it synchronises before returning).
struct WsBump {
    char* base = nullptr;
    size_t off = 0;
    template <typename T> T* take(size_t count) {
        const size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};
inline cudaError_t vqb_ws_reserve(vqb_ctx* ctx, size_t bytes) {
    if (ctx->ws_bytes >= bytes) return cudaSuccess;
    if (ctx->ws) cudaFree(ctx->ws);
    ctx->ws = nullptr; ctx->ws_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->ws, bytes);
    if (e == cudaSuccess) ctx->ws_bytes = bytes;
    return e;
}

// multi-GPU plumbing (comm.cu): in-place float sum over the ranks of the context's communicator on ctx->stream
int vqb_comm_allreduce_f32(vqb_ctx* ctx, float* buf, size_t count);
void vqb_comm_release(vqb_ctx* ctx);

// Returns ctx->stage[i] grown to at least `bytes` (callers hold ctx->mu; previous users of the slot are
// complete because every host-pointer call synchronises its streams before returning).
inline cudaError_t vqb_stage(vqb_ctx* ctx, int i, size_t bytes, void** out) {
    if (ctx->stage_bytes[i] < bytes) {
        if (ctx->stage[i]) cudaFree(ctx->stage[i]);
        ctx->stage[i] = nullptr; ctx->stage_bytes[i] = 0;
        cudaError_t e = cudaMalloc(&ctx->stage[i], bytes);
        if (e != cudaSuccess) return e;
        ctx->stage_bytes[i] = bytes;
    }
    *out = ctx->stage[i];
    return cudaSuccess;
}

inline int vqb_fail(vqb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return code;
}

#define VQB_CUDA(ctx, expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return vqb_fail((ctx), VQB_FAILURE, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,   \
                            cudaGetErrorString(_e));                                           \
    } while (0)

#define VQB_TRY(expr)                    \
    do {                                 \
        int _s = (expr);                 \
        if (_s != VQB_SUCCESS) return _s; \
    } while (0)

// Kernel launch bookkeeping: counts launches (bench.py `gpu_launches`) and surfaces
// configuration errors immediately.
#define VQB_LAUNCHED(ctx)                                                                      \
    do {                                                                                       \
        (ctx)->launches++;                                                                     \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return vqb_fail((ctx), VQB_FAILURE, "%s:%d kernel launch -> %s", __FILE__,         \
                            __LINE__, cudaGetErrorString(_e));                                 \
    } while (0)

// true when `p` can be dereferenced by kernels running on ctx->device
inline bool vqb_is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// RAII device buffer tied to a stream-ordered free
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t b) {
        release();
        bytes = b;
        if (b == 0) return cudaSuccess;
        return cudaMalloc(&p, b);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Brings a caller buffer onto the device if it is a host pointer (synchronous w.r.t. ctx->stream
// ordering: the copy is enqueued on ctx->stream).  `dev` receives the usable device pointer.
struct InputView {
    DevBuf staging;
    const void* dev = nullptr;
    bool was_host = false;
    // slot >= 0: the device copy lives in the context's grow-only staging slot `slot` (no cudaMalloc / cudaFree per call;
    // the caller holds ctx->mu and synchronises before returning); slot < 0: a buffer owned by this view
    int bind(vqb_ctx* ctx, const void* p, size_t bytes, int slot = -1) {
        if (bytes == 0) { dev = p; return VQB_SUCCESS; }
        if (vqb_is_device_ptr(p)) { dev = p; return VQB_SUCCESS; }
        was_host = true;
        void* d = nullptr;
        if (slot >= 0) VQB_CUDA(ctx, vqb_stage(ctx, slot, bytes, &d));
        else { VQB_CUDA(ctx, staging.alloc(bytes)); d = staging.p; }
        VQB_CUDA(ctx, cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
        dev = d;
        return VQB_SUCCESS;
    }
};

// Device-side destination for a caller buffer; `finish` copies back to the host when needed.
struct OutputView {
    DevBuf staging;
    void* dev = nullptr;
    void* host = nullptr;
    size_t bytes = 0;
    int bind(vqb_ctx* ctx, void* p, size_t b) {
        bytes = b;
        if (!p || b == 0) { dev = p; return VQB_SUCCESS; }
        if (vqb_is_device_ptr(p)) { dev = p; return VQB_SUCCESS; }
        host = p;
        VQB_CUDA(ctx, staging.alloc(b));
        dev = staging.p;
        return VQB_SUCCESS;
    }
    int finish(vqb_ctx* ctx) {
        if (host && bytes)
            VQB_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return VQB_SUCCESS;
    }
};

// ---- chunked three-stream pipeline for calls whose buffers live on the HOST ------------------------------------------
// Units [0, n) (rows or elements) are processed in chunks: H2D on ctx->copy_in | kernels on ctx->stream | D2H on
// ctx->copy_out, double-buffered through the context's grow-only staging slots (0/1 input, 2/3 first output, 4/5 second
// output), so the PCIe copies of chunk c+1 and c-1 overlap the kernels of chunk c and no call allocates after the first.
// Any of the three buffers may also be a device pointer (then it is used in place).  Whatever happens -- including an
// error return half-way -- all three streams are drained before the call returns, so no copy is left writing into the
// caller's buffers or the staging slots.
struct ChunkIo {
    const void* in = nullptr;  size_t in_unit = 0;     // bytes per unit
    void* out0 = nullptr;      size_t out0_unit = 0;   // may be null
    void* out1 = nullptr;      size_t out1_unit = 0;   // may be null
};
inline size_t vqb_chunk_bytes() {
    static const size_t mb = [] { const char* e = std::getenv("VQB_CHUNK_MB"); long v = e ? std::atol(e) : 0; return (size_t)(v > 0 ? v : 32); }();
    return mb << 20;
}
template <typename Launch>   // int launch(const void* in_dev, void* out0_dev, void* out1_dev, size_t units, size_t unit0)
int vqb_chunk_pipeline(vqb_ctx* ctx, size_t n, const ChunkIo& io, Launch&& launch) {
    if (n == 0) return VQB_SUCCESS;
    const bool in_dev = vqb_is_device_ptr(io.in);
    const bool o0_dev = !io.out0 || vqb_is_device_ptr(io.out0);
    const bool o1_dev = !io.out1 || vqb_is_device_ptr(io.out1);
    if (in_dev && o0_dev && o1_dev) return launch(io.in, io.out0, io.out1, n, 0);   // all on the device: one asynchronous enqueue
    // chunk = 32 MB of the widest of the three streams of bytes (decode grows 32-fold on the way out)
    const size_t widest = std::max(std::max(io.in_unit, io.out0 ? io.out0_unit : 0), std::max<size_t>(io.out1 ? io.out1_unit : 0, 1));
    size_t chunk = std::max<size_t>(1, vqb_chunk_bytes() / widest);
    chunk = std::min(chunk, n);
    if (chunk == n && n * widest <= (size_t)256 * 1024) {
        // small call (the reference's single-vector quantize): nothing to overlap -- upload, kernels and downloads in
        // order on the context stream, one synchronisation, no events and no second / third stream
        void *s_in = nullptr, *s0 = nullptr, *s1 = nullptr;
        if (!in_dev) VQB_CUDA(ctx, vqb_stage(ctx, 0, n * io.in_unit, &s_in));
        if (io.out0 && !o0_dev) VQB_CUDA(ctx, vqb_stage(ctx, 2, n * io.out0_unit, &s0));
        if (io.out1 && !o1_dev) VQB_CUDA(ctx, vqb_stage(ctx, 4, n * io.out1_unit, &s1));
        struct Drain1 { vqb_ctx* c; ~Drain1() { cudaStreamSynchronize(c->stream); } } drain1{ctx};
        const void* din = io.in;
        if (!in_dev) { VQB_CUDA(ctx, cudaMemcpyAsync(s_in, io.in, n * io.in_unit, cudaMemcpyHostToDevice, ctx->stream)); din = s_in; }
        void* d0 = io.out0 ? (o0_dev ? io.out0 : s0) : nullptr;
        void* d1 = io.out1 ? (o1_dev ? io.out1 : s1) : nullptr;
        VQB_TRY(launch(din, d0, d1, n, 0));
        if (io.out0 && !o0_dev) VQB_CUDA(ctx, cudaMemcpyAsync(io.out0, s0, n * io.out0_unit, cudaMemcpyDeviceToHost, ctx->stream));
        if (io.out1 && !o1_dev) VQB_CUDA(ctx, cudaMemcpyAsync(io.out1, s1, n * io.out1_unit, cudaMemcpyDeviceToHost, ctx->stream));
        return VQB_SUCCESS;   // ~Drain1 synchronises the stream
    }
    for (int i = 0; i < 7; ++i)
        if (!ctx->stage_ev[i]) VQB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
    void *sin[2] = {nullptr, nullptr}, *so0[2] = {nullptr, nullptr}, *so1[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; ++b) {
        if (!in_dev) VQB_CUDA(ctx, vqb_stage(ctx, b, chunk * io.in_unit, &sin[b]));
        if (io.out0 && !o0_dev) VQB_CUDA(ctx, vqb_stage(ctx, 2 + b, chunk * io.out0_unit, &so0[b]));
        if (io.out1 && !o1_dev) VQB_CUDA(ctx, vqb_stage(ctx, 4 + b, chunk * io.out1_unit, &so1[b]));
    }
    cudaEvent_t* ev_in = &ctx->stage_ev[0];     // [2] chunk uploaded
    cudaEvent_t* ev_comp = &ctx->stage_ev[2];   // [2] chunk computed
    cudaEvent_t* ev_out = &ctx->stage_ev[4];    // [2] chunk downloaded
    struct Drain {   // no copy or kernel of this call outlives it, on any return path
        vqb_ctx* c;
        ~Drain() { cudaStreamSynchronize(c->copy_in); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->copy_out); }
    } drain{ctx};
    // order the pipeline after whatever is already queued on the context stream
    VQB_CUDA(ctx, cudaEventRecord(ctx->stage_ev[6], ctx->stream));
    VQB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->stage_ev[6], 0));
    const size_t n_chunks = (n + chunk - 1) / chunk;
    for (size_t c = 0; c < n_chunks; ++c) {
        const int b = (int)(c & 1);
        const size_t u0 = c * chunk, units = std::min(chunk, n - u0);
        const void* din = static_cast<const char*>(io.in) + u0 * io.in_unit;
        if (!in_dev) {
            if (c >= 2) VQB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ev_comp[b], 0));   // staging b consumed
            VQB_CUDA(ctx, cudaMemcpyAsync(sin[b], din, units * io.in_unit, cudaMemcpyHostToDevice, ctx->copy_in));
            VQB_CUDA(ctx, cudaEventRecord(ev_in[b], ctx->copy_in));
            VQB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_in[b], 0));
            din = sin[b];
        }
        if (c >= 2) VQB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_out[b], 0));           // output staging b drained
        void* d0 = io.out0 ? (o0_dev ? static_cast<char*>(io.out0) + u0 * io.out0_unit : so0[b]) : nullptr;
        void* d1 = io.out1 ? (o1_dev ? static_cast<char*>(io.out1) + u0 * io.out1_unit : so1[b]) : nullptr;
        VQB_TRY(launch(din, d0, d1, units, u0));
        VQB_CUDA(ctx, cudaEventRecord(ev_comp[b], ctx->stream));
        VQB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ev_comp[b], 0));
        if (io.out0 && !o0_dev)
            VQB_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(io.out0) + u0 * io.out0_unit, so0[b], units * io.out0_unit,
                                          cudaMemcpyDeviceToHost, ctx->copy_out));
        if (io.out1 && !o1_dev)
            VQB_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(io.out1) + u0 * io.out1_unit, so1[b], units * io.out1_unit,
                                          cudaMemcpyDeviceToHost, ctx->copy_out));
        VQB_CUDA(ctx, cudaEventRecord(ev_out[b], ctx->copy_out));
    }
    return VQB_SUCCESS;   // ~Drain synchronises the three streams
}

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// internal entry points shared between translation units
int vqb_pq_assign_exact_launch(vqb_ctx* ctx, int metric_kind, const float* x, size_t n, size_t dim,
                               size_t m, size_t k, size_t sub_dim, const float* codebooks,
                               const int* sub_list_dev, int n_sub, void* codes, uint32_t code_bytes,
                               size_t code_stride_row, size_t code_stride_sub, __half* recon,
                               const int* n_sub_dev = nullptr, const uint32_t* go = nullptr);

// tiled Manhattan assignment for sub_dim 8, k <= 256 (pq_assign.cu): coalesced row tiles, codebooks resident in smem
bool vqb_l1_tiles_supported(int metric_kind, size_t sub_dim, size_t k);
int vqb_assign_l1_tiles_launch(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k, const float* codebooks,
                               void* codes, uint32_t code_bytes, size_t code_stride_row, size_t code_stride_sub, __half* recon);

// tensor-core (tcgen05) GEMM-form assignment, pq_tc.cu.  `prep` is a device workspace of
// vqb_tc_prep_bytes(m, sub_dim) bytes filled by vqb_tc_prepare from the current codebooks.  sub_dim 8, 16, 24 or 32.
size_t vqb_tc_prep_bytes(size_t m, size_t sub_dim);
bool vqb_tc_supported(int metric_kind, const float* x, size_t n, size_t dim, size_t m, size_t k, size_t sub_dim);
int vqb_tc_prepare(vqb_ctx* ctx, int metric_kind, const float* codebooks, size_t m, size_t k, size_t sub_dim, void* prep,
                   const uint32_t* go = nullptr);
// the 2-D tensor map over a row-major f32 matrix x[n, dim] used by the TMA-fed kernels: box = box_cols floats x 128 rows
// (32 columns: SWIZZLE_128B, one 128-byte line per row; otherwise unswizzled rows of box_cols floats)
struct CUtensorMap_st;
// swizzle: -1 = by width (32 columns swizzled), 0 = plain rows, 1 = SWIZZLE_128B (32 columns only)
int vqb_make_x_tensormap(vqb_ctx* ctx, const float* x, size_t n, size_t dim, CUtensorMap_st* out, int box_cols = 32,
                         int swizzle = -1);
int vqb_tc_assign_launch(vqb_ctx* ctx, int metric_kind, const float* x, size_t n, size_t dim, size_t m, size_t k,
                         const void* prep, const int* active_dev, void* codes, uint32_t code_bytes,
                         size_t code_stride_row, size_t code_stride_sub, __half* recon,
                         float* dbg_scores = nullptr, unsigned long long* dbg_stats = nullptr, int dbg_sub = 0,
                         unsigned long long* dbg_ts = nullptr, int dbg_ts_units = 0, const uint32_t* go = nullptr);
// rows below which VQB_ASSIGN_AUTO keeps the CUDA-core kernel (the tensor kernel stages 128 KB of
// codebooks per CTA before its first tile)
constexpr size_t VQB_TC_MIN_ROWS = 1024;

// metric_kind for the exact kernel: the four Distance variants + the training distance
enum { MK_SQEUCLID = 0, MK_EUCLID = 1, MK_MANHATTAN = 2, MK_COSINE = 3, MK_TRAIN = 4, MK_CHEBYSHEV = 5 };
