// ctx.cu -- context lifecycle, memory helpers and the backend string of libvqb200.
#include "common.cuh"

extern "C" {

int vqb_ctx_create(int device, vqb_ctx** out) {
    if (!out) return VQB_ERR_NULL_PTR;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return VQB_ERR_UNSUPPORTED_DEVICE;  // no CPU fallback: fail loudly
    }
    if (device < 0 || device >= count) return VQB_ERR_INVALID_INPUT;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VQB_FAILURE;
    if (prop.major != 10) return VQB_ERR_UNSUPPORTED_DEVICE;  // kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return VQB_FAILURE;
    vqb_ctx* c = new vqb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return VQB_FAILURE;
    }
    c->mailbox_bytes = 1 << 20;
    if (cudaHostAlloc(&c->mailbox, c->mailbox_bytes, cudaHostAllocDefault) != cudaSuccess) {
        delete c;
        return VQB_FAILURE;
    }
    *out = c;
    return VQB_SUCCESS;
}

int vqb_ctx_destroy(vqb_ctx* ctx) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
    vqb_comm_release(ctx);
    if (ctx->ws) cudaFree(ctx->ws);
    for (int i = 0; i < 6; ++i) if (ctx->stage[i]) cudaFree(ctx->stage[i]);
    for (int i = 0; i < 7; ++i) if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    delete ctx;
    return VQB_SUCCESS;
}

int vqb_ctx_synchronize(vqb_ctx* ctx) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}

void* vqb_ctx_stream(vqb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int vqb_ctx_set_stream(vqb_ctx* ctx, void* s) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)s;
    ctx->own_stream = false;
    return VQB_SUCCESS;
}

const char* vqb_last_error(vqb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

const char* vqb_backend_name(void) { return "vqb200 (sm_100a CUDA: tcgen05/TMEM + CUDA-core kernels)"; }

uint64_t vqb_ctx_launch_count(vqb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int vqb_malloc(vqb_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return VQB_ERR_NULL_PTR;
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    VQB_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return VQB_SUCCESS;
}

int vqb_free(vqb_ctx* ctx, void* dptr) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if (dptr) VQB_CUDA(ctx, cudaFree(dptr));
    return VQB_SUCCESS;
}

int vqb_host_alloc(vqb_ctx* ctx, size_t bytes, void** hptr) {
    if (!ctx || !hptr) return VQB_ERR_NULL_PTR;
    VQB_CUDA(ctx, cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return VQB_SUCCESS;
}

int vqb_host_free(vqb_ctx* ctx, void* hptr) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if (hptr) VQB_CUDA(ctx, cudaFreeHost(hptr));
    return VQB_SUCCESS;
}

int vqb_memcpy(vqb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if (bytes == 0) return VQB_SUCCESS;
    if (!dst || !src) return VQB_ERR_NULL_PTR;
    VQB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}

}  // extern "C"
