// comm.cu -- multi-GPU exchange owned by the library: one process per GPU, one NCCL communicator per context.
//
// PQ training shards rows over the GPUs of a box and all-reduces ONE fused buffer
// [sums | count_lo | count_hi] per k-means iteration (src/core/vector.rs:432-447 computed on row shards); encode /
// BQ / SQ shard rows with no collective.  The library owns the communicator so that every host language gets the
// multi-GPU path through the C ABI (vqb_comm_unique_id on rank 0 -> ship the 128 bytes to the other ranks by any
// means -> vqb_comm_init_rank on every rank); the vqb_allreduce_fn callback of vqb_train_opts stays as an override.
//
// libnccl.so.2 is opened at run time (dlopen): a single-GPU host needs no NCCL, and inside a process that already
// carries an NCCL (PyTorch) the same library instance is reused instead of a second copy being linked in.
#include "common.cuh"

#include <dlfcn.h>

namespace {

// the handful of NCCL declarations used here (nccl.h, stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
enum { NCCL_SUM = 0, NCCL_FLOAT32 = 7 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char* s) { void* f = dlsym(api.lib, s); if (!f && api.error.empty()) api.error = std::string("libnccl lacks ") + s; return f; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

}  // namespace

int vqb_comm_allreduce_f32(vqb_ctx* ctx, float* buf, size_t count) {
    if (!ctx->nccl_comm) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "the context has no communicator (vqb_comm_init_rank)");
    if (count == 0) return VQB_SUCCESS;
    NcclApi& a = nccl();
    ncclResult_t r = a.AllReduce(buf, buf, count, NCCL_FLOAT32, NCCL_SUM, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream);
    if (r != 0) return vqb_fail(ctx, VQB_FAILURE, "ncclAllReduce failed: %s", a.GetErrorString ? a.GetErrorString(r) : "?");
    return VQB_SUCCESS;
}

void vqb_comm_release(vqb_ctx* ctx) {
    if (ctx->nccl_comm) {
        NcclApi& a = nccl();
        if (a.CommDestroy) a.CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
        ctx->nccl_comm = nullptr;
    }
    ctx->comm_rank = 0; ctx->comm_world = 1;
}

extern "C" {

int vqb_comm_unique_id(void* id_out) {
    if (!id_out) return VQB_ERR_NULL_PTR;
    NcclApi& a = nccl();
    if (!a.error.empty() || !a.GetUniqueId) return VQB_ERR_UNSUPPORTED_DEVICE;
    ncclUniqueId id;
    if (a.GetUniqueId(&id) != 0) return VQB_FAILURE;
    std::memcpy(id_out, id.internal, sizeof(id.internal));
    return VQB_SUCCESS;
}

int vqb_comm_init_rank(vqb_ctx* ctx, const void* id, int rank, int world) {
    if (!ctx || !id) return VQB_ERR_NULL_PTR;
    if (world < 1 || rank < 0 || rank >= world) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "bad rank %d / world %d", rank, world);
    std::lock_guard<std::mutex> lk(ctx->mu);
    NcclApi& a = nccl();
    if (!a.error.empty()) return vqb_fail(ctx, VQB_ERR_UNSUPPORTED_DEVICE, "%s", a.error.c_str());
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    vqb_comm_release(ctx);
    ncclUniqueId uid;
    std::memcpy(uid.internal, id, sizeof(uid.internal));
    ncclComm_t comm = nullptr;
    ncclResult_t r = a.CommInitRank(&comm, world, uid, rank);
    if (r != 0) return vqb_fail(ctx, VQB_FAILURE, "ncclCommInitRank failed: %s", a.GetErrorString(r));
    ctx->nccl_comm = comm; ctx->comm_rank = rank; ctx->comm_world = world;
    return VQB_SUCCESS;
}

int vqb_comm_destroy(vqb_ctx* ctx) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStreamSynchronize(ctx->stream);
    vqb_comm_release(ctx);
    return VQB_SUCCESS;
}

int vqb_comm_info(vqb_ctx* ctx, int* rank, int* world) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if (rank) *rank = ctx->comm_rank;
    if (world) *world = ctx->nccl_comm ? ctx->comm_world : 1;
    return VQB_SUCCESS;
}

/* In-place float sum over the communicator's ranks (device pointer), enqueued on the context stream: the exchange the
 * training loop performs, exposed so hosts and tests can use the library-owned communicator directly. */
int vqb_comm_allreduce(vqb_ctx* ctx, float* buf, size_t count) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    if (count && !buf) return VQB_ERR_NULL_PTR;
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    return vqb_comm_allreduce_f32(ctx, buf, count);
}

}  // extern "C"
