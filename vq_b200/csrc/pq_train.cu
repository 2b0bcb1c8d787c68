// pq_train.cu -- LBG / k-means codebook training for all m subspaces at once.
//
// Reference: lbg_quantize (src/core/vector.rs:390-461) called serially per subspace by
// ProductQuantizer::new (src/pq.rs:121-132).  Here every active subspace advances one
// iteration per pass ("level-synchronous" over subspaces); a subspace that reports
// `!changed` is frozen exactly where the reference would `break` (vector.rs:455-457).
//
// One iteration =
//   assign      nearest centroid per (row, subspace)                 [pq_assign.cu / pq_tc.cu]
//   group       stable 8-bit radix pass of the row ids by code       (vector.rs:432-435)
//   sum         per (subspace, cluster, component) chain over the member ids in ascending
//               order, sequential f32 adds                           (vector.rs:368-384)
//   [allreduce] one fused buffer [sums | count_lo | count_hi] when rows are sharded over GPUs
//   finalize    mean = sum / count, epsilon test, changed flags      (vector.rs:438-447)
//   reseed      host callback per empty cluster, ascending j         (vector.rs:448-452)
//
// HBM layout: X row-major [n, dim] f32 stays resident; codes [m][n] (1/2/4 bytes);
// sorted ids [m][n] u32; sums [m][k][sub_dim] f32.  All per-subspace arrays are indexed by
// the subspace id; kernels run over the list of still-active subspaces (grid.y).
#include "common.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>

namespace {

// VQB_TRACE=1: per-iteration host-side phase times on stderr (diagnostics only)
inline bool trace_on() {
    static const bool on = [] { const char* e = std::getenv("VQB_TRACE"); return e && *e && *e != '0'; }();
    return on;
}
inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

constexpr int RX_THREADS = 256;   // one radix step = 256 consecutive rows
constexpr int RX_STEPS = 32;      // steps per chunk: (chunk, bin) runs of ~32 ids = 128 B keep the scatter's writes line-sized
constexpr int RX_CHUNK = RX_THREADS * RX_STEPS;
constexpr int RX_WARPS = RX_THREADS / 32;

__device__ __forceinline__ uint32_t load_code(const void* codes, uint32_t code_bytes, size_t off) {
    if (code_bytes == 1) return static_cast<const uint8_t*>(codes)[off];
    if (code_bytes == 2) return static_cast<const uint16_t*>(codes)[off];
    return static_cast<const uint32_t*>(codes)[off];
}

// ---- stable radix pass over digit (code >> shift) & 255 -------------------------------------
// grid (n_chunks, n_active).  ids_in == nullptr means the identity permutation.
__global__ void __launch_bounds__(RX_THREADS)
k_radix_hist(const void* __restrict__ codes, uint32_t code_bytes, size_t n, int shift,
             const int* __restrict__ sub_list, const uint32_t* __restrict__ ids_in,
             uint32_t* __restrict__ chunk_hist, int n_chunks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t s = sub_list[blockIdx.y];
    const size_t base = (size_t)blockIdx.x * RX_CHUNK;
    for (int st = 0; st < RX_STEPS; ++st) {
        size_t i = base + (size_t)st * RX_THREADS + threadIdx.x;
        if (i < n) {
            size_t row = ids_in ? ids_in[s * n + i] : i;
            uint32_t dgt = (load_code(codes, code_bytes, s * n + row) >> shift) & 255u;
            atomicAdd(&h[dgt], 1u);
        }
    }
    __syncthreads();
    chunk_hist[(s * n_chunks + blockIdx.x) * 256 + threadIdx.x] = h[threadIdx.x];
}

// grid (n_active), 256 threads (thread = bin): exclusive scan over chunks, then over bins.
// When seg_beg != nullptr (single-pass case, digit == cluster id) also emits the member ranges.
__global__ void __launch_bounds__(256)
k_radix_scan(uint32_t* __restrict__ chunk_hist, int n_chunks, const int* __restrict__ sub_list,
             uint32_t* __restrict__ bin_off, uint32_t* __restrict__ seg_beg,
             uint32_t* __restrict__ seg_end, int k) {
    __shared__ uint32_t tot[256];
    const size_t s = sub_list[blockIdx.x];
    uint32_t run = 0;
    for (int c = 0; c < n_chunks; ++c) {
        size_t idx = (s * n_chunks + c) * 256 + threadIdx.x;
        uint32_t v = chunk_hist[idx];
        chunk_hist[idx] = run;  // becomes the chunk's base within its bin
        run += v;
    }
    tot[threadIdx.x] = run;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int b = 0; b < 256; ++b) { uint32_t v = tot[b]; tot[b] = acc; acc += v; }
    }
    __syncthreads();
    bin_off[s * 256 + threadIdx.x] = tot[threadIdx.x];
    if (seg_beg && (int)threadIdx.x < k) {
        seg_beg[s * k + threadIdx.x] = tot[threadIdx.x];
        seg_end[s * k + threadIdx.x] = tot[threadIdx.x] + run;
    }
}

__global__ void __launch_bounds__(RX_THREADS)
k_radix_scatter(const void* __restrict__ codes, uint32_t code_bytes, size_t n, int shift,
                const int* __restrict__ sub_list, const uint32_t* __restrict__ ids_in,
                const uint32_t* __restrict__ chunk_base, const uint32_t* __restrict__ bin_off,
                int n_chunks, uint32_t* __restrict__ ids_out) {
    __shared__ uint32_t running[256];
    __shared__ uint16_t wcnt[RX_WARPS][256];
    const size_t s = sub_list[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    running[threadIdx.x] = bin_off[s * 256 + threadIdx.x] +
                           chunk_base[(s * n_chunks + blockIdx.x) * 256 + threadIdx.x];
    const size_t base = (size_t)blockIdx.x * RX_CHUNK;
    for (int st = 0; st < RX_STEPS; ++st) {
        for (int w = 0; w < RX_WARPS; ++w) wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        size_t i = base + (size_t)st * RX_THREADS + threadIdx.x;
        bool live = i < n;
        uint32_t row = 0, dgt = 0xFFFFFFFFu;
        if (live) {
            row = ids_in ? ids_in[s * n + i] : (uint32_t)i;
            dgt = (load_code(codes, code_bytes, s * n + row) >> shift) & 255u;
        }
        // rank among the lanes of this warp holding the same digit (dead lanes share 0xFFFFFFFF)
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, dgt);
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (live && rank == 0) wcnt[warp][dgt] = (uint16_t)__popc(peers);
        __syncthreads();
        if (live) {
            uint32_t before = 0;
            for (int w = 0; w < warp; ++w) before += wcnt[w][dgt];
            ids_out[s * n + running[dgt] + before + rank] = row;
        }
        __syncthreads();
        uint32_t add = 0;
        for (int w = 0; w < RX_WARPS; ++w) add += wcnt[w][threadIdx.x];
        running[threadIdx.x] += add;
        __syncthreads();
    }
}

// Member ranges from the fully sorted order (multi-pass case, k > 256).  seg_beg/seg_end must be
// zeroed first; grid (cdiv(n,256), n_active).
__global__ void k_cluster_ranges(const void* __restrict__ codes, uint32_t code_bytes, size_t n, int k,
                                 const int* __restrict__ sub_list, const uint32_t* __restrict__ ids,
                                 uint32_t* __restrict__ seg_beg, uint32_t* __restrict__ seg_end) {
    const size_t s = sub_list[blockIdx.y];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = load_code(codes, code_bytes, s * n + ids[s * n + i]);
    uint32_t prev = i ? load_code(codes, code_bytes, s * n + ids[s * n + i - 1]) : 0xFFFFFFFFu;
    uint32_t next = (i + 1 < n) ? load_code(codes, code_bytes, s * n + ids[s * n + i + 1]) : 0xFFFFFFFFu;
    if (c != prev) seg_beg[s * k + c] = (uint32_t)i;
    if (c != next) seg_end[s * k + c] = (uint32_t)i + 1;
}

// ---- per-(subspace, cluster, component) ordered sums ------------------------------------------
// thread = (active pos, cluster j, segment, component).  segs == 1 reproduces the reference's
// summation order exactly; segs > 1 splits each member list into equal contiguous pieces.
__global__ void __launch_bounds__(256)
k_chain_sums(const float* __restrict__ x, size_t n, size_t row_stride, size_t sub_stride, int d, int k, int segs,
             const int* __restrict__ sub_list, int n_active, const uint32_t* __restrict__ ids,
             const uint32_t* __restrict__ seg_beg, const uint32_t* __restrict__ seg_end,
             float* __restrict__ partial /* [m][k][segs][d] */) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n_active * k * segs * d;
    if (t >= total) return;
    int comp = (int)(t % d);
    size_t c = t / d;
    int seg = (int)(c % segs); c /= segs;
    int j = (int)(c % k);
    size_t s = sub_list[c / k];
    uint32_t off = seg_beg[s * k + j], cnt = seg_end[s * k + j] - off;
    uint32_t b = off + (uint32_t)(((uint64_t)cnt * seg) / segs);
    uint32_t e = off + (uint32_t)(((uint64_t)cnt * (seg + 1)) / segs);
    const uint32_t* idp = ids + s * n;
    // x is either the caller's row-major matrix (row_stride = dim, sub_stride = d) or the subspace-major
    // copy [m][n][d] (row_stride = d, sub_stride = n*d): the latter keeps the rows of one subspace adjacent,
    // so the member gathers of neighbouring chains land in the same DRAM pages / L2 lines
    const float* xs = x + s * sub_stride + comp;
    float acc = 0.0f;
    uint32_t i = b;
    for (; i + 8 <= e; i += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(xs + (size_t)__ldg(idp + i + u) * row_stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, v[u]);
    }
    for (; i < e; ++i) acc = __fadd_rn(acc, __ldg(xs + (size_t)__ldg(idp + i) * row_stride));
    partial[((s * k + j) * segs + seg) * d + comp] = acc;
}

// Packs [sums | count_lo | count_hi] (all f32) ; frozen subspaces are zeroed so that a cross-rank
// sum leaves them inert.
__global__ void k_pack(const float* __restrict__ partial, int segs, int d, int k, int m,
                       const uint32_t* __restrict__ seg_beg, const uint32_t* __restrict__ seg_end,
                       const int* __restrict__ is_active, float* __restrict__ pack) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n_sum = (size_t)m * k * d, n_c = (size_t)m * k;
    if (t < n_sum) {
        size_t c = t / d; int comp = (int)(t % d);
        int s = (int)(c / k);
        float acc = 0.0f;
        if (is_active[s]) {
            const float* p = partial + c * segs * d + comp;
            acc = p[0];
            for (int g = 1; g < segs; ++g) acc = __fadd_rn(acc, p[(size_t)g * d]);
        }
        pack[t] = acc;
    } else if (t < n_sum + n_c) {
        size_t c = t - n_sum;
        int s = (int)(c / k);
        uint32_t cnt = is_active[s] ? seg_end[c] - seg_beg[c] : 0u;
        pack[n_sum + c] = (float)(cnt & 0xFFFFu);
        pack[n_sum + n_c + c] = (float)(cnt >> 16);
    }
}

// mean, epsilon test (vector.rs:440-447, approx_eq :232-240).  thread = (s, j, comp).
__global__ void k_finalize(const float* __restrict__ pack, int d, int k, int m,
                           const int* __restrict__ is_active, float* __restrict__ codebooks,
                           uint32_t* __restrict__ changed, uint32_t* __restrict__ counts_out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n_sum = (size_t)m * k * d, n_c = (size_t)m * k;
    if (t >= n_sum) return;
    size_t c = t / d; int comp = (int)(t % d);
    int s = (int)(c / k);
    if (!is_active[s]) return;
    // counts travel as two exact small-integer floats (a sum over <= 256 ranks stays < 2^24)
    unsigned long long cnt = (unsigned long long)pack[n_sum + n_c + c] * 65536ull +
                             (unsigned long long)pack[n_sum + c];
    if (comp == 0) counts_out[c] = cnt > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cnt;
    if (cnt == 0) return;  // empty: re-seeded by the host, never sets `changed` (vector.rs:448-452)
    float nv = __fdiv_rn(pack[t], __ull2float_rn(cnt));  // `indices.len() as f32`
    float old = codebooks[t];
    if (!(fabsf(__fsub_rn(nv, old)) < 1e-6f)) atomicOr(&changed[s], 1u);
    codebooks[t] = nv;
}

// gathers sub-vectors x[row][s*d .. s*d+d) for a list of (row or -1, s); absent rows give zeros
__global__ void k_gather_subvecs(const float* __restrict__ x, int dim, int d, const long long* __restrict__ rows,
                                 const int* __restrict__ subs, size_t count, float* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * d) return;
    size_t e = t / d; int comp = (int)(t % d);
    long long r = rows[e];
    out[t] = r >= 0 ? x[(size_t)r * dim + (size_t)subs[e] * d + comp] : 0.0f;
}
__global__ void k_scatter_subvecs(const float* __restrict__ src, int d, const long long* __restrict__ dst_idx,
                                  size_t count, float* __restrict__ codebooks) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * d) return;
    size_t e = t / d; int comp = (int)(t % d);
    codebooks[(size_t)dst_idx[e] * d + comp] = src[t];
}

// One-time subspace-major copy xt[s][row][0..d) = x[row][s*d..] for the update's member gathers.
// A warp takes 32 consecutive rows of one subspace: its writes are one contiguous run, its reads are
// full 32-byte sectors; the other warps of the block read the neighbouring subspaces of the same rows.
__global__ void __launch_bounds__(256)
k_subspace_major(const float* __restrict__ x, size_t n, int dim, int d, int m, float* __restrict__ xt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row = (size_t)blockIdx.x * 32 + lane;
    if (row >= n) return;
    for (int s = warp; s < m; s += 8) {
        const float* src = x + row * dim + (size_t)s * d;
        float* dst = xt + ((size_t)s * n + row) * d;
        if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
            for (int c = 0; c < d; c += 4) *reinterpret_cast<float4*>(dst + c) = __ldg(reinterpret_cast<const float4*>(src + c));
        } else {
            for (int c = 0; c < d; ++c) dst[c] = __ldg(src + c);
        }
    }
}

struct TrainWs {
    size_t n = 0, dim = 0, m = 0, k = 0, d = 0;
    uint32_t code_bytes = 1;
    int n_chunks = 0, segs = 1, passes = 1;
    DevBuf codes, ids_a, ids_b, chunk_hist, bin_off, seg_beg, seg_end, partial, pack, cb;
    DevBuf sub_list, is_active, changed, counts, g_rows, g_subs, g_dst, g_vals, tc_prep;
    DevBuf xt;  // optional subspace-major copy of x (see k_subspace_major); absent when memory is short
    int alloc(vqb_ctx* ctx, size_t n_, size_t dim_, size_t m_, size_t k_, int segs_) {
        n = n_; dim = dim_; m = m_; k = k_; d = dim / m; segs = segs_;
        code_bytes = k <= 256 ? 1 : (k <= 65536 ? 2 : 4);
        passes = (int)code_bytes;
        n_chunks = (int)cdiv(n, RX_CHUNK);
        size_t gcap = std::max(m * k, (size_t)1);
        VQB_CUDA(ctx, codes.alloc(m * n * code_bytes));
        VQB_CUDA(ctx, ids_a.alloc(m * n * 4));
        if (passes > 1) VQB_CUDA(ctx, ids_b.alloc(m * n * 4));
        VQB_CUDA(ctx, chunk_hist.alloc(m * (size_t)n_chunks * 256 * 4));
        VQB_CUDA(ctx, bin_off.alloc(m * 256 * 4));
        VQB_CUDA(ctx, seg_beg.alloc(m * k * 4));
        VQB_CUDA(ctx, seg_end.alloc(m * k * 4));
        VQB_CUDA(ctx, partial.alloc(m * k * (size_t)segs * d * 4));
        VQB_CUDA(ctx, pack.alloc((m * k * d + 2 * m * k) * 4));
        VQB_CUDA(ctx, cb.alloc(m * k * d * 4));
        VQB_CUDA(ctx, sub_list.alloc(m * 4));
        VQB_CUDA(ctx, is_active.alloc(m * 4));
        VQB_CUDA(ctx, changed.alloc(m * 4));
        VQB_CUDA(ctx, counts.alloc(m * k * 4));
        VQB_CUDA(ctx, g_rows.alloc(gcap * 8));
        VQB_CUDA(ctx, g_subs.alloc(gcap * 4));
        VQB_CUDA(ctx, g_dst.alloc(gcap * 8));
        VQB_CUDA(ctx, g_vals.alloc(gcap * d * 4));
        VQB_CUDA(ctx, tc_prep.alloc(vqb_tc_prep_bytes(m)));
        return VQB_SUCCESS;
    }
};

struct TrainArgs {
    const float* x; size_t n, dim, m, k, d;
    uint32_t assign_mode;
    vqb_allreduce_fn allreduce; void* allreduce_user;
    uint64_t row_offset;
};

// `is_active` (device, [m] 0/1, may be null = all) and `sub_list` (device, the same set as a list, may be
// null = all) describe the subspaces to assign; `tc_prep` is the tensor-core workspace (may be null).
int assign_train(vqb_ctx* ctx, const TrainArgs& a, const float* cb, const int* sub_list, const int* is_active, int na,
                 void* codes, uint32_t code_bytes, void* tc_prep) {
    // codes[s*n + row]
    const bool tc_ok = tc_prep && vqb_tc_supported(MK_TRAIN, a.x, a.n, a.dim, a.m, a.k, a.d);
    if (a.assign_mode == VQB_ASSIGN_TENSOR && !tc_ok)
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT,
                        "tensor-core assignment needs sub_dim 8, k <= 256 and 16-byte aligned rows");
    if (tc_ok && (a.assign_mode == VQB_ASSIGN_TENSOR || (a.assign_mode == VQB_ASSIGN_AUTO && a.n >= VQB_TC_MIN_ROWS))) {
        VQB_TRY(vqb_tc_prepare(ctx, MK_TRAIN, cb, a.m, a.k, tc_prep));
        return vqb_tc_assign_launch(ctx, MK_TRAIN, a.x, a.n, a.dim, a.m, a.k, tc_prep, is_active, codes, code_bytes,
                                    /*stride_row=*/1, /*stride_sub=*/a.n, nullptr);
    }
    return vqb_pq_assign_exact_launch(ctx, MK_TRAIN, a.x, a.n, a.dim, a.m, a.k, a.d, cb, sub_list, na, codes,
                                      code_bytes, /*stride_row=*/1, /*stride_sub=*/a.n, nullptr);
}

// Uploads a small host table through the pinned mailbox (keeps the copy asynchronous and ordered).
int upload_small(vqb_ctx* ctx, void* dst, const void* src, size_t bytes, size_t& mb_off) {
    if (bytes == 0) return VQB_SUCCESS;
    size_t aligned = (bytes + 15) & ~size_t(15);
    if (mb_off + aligned > ctx->mailbox_bytes) {  // table larger than the mailbox: plain (synchronous) copy
        VQB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return VQB_SUCCESS;
    }
    char* p = static_cast<char*>(ctx->mailbox) + mb_off;
    std::memcpy(p, src, bytes);
    VQB_CUDA(ctx, cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    mb_off += aligned;
    return VQB_SUCCESS;
}

// One iteration for the subspaces in `active`.  Leaves ws.changed / ws.counts on the device.
int train_iteration(vqb_ctx* ctx, TrainWs& ws, const TrainArgs& a, const std::vector<int>& active) {
    const int na = (int)active.size();
    const size_t n = a.n, k = a.k, d = a.d, m = a.m;
    std::vector<int> act(m, 0);
    for (int s : active) act[s] = 1;
    size_t mb = 0;
    VQB_TRY(upload_small(ctx, ws.sub_list.p, active.data(), na * sizeof(int), mb));
    VQB_TRY(upload_small(ctx, ws.is_active.p, act.data(), m * sizeof(int), mb));
    VQB_CUDA(ctx, cudaMemsetAsync(ws.changed.p, 0, m * 4, ctx->stream));
    const int* sl = ws.sub_list.as<int>();

    VQB_TRY(assign_train(ctx, a, ws.cb.as<float>(), sl, ws.is_active.as<int>(), na, ws.codes.p, ws.code_bytes,
                         ws.tc_prep.p));

    // group: LSD radix, 8 bits per stable pass
    const uint32_t* ids_in = nullptr;
    uint32_t* ids_out = ws.ids_a.as<uint32_t>();
    dim3 gchunks(ws.n_chunks, na);
    for (int pass = 0; pass < ws.passes; ++pass) {
        bool last = pass == ws.passes - 1;
        bool ranges_here = last && ws.passes == 1;
        k_radix_hist<<<gchunks, RX_THREADS, 0, ctx->stream>>>(ws.codes.p, ws.code_bytes, n, pass * 8, sl, ids_in,
                                                              ws.chunk_hist.as<uint32_t>(), ws.n_chunks);
        VQB_LAUNCHED(ctx);
        k_radix_scan<<<na, 256, 0, ctx->stream>>>(ws.chunk_hist.as<uint32_t>(), ws.n_chunks, sl,
                                                  ws.bin_off.as<uint32_t>(),
                                                  ranges_here ? ws.seg_beg.as<uint32_t>() : nullptr,
                                                  ws.seg_end.as<uint32_t>(), (int)k);
        VQB_LAUNCHED(ctx);
        k_radix_scatter<<<gchunks, RX_THREADS, 0, ctx->stream>>>(ws.codes.p, ws.code_bytes, n, pass * 8, sl, ids_in,
                                                                 ws.chunk_hist.as<uint32_t>(),
                                                                 ws.bin_off.as<uint32_t>(), ws.n_chunks, ids_out);
        VQB_LAUNCHED(ctx);
        ids_in = ids_out;
        ids_out = (ids_out == ws.ids_a.as<uint32_t>()) ? ws.ids_b.as<uint32_t>() : ws.ids_a.as<uint32_t>();
    }
    const uint32_t* sorted = ids_in;
    if (ws.passes > 1) {
        VQB_CUDA(ctx, cudaMemsetAsync(ws.seg_beg.p, 0, m * k * 4, ctx->stream));
        VQB_CUDA(ctx, cudaMemsetAsync(ws.seg_end.p, 0, m * k * 4, ctx->stream));
        k_cluster_ranges<<<dim3(cdiv(n, 256), na), 256, 0, ctx->stream>>>(ws.codes.p, ws.code_bytes, n, (int)k, sl,
                                                                         sorted, ws.seg_beg.as<uint32_t>(),
                                                                         ws.seg_end.as<uint32_t>());
        VQB_LAUNCHED(ctx);
    }

    // sum
    size_t chains = (size_t)na * k * ws.segs * d;
    const bool use_xt = ws.xt.p != nullptr;
    k_chain_sums<<<cdiv(chains, 256), 256, 0, ctx->stream>>>(use_xt ? ws.xt.as<float>() : a.x, n, use_xt ? d : a.dim,
                                                            use_xt ? n * d : d, (int)d, (int)k, ws.segs, sl, na,
                                                            sorted, ws.seg_beg.as<uint32_t>(),
                                                            ws.seg_end.as<uint32_t>(), ws.partial.as<float>());
    VQB_LAUNCHED(ctx);
    size_t n_pack = m * k * d + 2 * m * k;
    k_pack<<<cdiv(n_pack, 256), 256, 0, ctx->stream>>>(ws.partial.as<float>(), ws.segs, (int)d, (int)k, (int)m,
                                                      ws.seg_beg.as<uint32_t>(), ws.seg_end.as<uint32_t>(),
                                                      ws.is_active.as<int>(), ws.pack.as<float>());
    VQB_LAUNCHED(ctx);
    if (a.allreduce) {
        int rc = a.allreduce(a.allreduce_user, ws.pack.as<float>(), n_pack, (void*)ctx->stream);
        if (rc != 0) return vqb_fail(ctx, VQB_FAILURE, "allreduce callback failed (%d)", rc);
    }
    k_finalize<<<cdiv(m * k * d, 256), 256, 0, ctx->stream>>>(ws.pack.as<float>(), (int)d, (int)k, (int)m,
                                                             ws.is_active.as<int>(), ws.cb.as<float>(),
                                                             ws.changed.as<uint32_t>(), ws.counts.as<uint32_t>());
    VQB_LAUNCHED(ctx);
    return VQB_SUCCESS;
}

// Writes x[row_e][s_e*d..] into codebooks[dst_e] for a host list; rows outside this rank's shard
// contribute zeros and the patch is summed across ranks.
int apply_rows(vqb_ctx* ctx, TrainWs& ws, const TrainArgs& a, const std::vector<long long>& rows,
               const std::vector<int>& subs, const std::vector<long long>& dst) {
    size_t cnt = rows.size();
    if (!cnt) return VQB_SUCCESS;
    size_t mb = 0;
    // tables can exceed the mailbox (m*k entries): upload_small falls back to a blocking copy
    VQB_TRY(upload_small(ctx, ws.g_rows.p, rows.data(), cnt * 8, mb));
    VQB_TRY(upload_small(ctx, ws.g_subs.p, subs.data(), cnt * 4, mb));
    VQB_TRY(upload_small(ctx, ws.g_dst.p, dst.data(), cnt * 8, mb));
    k_gather_subvecs<<<cdiv(cnt * a.d, 256), 256, 0, ctx->stream>>>(a.x, (int)a.dim, (int)a.d,
                                                                   ws.g_rows.as<long long>(), ws.g_subs.as<int>(),
                                                                   cnt, ws.g_vals.as<float>());
    VQB_LAUNCHED(ctx);
    if (a.allreduce) {
        int rc = a.allreduce(a.allreduce_user, ws.g_vals.as<float>(), cnt * a.d, (void*)ctx->stream);
        if (rc != 0) return vqb_fail(ctx, VQB_FAILURE, "allreduce callback failed (%d)", rc);
    }
    k_scatter_subvecs<<<cdiv(cnt * a.d, 256), 256, 0, ctx->stream>>>(ws.g_vals.as<float>(), (int)a.d,
                                                                    ws.g_dst.as<long long>(), cnt,
                                                                    ws.cb.as<float>());
    VQB_LAUNCHED(ctx);
    // the host vectors may go out of scope: make sure the mailbox copies are done
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}

// pq.rs:91-117 + vector.rs:396-410 validation, in the reference's order.
int validate_train(vqb_ctx* ctx, const float* x, size_t n_global, size_t dim, size_t m, size_t k) {
    if (n_global == 0) return vqb_fail(ctx, VQB_ERR_EMPTY_INPUT, "Empty input: at least one vector is required");
    if (!x) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null data pointer");
    if (m == 0) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "Invalid parameter 'm': must be greater than 0");
    if (dim < m)
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "Invalid parameter 'm': must be at most the data dimension (%zu)", dim);
    if (dim % m)
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "Invalid parameter 'm': dimension (%zu) must be divisible by m", dim);
    if (k == 0) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "Invalid parameter 'k': must be greater than 0");
    if (n_global < k)
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT,
                        "Invalid parameter 'k': not enough data points (%zu) for %zu clusters", n_global, k);
    return VQB_SUCCESS;
}

vqb_train_opts default_opts() {
    vqb_train_opts o;
    std::memset(&o, 0, sizeof(o));
    o.struct_size = sizeof(o);
    return o;
}

}  // namespace

extern "C" {

int vqb_pq_train(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k, size_t max_iters,
                 const uint64_t* init_idx, const vqb_train_opts* opts_in, float* codebooks, uint32_t* iters_run) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    vqb_train_opts o = default_opts();
    if (opts_in) std::memcpy(&o, opts_in, std::min<size_t>(sizeof(o), opts_in->struct_size ? opts_in->struct_size : sizeof(o)));
    const size_t n_global = o.n_global ? (size_t)o.n_global : n;
    VQB_TRY(validate_train(ctx, n ? x : (const float*)1, n_global, dim, m, k));
    if (n && !x) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null data pointer");
    if (!init_idx || !codebooks) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null init_idx / codebooks");
    if (n > 0xFFFFFFFFull) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "more than 2^32 rows per GPU");
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t d = dim / m;

    const double t_enter = trace_on() ? now_ms() : 0.0;
    InputView xin;
    VQB_TRY(xin.bind(ctx, x, n * dim * sizeof(float)));
    TrainWs ws;
    static const int fast_segs = [] { const char* e = std::getenv("VQB_SEGS"); int v = e ? std::atoi(e) : 0; return v > 0 ? v : 128; }();
    VQB_TRY(ws.alloc(ctx, std::max<size_t>(n, 1), dim, m, k, o.update_mode == VQB_UPDATE_FAST ? fast_segs : 1));
    ws.n = n;
    if (trace_on()) std::fprintf(stderr, "[vqb trace] bind + workspace alloc: %.3f ms\n", now_ms() - t_enter);
    TrainArgs a{static_cast<const float*>(xin.dev), n, dim, m, k, d, o.assign_mode, o.allreduce, o.allreduce_user,
                o.row_offset};
    // subspace-major copy for the update's gathers: pays for itself after about two iterations; skipped for
    // small inputs, single-iteration calls, m == 1 (already contiguous) or when the memory is not there
    static const bool no_xt = [] { const char* e = std::getenv("VQB_NO_XT"); return e && *e && *e != '0'; }();
    if (!no_xt && max_iters >= 2 && m > 1 && n >= 65536) {
        if (ws.xt.alloc(n * dim * sizeof(float)) == cudaSuccess) {
            k_subspace_major<<<cdiv(n, 32), 256, 0, ctx->stream>>>(a.x, n, (int)dim, (int)d, (int)m, ws.xt.as<float>());
            VQB_LAUNCHED(ctx);
        } else {
            cudaGetLastError();  // not enough memory: gather from the caller's layout instead
            ws.xt.p = nullptr; ws.xt.bytes = 0;
        }
    }

    // vector.rs:413: the sampled rows become the initial centroids
    {
        std::vector<long long> rows(m * k), dst(m * k);
        std::vector<int> subs(m * k);
        for (size_t s = 0; s < m; ++s)
            for (size_t j = 0; j < k; ++j) {
                uint64_t g = init_idx[s * k + j];
                if (g >= n_global) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "init_idx[%zu] = %llu out of range", s * k + j, (unsigned long long)g);
                bool local = g >= o.row_offset && g < o.row_offset + n;
                rows[s * k + j] = local ? (long long)(g - o.row_offset) : -1;
                subs[s * k + j] = (int)s;
                dst[s * k + j] = (long long)(s * k + j);
            }
        VQB_TRY(apply_rows(ctx, ws, a, rows, subs, dst));
    }

    std::vector<int> active(m);
    for (size_t s = 0; s < m; ++s) active[s] = (int)s;
    std::vector<uint32_t> iters(m, 0), h_changed(m), h_counts(m * k);
    cudaEvent_t ev_it[2] = {nullptr, nullptr};
    if (o.iter_ms) {
        VQB_CUDA(ctx, cudaEventCreate(&ev_it[0]));
        VQB_CUDA(ctx, cudaEventCreate(&ev_it[1]));
    }
    struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 2; ++i) if (e[i]) cudaEventDestroy(e[i]); } } ev_guard{ev_it};
    for (size_t it = 0; it < max_iters && !active.empty(); ++it) {
        const double t0 = trace_on() ? now_ms() : 0.0;
        if (o.iter_ms) VQB_CUDA(ctx, cudaEventRecord(ev_it[0], ctx->stream));
        VQB_TRY(train_iteration(ctx, ws, a, active));
        if (o.iter_ms) VQB_CUDA(ctx, cudaEventRecord(ev_it[1], ctx->stream));
        const double t1 = trace_on() ? now_ms() : 0.0;
        VQB_CUDA(ctx, cudaMemcpyAsync(h_changed.data(), ws.changed.p, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        VQB_CUDA(ctx, cudaMemcpyAsync(h_counts.data(), ws.counts.p, m * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
        VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (o.iter_ms) VQB_CUDA(ctx, cudaEventElapsedTime(&o.iter_ms[it], ev_it[0], ev_it[1]));
        if (trace_on())
            std::fprintf(stderr, "[vqb trace] iter %zu: active=%zu enqueue=%.3f ms, wait=%.3f ms\n", it, active.size(),
                         t1 - t0, now_ms() - t1);
        std::vector<long long> rows, dst;
        std::vector<int> subs;
        std::vector<int> next;
        for (int s : active) {
            iters[s]++;
            if (o.reseed)
                for (size_t j = 0; j < k; ++j)  // vector.rs:448-452, ascending j
                    if (h_counts[(size_t)s * k + j] == 0) {
                        uint64_t g = o.reseed(o.reseed_user, (uint32_t)s);
                        if (g >= n_global) g %= n_global;
                        bool local = g >= o.row_offset && g < o.row_offset + n;
                        rows.push_back(local ? (long long)(g - o.row_offset) : -1);
                        subs.push_back(s);
                        dst.push_back((long long)((size_t)s * k + j));
                    }
            if (h_changed[s]) next.push_back(s);  // vector.rs:455-457
        }
        VQB_TRY(apply_rows(ctx, ws, a, rows, subs, dst));
        active.swap(next);
    }
    if (iters_run) std::memcpy(iters_run, iters.data(), m * 4);
    VQB_CUDA(ctx, cudaMemcpyAsync(codebooks, ws.cb.p, m * k * d * 4, cudaMemcpyDefault, ctx->stream));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (trace_on()) std::fprintf(stderr, "[vqb trace] train total (before workspace free): %.3f ms\n", now_ms() - t_enter);
    return VQB_SUCCESS;
}

int vqb_pq_assign_train(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k,
                        const float* codebooks, uint32_t assign_mode, uint32_t* codes_out) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    VQB_TRY(validate_train(ctx, x, n, dim, m, std::min(k, n)));
    if (!codebooks || !codes_out) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t d = dim / m;
    InputView xin, cin; OutputView ov;
    VQB_TRY(xin.bind(ctx, x, n * dim * 4));
    VQB_TRY(cin.bind(ctx, codebooks, m * k * d * 4));
    VQB_TRY(ov.bind(ctx, codes_out, m * n * 4));
    TrainArgs a{static_cast<const float*>(xin.dev), n, dim, m, k, d, assign_mode, nullptr, nullptr, 0};
    DevBuf tcp;
    VQB_CUDA(ctx, tcp.alloc(vqb_tc_prep_bytes(m)));
    VQB_TRY(assign_train(ctx, a, static_cast<const float*>(cin.dev), nullptr, nullptr, (int)m, ov.dev, 4, tcp.p));
    VQB_TRY(ov.finish(ctx));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}

int vqb_pq_train_step(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k,
                      float* codebooks_inout, const vqb_train_opts* opts_in, uint32_t* changed_out,
                      uint32_t* counts_out) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    vqb_train_opts o = default_opts();
    if (opts_in) std::memcpy(&o, opts_in, std::min<size_t>(sizeof(o), opts_in->struct_size ? opts_in->struct_size : sizeof(o)));
    VQB_TRY(validate_train(ctx, x, n, dim, m, std::min(k, n)));
    if (!codebooks_inout || !changed_out) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t d = dim / m;
    InputView xin;
    VQB_TRY(xin.bind(ctx, x, n * dim * 4));
    TrainWs ws;
    VQB_TRY(ws.alloc(ctx, n, dim, m, k, o.update_mode == VQB_UPDATE_FAST ? 128 : 1));
    VQB_CUDA(ctx, cudaMemcpyAsync(ws.cb.p, codebooks_inout, m * k * d * 4, cudaMemcpyDefault, ctx->stream));
    TrainArgs a{static_cast<const float*>(xin.dev), n, dim, m, k, d, o.assign_mode, o.allreduce, o.allreduce_user,
                o.row_offset};
    std::vector<int> active(m);
    for (size_t s = 0; s < m; ++s) active[s] = (int)s;
    VQB_TRY(train_iteration(ctx, ws, a, active));
    VQB_CUDA(ctx, cudaMemcpyAsync(changed_out, ws.changed.p, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (counts_out)
        VQB_CUDA(ctx, cudaMemcpyAsync(counts_out, ws.counts.p, m * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VQB_CUDA(ctx, cudaMemcpyAsync(codebooks_inout, ws.cb.p, m * k * d * 4, cudaMemcpyDefault, ctx->stream));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}

}  // extern "C"
