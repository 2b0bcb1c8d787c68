// pq_train.cu -- LBG / k-means codebook training for all m subspaces at once.
//
// Reference: lbg_quantize (src/core/vector.rs:390-461) called serially per subspace by
// ProductQuantizer::new (src/pq.rs:121-132).  Here every active subspace advances one
// iteration per pass ("level-synchronous" over subspaces); a subspace that reports
// `!changed` is frozen exactly where the reference would `break` (vector.rs:455-457).
//
// One iteration =
//   assign      nearest centroid per (row, subspace)                 [pq_assign.cu / pq_tc.cu]
//   update      per (subspace, cluster, component) sums + counts     (vector.rs:432-435, 368-384)
//                 FAST    k_update_tiles: one pass over X in 512-row tiles (TMA); every row sets its bit in the membership
//                         mask of its cluster (shared memory), then one thread per (subspace, cluster) adds its members'
//                         sub-vectors, lowest row first, into registers; per-CTA partial sums are combined in a fixed order
//                 ORDERED stable 8-bit radix grouping of the row ids + one chain per (subspace, cluster, component)
//                         over the member ids in ascending order with sequential f32 adds: the reference's exact order
//   [allreduce] one fused buffer [sums | count_lo | count_hi] when rows are sharded over GPUs (NCCL owned by the
//               context, comm.cu, or the host's vqb_allreduce_fn)
//   finalize    mean = sum / count, epsilon test, changed flags      (vector.rs:438-447)
//   epilogue    freezes converged subspaces, counts iterations, lists empty clusters, publishes a status record in pinned
//               host memory and decides whether the NEXT iteration may run
//   reseed      host callback per empty cluster, ascending j         (vector.rs:448-452)
//
// The host stays out of the loop: the list of active subspaces, the per-subspace iteration counts and the "go" flag live
// on the device; iteration t+1 is enqueued before the status of iteration t has been read, and all of its kernels are
// no-ops unless iteration t allowed it (no empty cluster to re-seed, some subspace still moving).  The host synchronises
// only on the small status record, and re-issues an iteration after re-seeding.
//
// HBM layout: X row-major [n, dim] f32 stays resident and is never copied in FAST mode; codes [m][n] (1/2/4 bytes);
// sums [m][k][sub_dim] f32.  The workspace is a grow-only slab owned by the context (no cudaMalloc per call).
#include "common.cuh"

#include <cuda.h>

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

// device-side loop state
struct TrainCtrl {
    uint32_t go;         // 1: the next enqueued iteration runs; 0: its kernels return at once
    int n_active;        // length of sub_list
};
// status record of one iteration in pinned host memory (one slot per iteration modulo STATUS_SLOTS)
constexpr int STATUS_SLOTS = 4;
constexpr int STATUS_CAP = 1000;                         // (subspace, cluster) pairs listed explicitly
constexpr size_t STATUS_WORDS = 8 + 2 * STATUS_CAP;      // seq, n_active_next, n_empty, ran, pad[4], pairs
constexpr size_t STATUS_BYTES = STATUS_SLOTS * STATUS_WORDS * 4;

__device__ __forceinline__ uint32_t load_code(const void* codes, uint32_t code_bytes, size_t off) {
    if (code_bytes == 1) return static_cast<const uint8_t*>(codes)[off];
    if (code_bytes == 2) return static_cast<const uint16_t*>(codes)[off];
    return static_cast<const uint32_t*>(codes)[off];
}

// ---- stable radix pass over digit (code >> shift) & 255 (ORDERED update, and FAST for shapes the tile kernel does not
// cover).  grid (n_chunks, m): block y handles sub_list[y] and returns when y >= *n_active.  ids_in == nullptr: identity.
__global__ void __launch_bounds__(RX_THREADS)
k_radix_hist(const void* __restrict__ codes, uint32_t code_bytes, size_t n, int shift,
             const int* __restrict__ sub_list, const TrainCtrl* __restrict__ ctrl, const uint32_t* __restrict__ ids_in,
             uint32_t* __restrict__ chunk_hist, int n_chunks) {
    __shared__ uint32_t h[256];
    if (ctrl->go == 0 || (int)blockIdx.y >= ctrl->n_active) return;
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

// grid (m), 256 threads (thread = bin): exclusive scan over chunks, then over bins.
// When seg_beg != nullptr (single-pass case, digit == cluster id) also emits the member ranges.
__global__ void __launch_bounds__(256)
k_radix_scan(uint32_t* __restrict__ chunk_hist, int n_chunks, const int* __restrict__ sub_list,
             const TrainCtrl* __restrict__ ctrl, uint32_t* __restrict__ bin_off, uint32_t* __restrict__ seg_beg,
             uint32_t* __restrict__ seg_end, int k) {
    __shared__ uint32_t tot[256];
    if (ctrl->go == 0 || (int)blockIdx.x >= ctrl->n_active) return;
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
                const int* __restrict__ sub_list, const TrainCtrl* __restrict__ ctrl, const uint32_t* __restrict__ ids_in,
                const uint32_t* __restrict__ chunk_base, const uint32_t* __restrict__ bin_off,
                int n_chunks, uint32_t* __restrict__ ids_out) {
    __shared__ uint32_t running[256];
    __shared__ uint16_t wcnt[RX_WARPS][256];
    if (ctrl->go == 0 || (int)blockIdx.y >= ctrl->n_active) return;
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
// zeroed first; grid (cdiv(n,256), m).
__global__ void k_cluster_ranges(const void* __restrict__ codes, uint32_t code_bytes, size_t n, int k,
                                 const int* __restrict__ sub_list, const TrainCtrl* __restrict__ ctrl,
                                 const uint32_t* __restrict__ ids, uint32_t* __restrict__ seg_beg,
                                 uint32_t* __restrict__ seg_end) {
    if (ctrl->go == 0 || (int)blockIdx.y >= ctrl->n_active) return;
    const size_t s = sub_list[blockIdx.y];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = load_code(codes, code_bytes, s * n + ids[s * n + i]);
    uint32_t prev = i ? load_code(codes, code_bytes, s * n + ids[s * n + i - 1]) : 0xFFFFFFFFu;
    uint32_t next = (i + 1 < n) ? load_code(codes, code_bytes, s * n + ids[s * n + i + 1]) : 0xFFFFFFFFu;
    if (c != prev) seg_beg[s * k + c] = (uint32_t)i;
    if (c != next) seg_end[s * k + c] = (uint32_t)i + 1;
}
__global__ void k_zero_ranges(uint32_t* __restrict__ seg_beg, uint32_t* __restrict__ seg_end, size_t count,
                              const TrainCtrl* __restrict__ ctrl) {
    if (ctrl->go == 0) return;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) { seg_beg[t] = 0; seg_end[t] = 0; }
}

// ---- per-(subspace, cluster, component) ordered sums ------------------------------------------
// thread = (active pos, cluster j, segment, component).  segs == 1 reproduces the reference's
// summation order exactly; segs > 1 splits each member list into equal contiguous pieces.
__global__ void __launch_bounds__(256)
k_chain_sums(const float* __restrict__ x, size_t n, size_t row_stride, size_t sub_stride, int d, int k, int segs,
             const int* __restrict__ sub_list, const TrainCtrl* __restrict__ ctrl, const uint32_t* __restrict__ ids,
             const uint32_t* __restrict__ seg_beg, const uint32_t* __restrict__ seg_end,
             float* __restrict__ partial /* [m][k][segs][d] */) {
    if (ctrl->go == 0) return;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)ctrl->n_active * k * segs * d;
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

// ---- FAST update: tile kernel ------------------------------------------------------------------------------------
// A CTA owns UT_G = 32 / D consecutive subspaces (one 128-byte slab of every row) and every `parts`-th 512-row tile.
// Per tile: (A) every row sets its bit in the membership mask of its cluster -- mask[subspace][row warp][cluster], one
// 32-bit word per 32 consecutive rows, shared-memory atomicOr; (B) thread (q, j) walks the 16 mask words of cluster j of
// the subspace that owns 32-byte chunk q of the slab, lowest bit first, and adds the members' chunks into 8 registers:
// ascending row order, no sort, no prefix sums, two block barriers per tile.  The x tile arrives by TMA (same tensor map
// as the assignment kernel) one tile ahead; the next tile's codes are fetched while this tile is summed.
// With parts == 1 the sums are accumulated in ascending row order, i.e. in the reference's order.
constexpr int UT_ROWS = 512, UT_THREADS = 1024, UT_RW = UT_ROWS / 32;
constexpr uint32_t UT_TILE_BYTES = UT_ROWS * 128;
constexpr uint32_t UT_OFF_MASK = 2 * UT_TILE_BYTES;                       // [4][16][256] u32
constexpr uint32_t UT_OFF_NZ = UT_OFF_MASK + 4 * UT_RW * 256 * 4;         // [4][256] u32: which of a cluster's 16 mask words are non-empty
constexpr uint32_t UT_OFF_BAR = UT_OFF_NZ + 4 * 256 * 4;
constexpr uint32_t UT_SMEM = UT_OFF_BAR + 64 + 1024;
static_assert(UT_SMEM <= 232448, "exceeds the 227 KB opt-in shared memory of sm_100");

struct UtParams {
    const uint8_t* codes;       // [m][n] u8
    unsigned long long n;
    int m, k, n_groups, parts, num_tiles;
    const int* is_active;       // [m]
    const TrainCtrl* ctrl;
    float* partial;             // [m][k][parts][D]
    uint32_t* cnt_partial;      // [m][k][parts]
};

__device__ __forceinline__ uint32_t ut_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int D>
__global__ void __launch_bounds__(UT_THREADS, 1) k_update_tiles(const __grid_constant__ CUtensorMap xmap, const UtParams p) {
    constexpr int G = 32 / D;        // subspaces per CTA
    constexpr int TPS = D / 8;       // threads (32-byte chunks) per (subspace, cluster)
    extern __shared__ uint8_t ut_raw[];
    if (p.ctrl->go == 0) return;
    const uint32_t sbase = (ut_smem_u32(ut_raw) + 1023u) & ~1023u;
    uint8_t* sm = ut_raw + (sbase - ut_smem_u32(ut_raw));
    const int tid = threadIdx.x;
    const int grp = blockIdx.x % p.n_groups, part = blockIdx.x / p.n_groups;
    const int s0 = grp * G;
    const int g_cnt = min(G, p.m - s0);
    uint32_t act_mask = 0;
    for (int i = 0; i < g_cnt; ++i)
        if (p.is_active[s0 + i]) act_mask |= 1u << i;
    if (act_mask == 0) return;
    uint32_t* mask = reinterpret_cast<uint32_t*>(sm + UT_OFF_MASK);
    uint32_t* nzmap = reinterpret_cast<uint32_t*>(sm + UT_OFF_NZ);
    const uint32_t bar0 = sbase + UT_OFF_BAR;
    const int my_tiles = (p.num_tiles - part + p.parts - 1) / p.parts;
    for (int t = tid; t < G * UT_RW * 256 / 4; t += UT_THREADS) reinterpret_cast<uint4*>(mask)[t] = make_uint4(0, 0, 0, 0);
    for (int t = tid; t < 4 * 256; t += UT_THREADS) nzmap[t] = 0;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * i), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto load_tile = [&](int it) {   // thread 0 only
        const int st = it & 1;
        const int tile = part + it * p.parts;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * st), "r"(UT_TILE_BYTES) : "memory");
        for (int b = 0; b < UT_ROWS / 128; ++b)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                "r"(sbase + st * UT_TILE_BYTES + b * 128 * 128), "l"((uint64_t)&xmap), "r"(s0 * D), "r"(tile * UT_ROWS + b * 128),
                "r"(bar0 + 8 * st)
                : "memory");
    };
    if (tid == 0 && my_tiles > 0) load_tile(0);

    // phase B identity of this thread: 32-byte chunk q of the slab, cluster j
    const int q = tid >> 8, j = tid & 255;
    const int sb = q / TPS;                       // subspace (within the group) that owns chunk q
    const bool b_on = sb < g_cnt && (act_mask >> sb & 1);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
    uint32_t count = 0;
    // phase A identity: row of the tile, subspaces sp and sp + 2
    const int row = tid & (UT_ROWS - 1), sp = tid >> 9, rw = row >> 5;
    const uint32_t bit = 1u << (row & 31);
    bool a_on[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) { const int si = sp + 2 * a; a_on[a] = si < G && si < g_cnt && (act_mask >> si & 1); }
    auto fetch_codes = [&](int it, uint32_t (&c)[2]) {
        const unsigned long long rowg = (unsigned long long)(part + it * p.parts) * UT_ROWS + row;
#pragma unroll
        for (int a = 0; a < 2; ++a)
            c[a] = (a_on[a] && rowg < p.n) ? (uint32_t)__ldg(p.codes + (size_t)(s0 + sp + 2 * a) * p.n + rowg) : 0xFFFFFFFFu;
    };
    uint32_t cde[2];
    if (my_tiles > 0) fetch_codes(0, cde);

    for (int it = 0; it < my_tiles; ++it) {
        const int st = it & 1;
        if (tid == 0 && it + 1 < my_tiles) load_tile(it + 1);   // its stage was released by the barrier that ended tile it-1
        // ---- A: membership bits (the masks are all zero here)
#pragma unroll
        for (int a = 0; a < 2; ++a)
            if (cde[a] != 0xFFFFFFFFu) {
                atomicOr(&mask[((sp + 2 * a) * UT_RW + rw) * 256 + cde[a]], bit);
                atomicOr(&nzmap[(sp + 2 * a) * 256 + cde[a]], 1u << rw);
            }
        __syncthreads();
        if (it + 1 < my_tiles) fetch_codes(it + 1, cde);         // in flight while this tile is summed
        // ---- B: the x tile has landed; add the members' chunks, lowest row first
        {
            uint32_t ok = 0;
            const uint32_t parity = (it >> 1) & 1;
            while (!ok)
                asm volatile("{\n\t.reg .pred pp;\n\tmbarrier.try_wait.parity.shared::cta.b64 pp, [%1], %2;\n\tselp.u32 %0, 1, 0, pp;\n\t}"
                             : "=r"(ok) : "r"(bar0 + 8 * st), "r"(parity) : "memory");
        }
        if (b_on) {
            const uint8_t* xt = sm + st * UT_TILE_BYTES;
            // which of the 16 mask words of this cluster are non-empty (about two are), then only those are walked:
            // the warp iterates max-over-lanes(non-empty words) times instead of once per word
            uint32_t nz = nzmap[sb * 256 + j];
            if (TPS == 1) nzmap[sb * 256 + j] = 0;
            while (nz) {
                const int w = __ffs(nz) - 1;
                nz &= nz - 1;
                uint32_t* mp = &mask[(sb * UT_RW + w) * 256 + j];
                uint32_t mm = *mp;
                count += __popc(mm);
                if (TPS == 1) *mp = 0;   // this thread is the word's only reader: clear it for the next tile
                while (mm) {
                    const uint32_t rr = (uint32_t)(w * 32 + __ffs(mm) - 1);
                    mm &= mm - 1;
                    const uint8_t* rp = xt + rr * 128;
                    // SWIZZLE_128B: 16-byte chunk c of row r lives at chunk c ^ (r & 7)
                    const float4 v0 = *reinterpret_cast<const float4*>(rp + (((2 * q) ^ (rr & 7)) << 4));
                    const float4 v1 = *reinterpret_cast<const float4*>(rp + (((2 * q + 1) ^ (rr & 7)) << 4));
                    acc[0] = __fadd_rn(acc[0], v0.x); acc[1] = __fadd_rn(acc[1], v0.y); acc[2] = __fadd_rn(acc[2], v0.z); acc[3] = __fadd_rn(acc[3], v0.w);
                    acc[4] = __fadd_rn(acc[4], v1.x); acc[5] = __fadd_rn(acc[5], v1.y); acc[6] = __fadd_rn(acc[6], v1.z); acc[7] = __fadd_rn(acc[7], v1.w);
                }
            }
        }
        __syncthreads();   // the tile and the masks may be overwritten
        if (TPS > 1) {     // several threads read each mask word: clear them together
            for (int t = tid; t < 4 * 256; t += UT_THREADS) nzmap[t] = 0;
            for (int t = tid; t < G * UT_RW * 256 / 4; t += UT_THREADS) reinterpret_cast<uint4*>(mask)[t] = make_uint4(0, 0, 0, 0);
            __syncthreads();
        }
    }
    if (b_on && j < p.k) {
        const size_t cj = (size_t)(s0 + sb) * p.k + j;
        float* dst = p.partial + (cj * p.parts + part) * D + (q % TPS) * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        if (q % TPS == 0) p.cnt_partial[cj * p.parts + part] = count;
    }
}

// ---- ORDERED update: one CTA per subspace, rows in order ----------------------------------------------------------
// The reference adds the members of a cluster in ascending row order with sequential f32 adds (vector.rs:368-384), so a
// chain (subspace, cluster, component) cannot be split -- but the 256 clusters x D components of a subspace are independent
// chains.  One CTA owns a subspace and walks ALL row tiles in order: membership masks exactly as in k_update_tiles, thread j
// adds the members of cluster j, lowest row first, into D registers.  Bit-identical with the reference; X is read once
// (D*4-byte pieces of every row through a D-column TMA box); no row-id arrays, no radix passes.  Measured on 1M x 768
// (m = 96): 2.7 ms per pass -- bound by DRAM over-fetch of the 32-byte pieces at a 3 KB stride (the same pieces through
// 16-byte cp.async: 3.4 ms; 512-row tiles: 3.7 ms).
constexpr int UO_THREADS = 256;
struct UoParams {
    const uint8_t* codes;       // [m][n] u8
    unsigned long long n;
    int m, k, num_tiles;
    const int* sub_list;        // device-side list of the active subspaces
    const TrainCtrl* ctrl;
    float* partial;             // [m][k][1][D]
    uint32_t* cnt_partial;      // [m][k][1]
};
// rows per tile: as many as keep two tiles + the masks under ~200 KB (fewer, longer tiles: the per-tile barriers and
// latencies are the cost at 8 warps per SM)
template <int D> struct UoShape { static constexpr int ROWS = D <= 8 ? 2048 : (D <= 16 ? 1024 : 512); };

template <int D>
__global__ void __launch_bounds__(UO_THREADS) k_update_ordered(const __grid_constant__ CUtensorMap xmap, const UoParams p) {
    constexpr int ROWS = UoShape<D>::ROWS, RW = ROWS / 32, RPT = ROWS / UO_THREADS;
    constexpr uint32_t TILE_BYTES = ROWS * D * 4;
    extern __shared__ uint8_t uo_raw[];
    if (p.ctrl->go == 0 || (int)blockIdx.x >= p.ctrl->n_active) return;
    const int s = p.sub_list[blockIdx.x];
    const uint32_t sbase = (ut_smem_u32(uo_raw) + 127u) & ~127u;
    uint8_t* sm = uo_raw + (sbase - ut_smem_u32(uo_raw));
    uint32_t* mask = reinterpret_cast<uint32_t*>(sm + 2 * TILE_BYTES);          // [RW][256]
    const uint32_t bar0 = sbase + 2 * TILE_BYTES + RW * 256 * 4;
    const int tid = threadIdx.x;
    for (int t = tid; t < RW * 256 / 4; t += UO_THREADS) reinterpret_cast<uint4*>(mask)[t] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * i), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto load_tile = [&](int it) {   // thread 0 only
        const int st = it & 1;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * st), "r"(TILE_BYTES) : "memory");
        for (int b = 0; b < ROWS / 128; ++b)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                "r"(sbase + st * TILE_BYTES + b * 128 * D * 4), "l"((uint64_t)&xmap), "r"(s * D), "r"(it * ROWS + b * 128),
                "r"(bar0 + 8 * st)
                : "memory");
    };
    if (tid == 0 && p.num_tiles > 0) load_tile(0);
    float acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.0f;
    uint32_t count = 0;
    const int j = tid;
    auto fetch_codes = [&](int it, uint32_t (&c)[RPT]) {
#pragma unroll
        for (int a = 0; a < RPT; ++a) {
            const unsigned long long rowg = (unsigned long long)it * ROWS + tid + a * UO_THREADS;
            c[a] = rowg < p.n ? (uint32_t)__ldg(p.codes + (size_t)s * p.n + rowg) : 0xFFFFFFFFu;
        }
    };
    uint32_t cde[RPT];
    if (p.num_tiles > 0) fetch_codes(0, cde);
    for (int it = 0; it < p.num_tiles; ++it) {
        const int st = it & 1;
        if (tid == 0 && it + 1 < p.num_tiles) load_tile(it + 1);   // its stage was released by the barrier that ended tile it-1
#pragma unroll
        for (int a = 0; a < RPT; ++a) {
            const int row = tid + a * UO_THREADS;
            if (cde[a] != 0xFFFFFFFFu) atomicOr(&mask[(row >> 5) * 256 + cde[a]], 1u << (row & 31));
        }
        __syncthreads();
        if (it + 1 < p.num_tiles) fetch_codes(it + 1, cde);
        {
            uint32_t ok = 0;
            const uint32_t parity = (it >> 1) & 1;
            while (!ok)
                asm volatile("{\n\t.reg .pred pp;\n\tmbarrier.try_wait.parity.shared::cta.b64 pp, [%1], %2;\n\tselp.u32 %0, 1, 0, pp;\n\t}"
                             : "=r"(ok) : "r"(bar0 + 8 * st), "r"(parity) : "memory");
        }
        const uint8_t* xt = sm + st * TILE_BYTES;
        unsigned long long nz = 0;
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            const uint32_t m = mask[w * 256 + j];
            nz |= (unsigned long long)(m != 0u ? 1u : 0u) << w;
            count += __popc(m);
        }
        while (nz) {   // non-empty words, lowest rows first: ascending row order inside the cluster
            const int w = __ffsll((long long)nz) - 1;
            nz &= nz - 1;
            uint32_t mm = mask[w * 256 + j];
            mask[w * 256 + j] = 0;
            while (mm) {
                const uint32_t rr = (uint32_t)(w * 32 + __ffs(mm) - 1);
                mm &= mm - 1;
                const float4* rp = reinterpret_cast<const float4*>(xt + rr * D * 4);
#pragma unroll
                for (int c4 = 0; c4 < D / 4; ++c4) {
                    const float4 v = rp[c4];
                    acc[4 * c4] = __fadd_rn(acc[4 * c4], v.x); acc[4 * c4 + 1] = __fadd_rn(acc[4 * c4 + 1], v.y);
                    acc[4 * c4 + 2] = __fadd_rn(acc[4 * c4 + 2], v.z); acc[4 * c4 + 3] = __fadd_rn(acc[4 * c4 + 3], v.w);
                }
            }
        }
        __syncthreads();
    }
    if (j < p.k) {
        const size_t cj = (size_t)s * p.k + j;
#pragma unroll
        for (int c = 0; c < D; ++c) p.partial[cj * D + c] = acc[c];
        p.cnt_partial[cj] = count;
    }
}

// Packs [sums | count_lo | count_hi] (all f32); frozen subspaces are zeroed so that a cross-rank sum leaves them inert.
// partial [m][k][segs][d]; counts either from the member ranges (cnt_partial == nullptr) or from [m][k][segs] partial counts.
__global__ void k_pack(const float* __restrict__ partial, int segs, int d, int k, int m,
                       const uint32_t* __restrict__ seg_beg, const uint32_t* __restrict__ seg_end,
                       const uint32_t* __restrict__ cnt_partial, const int* __restrict__ is_active,
                       const TrainCtrl* __restrict__ ctrl, float* __restrict__ pack) {
    if (ctrl->go == 0) return;
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
        uint32_t cnt = 0;
        if (is_active[s]) {
            if (cnt_partial) for (int g = 0; g < segs; ++g) cnt += cnt_partial[c * segs + g];
            else cnt = seg_end[c] - seg_beg[c];
        }
        pack[n_sum + c] = (float)(cnt & 0xFFFFu);
        pack[n_sum + n_c + c] = (float)(cnt >> 16);
    }
}

// mean, epsilon test (vector.rs:440-447, approx_eq :232-240).  thread = (s, j, comp).
__global__ void k_finalize(const float* __restrict__ pack, int d, int k, int m,
                           const int* __restrict__ is_active, const TrainCtrl* __restrict__ ctrl,
                           float* __restrict__ codebooks, uint32_t* __restrict__ changed, uint32_t* __restrict__ counts_out) {
    if (ctrl->go == 0) return;
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
    if (!(fabsf(__fsub_rn(nv, old)) < 1e-6f) && changed[s] == 0) changed[s] = 1u;  // benign race: every writer stores 1
    codebooks[t] = nv;
}

// End of an iteration, one block: lists the empty clusters of the subspaces that ran (vector.rs:448-452), counts the
// iteration, freezes the subspaces that did not move (vector.rs:455-457), rebuilds the active list and decides whether the
// next (already enqueued) iteration runs.  The status record goes to pinned host memory.
__global__ void __launch_bounds__(1024)
k_epilogue(int m, int k, int* __restrict__ is_active, int* __restrict__ sub_list, TrainCtrl* __restrict__ ctrl,
           uint32_t* __restrict__ changed, const uint32_t* __restrict__ counts, uint32_t* __restrict__ iters,
           uint32_t* __restrict__ status, int has_reseed, uint32_t iter_idx, uint32_t max_iters) {
    __shared__ uint32_t n_empty;
    if (ctrl->go == 0) return;
    if (threadIdx.x == 0) n_empty = 0;
    __syncthreads();
    extern __shared__ int still[];   // [m]: first the active flags, then 1 = the subspace moved and stays on the active list
    for (int s = threadIdx.x; s < m; s += blockDim.x) still[s] = is_active[s];
    __syncthreads();
    for (int idx = threadIdx.x; idx < m * k; idx += blockDim.x) {
        const int s = idx / k;
        if (counts[idx] == 0 && still[s]) {
            const uint32_t pos = atomicAdd(&n_empty, 1u);
            if (pos < (uint32_t)STATUS_CAP) { status[8 + 2 * pos] = (uint32_t)s; status[9 + 2 * pos] = (uint32_t)(idx - s * k); }
        }
    }
    __syncthreads();   // every reader of the active flags is done before they are overwritten
    for (int s = threadIdx.x; s < m; s += blockDim.x) {
        int keep = 0;
        if (still[s]) {
            iters[s]++;
            keep = changed[s] ? 1 : 0;
            if (!keep) is_active[s] = 0;
            changed[s] = 0;
        }
        still[s] = keep;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int na = 0;
        for (int s = 0; s < m; ++s)
            if (still[s]) sub_list[na++] = s;
        ctrl->n_active = na;
        ctrl->go = (na > 0 && iter_idx + 1 < max_iters && (n_empty == 0 || !has_reseed)) ? 1u : 0u;
        status[1] = (uint32_t)na; status[2] = n_empty; status[3] = 1u;
        __threadfence_system();
        status[0] = iter_idx + 1;
    }
}
// start state of a training call: every subspace active, nothing changed, iteration counts zero, go
__global__ void k_train_reset(int m, int* __restrict__ is_active, int* __restrict__ sub_list, TrainCtrl* __restrict__ ctrl,
                              uint32_t* __restrict__ changed, uint32_t* __restrict__ iters) {
    for (int s = threadIdx.x; s < m; s += blockDim.x) { is_active[s] = 1; sub_list[s] = s; changed[s] = 0; iters[s] = 0; }
    if (threadIdx.x == 0) { ctrl->go = 1; ctrl->n_active = m; }
}
__global__ void k_set_go(TrainCtrl* ctrl, uint32_t v) { ctrl->go = v; }

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

// One-time subspace-major copy xt[s][row][0..d) = x[row][s*d..] for the chain-sum gathers of the segmented FAST update
// (shapes the tile kernel does not cover).  A warp takes 32 consecutive rows of one subspace.
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

// UPD_ORDERED / UPD_SEGMENTED: radix grouping + chain sums (any shape).  UPD_TILES: FAST, k_update_tiles.
// UPD_ORDERED_TILES: ORDERED, k_update_ordered (one CTA per subspace, the reference's summation order).
enum UpdateKind { UPD_ORDERED = 0, UPD_SEGMENTED = 1, UPD_TILES = 2, UPD_ORDERED_TILES = 3 };

struct TrainWs {
    size_t n = 0, dim = 0, m = 0, k = 0, d = 0;
    uint32_t code_bytes = 1;
    int n_chunks = 0, segs = 1, passes = 1, kind = UPD_ORDERED;
    int ut_groups = 0, ut_parts = 0;
    void* codes = nullptr;
    uint32_t *ids_a = nullptr, *ids_b = nullptr, *chunk_hist = nullptr, *bin_off = nullptr, *seg_beg = nullptr, *seg_end = nullptr;
    uint32_t *cnt_partial = nullptr, *changed = nullptr, *counts = nullptr, *iters = nullptr;
    float *partial = nullptr, *pack = nullptr, *cb = nullptr, *g_vals = nullptr;
    int *sub_list = nullptr, *is_active = nullptr, *g_subs = nullptr;
    long long *g_rows = nullptr, *g_dst = nullptr;
    TrainCtrl* ctrl = nullptr;
    void* tc_prep = nullptr;
    DevBuf xt;  // optional subspace-major copy of x (segmented update only); absent when memory is short

    void carve(WsBump& b) {
        const size_t gcap = std::max(m * k, (size_t)1);
        codes = b.take<uint8_t>(m * n * code_bytes);
        const bool radix = kind != UPD_TILES && kind != UPD_ORDERED_TILES;
        ids_a = b.take<uint32_t>(radix ? m * n : 1);
        ids_b = b.take<uint32_t>(radix && passes > 1 ? m * n : 1);
        chunk_hist = b.take<uint32_t>(radix ? m * (size_t)n_chunks * 256 : 1);
        bin_off = b.take<uint32_t>(m * 256);
        seg_beg = b.take<uint32_t>(m * k);
        seg_end = b.take<uint32_t>(m * k);
        partial = b.take<float>(m * k * (size_t)segs * d);
        cnt_partial = b.take<uint32_t>(!radix ? m * k * (size_t)segs : 1);
        pack = b.take<float>(m * k * d + 2 * m * k);
        cb = b.take<float>(m * k * d);
        sub_list = b.take<int>(m);
        is_active = b.take<int>(m);
        changed = b.take<uint32_t>(m);
        counts = b.take<uint32_t>(m * k);
        iters = b.take<uint32_t>(m);
        ctrl = b.take<TrainCtrl>(1);
        g_rows = b.take<long long>(gcap);
        g_subs = b.take<int>(gcap);
        g_dst = b.take<long long>(gcap);
        g_vals = b.take<float>(gcap * d);
        tc_prep = b.take<uint8_t>(vqb_tc_prep_bytes(m, d));
    }
    // Picks the update kernels for this shape and carves the arrays out of the context's grow-only slab.
    int setup(vqb_ctx* ctx, const float* x, size_t n_, size_t dim_, size_t m_, size_t k_, uint32_t update_mode,
              bool assign_only = false) {
        n = std::max<size_t>(n_, 1); dim = dim_; m = m_; k = k_; d = dim / m;
        code_bytes = k <= 256 ? 1 : (k <= 65536 ? 2 : 4);
        passes = (int)code_bytes;
        n_chunks = (int)cdiv(n, RX_CHUNK);
        kind = UPD_ORDERED; segs = 1;
        if (assign_only) {   // no update arrays (vqb_pq_assign_train)
            kind = UPD_TILES; ut_groups = 1; ut_parts = 1;
        } else if (update_mode != VQB_UPDATE_FAST) {
            // the reference's order without row-id arrays, when the shape allows the TMA-fed tile walk
            static const bool old_ordered = [] { const char* e = std::getenv("VQB_ORDERED_RADIX"); return e && *e && *e != '0'; }();
            const bool ok = !old_ordered && k <= 256 && d % 4 == 0 && d >= 4 && d <= 32 && dim % 4 == 0 &&
                            (reinterpret_cast<uintptr_t>(x) & 15) == 0 && n_ >= 4096 && n_ < ((size_t)1 << 31) && m <= 65535;
            if (ok) kind = UPD_ORDERED_TILES;
        } else {
            static const int fast_segs = [] { const char* e = std::getenv("VQB_SEGS"); int v = e ? std::atoi(e) : 0; return v > 0 ? v : 128; }();
            const bool tiles_ok = k <= 256 && (d == 8 || d == 16 || d == 32) && dim % 4 == 0 &&
                                  (reinterpret_cast<uintptr_t>(x) & 15) == 0 && n_ >= 4096 && n_ < ((size_t)1 << 31);
            if (tiles_ok) {
                kind = UPD_TILES;
                ut_groups = (int)cdiv(m, 32 / d);
                const int tiles = (int)cdiv(n_, UT_ROWS);
                ut_parts = std::max(1, std::min(tiles, ctx->sm_count / ut_groups));
                segs = ut_parts;
            } else {
                kind = UPD_SEGMENTED; segs = fast_segs;
            }
        }
        WsBump size_pass;
        carve(size_pass);
        VQB_CUDA(ctx, vqb_ws_reserve(ctx, size_pass.off));
        WsBump real;
        real.base = static_cast<char*>(ctx->ws);
        carve(real);
        n = n_;
        return VQB_SUCCESS;
    }
};

struct TrainArgs {
    const float* x; size_t n, dim, m, k, d;
    uint32_t assign_mode;
    vqb_allreduce_fn allreduce; void* allreduce_user;
    bool use_comm;
    uint64_t row_offset;
};

int exchange(vqb_ctx* ctx, const TrainArgs& a, float* buf, size_t count) {
    if (a.allreduce) {
        int rc = a.allreduce(a.allreduce_user, buf, count, (void*)ctx->stream);
        if (rc != 0) return vqb_fail(ctx, VQB_FAILURE, "allreduce callback failed (%d)", rc);
    } else if (a.use_comm) {
        VQB_TRY(vqb_comm_allreduce_f32(ctx, buf, count));
    }
    return VQB_SUCCESS;
}

// codes[s*n + row] for the subspaces of ws.sub_list (device-side list; gated by ws.ctrl->go)
int assign_train(vqb_ctx* ctx, const TrainArgs& a, TrainWs& ws, void* codes, uint32_t code_bytes, bool gated) {
    const uint32_t* go = gated ? &ws.ctrl->go : nullptr;
    const bool tc_ok = ws.tc_prep && vqb_tc_supported(MK_TRAIN, a.x, a.n, a.dim, a.m, a.k, a.d);
    if (a.assign_mode == VQB_ASSIGN_TENSOR && !tc_ok)
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT,
                        "tensor-core assignment needs sub_dim 8, k <= 256 and 16-byte aligned rows");
    if (tc_ok && (a.assign_mode == VQB_ASSIGN_TENSOR || (a.assign_mode == VQB_ASSIGN_AUTO && a.n >= VQB_TC_MIN_ROWS))) {
        VQB_TRY(vqb_tc_prepare(ctx, MK_TRAIN, ws.cb, a.m, a.k, a.d, ws.tc_prep, go));
        return vqb_tc_assign_launch(ctx, MK_TRAIN, a.x, a.n, a.dim, a.m, a.k, ws.tc_prep, ws.is_active, codes, code_bytes,
                                    /*stride_row=*/1, /*stride_sub=*/a.n, nullptr, nullptr, nullptr, 0, nullptr, 0, go);
    }
    return vqb_pq_assign_exact_launch(ctx, MK_TRAIN, a.x, a.n, a.dim, a.m, a.k, a.d, ws.cb, ws.sub_list, (int)a.m, codes,
                                      code_bytes, /*stride_row=*/1, /*stride_sub=*/a.n, nullptr, &ws.ctrl->n_active, go);
}

// Enqueues one iteration for the subspaces on the device-side active list.  Leaves ws.changed / ws.counts / the status
// record behind; every kernel is a no-op when ws.ctrl->go == 0.
int train_iteration(vqb_ctx* ctx, TrainWs& ws, const TrainArgs& a, uint32_t* status_slot, bool has_reseed, uint32_t iter_idx,
                    uint32_t max_iters) {
    const size_t n = a.n, k = a.k, d = a.d, m = a.m;
    if (n > 0) VQB_TRY(assign_train(ctx, a, ws, ws.codes, ws.code_bytes, true));

    if (ws.kind == UPD_TILES) {
        if (n > 0) {
            CUtensorMap map;
            VQB_TRY(vqb_make_x_tensormap(ctx, a.x, n, a.dim, &map));
            UtParams p;
            p.codes = static_cast<const uint8_t*>(ws.codes); p.n = n; p.m = (int)m; p.k = (int)k;
            p.n_groups = ws.ut_groups; p.parts = ws.ut_parts; p.num_tiles = (int)cdiv(n, UT_ROWS);
            p.is_active = ws.is_active; p.ctrl = ws.ctrl; p.partial = ws.partial; p.cnt_partial = ws.cnt_partial;
            const int grid = ws.ut_groups * ws.ut_parts;
#define VQB_UT_LAUNCH(DD)                                                                                              \
    do {                                                                                                               \
        auto kern = k_update_tiles<DD>;                                                                                \
        VQB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UT_SMEM));          \
        kern<<<grid, UT_THREADS, UT_SMEM, ctx->stream>>>(map, p);                                                      \
    } while (0)
            if (d == 8) VQB_UT_LAUNCH(8);
            else if (d == 16) VQB_UT_LAUNCH(16);
            else VQB_UT_LAUNCH(32);
#undef VQB_UT_LAUNCH
            VQB_LAUNCHED(ctx);
        } else {  // a rank without rows contributes zeros
            VQB_CUDA(ctx, cudaMemsetAsync(ws.partial, 0, m * k * (size_t)ws.segs * d * 4, ctx->stream));
            VQB_CUDA(ctx, cudaMemsetAsync(ws.cnt_partial, 0, m * k * (size_t)ws.segs * 4, ctx->stream));
        }
    } else if (ws.kind == UPD_ORDERED_TILES) {
        CUtensorMap map;
        VQB_TRY(vqb_make_x_tensormap(ctx, a.x, n, a.dim, &map, (int)d, /*swizzle=*/0));   // the kernel reads plain rows of d floats
        UoParams p;
        p.codes = static_cast<const uint8_t*>(ws.codes); p.n = n; p.m = (int)m; p.k = (int)k;
        p.sub_list = ws.sub_list; p.ctrl = ws.ctrl; p.partial = ws.partial; p.cnt_partial = ws.cnt_partial;
#define VQB_UO_LAUNCH(DD)                                                                                              \
    case DD: {                                                                                                         \
        auto kern = k_update_ordered<DD>;                                                                              \
        constexpr int rows_t = UoShape<DD>::ROWS;                                                                      \
        const int smem = 2 * rows_t * DD * 4 + (rows_t / 32) * 256 * 4 + 64 + 128;                                     \
        p.num_tiles = (int)cdiv(n, rows_t);                                                                            \
        VQB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                  \
        kern<<<(unsigned)m, UO_THREADS, smem, ctx->stream>>>(map, p);                                                  \
        break;                                                                                                         \
    }
        switch (d) {
            VQB_UO_LAUNCH(4) VQB_UO_LAUNCH(8) VQB_UO_LAUNCH(12) VQB_UO_LAUNCH(16) VQB_UO_LAUNCH(20) VQB_UO_LAUNCH(24)
            VQB_UO_LAUNCH(28) VQB_UO_LAUNCH(32)
            default: return vqb_fail(ctx, VQB_FAILURE, "internal: ordered tile update for sub_dim %zu", d);
        }
#undef VQB_UO_LAUNCH
        VQB_LAUNCHED(ctx);
    } else {
        // group: LSD radix, 8 bits per stable pass
        const uint32_t* ids_in = nullptr;
        uint32_t* ids_out = ws.ids_a;
        dim3 gchunks(std::max(ws.n_chunks, 1), (unsigned)m);
        for (int pass = 0; pass < ws.passes; ++pass) {
            bool last = pass == ws.passes - 1;
            bool ranges_here = last && ws.passes == 1;
            k_radix_hist<<<gchunks, RX_THREADS, 0, ctx->stream>>>(ws.codes, ws.code_bytes, n, pass * 8, ws.sub_list, ws.ctrl,
                                                                  ids_in, ws.chunk_hist, ws.n_chunks);
            VQB_LAUNCHED(ctx);
            k_radix_scan<<<(unsigned)m, 256, 0, ctx->stream>>>(ws.chunk_hist, ws.n_chunks, ws.sub_list, ws.ctrl, ws.bin_off,
                                                              ranges_here ? ws.seg_beg : nullptr, ws.seg_end, (int)k);
            VQB_LAUNCHED(ctx);
            k_radix_scatter<<<gchunks, RX_THREADS, 0, ctx->stream>>>(ws.codes, ws.code_bytes, n, pass * 8, ws.sub_list, ws.ctrl,
                                                                     ids_in, ws.chunk_hist, ws.bin_off, ws.n_chunks, ids_out);
            VQB_LAUNCHED(ctx);
            ids_in = ids_out;
            ids_out = (ids_out == ws.ids_a) ? ws.ids_b : ws.ids_a;
        }
        const uint32_t* sorted = ids_in;
        if (ws.passes > 1) {
            k_zero_ranges<<<cdiv(m * k, 256), 256, 0, ctx->stream>>>(ws.seg_beg, ws.seg_end, m * k, ws.ctrl);
            VQB_LAUNCHED(ctx);
            if (n > 0) {   // a rank without rows has no ranges to mark (and a zero-sized grid is an invalid launch)
                k_cluster_ranges<<<dim3(cdiv(n, 256), (unsigned)m), 256, 0, ctx->stream>>>(ws.codes, ws.code_bytes, n, (int)k,
                                                                                          ws.sub_list, ws.ctrl, sorted,
                                                                                          ws.seg_beg, ws.seg_end);
                VQB_LAUNCHED(ctx);
            }
        }
        // sum
        size_t chains = m * k * ws.segs * d;
        const bool use_xt = ws.xt.p != nullptr;
        k_chain_sums<<<cdiv(chains, 256), 256, 0, ctx->stream>>>(use_xt ? ws.xt.as<float>() : a.x, n, use_xt ? d : a.dim,
                                                                use_xt ? n * d : d, (int)d, (int)k, ws.segs, ws.sub_list,
                                                                ws.ctrl, sorted, ws.seg_beg, ws.seg_end, ws.partial);
        VQB_LAUNCHED(ctx);
    }
    size_t n_pack = m * k * d + 2 * m * k;
    k_pack<<<cdiv(n_pack, 256), 256, 0, ctx->stream>>>(ws.partial, ws.segs, (int)d, (int)k, (int)m, ws.seg_beg, ws.seg_end,
                                                      (ws.kind == UPD_TILES || ws.kind == UPD_ORDERED_TILES) ? ws.cnt_partial : nullptr,
                                                      ws.is_active, ws.ctrl,
                                                      ws.pack);
    VQB_LAUNCHED(ctx);
    VQB_TRY(exchange(ctx, a, ws.pack, n_pack));
    k_finalize<<<cdiv(m * k * d, 256), 256, 0, ctx->stream>>>(ws.pack, (int)d, (int)k, (int)m, ws.is_active, ws.ctrl, ws.cb,
                                                             ws.changed, ws.counts);
    VQB_LAUNCHED(ctx);
    k_epilogue<<<1, 1024, m * sizeof(int), ctx->stream>>>((int)m, (int)k, ws.is_active, ws.sub_list, ws.ctrl, ws.changed, ws.counts, ws.iters,
                                           status_slot, has_reseed ? 1 : 0, iter_idx, max_iters);
    VQB_LAUNCHED(ctx);
    return VQB_SUCCESS;
}

// Uploads a small host table through the pinned mailbox behind the status records (keeps the copy asynchronous and ordered).
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

// Writes x[row_e][s_e*d..] into codebooks[dst_e] for a host list; rows outside this rank's shard
// contribute zeros and the patch is summed across ranks.
int apply_rows(vqb_ctx* ctx, TrainWs& ws, const TrainArgs& a, const std::vector<long long>& rows,
               const std::vector<int>& subs, const std::vector<long long>& dst) {
    size_t cnt = rows.size();
    if (!cnt) return VQB_SUCCESS;
    size_t mb = STATUS_BYTES;
    // tables can exceed the mailbox (m*k entries): upload_small falls back to a blocking copy
    VQB_TRY(upload_small(ctx, ws.g_rows, rows.data(), cnt * 8, mb));
    VQB_TRY(upload_small(ctx, ws.g_subs, subs.data(), cnt * 4, mb));
    VQB_TRY(upload_small(ctx, ws.g_dst, dst.data(), cnt * 8, mb));
    k_gather_subvecs<<<cdiv(cnt * a.d, 256), 256, 0, ctx->stream>>>(a.x, (int)a.dim, (int)a.d, ws.g_rows, ws.g_subs, cnt, ws.g_vals);
    VQB_LAUNCHED(ctx);
    VQB_TRY(exchange(ctx, a, ws.g_vals, cnt * a.d));
    k_scatter_subvecs<<<cdiv(cnt * a.d, 256), 256, 0, ctx->stream>>>(ws.g_vals, (int)a.d, ws.g_dst, cnt, ws.cb);
    VQB_LAUNCHED(ctx);
    // the mailbox is reused by the next table: make sure the copies are done
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

int reset_state(vqb_ctx* ctx, TrainWs& ws) {
    k_train_reset<<<1, 256, 0, ctx->stream>>>((int)ws.m, ws.is_active, ws.sub_list, ws.ctrl, ws.changed, ws.iters);
    VQB_LAUNCHED(ctx);
    return VQB_SUCCESS;
}

struct EventPool {
    std::vector<cudaEvent_t> ev;
    ~EventPool() { for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e); }
    cudaError_t make(size_t count, unsigned flags) {
        ev.assign(count, nullptr);
        for (size_t i = 0; i < count; ++i) { cudaError_t e = cudaEventCreateWithFlags(&ev[i], flags); if (e != cudaSuccess) return e; }
        return cudaSuccess;
    }
};

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
    if (max_iters > 0xFFFFFFF0ull) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "max_iters too large");
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t d = dim / m;
    const bool use_comm = !o.allreduce && (o.flags & VQB_TRAIN_USE_COMM) != 0;
    if (use_comm && !ctx->nccl_comm)
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "VQB_TRAIN_USE_COMM needs vqb_comm_init_rank on this context first");
    if (ctx->mailbox_bytes < STATUS_BYTES + 4096) return vqb_fail(ctx, VQB_FAILURE, "context mailbox too small");

    const double t_enter = trace_on() ? now_ms() : 0.0;
    InputView xin;
    VQB_TRY(xin.bind(ctx, x, n * dim * sizeof(float), /*slot=*/0));
    TrainWs ws;
    VQB_TRY(ws.setup(ctx, static_cast<const float*>(xin.dev), n, dim, m, k, o.update_mode));
    if (trace_on()) std::fprintf(stderr, "[vqb trace] bind + workspace: %.3f ms (update kind %d)\n", now_ms() - t_enter, ws.kind);
    TrainArgs a{static_cast<const float*>(xin.dev), n, dim, m, k, d, o.assign_mode, o.allreduce, o.allreduce_user, use_comm,
                o.row_offset};
    // subspace-major copy for the segmented update's gathers: pays for itself after about two iterations; skipped for
    // small inputs, single-iteration calls, m == 1 (already contiguous) or when the memory is not there
    static const bool no_xt = [] { const char* e = std::getenv("VQB_NO_XT"); return e && *e && *e != '0'; }();
    if (ws.kind == UPD_SEGMENTED && !no_xt && max_iters >= 2 && m > 1 && n >= 65536) {
        if (ws.xt.alloc(n * dim * sizeof(float)) == cudaSuccess) {
            k_subspace_major<<<cdiv(n, 32), 256, 0, ctx->stream>>>(a.x, n, (int)dim, (int)d, (int)m, ws.xt.as<float>());
            VQB_LAUNCHED(ctx);
        } else {
            cudaGetLastError();  // not enough memory: gather from the caller's layout instead
            ws.xt.p = nullptr; ws.xt.bytes = 0;
        }
    }
    VQB_TRY(reset_state(ctx, ws));

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

    uint32_t* status = static_cast<uint32_t*>(ctx->mailbox);
    std::memset(status, 0, STATUS_BYTES);
    const uint32_t iters_max = (uint32_t)max_iters;
    EventPool done, tick;   // done[t & 1]: end of iteration t;  tick[2t], tick[2t+1]: its device time (diagnostics)
    VQB_CUDA(ctx, done.make(2, cudaEventDisableTiming));
    if (o.iter_ms) VQB_CUDA(ctx, tick.make(2 * max_iters, cudaEventDefault));
    auto enqueue = [&](uint32_t t) -> int {
        if (o.iter_ms) VQB_CUDA(ctx, cudaEventRecord(tick.ev[2 * t], ctx->stream));
        VQB_TRY(train_iteration(ctx, ws, a, status + (t % STATUS_SLOTS) * STATUS_WORDS, o.reseed != nullptr, t, iters_max));
        if (o.iter_ms) VQB_CUDA(ctx, cudaEventRecord(tick.ev[2 * t + 1], ctx->stream));
        VQB_CUDA(ctx, cudaEventRecord(done.ev[t & 1], ctx->stream));
        return VQB_SUCCESS;
    };
    uint32_t ran = 0;
    if (iters_max > 0) VQB_TRY(enqueue(0));
    for (uint32_t t = 0; t < iters_max; ++t) {
        const double t0 = trace_on() ? now_ms() : 0.0;
        // iteration t+1 goes in behind t before t's outcome is known; it is a no-op on the device unless t allows it
        if (t + 1 < iters_max) VQB_TRY(enqueue(t + 1));
        const double t1 = trace_on() ? now_ms() : 0.0;
        VQB_CUDA(ctx, cudaEventSynchronize(done.ev[t & 1]));
        volatile uint32_t* st = status + (t % STATUS_SLOTS) * STATUS_WORDS;
        if (st[0] != t + 1) return vqb_fail(ctx, VQB_FAILURE, "training loop lost track of iteration %u (status %u)", t, st[0]);
        ran = t + 1;
        const uint32_t n_active_next = st[1], n_empty = st[2];
        if (trace_on())
            std::fprintf(stderr, "[vqb trace] iter %u: enqueue(next)=%.3f ms, wait=%.3f ms, active next=%u, empty=%u\n", t, t1 - t0,
                         now_ms() - t1, n_active_next, n_empty);
        if (n_empty > 0 && o.reseed) {
            // vector.rs:448-452: every empty cluster of a subspace that ran is re-seeded, ascending j within the subspace
            std::vector<std::pair<uint32_t, uint32_t>> pairs;
            if (n_empty <= (uint32_t)STATUS_CAP) {
                for (uint32_t e = 0; e < n_empty; ++e) pairs.emplace_back(st[8 + 2 * e], st[9 + 2 * e]);
            } else {   // too many to list: read the counts back; the subspaces that ran are those whose count went up by one
                std::vector<uint32_t> h_counts(m * k), h_iters(m);
                VQB_CUDA(ctx, cudaMemcpyAsync(h_counts.data(), ws.counts, m * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
                VQB_CUDA(ctx, cudaMemcpyAsync(h_iters.data(), ws.iters, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
                VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                for (size_t s = 0; s < m; ++s)
                    if (h_iters[s] == t + 1)   // ran in every iteration so far, including this one
                        for (size_t j = 0; j < k; ++j)
                            if (h_counts[s * k + j] == 0) pairs.emplace_back((uint32_t)s, (uint32_t)j);
            }
            std::sort(pairs.begin(), pairs.end());
            std::vector<long long> rows, dst;
            std::vector<int> subs;
            for (auto& pr : pairs) {
                uint64_t g = o.reseed(o.reseed_user, pr.first);
                if (g >= n_global) g %= n_global;
                bool local = g >= o.row_offset && g < o.row_offset + n;
                rows.push_back(local ? (long long)(g - o.row_offset) : -1);
                subs.push_back((int)pr.first);
                dst.push_back((long long)((size_t)pr.first * k + pr.second));
            }
            VQB_TRY(apply_rows(ctx, ws, a, rows, subs, dst));
            if (n_active_next > 0 && t + 1 < iters_max) {   // the speculative t+1 was a no-op: issue it for real
                k_set_go<<<1, 1, 0, ctx->stream>>>(ws.ctrl, 1u);
                VQB_LAUNCHED(ctx);
                VQB_TRY(enqueue(t + 1));
            }
        }
        if (n_active_next == 0) break;
    }
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (o.iter_ms)
        for (uint32_t t = 0; t < ran; ++t) VQB_CUDA(ctx, cudaEventElapsedTime(&o.iter_ms[t], tick.ev[2 * t], tick.ev[2 * t + 1]));
    if (iters_run) {
        VQB_CUDA(ctx, cudaMemcpyAsync(iters_run, ws.iters, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    VQB_CUDA(ctx, cudaMemcpyAsync(codebooks, ws.cb, m * k * d * 4, cudaMemcpyDefault, ctx->stream));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (trace_on()) std::fprintf(stderr, "[vqb trace] train total: %.3f ms\n", now_ms() - t_enter);
    return VQB_SUCCESS;
}

int vqb_pq_assign_train(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k,
                        const float* codebooks, uint32_t assign_mode, uint32_t* codes_out) {
    if (!ctx) return VQB_ERR_NULL_PTR;
    VQB_TRY(validate_train(ctx, x, n, dim, m, std::min(k, n)));
    if (!codebooks || !codes_out) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t d = dim / m;
    InputView xin; OutputView ov;
    VQB_TRY(xin.bind(ctx, x, n * dim * 4, /*slot=*/0));
    VQB_TRY(ov.bind(ctx, codes_out, m * n * 4));
    TrainWs ws;
    VQB_TRY(ws.setup(ctx, static_cast<const float*>(xin.dev), n, dim, m, k, VQB_UPDATE_ORDERED, /*assign_only=*/true));
    VQB_CUDA(ctx, cudaMemcpyAsync(ws.cb, codebooks, m * k * d * 4, cudaMemcpyDefault, ctx->stream));
    VQB_TRY(reset_state(ctx, ws));
    TrainArgs a{static_cast<const float*>(xin.dev), n, dim, m, k, d, assign_mode, nullptr, nullptr, false, 0};
    VQB_TRY(assign_train(ctx, a, ws, ov.dev, 4, false));
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
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t d = dim / m;
    const bool use_comm = !o.allreduce && (o.flags & VQB_TRAIN_USE_COMM) != 0;
    if (use_comm && !ctx->nccl_comm)
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "VQB_TRAIN_USE_COMM needs vqb_comm_init_rank on this context first");
    InputView xin;
    VQB_TRY(xin.bind(ctx, x, n * dim * 4, /*slot=*/0));
    TrainWs ws;
    VQB_TRY(ws.setup(ctx, static_cast<const float*>(xin.dev), n, dim, m, k, o.update_mode));
    VQB_CUDA(ctx, cudaMemcpyAsync(ws.cb, codebooks_inout, m * k * d * 4, cudaMemcpyDefault, ctx->stream));
    VQB_TRY(reset_state(ctx, ws));
    TrainArgs a{static_cast<const float*>(xin.dev), n, dim, m, k, d, o.assign_mode, o.allreduce, o.allreduce_user, use_comm,
                o.row_offset};
    uint32_t* status = static_cast<uint32_t*>(ctx->mailbox);
    std::memset(status, 0, STATUS_WORDS * 4);
    VQB_TRY(train_iteration(ctx, ws, a, status, false, 0, 1));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // `changed` of this iteration == the subspaces still on the active list
    std::vector<int> act(m);
    VQB_CUDA(ctx, cudaMemcpy(act.data(), ws.is_active, m * 4, cudaMemcpyDeviceToHost));
    for (size_t s = 0; s < m; ++s) changed_out[s] = act[s] ? 1u : 0u;
    if (counts_out)
        VQB_CUDA(ctx, cudaMemcpyAsync(counts_out, ws.counts, m * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VQB_CUDA(ctx, cudaMemcpyAsync(codebooks_inout, ws.cb, m * k * d * 4, cudaMemcpyDefault, ctx->stream));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}

}  // extern "C"
