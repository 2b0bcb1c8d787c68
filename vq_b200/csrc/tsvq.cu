// tsvq.cu -- TSVQ construction (level-synchronous) and batched greedy-descent encoding.
//
// Reference: TSVQNode::build (src/tsvq.rs:31-115) is a recursive mean / max-variance-dimension /
// median split; TSVQNode::find_leaf (:117-132) descends comparing Distance to the two children.
// Nodes are independent, so all nodes of one depth are processed together here:
//   mean       sequential f32 column sums over the node's rows in parent order, / n   (vector.rs:332-348)
//   variance   sequential sum of (x - mean)^2 per column                              (tsvq.rs:47-57)
//   split dim  arg-max over non-NaN variances, LAST maximum wins                      (tsvq.rs:59-66)
//   median     exact order statistics of the split column by 4x8-bit radix select     (tsvq.rs:68-81)
//   partition  stable `<= median` split of the row-id permutation                     (tsvq.rs:84-85)
// The sums keep the reference's order (one chain per (node, column)), so node centroids and
// split decisions are bit-identical with the CPU result; rows are streamed through a 4-stage
// cp.async shared-memory ring so a chain's loads are deep in flight while its adds stay serial.
//
// HBM layout: X row-major [n, dim] f32 resident; perm [n] u32 (row ids grouped by node, parent
// order preserved); node centroids [n_nodes][dim] f32, breadth-first numbering.
#include "common.cuh"
#include "distance.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

struct vqb_tsvq {
    vqb_ctx* ctx = nullptr;
    size_t dim = 0, n_nodes = 0;
    int metric = 0;
    DevBuf cent, left, right;
    DevBuf cent_lane;  // [node][dim/64][16 lanes][4]: element j + 16 (4 q + e) of a centroid at (q * 16 + j) * 4 + e (built on first encode)
    std::vector<int32_t> h_left, h_right, h_split;
    std::vector<float> h_median;
    std::vector<uint64_t> h_count;
    int max_levels = 0;
};

namespace {

constexpr int CS_THREADS = 256;
constexpr int CS_SLICE = 32;              // columns per CTA (one accumulating warp)
constexpr int CS_ROWS = 128;              // rows per ring stage
constexpr int CS_STAGES = 4;
constexpr int CS_SMEM = CS_STAGES * CS_ROWS * CS_SLICE * 4;  // 64 KB
constexpr int PT_THREADS = 256;
constexpr int PT_STEPS = 8;
constexpr int PT_CHUNK = PT_THREADS * PT_STEPS;

struct NodeSeg { uint32_t beg, len; };
struct Chunk { uint32_t node, beg, len; };  // node = index into the level's split list

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// MODE 0: out[node][col] = (sum_rows x) / len          (mean_vector)
// MODE 1: out[node][col] = sum_rows (x - mean[col])^2   (variances)
// grid (cdiv(dim, 32), n_nodes); one chain per (node, column), rows streamed through the ring.
template <int MODE>
__global__ void __launch_bounds__(CS_THREADS)
k_colsum(const float* __restrict__ x, int dim, const uint32_t* __restrict__ perm,
         const NodeSeg* __restrict__ nodes, const float* __restrict__ mean /* MODE 1: [n_nodes][dim] */,
         float* __restrict__ out, int vec_ok) {
    extern __shared__ __align__(16) float ring[];  // [stage][row][32]
    const NodeSeg ns = nodes[blockIdx.y];
    const int col0 = blockIdx.x * CS_SLICE;
    const int ncol = min(CS_SLICE, dim - col0);
    const int tid = threadIdx.x;
    const int n_tiles = (int)((ns.len + CS_ROWS - 1) / CS_ROWS);
    const uint32_t* ids = perm + ns.beg;

    auto issue = [&](int tile) {
        if (tile < n_tiles) {
            float* st = ring + (size_t)(tile % CS_STAGES) * CS_ROWS * CS_SLICE;
            const int r0 = tile * CS_ROWS;
            const int rows = min(CS_ROWS, (int)ns.len - r0);
            if (vec_ok && ncol == CS_SLICE) {
                // 8 threads x 16 B cover one 128 B row slice; 256 threads cover 32 rows per pass
                for (int r = tid >> 3; r < rows; r += CS_THREADS / 8) {
                    const float* src = x + (size_t)ids[r0 + r] * dim + col0 + (tid & 7) * 4;
                    cp_async16(st + r * CS_SLICE + (tid & 7) * 4, src);
                }
            } else {
                for (int e = tid; e < rows * CS_SLICE; e += CS_THREADS) {
                    int r = e / CS_SLICE, c = e % CS_SLICE;
                    if (c < ncol) cp_async4(st + r * CS_SLICE + c, x + (size_t)ids[r0 + r] * dim + col0 + c);
                }
            }
        }
        cp_async_commit();  // one group per call keeps the wait arithmetic uniform
    };

    float acc = 0.0f, mu = 0.0f;
    if (MODE == 1 && tid < ncol) mu = mean[(size_t)blockIdx.y * dim + col0 + tid];
#pragma unroll
    for (int s = 0; s < CS_STAGES - 1; ++s) issue(s);
    for (int t = 0; t < n_tiles; ++t) {
        issue(t + CS_STAGES - 1);
        cp_async_wait<CS_STAGES - 1>();
        __syncthreads();
        if (tid < ncol) {
            const float* st = ring + (size_t)(t % CS_STAGES) * CS_ROWS * CS_SLICE + tid;
            const int rows = min(CS_ROWS, (int)ns.len - t * CS_ROWS);
            if (MODE == 0) {
#pragma unroll 8
                for (int r = 0; r < rows; ++r) acc = __fadd_rn(acc, st[r * CS_SLICE]);
            } else {
#pragma unroll 8
                for (int r = 0; r < rows; ++r) {
                    float df = __fsub_rn(st[r * CS_SLICE], mu);
                    acc = __fadd_rn(acc, __fmul_rn(df, df));
                }
            }
        }
        __syncthreads();
    }
    if (tid < ncol) {
        if (MODE == 0) acc = __fdiv_rn(acc, __uint2float_rn(ns.len));  // `vectors.len()` as f32
        out[(size_t)blockIdx.y * dim + col0 + tid] = acc;
    }
}

// The same chains, one WARP per (node, 32-column slice) with its own cp.async ring and no block-wide barrier:
// the sequential f32 chain of a (node, column) pair cannot be split without changing its rounding, so the
// kernel's speed is rows per cycle per chain.  Here a lane's chain advances one row per LDS + dependent FADD
// while the warp's next tiles are in flight, and hundreds of independent warps keep HBM busy on the levels that
// have enough (node, slice) pairs; on the first levels (48, 96, ... chains of 32 columns) the chain itself binds.
// Used (four warps, 4-deep rings, 64 KB per CTA) on the levels with more than ~900 chains; k_colsum_pc below serves
// the levels with fewer.
constexpr int CW_ROWS = 32;
constexpr int CW_SMEM = 16 * CW_ROWS * CS_SLICE * 4;  // 64 KB = warps x stages x 4 KB

template <int MODE, int CW_WARPS, int CW_STAGES>
__global__ void __launch_bounds__(CW_WARPS * 32)
k_colsum_w(const float* __restrict__ x, int dim, const uint32_t* __restrict__ perm, const NodeSeg* __restrict__ nodes,
           const float* __restrict__ mean, float* __restrict__ out, int n_slices, unsigned total_warps) {
    extern __shared__ __align__(16) float ring[];  // [warp][stage][row][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned wg = blockIdx.x * CW_WARPS + warp;
    if (wg >= total_warps) return;
    const unsigned node = wg / n_slices;
    const int col0 = (int)(wg % n_slices) * CS_SLICE;
    const NodeSeg ns = nodes[node];
    const uint32_t* ids = perm + ns.beg;
    float* myring = ring + (size_t)warp * CW_STAGES * CW_ROWS * CS_SLICE;
    const int n_tiles = (int)((ns.len + CW_ROWS - 1) / CW_ROWS);
    const int sub = lane >> 3, part = (lane & 7) * 4;  // 8 lanes x 16 B cover one 128-byte row slice

    // the row ids of a tile are fetched one call ahead, so no issue() waits on a global load
    uint32_t idv_next = (lane < (int)min((uint32_t)CW_ROWS, ns.len)) ? __ldg(ids + lane) : 0u;
    auto issue = [&](int tile) {
        if (tile < n_tiles) {
            float* st = myring + (size_t)(tile % CW_STAGES) * CW_ROWS * CS_SLICE;
            const int r0 = tile * CW_ROWS;
            const int rows = min(CW_ROWS, (int)ns.len - r0);
            const uint32_t idv = idv_next;
            const int rn = r0 + CW_ROWS + lane;
            idv_next = rn < (int)ns.len ? __ldg(ids + rn) : 0u;
#pragma unroll
            for (int p = 0; p < CW_ROWS / 4; ++p) {
                const int r = p * 4 + sub;
                const uint32_t id = __shfl_sync(0xFFFFFFFFu, idv, r);
                if (r < rows) cp_async16(st + r * CS_SLICE + part, x + (size_t)id * dim + col0 + part);
            }
        }
        cp_async_commit();  // one group per call keeps the wait arithmetic uniform
    };

    float acc = 0.0f, mu = 0.0f;
    if (MODE == 1) mu = mean[(size_t)node * dim + col0 + lane];
#pragma unroll
    for (int s = 0; s < CW_STAGES - 1; ++s) issue(s);
    for (int t = 0; t < n_tiles; ++t) {
        issue(t + CW_STAGES - 1);
        cp_async_wait<CW_STAGES - 1>();
        __syncwarp();
        const float* st = myring + (size_t)(t % CW_STAGES) * CW_ROWS * CS_SLICE + lane;
        const int rows = min(CW_ROWS, (int)ns.len - t * CW_ROWS);
        if (rows == CW_ROWS) {
            float v[CW_ROWS];
#pragma unroll
            for (int r = 0; r < CW_ROWS; ++r) v[r] = st[r * CS_SLICE];
#pragma unroll
            for (int r = 0; r < CW_ROWS; ++r) {
                if (MODE == 0) acc = __fadd_rn(acc, v[r]);
                else { const float df = __fsub_rn(v[r], mu); acc = __fadd_rn(acc, __fmul_rn(df, df)); }
            }
        } else {
            for (int r = 0; r < rows; ++r) {
                const float vv = st[r * CS_SLICE];
                if (MODE == 0) acc = __fadd_rn(acc, vv);
                else { const float df = __fsub_rn(vv, mu); acc = __fadd_rn(acc, __fmul_rn(df, df)); }
            }
        }
        __syncwarp();  // every lane is done with this stage before the next issue overwrites it
    }
    if (MODE == 0) acc = __fdiv_rn(acc, __uint2float_rn(ns.len));  // `vectors.len()` as f32
    out[(size_t)node * dim + col0 + lane] = acc;
}

// Few-chain levels (the first four or five): one CTA per (node, slice) with ONE accumulating warp and SEVEN
// producer warps.  The chain's dependent FADD (4 cycles a row) is the floor there, so the accumulating warp does
// nothing else: producers gather the rows with cp.async into a 16-stage ring and publish each tile through a
// shared-memory sequence word; the consumer publishes its progress so a stage is only overwritten once read.
//
// ROWS x STAGES: the tile size and ring depth.  Removing the loads or the chain from this kernel leaves most of its time
// (experiment of round 2: 23.7 ms for a depth-1 build, 22.7 ms with the producers' loads skipped, 18.9 ms with the
// consumer's chain skipped): on the few-node levels the publish / poll / progress hand-shake per 32-row tile is the cost, not
// memory or the dependent adds.  Levels with at most one CTA per SM therefore use 128-row tiles (a quarter of the hand-shakes
// per row, 12 x 16 KB stages), the other levels this kernel serves 64-row tiles (12 x 8 KB).  The ring needs at least
// producers x DEPTH + 1 stages (a producer publishes tile i only after issuing tile i + DEPTH: 6 stages dead-lock).
constexpr int CP_PROD = 7;
template <int MODE, int ROWS, int STAGES, int DEPTH>
__global__ void __launch_bounds__((CP_PROD + 1) * 32)
k_colsum_pc(const float* __restrict__ x, int dim, const uint32_t* __restrict__ perm, const NodeSeg* __restrict__ nodes,
            const float* __restrict__ mean, float* __restrict__ out, int n_slices, int sleep_ns) {
    constexpr int SB = ROWS / 32;                  // 32-row batches per tile
    const int nprod = (int)(blockDim.x >> 5) - 1;  // producer warps of this launch (<= CP_PROD)
    static_assert(STAGES >= CP_PROD * DEPTH + 1, "ring too shallow: a producer issues tile i + DEPTH before it publishes tile i, so it would wait for a stage whose tile it has not published");
    extern __shared__ __align__(16) float ring[];  // [stage][row][32]
    __shared__ volatile uint32_t full[STAGES];     // tile index + 1 that currently fills the stage
    __shared__ volatile uint32_t done;             // tiles the consumer has finished
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned node = blockIdx.x / n_slices;
    const int col0 = (int)(blockIdx.x % n_slices) * CS_SLICE;
    const NodeSeg ns = nodes[node];
    const uint32_t* ids = perm + ns.beg;
    const int n_tiles = (int)((ns.len + ROWS - 1) / ROWS);
    if (threadIdx.x < STAGES) full[threadIdx.x] = 0;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();

    if (warp == 0) {
        // ---- consumer: the chain.  Inside a tile the 32 values of the next batch are loaded into a second register array
        // before the current batch's dependent adds start (4 cycles each, three issue slots in four free), so only the first
        // batch of a tile waits for shared memory.
        float acc = 0.0f, mu = 0.0f;
        if (MODE == 1) mu = mean[(size_t)node * dim + col0 + lane];
        auto load32 = [&](const float* base, float (&dst)[32]) {
#pragma unroll
            for (int r = 0; r < 32; ++r) dst[r] = base[r * CS_SLICE];
        };
        auto chain32 = [&](const float (&src)[32]) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                if (MODE == 0) acc = __fadd_rn(acc, src[r]);
                else { const float df = __fsub_rn(src[r], mu); acc = __fadd_rn(acc, __fmul_rn(df, df)); }
            }
        };
        for (int t = 0; t < n_tiles; ++t) {
            const int stg = t % STAGES;
            while (full[stg] != (uint32_t)(t + 1)) { }
            __syncwarp();
            const float* st = ring + (size_t)stg * ROWS * CS_SLICE + lane;
            const int rows = min(ROWS, (int)ns.len - t * ROWS);
            if (rows == ROWS) {
                float va[32], vb[32];
                load32(st, va);
                if (SB == 1) chain32(va);
                else {
#pragma unroll
                    // one load of the next batch behind every dependent add of the running one (the add has a 4-cycle latency,
                    // the load fills one of the free issue slots)
                    auto chain_load = [&](const float (&src)[32], float (&dst)[32], const float* nxt) {
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            if (MODE == 0) acc = __fadd_rn(acc, src[r]);
                            else { const float df = __fsub_rn(src[r], mu); acc = __fadd_rn(acc, __fmul_rn(df, df)); }
                            dst[r] = nxt[r * CS_SLICE];
                        }
                    };
#pragma unroll
                    for (int b = 0; b < SB; b += 2) {
                        chain_load(va, vb, st + (size_t)(b + 1) * 32 * CS_SLICE);
                        if (b + 2 < SB) chain_load(vb, va, st + (size_t)(b + 2) * 32 * CS_SLICE);
                        else chain32(vb);
                    }
                }
            } else {
                for (int r = 0; r < rows; ++r) {
                    const float vv = st[r * CS_SLICE];
                    if (MODE == 0) acc = __fadd_rn(acc, vv);
                    else { const float df = __fsub_rn(vv, mu); acc = __fadd_rn(acc, __fmul_rn(df, df)); }
                }
            }
            __syncwarp();
            if (lane == 0) done = (uint32_t)(t + 1);
        }
        if (MODE == 0) acc = __fdiv_rn(acc, __uint2float_rn(ns.len));  // `vectors.len()` as f32
        out[(size_t)node * dim + col0 + lane] = acc;
    } else {
        // ---- producers: tiles p, p + CP_PROD, p + 2 CP_PROD, ...; row ids fetched one tile ahead
        const int p = warp - 1;
        const int sub = lane >> 3, part = (lane & 7) * 4;
        const int my_tiles = (n_tiles - p + nprod - 1) / nprod;
        auto publish = [&](int i) {  // this producer's i-th tile has landed
            const int tile = p + i * nprod;
            __syncwarp();
            __threadfence_block();
            if (lane == 0) full[tile % STAGES] = (uint32_t)(tile + 1);
        };
        auto fetch_ids = [&](long long r0, uint32_t (&idv)[SB]) {   // lane l holds rows r0 + 32 b + l
#pragma unroll
            for (int b = 0; b < SB; ++b) {
                const long long rn = r0 + 32 * b + lane;
                idv[b] = rn < (long long)ns.len ? __ldg(ids + rn) : 0u;
            }
        };
        uint32_t idv_next[SB];
        fetch_ids((long long)p * ROWS, idv_next);
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = p + i * nprod;
            const int r0 = tile * ROWS;
            const int rows = min(ROWS, (int)ns.len - r0);
            uint32_t idv[SB];
#pragma unroll
            for (int b = 0; b < SB; ++b) idv[b] = idv_next[b];
            fetch_ids((long long)r0 + (long long)nprod * ROWS, idv_next);
            // the stage's previous tile must have been consumed (plain polling: a measured nanosleep back-off here made
            // the first levels 15 % slower)
            if (tile >= STAGES) while ((int)done < tile - STAGES + 1) { if (sleep_ns) __nanosleep(sleep_ns); }
            float* st = ring + (size_t)(tile % STAGES) * ROWS * CS_SLICE;
#pragma unroll
            for (int b = 0; b < SB; ++b) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int r = b * 32 + q * 4 + sub;
                    const uint32_t id = __shfl_sync(0xFFFFFFFFu, idv[b], q * 4 + sub);
                    if (r < rows) cp_async16(st + r * CS_SLICE + part, x + (size_t)id * dim + col0 + part);
                }
            }
            cp_async_commit();
            if (i >= DEPTH) { cp_async_wait<DEPTH>(); publish(i - DEPTH); }
        }
        cp_async_wait<0>();
        for (int i = max(0, my_tiles - DEPTH); i < my_tiles; ++i) publish(i);
    }
}

// arg-max over non-NaN variances, last maximum wins (Iterator::max_by); all NaN -> 0.  One CTA per node.
__global__ void __launch_bounds__(256) k_argmax_last(const float* __restrict__ var, int dim, int* __restrict__ split_dim) {
    __shared__ float sv[256];
    __shared__ int si[256];
    const float* v = var + (size_t)blockIdx.x * dim;
    float best = 0.f; int bi = -1;
    // each thread scans a contiguous range so that "last maximum" is preserved by an ordered merge
    int per = (dim + 255) / 256;
    int b = threadIdx.x * per, e = min(dim, b + per);
    for (int i = b; i < e; ++i) {
        float f = v[i];
        if (isnan(f)) continue;
        if (bi < 0 || !(f < best)) { best = f; bi = i; }
    }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
        float g = 0.f; int gi = -1;
        for (int t = 0; t < 256; ++t) {
            if (si[t] < 0) continue;
            if (gi < 0 || !(sv[t] < g)) { g = sv[t]; gi = si[t]; }
        }
        split_dim[blockIdx.x] = gi < 0 ? 0 : gi;
    }
}

__device__ __forceinline__ uint32_t order_key(float f) {  // unsigned order == f32::total_cmp order
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(b);
}

// vals[i] = x[perm[i]][split_dim[node]] for every position of every split node; counts non-NaN.
__global__ void __launch_bounds__(PT_THREADS)
k_gather_split(const float* __restrict__ x, int dim, const uint32_t* __restrict__ perm,
               const Chunk* __restrict__ chunks, const int* __restrict__ split_dim,
               float* __restrict__ vals, uint32_t* __restrict__ n_valid) {
    const Chunk ch = chunks[blockIdx.x];
    const int sd = split_dim[ch.node];
    uint32_t local = 0;
    for (uint32_t i = threadIdx.x; i < ch.len; i += PT_THREADS) {
        float v = x[(size_t)perm[ch.beg + i] * dim + sd];
        vals[ch.beg + i] = v;
        local += !isnan(v);
    }
    // block reduction
    __shared__ uint32_t red[PT_THREADS / 32];
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < PT_THREADS / 32; ++w) t += red[w];
        if (t) atomicAdd(&n_valid[ch.node], t);
    }
}

struct SelState { uint32_t prefix[2]; uint32_t rem[2]; };

__global__ void k_select_init(const uint32_t* __restrict__ n_valid, SelState* __restrict__ st, int n_nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    uint32_t nv = n_valid[i];
    st[i].prefix[0] = st[i].prefix[1] = 0;
    if (nv == 0) { st[i].rem[0] = st[i].rem[1] = 0; return; }
    st[i].rem[0] = (nv % 2 == 0) ? nv / 2 - 1 : nv / 2;  // tsvq.rs:77-81
    st[i].rem[1] = nv / 2;
}

// histogram of digit (key >> shift) & 255 over the keys that share each target's current prefix
__global__ void __launch_bounds__(PT_THREADS)
k_select_hist(const float* __restrict__ vals, const Chunk* __restrict__ chunks, const SelState* __restrict__ st,
              int shift, uint32_t* __restrict__ hist /* [n_nodes][2][256] */) {
    __shared__ uint32_t h[2][256];
    h[0][threadIdx.x] = 0; h[1][threadIdx.x] = 0;
    __syncthreads();
    const Chunk ch = chunks[blockIdx.x];
    const SelState s = st[ch.node];
    const uint32_t hi_mask = shift >= 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (uint32_t i = threadIdx.x; i < ch.len; i += PT_THREADS) {
        float v = vals[ch.beg + i];
        if (isnan(v)) continue;  // tsvq.rs:71
        uint32_t key = order_key(v), dgt = (key >> shift) & 255u;
        if ((key & hi_mask) == (s.prefix[0] & hi_mask)) atomicAdd(&h[0][dgt], 1u);
        if ((key & hi_mask) == (s.prefix[1] & hi_mask)) atomicAdd(&h[1][dgt], 1u);
    }
    __syncthreads();
    uint32_t a = h[0][threadIdx.x], b = h[1][threadIdx.x];
    if (a) atomicAdd(&hist[((size_t)ch.node * 2 + 0) * 256 + threadIdx.x], a);
    if (b) atomicAdd(&hist[((size_t)ch.node * 2 + 1) * 256 + threadIdx.x], b);
}

__global__ void k_select_pick(SelState* __restrict__ st, uint32_t* __restrict__ hist, int shift, int n_nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes * 2) return;
    int node = i >> 1, w = i & 1;
    uint32_t* h = hist + (size_t)i * 256;
    uint32_t rem = st[node].rem[w], acc = 0;
    int dg = 255;
    for (int b = 0; b < 256; ++b) {
        uint32_t c = h[b];
        if (acc + c > rem) { dg = b; break; }
        acc += c;
    }
    st[node].rem[w] = rem - acc;
    st[node].prefix[w] |= (uint32_t)dg << shift;
    for (int b = 0; b < 256; ++b) h[b] = 0;  // ready for the next pass
}

__global__ void k_select_finish(const SelState* __restrict__ st, const uint32_t* __restrict__ n_valid,
                                float* __restrict__ median, int n_nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    uint32_t nv = n_valid[i];
    if (nv == 0) { median[i] = nanf(""); return; }
    float a = key_to_float(st[i].prefix[0]), b = key_to_float(st[i].prefix[1]);
    median[i] = (nv % 2 == 0) ? __fdiv_rn(__fadd_rn(a, b), 2.0f) : b;
}

// stable partition, pass 1: number of `<= median` elements per chunk
__global__ void __launch_bounds__(PT_THREADS)
k_part_count(const float* __restrict__ vals, const Chunk* __restrict__ chunks, const float* __restrict__ median,
             uint32_t* __restrict__ chunk_left) {
    const Chunk ch = chunks[blockIdx.x];
    const float med = median[ch.node];
    uint32_t local = 0;
    for (uint32_t i = threadIdx.x; i < ch.len; i += PT_THREADS) local += (vals[ch.beg + i] <= med);
    __shared__ uint32_t red[PT_THREADS / 32];
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xFFFFFFFFu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < PT_THREADS / 32; ++w) t += red[w];
        chunk_left[blockIdx.x] = t;
    }
}

// pass 2 (one thread per node): exclusive scan of its chunks' left counts; emits the node's left total
__global__ void k_part_scan(const uint32_t* __restrict__ node_chunk_beg, uint32_t* __restrict__ chunk_left,
                            uint32_t* __restrict__ node_left, int n_nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    uint32_t acc = 0;
    for (uint32_t c = node_chunk_beg[i]; c < node_chunk_beg[i + 1]; ++c) {
        uint32_t v = chunk_left[c];
        chunk_left[c] = acc;
        acc += v;
    }
    node_left[i] = acc;
}

// pass 3: stable scatter into perm_out (order inside each side preserved, tsvq.rs:84-85)
__global__ void __launch_bounds__(PT_THREADS)
k_part_scatter(const float* __restrict__ vals, const uint32_t* __restrict__ perm, const Chunk* __restrict__ chunks,
               const NodeSeg* __restrict__ nodes, const float* __restrict__ median,
               const uint32_t* __restrict__ chunk_left, const uint32_t* __restrict__ node_left,
               uint32_t* __restrict__ perm_out) {
    __shared__ uint32_t wl[PT_THREADS / 32];
    __shared__ uint32_t run_l, run_r;
    const Chunk ch = chunks[blockIdx.x];
    const NodeSeg ns = nodes[ch.node];
    const float med = median[ch.node];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nl_total = node_left[ch.node];
    if (threadIdx.x == 0) {
        run_l = chunk_left[blockIdx.x];
        run_r = (ch.beg - ns.beg) - chunk_left[blockIdx.x];  // elements before this chunk that went right
    }
    __syncthreads();
    for (uint32_t base = 0; base < ch.len; base += PT_THREADS) {
        uint32_t i = base + threadIdx.x;
        bool live = i < ch.len;
        bool goes_left = live && (vals[ch.beg + i] <= med);
        uint32_t bl = __ballot_sync(0xFFFFFFFFu, goes_left);
        if (lane == 0) wl[warp] = __popc(bl);
        __syncthreads();
        uint32_t before_l = 0;
        for (int w = 0; w < warp; ++w) before_l += wl[w];
        uint32_t lrank = before_l + __popc(bl & ((1u << lane) - 1u));
        uint32_t pos_in_step = threadIdx.x;             // live lanes are a prefix of the step
        uint32_t rrank = pos_in_step - lrank;           // earlier elements of this step that went right
        if (live) {
            uint32_t id = perm[ch.beg + i];
            if (goes_left) perm_out[ns.beg + run_l + lrank] = id;
            else perm_out[ns.beg + nl_total + run_r + rrank] = id;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tl = 0;
            for (int w = 0; w < PT_THREADS / 32; ++w) tl += wl[w];
            uint32_t step = min((uint32_t)PT_THREADS, ch.len - base);
            run_l += tl;
            run_r += step - tl;
        }
        __syncthreads();
    }
}

// ---- encode: one warp per vector, the two child distances on the two half-warps -----------------
// Lane l of a half-warp owns elements i == l (mod 16), i.e. exactly one AVX-512 lane of hsdlib's
// kernels, and the half-warp shuffle tree repeats _mm512_reduce_add_ps.
__device__ __forceinline__ float half_reduce16(float v) {
    v = __fadd_rn(v, __shfl_down_sync(0xFFFFFFFFu, v, 8, 16));
    v = __fadd_rn(v, __shfl_down_sync(0xFFFFFFFFu, v, 4, 16));
    v = __fadd_rn(v, __shfl_down_sync(0xFFFFFFFFu, v, 2, 16));
    v = __fadd_rn(v, __shfl_down_sync(0xFFFFFFFFu, v, 1, 16));
    return v;  // valid on lane 0 of the half-warp
}

template <int METRIC>
__device__ __forceinline__ float pair_distance_halfwarp(const float* __restrict__ v, const float* __restrict__ c,
                                                        int n, int hl) {
    const int nb = n & ~15;
    PtrAcc pa{v}, pb{c};
    float result = 0.f;
    if (METRIC == VQB_COSINE) {
        float d = 0.f, xa = 0.f, xb = 0.f;
        for (int i = hl; i < nb; i += 16) {
            float u = v[i], w = c[i];
            d = __fmaf_rn(u, w, d); xa = __fmaf_rn(u, u, xa); xb = __fmaf_rn(w, w, xb);
        }
        d = half_reduce16(d); xa = half_reduce16(xa); xb = half_reduce16(xb);
        if (hl == 0) {
            bool tail_ok = true;
            if (n < 16) { d = xa = xb = 0.f; }
            for (int i = nb; i < n; ++i) {
                float u = v[i], w = c[i];
                if (vqb_bad(u) || vqb_bad(w)) { tail_ok = false; break; }
                d = __fadd_rn(d, __fmul_rn(u, w));
                xa = __fadd_rn(xa, __fmul_rn(u, u));
                xb = __fadd_rn(xb, __fmul_rn(w, w));
            }
            bool ok = tail_ok;
            float sim = 0.f;
            if (ok) sim = hsd_cosine_from_sums(d, xa, xb, __fsqrt_rn(xa), __fsqrt_rn(xb), ok);
            result = ok ? __fsub_rn(1.0f, sim) : rust_cos<0>(pa, pb, n);
            if (n == 0) result = 0.f;
        }
    } else {
        float acc = 0.f;
        for (int i = hl; i < nb; i += 16) {
            float df = __fsub_rn(v[i], c[i]);
            acc = (METRIC == VQB_MANHATTAN) ? __fadd_rn(acc, fabsf(df)) : __fmaf_rn(df, df, acc);
        }
        acc = half_reduce16(acc);
        if (hl == 0) {
            bool ok = true;
            if (n < 16) acc = 0.f;
            for (int i = nb; i < n; ++i) {
                float u = v[i], w = c[i];
                if (vqb_bad(u) || vqb_bad(w)) { ok = false; break; }
                float df = __fsub_rn(u, w);
                acc = (METRIC == VQB_MANHATTAN) ? __fadd_rn(acc, fabsf(df)) : __fadd_rn(acc, __fmul_rn(df, df));
            }
            if (ok && vqb_bad(acc)) ok = false;
            if (!ok) acc = (METRIC == VQB_MANHATTAN) ? rust_l1<0>(pa, pb, n) : dist2_seq<0>(pa, pb, n);
            result = (METRIC == VQB_EUCLIDEAN) ? __fsqrt_rn(acc) : acc;
        }
    }
    return result;
}

template <int METRIC>
__global__ void __launch_bounds__(256)
k_tsvq_encode(const float* __restrict__ x, size_t n, int dim, const float* __restrict__ cent,
              const int* __restrict__ left, const int* __restrict__ right, uint32_t* __restrict__ leaf_out,
              __half* __restrict__ recon) {
    const size_t row = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= n) return;  // warp-uniform
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    const float* v = x + row * (size_t)dim;
    int node = 0;
    for (;;) {
        int l = left[node], r = right[node];
        if (l >= 0 && r >= 0) {
            const float* c = cent + (size_t)(half ? r : l) * dim;
            float dmine = pair_distance_halfwarp<METRIC>(v, c, dim, hl);
            float dl = __shfl_sync(0xFFFFFFFFu, dmine, 0), dr = __shfl_sync(0xFFFFFFFFu, dmine, 16);
            node = (dl <= dr) ? l : r;  // tsvq.rs:122
        } else if (l >= 0) node = l;
        else if (r >= 0) node = r;
        else break;
    }
    if (leaf_out && lane == 0) leaf_out[row] = (uint32_t)node;
    if (recon) {  // tsvq.rs:248-254
        const float* c = cent + (size_t)node * dim;
        __half* o = recon + row * (size_t)dim;
        for (int i = lane; i < dim; i += 32) o[i] = __float2half_rn(c[i]);
    }
}

// Lane-major form for the L2 / L1 metrics when dim is a multiple of 128.  hsdlib's kernel gives AVX lane j the elements
// j, j + 16, j + 32, ... and accumulates them in that order; the generic kernel above therefore issues one 4-byte load
// per element and lane and is bound by how many of those it keeps in flight (10.6 ms per 1M x 1536, depth 8).  Here the
// centroids are stored a second time as [dim/64][16 lanes][4] (`cent_lane`: four consecutive elements of a lane form one
// 16-byte unit, the units of the 16 lanes are adjacent), the vector is staged once in shared memory in the same order, and
// a lane reads both with 16-byte loads that are contiguous across the half-warp: a quarter of the load instructions, four
// times the bytes in flight, 256 contiguous bytes per half-warp and instruction.  The per-lane summation order -- and
// with it every bit of the distances -- is unchanged.
__global__ void k_lane_major(const float* __restrict__ cent, size_t n_nodes, int dim, float* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_nodes * dim) return;
    const size_t node = t / dim;
    const int i = (int)(t - node * dim), j = i & 15, tt = i >> 4;
    out[node * dim + (size_t)(((tt >> 2) * 16 + j) * 4 + (tt & 3))] = cent[t];
}

constexpr int TL_WARPS = 4;
template <int METRIC>
__global__ void __launch_bounds__(TL_WARPS * 32)
k_tsvq_encode_lm(const float* __restrict__ x, size_t n, int dim, const float* __restrict__ cent,
                 const float* __restrict__ cent_lane, const int* __restrict__ left, const int* __restrict__ right,
                 uint32_t* __restrict__ leaf_out, __half* __restrict__ recon) {
    extern __shared__ __align__(16) float xs_all[];   // [warp][dim/64][16 lanes][4], the layout of cent_lane
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    const size_t row = (size_t)blockIdx.x * TL_WARPS + warp;
    if (row >= n) return;  // warp-uniform
    const int nt = dim >> 4;
    float* xs = xs_all + (size_t)warp * dim;
    const float* v = x + row * (size_t)dim;
    for (int q = lane; q < dim / 4; q += 32) {   // element i = 4q + e belongs to lane i & 15, position tt = i >> 4
        const float4 a = __ldg(reinterpret_cast<const float4*>(v) + q);
        const int j0 = (4 * q) & 15, tt = (4 * q) >> 4;
        float* d = xs + ((tt >> 2) * 16 + j0) * 4 + (tt & 3);
        d[0] = a.x; d[4] = a.y; d[8] = a.z; d[12] = a.w;
    }
    __syncwarp();
    const float4* x4 = reinterpret_cast<const float4*>(xs) + hl;
    int node = 0;
    for (;;) {
        const int l = left[node], r = right[node];
        if (l >= 0 && r >= 0) {
            const int child = half ? r : l;
            const float4* c4 = reinterpret_cast<const float4*>(cent_lane + (size_t)child * dim) + hl;
            float acc = 0.f;
#pragma unroll 2
            for (int t0 = 0; t0 < nt / 4; t0 += 4) {   // 16 elements of this lane per step (nt is a multiple of 8)
                float4 cv[4], xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) if (t0 + u < nt / 4) cv[u] = __ldg(c4 + (t0 + u) * 16);
#pragma unroll
                for (int u = 0; u < 4; ++u) if (t0 + u < nt / 4) xv[u] = x4[(t0 + u) * 16];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (t0 + u < nt / 4) {
                        const float d0 = __fsub_rn(xv[u].x, cv[u].x), d1 = __fsub_rn(xv[u].y, cv[u].y);
                        const float d2 = __fsub_rn(xv[u].z, cv[u].z), d3 = __fsub_rn(xv[u].w, cv[u].w);
                        if (METRIC == VQB_MANHATTAN) {
                            acc = __fadd_rn(acc, fabsf(d0)); acc = __fadd_rn(acc, fabsf(d1));
                            acc = __fadd_rn(acc, fabsf(d2)); acc = __fadd_rn(acc, fabsf(d3));
                        } else {
                            acc = __fmaf_rn(d0, d0, acc); acc = __fmaf_rn(d1, d1, acc);
                            acc = __fmaf_rn(d2, d2, acc); acc = __fmaf_rn(d3, d3, acc);
                        }
                    }
                }
            }
            acc = half_reduce16(acc);
            float dmine = (METRIC == VQB_EUCLIDEAN) ? __fsqrt_rn(acc) : acc;
            // hsdlib rejects a non-finite result (the Rust side then recomputes sequentially): same route as the generic kernel
            const bool bad = (hl == 0) && vqb_bad(acc);
            if (__any_sync(0xFFFFFFFFu, bad)) dmine = pair_distance_halfwarp<METRIC>(v, cent + (size_t)child * dim, dim, hl);
            const float dl = __shfl_sync(0xFFFFFFFFu, dmine, 0), dr = __shfl_sync(0xFFFFFFFFu, dmine, 16);
            node = (dl <= dr) ? l : r;  // tsvq.rs:122
        } else if (l >= 0) node = l;
        else if (r >= 0) node = r;
        else break;
    }
    if (leaf_out && lane == 0) leaf_out[row] = (uint32_t)node;
    if (recon) {  // tsvq.rs:248-254
        const float4* c4 = reinterpret_cast<const float4*>(cent + (size_t)node * dim);
        uint2* o = reinterpret_cast<uint2*>(recon + row * (size_t)dim);
        for (int i = lane; i < dim / 4; i += 32) {
            const float4 cv = __ldg(c4 + i);
            __half2 a = __floats2half2_rn(cv.x, cv.y), b = __floats2half2_rn(cv.z, cv.w);
            uint2 w; w.x = *reinterpret_cast<uint32_t*>(&a); w.y = *reinterpret_cast<uint32_t*>(&b);
            o[i] = w;
        }
    }
}

__global__ void k_iota(uint32_t* __restrict__ p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

struct LevelNode { uint32_t id, beg, len, depth_left; };

// Per-level scratch carved out of one grow-only slab (a level used to cost ~14 cudaMalloc/cudaFree pairs = 3 ms).
struct SlabView {
    void* p = nullptr;
    template <typename T> T* as() const { return static_cast<T*>(p); }
};
// The per-level tables live behind the call's fixed arrays in the context's grow-only workspace (no cudaMalloc per call
// once the workspace has reached its size; the context mutex serialises its users).
struct Slab {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    static size_t pad(size_t b) { return (b + 255) & ~size_t(255); }
    bool fits(size_t bytes) { off = 0; return bytes <= cap; }
    SlabView take(size_t bytes) {
        SlabView v;
        v.p = base + off;
        off += pad(bytes);
        return v;
    }
};

}  // namespace

extern "C" {

int vqb_tsvq_destroy(vqb_tsvq* t) {
    if (!t) return VQB_ERR_NULL_PTR;
    cudaStreamSynchronize(t->ctx->stream);
    delete t;
    return VQB_SUCCESS;
}

int vqb_tsvq_num_nodes(vqb_tsvq* t, size_t* n_nodes, size_t* dim) {
    if (!t) return VQB_ERR_NULL_PTR;
    if (n_nodes) *n_nodes = t->n_nodes;
    if (dim) *dim = t->dim;
    return VQB_SUCCESS;
}

int vqb_tsvq_export(vqb_tsvq* t, float* centroids, int32_t* left, int32_t* right, int32_t* split_dim,
                    float* median, uint64_t* count) {
    if (!t) return VQB_ERR_NULL_PTR;
    vqb_ctx* ctx = t->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (centroids) {
        VQB_CUDA(ctx, cudaMemcpyAsync(centroids, t->cent.p, t->n_nodes * t->dim * 4, cudaMemcpyDefault, ctx->stream));
        VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (left) std::memcpy(left, t->h_left.data(), t->n_nodes * 4);
    if (right) std::memcpy(right, t->h_right.data(), t->n_nodes * 4);
    if (split_dim) std::memcpy(split_dim, t->h_split.data(), t->n_nodes * 4);
    if (median) std::memcpy(median, t->h_median.data(), t->n_nodes * 4);
    if (count) std::memcpy(count, t->h_count.data(), t->n_nodes * 8);
    return VQB_SUCCESS;
}

int vqb_tsvq_create(vqb_ctx* ctx, const float* centroids, const int32_t* left, const int32_t* right,
                    size_t n_nodes, size_t dim, int metric, vqb_tsvq** out) {
    if (!ctx || !out) return VQB_ERR_NULL_PTR;
    *out = nullptr;
    if (!centroids || !left || !right) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null tree arrays");
    if (n_nodes == 0 || dim == 0) return vqb_fail(ctx, VQB_ERR_EMPTY_INPUT, "empty tree");
    if (metric < 0 || metric > 3) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "unknown metric %d", metric);
    if (vqb_is_device_ptr(left) || vqb_is_device_ptr(right))
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "left/right must be host arrays");
    for (size_t i = 0; i < n_nodes; ++i)  // children must point forward (breadth-first numbering): no cycles
        if ((left[i] >= 0 && ((size_t)left[i] <= i || (size_t)left[i] >= n_nodes)) ||
            (right[i] >= 0 && ((size_t)right[i] <= i || (size_t)right[i] >= n_nodes)))
            return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "node %zu has an invalid child", i);
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    vqb_tsvq* t = new vqb_tsvq();
    t->ctx = ctx; t->dim = dim; t->n_nodes = n_nodes; t->metric = metric;
    t->h_left.assign(left, left + n_nodes);
    t->h_right.assign(right, right + n_nodes);
    t->h_split.assign(n_nodes, -1);
    t->h_median.assign(n_nodes, std::nanf(""));
    t->h_count.assign(n_nodes, 0);
    cudaError_t e = t->cent.alloc(n_nodes * dim * 4);
    if (e == cudaSuccess) e = t->left.alloc(n_nodes * 4);
    if (e == cudaSuccess) e = t->right.alloc(n_nodes * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->cent.p, centroids, n_nodes * dim * 4, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->left.p, left, n_nodes * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->right.p, right, n_nodes * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { delete t; return vqb_fail(ctx, VQB_FAILURE, "tree upload failed: %s", cudaGetErrorString(e)); }
    *out = t;
    return VQB_SUCCESS;
}

int vqb_tsvq_train(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t max_depth, int metric,
                   vqb_tsvq** out) {
    if (!ctx || !out) return VQB_ERR_NULL_PTR;
    *out = nullptr;
    if (n == 0) return vqb_fail(ctx, VQB_ERR_EMPTY_INPUT, "Empty input: at least one vector is required");
    if (!x) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null data pointer");
    if (dim == 0) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "dimension must be > 0");
    if (metric < 0 || metric > 3) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "unknown metric %d", metric);
    if (n > 0xFFFFFFF0ull || dim > (size_t)INT32_MAX) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "shape too large");
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    InputView xin;
    VQB_TRY(xin.bind(ctx, x, n * dim * 4, /*slot=*/0));
    const float* xd = static_cast<const float*>(xin.dev);
    const int vec_ok = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(xd) & 15) == 0);
    VQB_CUDA(ctx, cudaFuncSetAttribute(k_colsum<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
    VQB_CUDA(ctx, cudaFuncSetAttribute(k_colsum<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
    VQB_CUDA(ctx, (cudaFuncSetAttribute(k_colsum_w<0, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM)));
    VQB_CUDA(ctx, (cudaFuncSetAttribute(k_colsum_w<1, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM)));
    const unsigned cta_slots = (unsigned)ctx->sm_count * 3;  // 64 KB CTAs resident at once
    constexpr int PC_SMEM_128 = 12 * 128 * CS_SLICE * 4, PC_SMEM_64 = 12 * 64 * CS_SLICE * 4;
    VQB_CUDA(ctx, (cudaFuncSetAttribute(k_colsum_pc<0, 32, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM)));
    VQB_CUDA(ctx, (cudaFuncSetAttribute(k_colsum_pc<1, 32, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM)));
    VQB_CUDA(ctx, (cudaFuncSetAttribute(k_colsum_pc<0, 64, 12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM_64)));
    VQB_CUDA(ctx, (cudaFuncSetAttribute(k_colsum_pc<1, 64, 12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM_64)));
    VQB_CUDA(ctx, (cudaFuncSetAttribute(k_colsum_pc<0, 128, 12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM_128)));
    VQB_CUDA(ctx, (cudaFuncSetAttribute(k_colsum_pc<1, 128, 12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM_128)));
    static const int pc_prod = [] { const char* e = std::getenv("VQB_TSVQ_PROD"); int v = e ? std::atoi(e) : 3; return v >= 1 && v <= CP_PROD ? v : 3; }();   // 3 producer warps: 28.6 ms against 31.3 with 7 (1M x 1536, depth 8)
    static const int pc_sleep = [] { const char* e = std::getenv("VQB_TSVQ_PSLEEP"); return e ? std::atoi(e) : 0; }();
    static const int pc_force = [] { const char* e = std::getenv("VQB_TSVQ_PC_ROWS"); return e ? std::atoi(e) : 0; }();
    auto pc_rows = [&](unsigned ctas) -> int {   // tile rows of the producer / consumer kernel for a level of `ctas` CTAs
        if (pc_force == 32 || pc_force == 64 || pc_force == 128) return pc_force;
        if (ctas <= (unsigned)ctx->sm_count) return 128;
        return 64;   // measured per level (1M x 1536): 64-row tiles beat 32-row ones on every level this kernel serves
    };
    static const bool old_colsum = [] { const char* e = std::getenv("VQB_TSVQ_OLD_COLSUM"); return e && *e && *e != '0'; }();
    const bool warp_chains = vec_ok && dim % CS_SLICE == 0 && !old_colsum;  // else: block-wide ring kernel (any dim / alignment)
    const int n_slices = (int)(dim / CS_SLICE);

    // workspace: two row permutations, the split-coordinate values, then the per-level tables.  The widest level has at
    // most min(2^max_depth, n) nodes; its tables are bounded by level_need() below.
    const size_t max_level_nodes = (size_t)std::min<double>(std::ldexp(1.0, (int)std::min<size_t>(max_depth, 40)), (double)n);
    auto level_need = [&](size_t ln, size_t cmax) {
        return 3 * Slab::pad(ln * dim * 4) + 10 * Slab::pad(ln * 16) + Slab::pad(ln * 2 * 256 * 4) +
               Slab::pad(ln * sizeof(SelState)) + 2 * Slab::pad((cmax + 1) * sizeof(Chunk)) + 4096;
    };
    const size_t fixed = 3 * Slab::pad(n * 4);
    const size_t level_cap = level_need(max_level_nodes, n / PT_CHUNK + max_level_nodes + 1);
    VQB_CUDA(ctx, vqb_ws_reserve(ctx, fixed + level_cap));
    char* wsb = static_cast<char*>(ctx->ws);
    uint32_t* perm = reinterpret_cast<uint32_t*>(wsb);
    uint32_t* perm_next = reinterpret_cast<uint32_t*>(wsb + Slab::pad(n * 4));
    SlabView vals;
    vals.p = wsb + 2 * Slab::pad(n * 4);
    k_iota<<<cdiv(n, 256), 256, 0, st>>>(perm, n);
    VQB_LAUNCHED(ctx);

    // node centroids stay on the device: the levels come out in breadth-first id order, so every level's means are
    // written straight behind the previous level's in the tree's own centroid table (bounded by 2^(depth+1)-1 and 2n-1
    // nodes; a table that would not fit 1 GB is assembled through the host instead)
    const double ub_nodes_d = std::min(std::ldexp(1.0, (int)std::min<size_t>(max_depth + 1, 41)) - 1.0, 2.0 * (double)n - 1.0);
    const bool cent_on_device = ub_nodes_d * (double)dim * 4.0 <= 1024.0 * 1024.0 * 1024.0;
    DevBuf tree_cent;
    if (cent_on_device) VQB_CUDA(ctx, tree_cent.alloc((size_t)ub_nodes_d * dim * 4));

    std::vector<int32_t> h_left, h_right, h_split;
    std::vector<float> h_median;
    std::vector<uint64_t> h_count;
    std::vector<std::vector<float>> level_cent;  // host copies per level (<= a few MB each)
    std::vector<LevelNode> level{{0u, 0u, (uint32_t)n, (uint32_t)std::min<size_t>(max_depth, 0xFFFFFFFFu)}};
    h_left.push_back(-1); h_right.push_back(-1); h_split.push_back(-1);
    h_median.push_back(std::nanf("")); h_count.push_back(n);
    size_t n_nodes = 1;

    Slab slab;
    slab.base = wsb + fixed; slab.cap = ctx->ws_bytes - fixed;
    while (!level.empty()) {
        const size_t ln = level.size();
        // ---- means of every node of the level (tsvq.rs:36) ----
        std::vector<NodeSeg> segs(ln);
        for (size_t i = 0; i < ln; ++i) segs[i] = {level[i].beg, level[i].len};
        {   // everything this level can need: ln/sn-sized tables, three [nodes][dim] matrices, chunk tables
            size_t cmax = 0;
            for (size_t i = 0; i < ln; ++i) cmax += (level[i].len + PT_CHUNK - 1) / PT_CHUNK;
            if (!slab.fits(level_need(ln, cmax))) return vqb_fail(ctx, VQB_FAILURE, "internal: TSVQ level tables exceed the workspace");
        }
        SlabView d_segs = slab.take(ln * sizeof(NodeSeg)), d_mean = slab.take(ln * dim * 4);
        if (cent_on_device) d_mean.p = tree_cent.as<float>() + (size_t)level[0].id * dim;   // this level's rows of the tree table
        VQB_CUDA(ctx, cudaMemcpyAsync(d_segs.p, segs.data(), ln * sizeof(NodeSeg), cudaMemcpyHostToDevice, st));
        dim3 gcs(cdiv(dim, CS_SLICE), (unsigned)ln);
        if (gcs.y > 65535) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "too many nodes on one level");
        if (warp_chains) {
            const unsigned tw = (unsigned)(ln * n_slices);
            if (tw <= 2 * cta_slots) {
                const int pr = pc_rows(tw);
                if (pr == 128) k_colsum_pc<0, 128, 12, 1><<<tw, (pc_prod + 1) * 32, PC_SMEM_128, st>>>(xd, (int)dim, perm, d_segs.as<NodeSeg>(), nullptr, d_mean.as<float>(), n_slices, pc_sleep);
                else if (pr == 64) k_colsum_pc<0, 64, 12, 1><<<tw, (pc_prod + 1) * 32, PC_SMEM_64, st>>>(xd, (int)dim, perm, d_segs.as<NodeSeg>(), nullptr, d_mean.as<float>(), n_slices, pc_sleep);
                else k_colsum_pc<0, 32, 16, 2><<<tw, (CP_PROD + 1) * 32, CW_SMEM, st>>>(xd, (int)dim, perm, d_segs.as<NodeSeg>(), nullptr, d_mean.as<float>(), n_slices, 0);
            } else
                k_colsum_w<0, 4, 4><<<cdiv(tw, 4), 128, CW_SMEM, st>>>(xd, (int)dim, perm, d_segs.as<NodeSeg>(), nullptr, d_mean.as<float>(), n_slices, tw);
        } else
        k_colsum<0><<<gcs, CS_THREADS, CS_SMEM, st>>>(xd, (int)dim, perm, d_segs.as<NodeSeg>(), nullptr,
                                                     d_mean.as<float>(), vec_ok);
        VQB_LAUNCHED(ctx);
        if (!cent_on_device) {
            level_cent.emplace_back(ln * dim);
            VQB_CUDA(ctx, cudaMemcpyAsync(level_cent.back().data(), d_mean.p, ln * dim * 4, cudaMemcpyDeviceToHost, st));
        }

        // ---- nodes that try to split (tsvq.rs:38) ----
        std::vector<uint32_t> split_idx;
        for (size_t i = 0; i < ln; ++i)
            if (level[i].depth_left > 0 && level[i].len > 1) split_idx.push_back((uint32_t)i);
        if (split_idx.empty()) { VQB_CUDA(ctx, cudaStreamSynchronize(st)); break; }
        const size_t sn = split_idx.size();
        std::vector<NodeSeg> ssegs(sn);
        std::vector<Chunk> chunks;
        std::vector<uint32_t> node_chunk_beg(sn + 1);
        for (size_t q = 0; q < sn; ++q) {
            const LevelNode& nd = level[split_idx[q]];
            ssegs[q] = {nd.beg, nd.len};
            node_chunk_beg[q] = (uint32_t)chunks.size();
            for (uint32_t o = 0; o < nd.len; o += PT_CHUNK)
                chunks.push_back({(uint32_t)q, nd.beg + o, std::min<uint32_t>(PT_CHUNK, nd.len - o)});
        }
        node_chunk_beg[sn] = (uint32_t)chunks.size();
        const size_t cn = chunks.size();
        const bool all_split = sn == ln;  // then the level's means are already the splitting nodes' means, in order
        SlabView d_ssegs = slab.take(sn * sizeof(NodeSeg));
        SlabView d_smean = all_split ? d_mean : slab.take(sn * dim * 4);
        SlabView d_var = slab.take(sn * dim * 4), d_sd = slab.take(sn * 4), d_chunks = slab.take(cn * sizeof(Chunk));
        SlabView d_ncb = slab.take((sn + 1) * 4), d_nvalid = slab.take(sn * 4), d_sel = slab.take(sn * sizeof(SelState));
        SlabView d_hist = slab.take(sn * 2 * 256 * 4), d_median = slab.take(sn * 4), d_cleft = slab.take(cn * 4);
        SlabView d_nleft = slab.take(sn * 4);
        VQB_CUDA(ctx, cudaMemcpyAsync(d_ssegs.p, ssegs.data(), sn * sizeof(NodeSeg), cudaMemcpyHostToDevice, st));
        VQB_CUDA(ctx, cudaMemcpyAsync(d_chunks.p, chunks.data(), cn * sizeof(Chunk), cudaMemcpyHostToDevice, st));
        VQB_CUDA(ctx, cudaMemcpyAsync(d_ncb.p, node_chunk_beg.data(), (sn + 1) * 4, cudaMemcpyHostToDevice, st));
        // compact the means of the splitting nodes (device-to-device row copies)
        for (size_t q = 0; q < sn && !all_split; ++q)
            VQB_CUDA(ctx, cudaMemcpyAsync(d_smean.as<float>() + q * dim, d_mean.as<float>() + (size_t)split_idx[q] * dim,
                                          dim * 4, cudaMemcpyDeviceToDevice, st));
        VQB_CUDA(ctx, cudaMemsetAsync(d_nvalid.p, 0, sn * 4, st));
        VQB_CUDA(ctx, cudaMemsetAsync(d_hist.p, 0, sn * 2 * 256 * 4, st));

        dim3 gvs(cdiv(dim, CS_SLICE), (unsigned)sn);
        if (warp_chains) {
            const unsigned tw = (unsigned)(sn * n_slices);
            if (tw <= 2 * cta_slots) {
                const int pr = pc_rows(tw);
                if (pr == 128) k_colsum_pc<1, 128, 12, 1><<<tw, (pc_prod + 1) * 32, PC_SMEM_128, st>>>(xd, (int)dim, perm, d_ssegs.as<NodeSeg>(), d_smean.as<float>(), d_var.as<float>(), n_slices, pc_sleep);
                else if (pr == 64) k_colsum_pc<1, 64, 12, 1><<<tw, (pc_prod + 1) * 32, PC_SMEM_64, st>>>(xd, (int)dim, perm, d_ssegs.as<NodeSeg>(), d_smean.as<float>(), d_var.as<float>(), n_slices, pc_sleep);
                else k_colsum_pc<1, 32, 16, 2><<<tw, (CP_PROD + 1) * 32, CW_SMEM, st>>>(xd, (int)dim, perm, d_ssegs.as<NodeSeg>(), d_smean.as<float>(), d_var.as<float>(), n_slices, 0);
            } else
                k_colsum_w<1, 4, 4><<<cdiv(tw, 4), 128, CW_SMEM, st>>>(xd, (int)dim, perm, d_ssegs.as<NodeSeg>(), d_smean.as<float>(), d_var.as<float>(), n_slices, tw);
        } else
        k_colsum<1><<<gvs, CS_THREADS, CS_SMEM, st>>>(xd, (int)dim, perm, d_ssegs.as<NodeSeg>(), d_smean.as<float>(),
                                                     d_var.as<float>(), vec_ok);
        VQB_LAUNCHED(ctx);
        k_argmax_last<<<(unsigned)sn, 256, 0, st>>>(d_var.as<float>(), (int)dim, d_sd.as<int>());
        VQB_LAUNCHED(ctx);
        k_gather_split<<<(unsigned)cn, PT_THREADS, 0, st>>>(xd, (int)dim, perm, d_chunks.as<Chunk>(), d_sd.as<int>(),
                                                           vals.as<float>(), d_nvalid.as<uint32_t>());
        VQB_LAUNCHED(ctx);
        k_select_init<<<cdiv(sn, 128), 128, 0, st>>>(d_nvalid.as<uint32_t>(), d_sel.as<SelState>(), (int)sn);
        VQB_LAUNCHED(ctx);
        for (int shift = 24; shift >= 0; shift -= 8) {
            k_select_hist<<<(unsigned)cn, PT_THREADS, 0, st>>>(vals.as<float>(), d_chunks.as<Chunk>(),
                                                              d_sel.as<SelState>(), shift, d_hist.as<uint32_t>());
            VQB_LAUNCHED(ctx);
            k_select_pick<<<cdiv(sn * 2, 64), 64, 0, st>>>(d_sel.as<SelState>(), d_hist.as<uint32_t>(), shift, (int)sn);
            VQB_LAUNCHED(ctx);
        }
        k_select_finish<<<cdiv(sn, 128), 128, 0, st>>>(d_sel.as<SelState>(), d_nvalid.as<uint32_t>(),
                                                      d_median.as<float>(), (int)sn);
        VQB_LAUNCHED(ctx);
        k_part_count<<<(unsigned)cn, PT_THREADS, 0, st>>>(vals.as<float>(), d_chunks.as<Chunk>(), d_median.as<float>(),
                                                         d_cleft.as<uint32_t>());
        VQB_LAUNCHED(ctx);
        k_part_scan<<<cdiv(sn, 64), 64, 0, st>>>(d_ncb.as<uint32_t>(), d_cleft.as<uint32_t>(), d_nleft.as<uint32_t>(), (int)sn);
        VQB_LAUNCHED(ctx);
        // rows of nodes that do not split keep their place: start from a copy of the permutation
        VQB_CUDA(ctx, cudaMemcpyAsync(perm_next, perm, n * 4, cudaMemcpyDeviceToDevice, st));
        k_part_scatter<<<(unsigned)cn, PT_THREADS, 0, st>>>(vals.as<float>(), perm, d_chunks.as<Chunk>(),
                                                           d_ssegs.as<NodeSeg>(), d_median.as<float>(),
                                                           d_cleft.as<uint32_t>(), d_nleft.as<uint32_t>(), perm_next);
        VQB_LAUNCHED(ctx);

        std::vector<uint32_t> nleft(sn), nvalid(sn);
        std::vector<int> sd(sn);
        std::vector<float> med(sn);
        VQB_CUDA(ctx, cudaMemcpyAsync(nleft.data(), d_nleft.p, sn * 4, cudaMemcpyDeviceToHost, st));
        VQB_CUDA(ctx, cudaMemcpyAsync(nvalid.data(), d_nvalid.p, sn * 4, cudaMemcpyDeviceToHost, st));
        VQB_CUDA(ctx, cudaMemcpyAsync(sd.data(), d_sd.p, sn * 4, cudaMemcpyDeviceToHost, st));
        VQB_CUDA(ctx, cudaMemcpyAsync(med.data(), d_median.p, sn * 4, cudaMemcpyDeviceToHost, st));
        VQB_CUDA(ctx, cudaStreamSynchronize(st));

        std::vector<LevelNode> next;
        for (size_t q = 0; q < sn; ++q) {
            const LevelNode nd = level[split_idx[q]];
            if (nvalid[q] == 0)  // every split coordinate is NaN: the reference indexes an empty Vec and panics
                return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "node %u: split column holds only NaN", nd.id);
            h_split[nd.id] = sd[q];
            h_median[nd.id] = med[q];
            uint32_t nl = nleft[q], nr = nd.len - nl;
            if (nl > 0 && nl < nd.len) {  // tsvq.rs:88
                h_left[nd.id] = (int32_t)n_nodes;
                next.push_back({(uint32_t)n_nodes, nd.beg, nl, nd.depth_left - 1});
                h_left.push_back(-1); h_right.push_back(-1); h_split.push_back(-1);
                h_median.push_back(std::nanf("")); h_count.push_back(nl);
                ++n_nodes;
            }
            if (nr > 0 && nr < nd.len) {  // tsvq.rs:99
                h_right[nd.id] = (int32_t)n_nodes;
                next.push_back({(uint32_t)n_nodes, nd.beg + nl, nr, nd.depth_left - 1});
                h_left.push_back(-1); h_right.push_back(-1); h_split.push_back(-1);
                h_median.push_back(std::nanf("")); h_count.push_back(nr);
                ++n_nodes;
            }
        }
        std::swap(perm, perm_next);
        level.swap(next);
    }
    VQB_CUDA(ctx, cudaStreamSynchronize(st));

    // the breadth-first centroid table (levels were produced in id order)
    std::vector<float> cent;
    if (!cent_on_device) {
        cent.resize(n_nodes * dim);
        size_t off = 0;
        for (auto& lc : level_cent) { std::memcpy(cent.data() + off, lc.data(), lc.size() * 4); off += lc.size(); }
        if (off != n_nodes * dim) return vqb_fail(ctx, VQB_FAILURE, "internal: centroid table size mismatch");
    }

    vqb_tsvq* t = new vqb_tsvq();
    t->ctx = ctx; t->dim = dim; t->n_nodes = n_nodes; t->metric = metric;
    t->h_left = h_left; t->h_right = h_right; t->h_split = h_split; t->h_median = h_median; t->h_count = h_count;
    cudaError_t e = cudaSuccess;
    if (cent_on_device) { std::swap(t->cent.p, tree_cent.p); std::swap(t->cent.bytes, tree_cent.bytes); }
    else e = t->cent.alloc(n_nodes * dim * 4);
    if (e == cudaSuccess) e = t->left.alloc(n_nodes * 4);
    if (e == cudaSuccess) e = t->right.alloc(n_nodes * 4);
    if (e == cudaSuccess && !cent_on_device) e = cudaMemcpyAsync(t->cent.p, cent.data(), n_nodes * dim * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->left.p, h_left.data(), n_nodes * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->right.p, h_right.data(), n_nodes * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { delete t; return vqb_fail(ctx, VQB_FAILURE, "tree upload failed: %s", cudaGetErrorString(e)); }
    *out = t;
    return VQB_SUCCESS;
}

int vqb_tsvq_encode(vqb_tsvq* t, const float* x, size_t n, uint32_t* leaf_out, uint16_t* recon_out) {
    if (!t) return VQB_ERR_NULL_PTR;
    vqb_ctx* ctx = t->ctx;
    if (n == 0) return VQB_SUCCESS;
    if (!x) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "null input");
    if (!leaf_out && !recon_out) return vqb_fail(ctx, VQB_ERR_NULL_PTR, "no output requested");
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t dim = t->dim;
    // host inputs / outputs: row chunks over three streams (common.cuh), so arbitrarily large batches fit and the copies
    // overlap the descent; device buffers: one asynchronous launch
    ChunkIo io;
    io.in = x; io.in_unit = dim * 4;
    io.out0 = leaf_out; io.out0_unit = 4;
    io.out1 = recon_out; io.out1_unit = dim * 2;
    return vqb_chunk_pipeline(ctx, n, io, [&](const void* din, void* d0, void* d1, size_t rows, size_t) -> int {
        const float* xd = static_cast<const float*>(din);
        uint32_t* ld = static_cast<uint32_t*>(d0);
        __half* rd = static_cast<__half*>(d1);
        unsigned grid = cdiv(rows, 8);
        static const bool old_enc = [] { const char* e = std::getenv("VQB_TSVQ_OLD_ENCODE"); return e && *e && *e != '0'; }();
        const size_t lm_smem = (size_t)TL_WARPS * dim * 4;
        const bool reg_ok = !old_enc && t->metric != VQB_COSINE && dim % 128 == 0 && lm_smem <= 200 * 1024 &&
                            (reinterpret_cast<uintptr_t>(rd) & 7) == 0 && (reinterpret_cast<uintptr_t>(xd) & 15) == 0;
        if (reg_ok) {
            if (!t->cent_lane.p) {  // derived copy of the immutable centroids, made once (calls are serialised by ctx->mu)
                VQB_CUDA(ctx, t->cent_lane.alloc(t->n_nodes * dim * 4));
                k_lane_major<<<cdiv(t->n_nodes * dim, 256), 256, 0, ctx->stream>>>(t->cent.as<float>(), t->n_nodes, (int)dim,
                                                                                  t->cent_lane.as<float>());
                VQB_LAUNCHED(ctx);
            }
            const unsigned g4 = cdiv(rows, TL_WARPS);
            const float* cd = t->cent.as<float>();
            const float* cl = t->cent_lane.as<float>();
            const int* lf = t->left.as<int>();
            const int* rt = t->right.as<int>();
            VQB_CUDA(ctx, cudaFuncSetAttribute(k_tsvq_encode_lm<VQB_SQUARED_EUCLIDEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lm_smem));
            VQB_CUDA(ctx, cudaFuncSetAttribute(k_tsvq_encode_lm<VQB_EUCLIDEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lm_smem));
            VQB_CUDA(ctx, cudaFuncSetAttribute(k_tsvq_encode_lm<VQB_MANHATTAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lm_smem));
            if (t->metric == VQB_SQUARED_EUCLIDEAN)
                k_tsvq_encode_lm<VQB_SQUARED_EUCLIDEAN><<<g4, TL_WARPS * 32, lm_smem, ctx->stream>>>(xd, rows, (int)dim, cd, cl, lf, rt, ld, rd);
            else if (t->metric == VQB_EUCLIDEAN)
                k_tsvq_encode_lm<VQB_EUCLIDEAN><<<g4, TL_WARPS * 32, lm_smem, ctx->stream>>>(xd, rows, (int)dim, cd, cl, lf, rt, ld, rd);
            else
                k_tsvq_encode_lm<VQB_MANHATTAN><<<g4, TL_WARPS * 32, lm_smem, ctx->stream>>>(xd, rows, (int)dim, cd, cl, lf, rt, ld, rd);
        } else
        switch (t->metric) {
            case VQB_SQUARED_EUCLIDEAN:
                k_tsvq_encode<VQB_SQUARED_EUCLIDEAN><<<grid, 256, 0, ctx->stream>>>(xd, rows, (int)dim, t->cent.as<float>(), t->left.as<int>(), t->right.as<int>(), ld, rd); break;
            case VQB_EUCLIDEAN:
                k_tsvq_encode<VQB_EUCLIDEAN><<<grid, 256, 0, ctx->stream>>>(xd, rows, (int)dim, t->cent.as<float>(), t->left.as<int>(), t->right.as<int>(), ld, rd); break;
            case VQB_MANHATTAN:
                k_tsvq_encode<VQB_MANHATTAN><<<grid, 256, 0, ctx->stream>>>(xd, rows, (int)dim, t->cent.as<float>(), t->left.as<int>(), t->right.as<int>(), ld, rd); break;
            default:
                k_tsvq_encode<VQB_COSINE><<<grid, 256, 0, ctx->stream>>>(xd, rows, (int)dim, t->cent.as<float>(), t->left.as<int>(), t->right.as<int>(), ld, rd); break;
        }
        VQB_LAUNCHED(ctx);
        return VQB_SUCCESS;
    });
}

}  // extern "C"
