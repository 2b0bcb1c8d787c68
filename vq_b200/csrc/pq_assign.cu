// pq_assign.cu -- exact nearest-centroid assignment on CUDA cores.
//
// Evaluates, for every (vector, subspace) pair, the reference's distance formula against
// every centroid with the reference's own operation order and its strict-'<' / first-minimum
// rule (src/core/vector.rs:352-363 for training, src/pq.rs:177-196 + src/core/distance.rs:48-120
// for encoding), so codes are bit-identical with the CPU result.  It is the path for
//   * Manhattan (not a contraction: stays on CUDA cores by design),
//   * every shape the tcgen05 kernel does not cover (sub_dim not in {8,16}, k > 256, ...),
//   * the exact re-check reference the tensor-core kernel is validated against.
//
// Layout: one thread = one vector row, its sub-vector lives in registers (compile-time sub_dim)
// ; the subspace codebook (k x sub_dim f32, + per-centroid cosine norms) is staged in shared
// memory once per CTA and read by warp-wide broadcast (all lanes read the same centroid word:
// conflict-free).  Bound: FP32 issue, ~3*sub_dim+2 non-FMA ops per (row, centroid).
#include "common.cuh"
#include "distance.cuh"

namespace {

constexpr int AS_THREADS = 256;

template <int D>
struct RegAcc {
    float v[D > 0 ? D : 1];
    VQB_DEV float operator()(int i) const { return v[i]; }
};

VQB_DEV void store_code_any(void* codes, uint32_t code_bytes, size_t off, uint32_t v) {
    if (code_bytes == 1) static_cast<uint8_t*>(codes)[off] = (uint8_t)v;
    else if (code_bytes == 2) static_cast<uint16_t*>(codes)[off] = (uint16_t)v;
    else static_cast<uint32_t*>(codes)[off] = v;
}

// extra per-centroid words kept after the codebook chunk in smem for cosine
struct CosAux { float nb, sb, tail_ok; };

template <int MK, int D>
__global__ void __launch_bounds__(AS_THREADS)
k_assign_exact(const float* __restrict__ x, size_t n, int dim, int k, int sub_dim_rt,
               const float* __restrict__ codebooks, const int* __restrict__ sub_list, int k_chunk,
               void* __restrict__ codes, uint32_t code_bytes, size_t stride_row, size_t stride_sub,
               __half* __restrict__ recon, const int* __restrict__ n_sub_dev, const uint32_t* __restrict__ go) {
    extern __shared__ __align__(16) float smem[];
    // training loop gating (pq_train.cu): a speculatively enqueued iteration is a no-op unless the previous one allowed
    // it, and the list of still-active subspaces lives on the device
    if (go && *go == 0) return;
    if (n_sub_dev && (int)blockIdx.y >= *n_sub_dev) return;
    const int d = D > 0 ? D : sub_dim_rt;
    const int s = sub_list ? sub_list[blockIdx.y] : (int)blockIdx.y;
    const size_t row = (size_t)blockIdx.x * AS_THREADS + threadIdx.x;
    const bool live = row < n;
    const float* cb = codebooks + (size_t)s * k * d;
    float* cs = smem;                                                     // [k_chunk][d]
    CosAux* aux = reinterpret_cast<CosAux*>(smem + (size_t)k_chunk * d);  // [k_chunk] (cosine only)

    // this thread's sub-vector
    RegAcc<D> xr;
    const float* xg = x + (live ? row : 0) * (size_t)dim + (size_t)s * d;
    if (D > 0) {
        if (D % 4 == 0 && (dim % 4) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < D / 4; ++i) {
                float4 t = __ldg(reinterpret_cast<const float4*>(xg) + i);
                xr.v[4 * i] = t.x; xr.v[4 * i + 1] = t.y; xr.v[4 * i + 2] = t.z; xr.v[4 * i + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i) xr.v[i] = __ldg(xg + i);
        }
    }
    PtrAcc xp{xg};

    // per-row cosine terms (depend on x only)
    float na = 0.f, sa = 0.f;
    bool a_tail_ok = true;
    if (MK == MK_COSINE) {
        na = (D > 0) ? hsd_cosine_norm<D>(xr, d, a_tail_ok) : hsd_cosine_norm<0>(xp, d, a_tail_ok);
        sa = __fsqrt_rn(na);
    }

    uint32_t best = 0;
    float best_dist = 0.f;
    for (int k0 = 0; k0 < k; k0 += k_chunk) {
        const int kc = min(k_chunk, k - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kc * d; i += AS_THREADS) cs[i] = __ldg(cb + (size_t)k0 * d + i);
        __syncthreads();
        if (MK == MK_COSINE) {
            for (int j = threadIdx.x; j < kc; j += AS_THREADS) {
                bool tok;
                PtrAcc cp{cs + (size_t)j * d};
                float nb = hsd_cosine_norm<D>(cp, d, tok);
                aux[j].nb = nb; aux[j].sb = __fsqrt_rn(nb); aux[j].tail_ok = tok ? 1.f : 0.f;
            }
            __syncthreads();
        }
        if (live) {
            auto eval = [&](int j) -> float {
                PtrAcc cp{cs + (size_t)j * d};
                if (MK == MK_TRAIN) return (D > 0) ? dist2_seq<D>(xr, cp, d) : dist2_seq<0>(xp, cp, d);
                if (MK == MK_COSINE) {
                    if (d == 0) return 0.f;
                    float dot = (D > 0) ? hsd_cosine_dot<D>(xr, cp, d) : hsd_cosine_dot<0>(xp, cp, d);
                    bool ok = a_tail_ok && aux[j].tail_ok != 0.f;
                    float sim = 0.f;
                    if (ok) sim = hsd_cosine_from_sums(dot, na, aux[j].nb, sa, aux[j].sb, ok);
                    return ok ? __fsub_rn(1.0f, sim) : ((D > 0) ? rust_cos<D>(xr, cp, d) : rust_cos<0>(xp, cp, d));
                }
                return (D > 0) ? vq_distance<D>(MK, xr, cp, d) : vq_distance<0>(MK, xp, cp, d);
            };
            // pq.rs:183-191 / vector.rs:354-361: index 0 seeds the minimum (whatever its value), then strict '<'
            int j = 0;
            if (k0 == 0) { best_dist = eval(0); best = 0; j = 1; }
#pragma unroll 4
            for (; j < kc; ++j) {
                const float dist = eval(j);
                if (dist < best_dist) { best_dist = dist; best = (uint32_t)(k0 + j); }
            }
        }
    }
    if (!live) return;
    if (codes) store_code_any(codes, code_bytes, row * stride_row + (size_t)s * stride_sub, best);
    if (recon) {  // pq.rs:193-195: f16::from_f32 of the chosen centroid (round-to-nearest-even)
        const float* c = cb + (size_t)best * d;
        __half* r = recon + row * (size_t)dim + (size_t)s * d;
        for (int i = 0; i < d; ++i) r[i] = __float2half_rn(__ldg(c + i));
    }
}

// A handful of rows (the reference's call shape is ONE vector per quantize call, src/pq.rs:167): the kernel above would run
// one live thread per CTA through all k centroids.  Here a warp serves one (row, subspace) pair, its lanes take the centroids
// j = lane, lane + 32, ... with the same evaluation functions, and the warp keeps the reference's choice: index 0 seeds the
// minimum whatever its value, then strict '<' in ascending index order (= smallest distance, lowest index on ties).
template <int MK>
__global__ void __launch_bounds__(32)
k_assign_small(const float* __restrict__ x, int dim, int k, int d, const float* __restrict__ codebooks,
               void* __restrict__ codes, uint32_t code_bytes, size_t stride_row, size_t stride_sub, __half* __restrict__ recon) {
    const int lane = threadIdx.x;
    const size_t row = blockIdx.x;
    const int s = blockIdx.y;
    const float* cb = codebooks + (size_t)s * k * d;
    PtrAcc xp{x + row * (size_t)dim + (size_t)s * d};
    float na = 0.f, sa = 0.f;
    bool a_tail_ok = true;
    if (MK == MK_COSINE) { na = hsd_cosine_norm<0>(xp, d, a_tail_ok); sa = __fsqrt_rn(na); }
    float bd = __int_as_float(0x7f800000);
    uint32_t bj = 0xFFFFFFFFu;
    bool d0nan = false;
    for (int j = lane; j < k; j += 32) {
        PtrAcc cp{cb + (size_t)j * d};
        float dd;
        if (MK == MK_TRAIN) dd = dist2_seq<0>(xp, cp, d);
        else if (MK == MK_COSINE) {
            if (d == 0) dd = 0.f;
            else {
                bool tok;
                const float nb = hsd_cosine_norm<0>(cp, d, tok);
                const float dot = hsd_cosine_dot<0>(xp, cp, d);
                bool ok = a_tail_ok && tok;
                float sim = 0.f;
                if (ok) sim = hsd_cosine_from_sums(dot, na, nb, sa, __fsqrt_rn(nb), ok);
                dd = ok ? __fsub_rn(1.0f, sim) : rust_cos<0>(xp, cp, d);
            }
        } else dd = vq_distance<0>(MK, xp, cp, d);
        if (j == 0) d0nan = isnan(dd);
        if (isnan(dd)) dd = __int_as_float(0x7f800000);
        if (bj == 0xFFFFFFFFu || dd < bd) { bd = dd; bj = (uint32_t)j; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float od = __shfl_xor_sync(0xFFFFFFFFu, bd, off);
        const uint32_t oj = __shfl_xor_sync(0xFFFFFFFFu, bj, off);
        if (oj != 0xFFFFFFFFu && (bj == 0xFFFFFFFFu || od < bd || (od == bd && oj < bj))) { bd = od; bj = oj; }
    }
    d0nan = __shfl_sync(0xFFFFFFFFu, d0nan ? 1 : 0, 0) != 0;
    const uint32_t best = d0nan ? 0u : bj;   // vector.rs:354-361 / pq.rs:183-191: a NaN at index 0 is never replaced
    if (lane == 0 && codes) store_code_any(codes, code_bytes, row * stride_row + (size_t)s * stride_sub, best);
    if (recon) {  // pq.rs:193-195
        const float* c = cb + (size_t)best * d;
        __half* r = recon + row * (size_t)dim + (size_t)s * d;
        for (int i = lane; i < d; i += 32) r[i] = __float2half_rn(__ldg(c + i));
    }
}

template <int MK>
int launch_mk(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t k, size_t d, const float* cb,
              const int* sub_list, int n_sub, void* codes, uint32_t code_bytes, size_t stride_row,
              size_t stride_sub, __half* recon, const int* n_sub_dev, const uint32_t* go) {
    if (n <= 64 && !sub_list && !n_sub_dev && !go && (size_t)n_sub <= 65535) {   // a few rows: warp per (row, subspace)
        k_assign_small<MK><<<dim3((unsigned)n, (unsigned)n_sub), 32, 0, ctx->stream>>>(x, (int)dim, (int)k, (int)d, cb, codes, code_bytes,
                                                                                   stride_row, stride_sub, recon);
        VQB_LAUNCHED(ctx);
        return VQB_SUCCESS;
    }
    // centroid chunk that fits a 96 KB dynamic smem budget
    size_t per = d * sizeof(float) + (MK == MK_COSINE ? sizeof(CosAux) : 0);
    if (per == 0) per = 4;
    size_t budget = 96 * 1024;
    int k_chunk = (int)std::min<size_t>(k, std::max<size_t>(1, budget / per));
    size_t smem = (size_t)k_chunk * per + 16;
    dim3 grid(cdiv(n, AS_THREADS), (unsigned)n_sub);
    if (grid.y > 65535) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "too many subspaces (%d)", n_sub);
#define VQB_AS_CASE(DD)                                                                                    \
    case DD: {                                                                                             \
        auto kern = k_assign_exact<MK, DD>;                                                                \
        VQB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, AS_THREADS, smem, ctx->stream>>>(x, n, (int)dim, (int)k, (int)d, cb, sub_list, k_chunk, \
                                                     codes, code_bytes, stride_row, stride_sub, recon, n_sub_dev, go); \
        break;                                                                                             \
    }
    switch (d) {
        VQB_AS_CASE(2)
        VQB_AS_CASE(4)
        VQB_AS_CASE(8)
        VQB_AS_CASE(12)
        VQB_AS_CASE(16)
        VQB_AS_CASE(20)
        VQB_AS_CASE(24)
        VQB_AS_CASE(32)
        VQB_AS_CASE(40)
        VQB_AS_CASE(48)
        VQB_AS_CASE(64)
        default: {
            auto kern = k_assign_exact<MK, 0>;
            VQB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, AS_THREADS, smem, ctx->stream>>>(x, n, (int)dim, (int)k, (int)d, cb, sub_list, k_chunk,
                                                         codes, code_bytes, stride_row, stride_sub, recon, n_sub_dev, go);
        }
    }
#undef VQB_AS_CASE
    VQB_LAUNCHED(ctx);
    return VQB_SUCCESS;
}

}  // namespace

int vqb_pq_assign_exact_launch(vqb_ctx* ctx, int mk, const float* x, size_t n, size_t dim, size_t m, size_t k,
                               size_t d, const float* cb, const int* sub_list, int n_sub, void* codes,
                               uint32_t code_bytes, size_t stride_row, size_t stride_sub, __half* recon,
                               const int* n_sub_dev, const uint32_t* go) {
    (void)m;
    if (n == 0 || n_sub == 0) return VQB_SUCCESS;
    if (dim > (size_t)INT32_MAX || k > (size_t)INT32_MAX) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "shape too large");
    switch (mk) {
        case MK_SQEUCLID: return launch_mk<MK_SQEUCLID>(ctx, x, n, dim, k, d, cb, sub_list, n_sub, codes, code_bytes, stride_row, stride_sub, recon, n_sub_dev, go);
        case MK_EUCLID: return launch_mk<MK_EUCLID>(ctx, x, n, dim, k, d, cb, sub_list, n_sub, codes, code_bytes, stride_row, stride_sub, recon, n_sub_dev, go);
        case MK_MANHATTAN: return launch_mk<MK_MANHATTAN>(ctx, x, n, dim, k, d, cb, sub_list, n_sub, codes, code_bytes, stride_row, stride_sub, recon, n_sub_dev, go);
        case MK_COSINE: return launch_mk<MK_COSINE>(ctx, x, n, dim, k, d, cb, sub_list, n_sub, codes, code_bytes, stride_row, stride_sub, recon, n_sub_dev, go);
        case MK_TRAIN: return launch_mk<MK_TRAIN>(ctx, x, n, dim, k, d, cb, sub_list, n_sub, codes, code_bytes, stride_row, stride_sub, recon, n_sub_dev, go);
        case MK_CHEBYSHEV: return launch_mk<MK_CHEBYSHEV>(ctx, x, n, dim, k, d, cb, sub_list, n_sub, codes, code_bytes, stride_row, stride_sub, recon, n_sub_dev, go);
    }
    return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "unknown metric kind %d", mk);
}

// =====================================================================================================================
// Manhattan assignment for sub_dim 8, k <= 256 (BASELINE config 5b): tiled variant of the kernel above.
//
// Same arithmetic, operation for operation (n < 16: hsdlib's manhattan kernel is its scalar loop `sum += fabsf(a - b)`,
// manhattan.c:132-163, the same sequential sum as the Rust fallback distance.rs:93-95), strict '<' / first minimum
// (pq.rs:183-191).  What changes is the memory side:
//   * a CTA owns 4 consecutive subspaces = one 128-byte line of every row and walks row tiles of 128 rows; the tile is
//     staged with coalesced 16-byte loads (a warp reads four full lines per instruction) instead of 32-byte pieces at a
//     3 KB stride (r01 profile: 15.2 GB of DRAM reads for 3.07 GB of input);
//   * the four codebooks (32 KB) stay in shared memory for the CTA's life; a warp works on ONE subspace, so every centroid
//     read is a broadcast; each thread carries two rows, so a centroid is read once per two distances;
//   * codes leave through shared memory as one 4-byte store per row (encode layout) or 32-byte runs (training layout)
//     instead of one-byte stores at a stride of m (3.05 GB written for 96 MB).
namespace {

constexpr int L1_D = 8, L1_G = 4, L1_ROWS = 128, L1_THREADS = 256;
constexpr int L1_CB_BYTES = L1_G * 256 * L1_D * 4, L1_XT_BYTES = L1_ROWS * (L1_G * L1_D + 4) * 4;
constexpr int L1_SMEM = L1_CB_BYTES + L1_XT_BYTES + L1_ROWS * 4;

__global__ void __launch_bounds__(L1_THREADS)
k_assign_l1_tiles(const float* __restrict__ x, size_t n, int dim, int m, int k, const float* __restrict__ codebooks,
                  void* __restrict__ codes, uint32_t code_bytes, size_t stride_row, size_t stride_sub,
                  __half* __restrict__ recon, int n_groups, int parts, int num_tiles) {
    extern __shared__ __align__(16) uint8_t l1_smem[];
    float (*cb)[256 * L1_D] = reinterpret_cast<float (*)[256 * L1_D]>(l1_smem);                       // [4][2048]: 32 KB
    float (*xt)[L1_G * L1_D + 4] = reinterpret_cast<float (*)[L1_G * L1_D + 4]>(l1_smem + L1_CB_BYTES); // rows padded to 144 B: conflict-free 16-byte reads
    uint32_t* ct = reinterpret_cast<uint32_t*>(l1_smem + L1_CB_BYTES + L1_XT_BYTES);                    // 4 codes (bytes) per row
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = blockIdx.x % n_groups, part = blockIdx.x / n_groups;
    const int s0 = grp * L1_G;
    const int g_cnt = min(L1_G, m - s0);
    for (int i = 0; i < g_cnt; ++i)
        for (int t = tid; t < k * L1_D / 4; t += L1_THREADS)
            reinterpret_cast<float4*>(cb[i])[t] = __ldg(reinterpret_cast<const float4*>(codebooks + (size_t)(s0 + i) * k * L1_D) + t);
    const int si = warp & 3;              // subspace of this warp
    const int rbase = (warp >> 2) * 64;   // its 64 rows: lane -> rows rbase + lane, rbase + 32 + lane
    const bool sub_ok = si < g_cnt;
    const bool vec_ok = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (s0 * L1_D + L1_G * L1_D <= dim);
    for (int it = part; it < num_tiles; it += parts) {
        const size_t row0 = (size_t)it * L1_ROWS;
        __syncthreads();   // previous tile's readers are done (and the codebooks are staged)
        // ---- stage the tile: 128 rows x 8 float4; consecutive threads take consecutive 16-byte pieces of a row
        for (int t = tid; t < L1_ROWS * 8; t += L1_THREADS) {
            const int r = t >> 3, c4 = t & 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < n) {
                const float* src = x + (row0 + r) * (size_t)dim + (size_t)s0 * L1_D + c4 * 4;
                if (vec_ok) v = __ldg(reinterpret_cast<const float4*>(src));
                else {
                    const int col = s0 * L1_D + c4 * 4;
                    v.x = col < dim ? __ldg(src) : 0.f; v.y = col + 1 < dim ? __ldg(src + 1) : 0.f;
                    v.z = col + 2 < dim ? __ldg(src + 2) : 0.f; v.w = col + 3 < dim ? __ldg(src + 3) : 0.f;
                }
            }
            *reinterpret_cast<float4*>(&xt[r][c4 * 4]) = v;
        }
        if (tid < L1_ROWS) ct[tid] = 0;
        __syncthreads();
        if (sub_ok) {
            float xa[L1_D], xb[L1_D];
            {
                const float4 a0 = *reinterpret_cast<const float4*>(&xt[rbase + lane][si * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&xt[rbase + lane][si * 8 + 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&xt[rbase + 32 + lane][si * 8]);
                const float4 b1 = *reinterpret_cast<const float4*>(&xt[rbase + 32 + lane][si * 8 + 4]);
                xa[0] = a0.x; xa[1] = a0.y; xa[2] = a0.z; xa[3] = a0.w; xa[4] = a1.x; xa[5] = a1.y; xa[6] = a1.z; xa[7] = a1.w;
                xb[0] = b0.x; xb[1] = b0.y; xb[2] = b0.z; xb[3] = b0.w; xb[4] = b1.x; xb[5] = b1.y; xb[6] = b1.z; xb[7] = b1.w;
            }
            const float* c = cb[si];
            // index 0 seeds the minimum whatever its value (a NaN there is never replaced), then strict '<'
            float best_a, best_b;
            uint32_t ja = 0, jb = 0;
            {
                float da = 0.f, db = 0.f;
#pragma unroll
                for (int i = 0; i < L1_D; ++i) {
                    da = __fadd_rn(da, fabsf(__fsub_rn(xa[i], c[i])));
                    db = __fadd_rn(db, fabsf(__fsub_rn(xb[i], c[i])));
                }
                best_a = da; best_b = db;
            }
#pragma unroll 4
            for (int j = 1; j < k; ++j) {
                const float4 c0 = *reinterpret_cast<const float4*>(c + j * L1_D);
                const float4 c1 = *reinterpret_cast<const float4*>(c + j * L1_D + 4);
                const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                float da = 0.f, db = 0.f;
#pragma unroll
                for (int i = 0; i < L1_D; ++i) {
                    da = __fadd_rn(da, fabsf(__fsub_rn(xa[i], cc[i])));
                    db = __fadd_rn(db, fabsf(__fsub_rn(xb[i], cc[i])));
                }
                if (da < best_a) { best_a = da; ja = (uint32_t)j; }
                if (db < best_b) { best_b = db; jb = (uint32_t)j; }
            }
            // ---- codes: through shared memory for the row-major layout, direct 32-byte runs for the subspace-major one
            const size_t ra = row0 + rbase + lane, rb = ra + 32;
            const size_t s = (size_t)(s0 + si);
            if (code_bytes == 1 && stride_sub == 1 && codes) {
                atomicOr(&ct[rbase + lane], ja << (8 * si));
                atomicOr(&ct[rbase + 32 + lane], jb << (8 * si));
            } else if (codes) {
                if (ra < n) store_code_any(codes, code_bytes, ra * stride_row + s * stride_sub, ja);
                if (rb < n) store_code_any(codes, code_bytes, rb * stride_row + s * stride_sub, jb);
            }
            if (recon) {  // pq.rs:193-195: f16::from_f32 of the chosen centroid
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const size_t rr = h ? rb : ra;
                    if (rr >= n) continue;
                    const float* cp = c + (h ? jb : ja) * L1_D;
                    __half2 hh[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) hh[q] = __floats2half2_rn(cp[2 * q], cp[2 * q + 1]);
                    __half* dst = recon + rr * (size_t)dim + s * L1_D;
                    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(hh);
                    else for (int q = 0; q < 4; ++q) { dst[2 * q] = __low2half(hh[q]); dst[2 * q + 1] = __high2half(hh[q]); }
                }
            }
        }
        if (code_bytes == 1 && stride_sub == 1 && codes) {
            __syncthreads();
            if (tid < L1_ROWS && row0 + tid < n) {
                uint8_t* dst = static_cast<uint8_t*>(codes) + (row0 + tid) * stride_row + s0;
                const uint32_t w = ct[tid];
                if (g_cnt == 4 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) *reinterpret_cast<uint32_t*>(dst) = w;
                else for (int i = 0; i < g_cnt; ++i) dst[i] = (uint8_t)(w >> (8 * i));
            }
        }
    }
}

}  // namespace

bool vqb_l1_tiles_supported(int mk, size_t d, size_t k) { return mk == MK_MANHATTAN && d == L1_D && k >= 1 && k <= 256; }

int vqb_assign_l1_tiles_launch(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k, const float* cb,
                               void* codes, uint32_t code_bytes, size_t stride_row, size_t stride_sub, __half* recon) {
    if (n == 0) return VQB_SUCCESS;
    const int n_groups = (int)((m + L1_G - 1) / L1_G);
    const int num_tiles = (int)cdiv(n, L1_ROWS);
    // 4 CTAs of 256 threads per SM (~52 KB of shared memory, 47 registers each): the kernel is FP32-issue bound
    const int parts = std::max(1, std::min(num_tiles, 4 * ctx->sm_count / n_groups));
    VQB_CUDA(ctx, cudaFuncSetAttribute(k_assign_l1_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, L1_SMEM));
    k_assign_l1_tiles<<<n_groups * parts, L1_THREADS, L1_SMEM, ctx->stream>>>(x, n, (int)dim, (int)m, (int)k, cb, codes, code_bytes,
                                                                        stride_row, stride_sub, recon, n_groups, parts, num_tiles);
    VQB_LAUNCHED(ctx);
    return VQB_SUCCESS;
}
