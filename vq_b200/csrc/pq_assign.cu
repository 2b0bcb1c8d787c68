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

template <int MK>
int launch_mk(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t k, size_t d, const float* cb,
              const int* sub_list, int n_sub, void* codes, uint32_t code_bytes, size_t stride_row,
              size_t stride_sub, __half* recon, const int* n_sub_dev, const uint32_t* go) {
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
        VQB_AS_CASE(16)
        VQB_AS_CASE(32)
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
    }
    return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "unknown metric kind %d", mk);
}
