// ub_sm100.cu -- micro-benchmarks that size the GEMM-form assignment kernel on B200:
//   (A) issue rates of the epilogue instructions (FMNMX3, LOP3, IMAD, FFMA.SAT imm, FSET ...)
//   (B) tcgen05.ld (TMEM -> registers) throughput with 4 and 8 warps
//   (C) one tcgen05.mma kind::tf32 128x256x8 with no-swizzle K-major smem descriptors:
//       numerical check against the host + cycles per MMA
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ub_sm100 ub_sm100.cu
// Not part of the product; results are recorded in profiles/.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------- (A) issue rates
enum { OP_FMNMX = 0, OP_FMNMX3, OP_LOP3, OP_IMAD, OP_FFMA3, OP_FFMA_IMM_SAT, OP_FSET, OP_VIMNMX3, OP_MIX_A, OP_MIX_B, OP_MIX_C, OP_COUNT };
static const char* OP_NAME[] = {"FMNMX (2-in)", "FMNMX3", "LOP3", "IMAD", "FFMA 3-reg", "FFMA.SAT imm", "FSET", "VIMNMX3.U32",
                                "mix: 1 FMNMX3 + 4 FFMA.SAT-imm (per 2 elems: rigorous elem-level)",
                                "mix: 2 LOP3 + 1 FMNMX3 (keyed top-1 per 2 elems)",
                                "mix: 2 IMAD + 1 VIMNMX3 (int keyed top-1 per 2 elems)"};

template <int OP>
__global__ void __launch_bounds__(1024) k_issue(float* out, int iters, float fa, float fb, unsigned ua, unsigned ub, long long* cyc) {
    float f[16];
    unsigned u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { f[i] = fa + i + threadIdx.x; u[i] = ua + i * 77 + threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == OP_FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb));
            if (OP == OP_FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(f[(i + 1) & 15]));
            if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(ub), "r"(0x55u + i));
            if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, 256, %1;" : "+r"(u[i]) : "r"(ub));
            if (OP == OP_FFMA3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fa));
            if (OP == OP_FFMA_IMM_SAT) asm volatile("fma.rn.sat.f32 %0, %0, 0fF149F2CA, %1;" : "+f"(f[i]) : "f"(fb));
            if (OP == OP_FSET) asm volatile("set.le.f32.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb));
            if (OP == OP_VIMNMX3) asm volatile("min.u32 %0, %0, %1; min.u32 %0, %0, %2;" : "+r"(u[i]) : "r"(ub), "r"(u[(i + 1) & 15]));
        }
        if (OP == OP_MIX_A) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float ind0, ind1;
                asm volatile("fma.rn.sat.f32 %0, %1, 0fF149F2CA, %2;" : "=f"(ind0) : "f"(f[i]), "f"(fb));
                asm volatile("fma.rn.sat.f32 %0, %1, 0fF149F2CA, %2;" : "=f"(ind1) : "f"(f[i + 1]), "f"(fb));
                asm volatile("fma.rn.f32 %0, %1, 0f44810000, %0;" : "+f"(f[(i + 2) & 15]) : "f"(ind0));
                asm volatile("fma.rn.f32 %0, %1, 0f44812000, %0;" : "+f"(f[(i + 3) & 15]) : "f"(ind1));
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(fa) : "f"(f[i]), "f"(f[i + 1]));
            }
        }
        if (OP == OP_MIX_B) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned k0, k1;
                asm volatile("lop3.b32 %0, %1, 0xFFFFFF00, %2, 0xEA;" : "=r"(k0) : "r"(u[i]), "r"(i));
                asm volatile("lop3.b32 %0, %1, 0xFFFFFF00, %2, 0xEA;" : "=r"(k1) : "r"(u[i + 1]), "r"(i + 1));
                asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(fa) : "f"(__uint_as_float(k0)), "f"(__uint_as_float(k1)));
                u[i] += ua;
            }
        }
        if (OP == OP_MIX_C) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned k0, k1;
                asm volatile("mad.lo.u32 %0, %1, 256, %2;" : "=r"(k0) : "r"(u[i]), "r"(i));
                asm volatile("mad.lo.u32 %0, %1, 256, %2;" : "=r"(k1) : "r"(u[i + 1]), "r"(i + 1));
                asm volatile("min.u32 %0, %0, %1; min.u32 %0, %0, %2;" : "+r"(ua) : "r"(k0), "r"(k1));
                u[i] += ub;
            }
        }
    }
    long long t1 = clock64();
    float acc = fa;
    unsigned uacc = ua;
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc += f[i]; uacc ^= u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(uacc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run_issue(int warps, float* dout, long long* dcyc) {
    const int iters = 2000;
    k_issue<OP><<<1, warps * 32>>>(dout, iters, 1.5f, 2.5f, 12345u, 777u, dcyc);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
    double instr = 16.0;  // per-thread instructions per iteration (issue slots)
    if (OP == OP_VIMNMX3) instr = 16.0;  // fused by ptxas into one VIMNMX3 (check SASS)
    if (OP == OP_MIX_A) instr = 8 * 5;
    if (OP == OP_MIX_B) instr = 8 * 4;  // + the u[i] += ua IADD
    if (OP == OP_MIX_C) instr = 8 * 4;
    double lane_ops = instr * iters * warps * 32;
    printf("  %-72s warps=%2d cycles=%9lld  lane-instr/cyc/SM=%7.1f  (warp-instr/cyc/SM=%5.2f)\n", OP_NAME[OP], warps, c,
           lane_ops / (double)c, lane_ops / 32.0 / (double)c);
}

// ---------------------------------------------------------------- PTX helpers for (B), (C)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// no-swizzle K-major descriptor: core matrix = 8 rows x 16 B contiguous; LBO = byte step between the two
// 16-byte K chunks of one MMA, SBO = byte step between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // layout_type 0 = SWIZZLE_NONE
}

// ---------------------------------------------------------------- (B) tcgen05.ld throughput
__global__ void __launch_bounds__(256) k_tmem_ld(int iters, int passes_cols, float* out, long long* cyc) {
    __shared__ uint32_t tbase;
    if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < passes_cols; c += 32) {
            uint32_t v[32];
            tmem_ld32(base + lane_base + (uint32_t)((c + (warp >> 2) * 256) & 511), v);
            tmem_ld_wait();
            float m = __uint_as_float(v[0]);
#pragma unroll
            for (int i = 1; i < 32; i += 2) m = fminf(fminf(m, __uint_as_float(v[i])), __uint_as_float(v[(i + 1) & 31]));
            acc += m;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

// ---------------------------------------------------------------- (B2) realistic epilogue on synthetic TMEM
// per chunk of 32 columns: 11 group minima (triples), chunk minimum, 11 saturated indicators, 11 FFMA
__device__ __forceinline__ float fsat_ind(float g, float thH) {  // sat((th - g) * 2^100)
    float r;
    asm("fma.rn.sat.f32 %0, %1, 0fF1800000, %2;" : "=f"(r) : "f"(g), "f"(thH));
    return r;
}
template <bool PIPE>
__global__ void __launch_bounds__(256) k_epi(int iters, float margin, float* out, long long* cyc) {
    __shared__ uint32_t tbase;
    if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col0 = (warp >> 2) * 256;
    // fill TMEM with something finite
    {
        uint32_t v[32];
        for (int c = 0; c < 256; c += 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(1.0f + 0.001f * ((threadIdx.x * 37 + (c + i) * 101) % 977));
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                ::"r"(base + lane_base + col0 + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
                "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    float total = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float cm[8], ac[8];
        uint32_t va[32], vb[32];
        if (PIPE) tmem_ld32(base + lane_base + col0, va);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t(&v)[32] = (PIPE && (c & 1)) ? vb : va;
            uint32_t(&nx)[32] = (PIPE && (c & 1)) ? va : vb;
            if (PIPE) {
                tmem_ld_wait();
                if (c < 7) tmem_ld32(base + lane_base + col0 + (c + 1) * 32, nx);
            } else {
                tmem_ld32(base + lane_base + col0 + c * 32, v);
                tmem_ld_wait();
            }
            float g[11];
#pragma unroll
            for (int t = 0; t < 10; ++t)
                g[t] = fminf(fminf(__uint_as_float(v[3 * t]), __uint_as_float(v[3 * t + 1])), __uint_as_float(v[3 * t + 2]));
            g[10] = fminf(__uint_as_float(v[30]), __uint_as_float(v[31]));
            float m0 = fminf(fminf(g[0], g[1]), g[2]), m1 = fminf(fminf(g[3], g[4]), g[5]), m2 = fminf(fminf(g[6], g[7]), g[8]);
            float m3 = fminf(g[9], g[10]);
            float m = fminf(fminf(fminf(m0, m1), m2), m3);
            float thH = fmaf(m, 1.2676506e30f, margin * 1.2676506e30f);  // (m + margin) * 2^100
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int t = 0; t < 11; ++t) {
                float w = (float)(32 + t), ind = fsat_ind(g[t], thH);
                if ((t & 3) == 0) a0 = fmaf(ind, w, a0);
                if ((t & 3) == 1) a1 = fmaf(ind, w, a1);
                if ((t & 3) == 2) a2 = fmaf(ind, w, a2);
                if ((t & 3) == 3) a3 = fmaf(ind, w, a3);
            }
            cm[c] = m; ac[c] = (a0 + a1) + (a2 + a3);
        }
        float mm = cm[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) mm = fminf(mm, cm[c]);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) s += (cm[c] < mm + margin) ? ac[c] : 0.f;
        total += s;
    }
    long long t1 = clock64();
    out[threadIdx.x] = total;
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

// pure TMEM read throughput: loads only, two in flight
__global__ void __launch_bounds__(256) k_tmem_bw(int iters, float* out, long long* cyc) {
    __shared__ uint32_t tbase;
    if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col0 = (warp >> 2) * 256;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t va[32], vb[32];
        tmem_ld32(base + lane_base + col0 + ((it * 64) & 255), va);
        tmem_ld32(base + lane_base + col0 + ((it * 64 + 32) & 255), vb);
        tmem_ld_wait();
        acc ^= va[0] ^ vb[31] ^ va[17];
    }
    long long t1 = clock64();
    out[threadIdx.x] = __uint_as_float(acc);
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

// ---------------------------------------------------------------- (C) MMA check
// A: 128 x K (K = 8*ksteps), B: 256 x K, both K-major, no-swizzle layout [row/8][chunk][row%8][16B]
__global__ void __launch_bounds__(128) k_mma(const float* __restrict__ A, const float* __restrict__ B, int ksteps, int reps,
                                             float* __restrict__ D, long long* cyc) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tbase;
    __shared__ __align__(8) uint64_t bar;
    const int K = ksteps * 8, nchunk = ksteps * 2;
    uint8_t* sA = smem;                          // 16 groups * nchunk * 128 B
    uint8_t* sB = smem + 16 * nchunk * 128;      // 32 groups * nchunk * 128 B
    for (int i = threadIdx.x; i < 128 * nchunk; i += blockDim.x) {
        int r = i / nchunk, c = i % nchunk;
        float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * K + c * 4);
        *reinterpret_cast<float4*>(sA + (r >> 3) * (nchunk * 128) + c * 128 + (r & 7) * 16) = v;
    }
    for (int i = threadIdx.x; i < 256 * nchunk; i += blockDim.x) {
        int r = i / nchunk, c = i % nchunk;
        float4 v = *reinterpret_cast<const float4*>(B + (size_t)r * K + c * 4);
        *reinterpret_cast<float4*>(sB + (r >> 3) * (nchunk * 128) + c * 128 + (r & 7) * 16) = v;
    }
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    if (threadIdx.x < 32) tmem_alloc(&tbase, 256);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
            for (int ks = 0; ks < ksteps; ++ks) {
                uint64_t ad = make_desc(smem_u32(sA) + ks * 256, 128, nchunk * 128);
                uint64_t bd = make_desc(smem_u32(sB) + ks * 256, 128, nchunk * 128);
                umma_tf32(base, ad, bd, idesc, ks > 0 ? 1u : 0u);
            }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (threadIdx.x == 0) { t1 = clock64(); cyc[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = warp * 32 + lane;
    for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        tmem_ld32(base + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) D[(size_t)row * 256 + c + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(base, 256);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device: %s  SMs=%d  cc=%d.%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.major, p.minor, p.clockRate);
    float* dout; long long* dcyc;
    CK(cudaMalloc(&dout, 1 << 20)); CK(cudaMalloc(&dcyc, 4096));

    printf("(A) issue rates, one CTA on one SM\n");
    for (int warps : {4, 8, 16, 32}) {
        if (getenv("UB_SKIP_A")) break;
        run_issue<OP_FMNMX>(warps, dout, dcyc); run_issue<OP_FMNMX3>(warps, dout, dcyc); run_issue<OP_LOP3>(warps, dout, dcyc);
        run_issue<OP_IMAD>(warps, dout, dcyc); run_issue<OP_FFMA3>(warps, dout, dcyc); run_issue<OP_FFMA_IMM_SAT>(warps, dout, dcyc);
        run_issue<OP_FSET>(warps, dout, dcyc); run_issue<OP_VIMNMX3>(warps, dout, dcyc);
        run_issue<OP_MIX_A>(warps, dout, dcyc); run_issue<OP_MIX_B>(warps, dout, dcyc); run_issue<OP_MIX_C>(warps, dout, dcyc);
    }

    printf("(B) tcgen05.ld 32x32b.x32 (+ 16 FMNMX3 per load), 256 columns per pass\n");
    for (int warps : {4, 8}) {
        const int iters = 500;
        k_tmem_ld<<<1, warps * 32>>>(iters, 256, dout, dcyc);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
        printf("  warps=%d: %lld cycles for %d passes -> %.1f cycles per 128x256 fp32 pass per warp-quarter, %.1f B/cyc/SM\n", warps, c,
               iters, (double)c / iters, (double)iters * warps * 32 * 256 * 4 / (double)c);
    }


    printf("(B1) pure tcgen05.ld throughput (2 x32 loads in flight per warp)\n");
    for (int warps : {4, 8}) {
        const int iters = 2000;
        k_tmem_bw<<<1, warps * 32>>>(iters, dout, dcyc);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
        printf("  warps=%d: %lld cycles, %.1f B/cyc/SM, %.1f cycles per x32 load per warp\n", warps, c,
               (double)iters * 2 * warps * 32 * 32 * 4 / (double)c, (double)c / (iters * 2));
    }
    printf("(B2) realistic epilogue (triples + indicators), per 128x256 accumulator\n");
    for (int warps : {4, 8}) {
        const int iters = 500;
        k_epi<false><<<1, warps * 32>>>(iters, 1e-3f, dout, dcyc);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
        printf("  serial   warps=%d: %.1f cycles per accumulator per warpgroup (effective %.1f per accumulator)\n", warps, (double)c / iters,
               (double)c / iters / (warps / 4));
        k_epi<true><<<1, warps * 32>>>(iters, 1e-3f, dout, dcyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
        printf("  pipelined warps=%d: %.1f cycles per accumulator per warpgroup (effective %.1f per accumulator)\n", warps, (double)c / iters,
               (double)c / iters / (warps / 4));
    }
    printf("(C) tcgen05.mma kind::tf32 M=128 N=256, no-swizzle K-major descriptors\n");
    for (int ksteps : {1, 4}) {
        const int K = ksteps * 8;
        std::vector<float> hA(128 * K), hB(256 * K), hD(128 * 256);
        srand(7);
        for (auto& v : hA) v = (float)rand() / RAND_MAX * 2.f - 1.f;
        for (auto& v : hB) v = (float)rand() / RAND_MAX * 2.f - 1.f;
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
        CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
        size_t smem = (16 + 32) * ksteps * 2 * 128;
        CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int reps : {1, 64}) {
            k_mma<<<1, 128, smem>>>(dA, dB, ksteps, reps, dD, dcyc);
            CK(cudaDeviceSynchronize());
            long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
            double max_err_trunc = 0, max_err_full = 0;
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 256; ++j) {
                    double st = 0, sf = 0;
                    for (int k = 0; k < K; ++k) {
                        st += (double)tf32_trunc(hA[i * K + k]) * tf32_trunc(hB[j * K + k]);
                        sf += (double)hA[i * K + k] * hB[j * K + k];
                    }
                    max_err_trunc = fmax(max_err_trunc, fabs(st - hD[i * 256 + j]));
                    max_err_full = fmax(max_err_full, fabs(sf - hD[i * 256 + j]));
                }
            printf("  K=%2d reps=%2d: cycles=%lld (%.1f per MMA)  max|D - tf32-truncated ref|=%.3e  max|D - f64 ref|=%.3e  D[0]=%f D[last]=%f\n", K,
                   reps, c, (double)c / (reps * ksteps), max_err_trunc, max_err_full, hD[0], hD[128 * 256 - 1]);
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return 0;
}
