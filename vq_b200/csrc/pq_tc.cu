// pq_tc.cu -- GEMM-form nearest-centroid assignment on the 5th-generation tensor cores.
//
// Replaces, for sub_dim == 8 and k <= 256, the per-pair distance loops of
//   find_nearest_centroid      src/core/vector.rs:352-363   (training, squared L2)
//   ProductQuantizer::quantize src/pq.rs:177-196            (encode: squared L2 / L2 / cosine)
// by a dense contraction whose result is accepted only where it provably equals the reference's:
//
//   scores  S[row, j] = ||c_j||^2 - 2 x.c_j  (L2 kinds)   or   - x.c_j/||c_j||  (cosine)
//           = one tcgen05.mma kind::tf32 M=128 N=256 chain per (128-row tile, subspace), K = 4 x 8:
//             x_hi.c_hi + x_hi.c_lo + x_lo.c_hi (3xTF32 split) + [1,1,1].[n1,n2,n3] (the squared norm as
//             three tf32 pieces); accumulators in TMEM (2 x 256 columns).  x_hi is not materialised: the
//             tensor core reads the fp32 TMA tile itself (SWIZZLE_128B descriptor) and truncates to tf32
//             in hardware; only x_lo = x - trunc_tf32(x) (exact) is written by the splitter.
//   scan    one thread per row (= TMEM lane) reads its 256 scores with tcgen05.ld (x16, double
//           buffered), keeps the minima of the 64 groups of four columns (FMNMX3 + FMNMX), then -- once
//           per row -- the row minimum and, with saturating FMAs and odd weights, the number and
//           position of groups within a rigorous error margin M of it.  Exactly one such group: its
//           index goes to the resolve stage; otherwise the row is "ambiguous".
//   resolve re-scores the four candidates of that group in fp32 (conflict-free lane-rotated gathers
//           of the raw codebook in shared memory) and accepts the best one only if it reproduces the
//           row minimum (within M/2) and beats the runner-up by more than M.  Anything else -- near-ties
//           inside M, NaN/Inf, degenerate norms, unsafe codebooks -- is decided by the warp re-scanning
//           all k centroids of that row with the reference's own formula, operation for operation
//           (distance.cuh), and its strict-'<' / lowest-index rule.  M bounds the tensor-core error plus
//           the reference's own rounding, so an accepted code is the reference's code.
//
// Decomposition (B-stationary): a CTA owns 4 consecutive subspaces (32 floats = one 128-byte line
// per row); their prepared codebooks stay in shared memory for the CTA's life and the CTA streams
// 128-row tiles of its column slab by TMA (SWIZZLE_128B box 32 x 128).  X is read once per pass.
//
// Warp roles (768 threads = 6 warpgroups, registers re-balanced with setmaxnreg):
//   WG0  w0 TMA producer | w1 MMA issuer + TMEM owner | w2-3 idle
//   WG1  splitter, once per row tile: the x_lo operand tile of all four subspaces (same SWIZZLE_128B layout as
//        the raw tile, so the tensor core reads it through the same descriptor form) + the rows' error margins
//   two scan warpgroups: one per TMEM accumulator; thread = row = TMEM lane; pure register work
//   two resolve warpgroups: fp32 re-score of the candidate group (or the full reference re-scan), code /
//        f16 reconstruction stores
// The per-unit chain split -> MMA -> scan -> resolve is a software pipeline over shared-memory rings;
// all hand-offs are mbarriers; tcgen05.commit releases smem / signals TMEM.  Every wait is a
// mbarrier.try_wait with a suspend-time hint: waiting warps sleep in hardware and take no issue slots.
#include "common.cuh"
#include "distance.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>

namespace {

#ifndef VQB_SPIN_FULL
#define VQB_SPIN_FULL 32
#endif
#ifndef VQB_SPIN_EMPTY
#define VQB_SPIN_EMPTY 32
#endif
constexpr int TC_N = 256;          // MMA N = centroid slots per subspace
constexpr int TC_ROWS = 128;       // MMA M = rows per tile = TMEM lanes
constexpr int RAW_STAGES = 3;      // TMA -> everyone: raw fp32 row tiles (128 rows x 128 B, SWIZZLE_128B)
constexpr int AT_STAGES = 2;       // splitter -> MMA: x_lo tiles, same shape and swizzle as the raw tile
constexpr int MG_STAGES = 2;       // splitter -> scan: margins of one row tile, [G][128 rows]
constexpr int RES_STAGES = 4;      // scan -> resolve: per-row candidate group of one unit
constexpr int TC_THREADS = 640;    // 5 warpgroups: control | 2 helpers (split + resolve) | 2 scan (one per TMEM accumulator)
// setmaxnreg budget: the pool is what the CTA was launched with (640 threads x 96 registers = 61440), so
// 128*24 + 256*72 + 256*152 = 60416 must not exceed it or the last setmaxnreg.inc never returns
constexpr int REGS_LAUNCH = 96, REGS_CTRL = 24, REGS_HELP = 72, REGS_SCAN = 152;
static_assert(128 * REGS_CTRL + 256 * REGS_HELP + 256 * REGS_SCAN <= TC_THREADS * REGS_LAUNCH,
              "setmaxnreg budget exceeds the registers the CTA owns");

constexpr uint32_t RAW_BYTES = TC_ROWS * 128;           // 16 KB per stage
constexpr uint32_t BP_PAD = 128;                        // zeros behind the last image (read by the last norm chunk pair)
constexpr uint32_t AUX_BYTES = TC_N * 8;                // 2 KB (nb, sb) per centroid (cosine, exact path)
constexpr uint32_t RINV_BYTES = TC_N * 4;               // 1 KB per centroid: -1/||c|| (cosine) or ||c||^2 (L2 kinds), fp32 re-score
constexpr uint32_t ONES_BYTES = 16 * 2 * 128;           // 4 KB: [16 row groups][2 k-chunks][8 rows][16 B]
constexpr uint32_t RES_BYTES = TC_ROWS * 8;             // float2 {row minimum, M | group index} per row
constexpr uint32_t RES_AMBIGUOUS = 0x80000000u;         // sign bit of the margin word

// Everything that depends on the sub-vector length D (8, 16, 24 or 32 floats).  A CTA owns the G subspaces that share one
// 128-byte line of each row (D = 24: one subspace, the last 32 bytes of the line belong to its neighbour and are ignored);
// a unit's scores are KC = D / 8 chained K = 8 MMAs per term.
template <int D>
struct Cfg {
    static_assert(D == 8 || D == 16 || D == 24 || D == 32, "sub_dim handled by the tensor kernel");
    static constexpr int G = (D == 24) ? 1 : 32 / D;          // subspaces per CTA
    static constexpr int KC = D / 8;                          // K = 8 MMAs per operand term
    static constexpr int NCH = D / 4;                         // 16-byte chunks per sub-vector / centroid
    // B image of one subspace: [32 row groups][2 NCH + 1 k-chunks: hi.. lo.. norm][8 rows][16 B]; the norm MMA's second
    // K chunk is the next row group's first hi chunk (finite) times the ones tile's zeros
    static constexpr uint32_t BP_CHUNKS = 2 * NCH + 1;
    static constexpr uint32_t BP_SBO = BP_CHUNKS * 128;       // bytes between 8-row groups
    static constexpr uint32_t BP_BYTES = 32 * BP_SBO;
    static constexpr uint32_t CB_BYTES = TC_N * D * 4;        // raw f32 codebook, row-major
    static constexpr uint32_t MG_BYTES = G * TC_ROWS * 8;     // float2 {H, M} per (subspace, row)
    static constexpr uint32_t PREP_BYTES = BP_BYTES + CB_BYTES + AUX_BYTES + RINV_BYTES;  // per-subspace prepared image in HBM
    static constexpr uint32_t OFF_RAW = 0;
    static constexpr uint32_t OFF_AT = OFF_RAW + RAW_STAGES * RAW_BYTES;
    static constexpr uint32_t OFF_BP = OFF_AT + AT_STAGES * RAW_BYTES;
    static constexpr uint32_t OFF_CB = OFF_BP + G * BP_BYTES + BP_PAD;
    static constexpr uint32_t OFF_AUX = OFF_CB + G * CB_BYTES;
    static constexpr uint32_t OFF_RINV = OFF_AUX + G * AUX_BYTES;
    static constexpr uint32_t OFF_ONES = OFF_RINV + G * RINV_BYTES;
    static constexpr uint32_t OFF_MG = OFF_ONES + ONES_BYTES;
    static constexpr uint32_t OFF_RES = OFF_MG + MG_STAGES * MG_BYTES;
    static constexpr uint32_t OFF_SINFO = OFF_RES + RES_STAGES * RES_BYTES;  // G x {sqrt(cmax2), unsafe}
    static constexpr uint32_t OFF_BAR = OFF_SINFO + 64;
    static constexpr uint32_t SMEM_BYTES = OFF_BAR + 512 + 1024;   // barriers + slack for the 1024-byte alignment
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB opt-in shared memory of sm_100");
    static_assert(OFF_AT % 1024 == 0 && OFF_BP % 128 == 0 && OFF_CB % 128 == 0 && OFF_ONES % 128 == 0, "operand tiles must be aligned");
};
// the prepared images of all sub_dims share one size formula on the host side
constexpr size_t prep_bytes_of(size_t d) { return 32 * ((2 * (d / 4) + 1) * 128) + TC_N * d * 4 + AUX_BYTES + RINV_BYTES; }

template <int D>
struct RegArr {
    float v[D];
    VQB_DEV float operator()(int i) const { return v[i]; }
};

struct SubInfo {        // per subspace, written by the prepare kernel
    float cmax2;        // max_j ||c_j||^2 (L2 kinds) ; unused for cosine
    uint32_t unsafe;    // 1: a centroid component is NaN/Inf or out of range -> every row takes the exact path
};

// margin constants: M = KAPPA * S, S = (||x|| + max||c||)^2 for the L2 kinds, ||x|| for cosine.
// Error budget behind it: x = x_hi + x_lo exactly, x_lo truncated to tf32 by the tensor
// core (<= 2^-21 |x_i|), c = c_hi + c_lo + (<= 2^-23 |c_i|), dropped x_lo.c_lo (<= 2^-21), fp32
// accumulation of <= 32 products, the norm pieces, and the reference's own rounding (<= 11 ulp of d).
// Measured on B200 (tests/test_gpu_tensor.py::test_tensor_scores_within_margin).
// Longer sub-vectors chain more MMAs and more products per score (and the reference's own sums are longer): the margin
// grows in proportion, KAPPA * (D / 8); tests/test_gpu_tensor.py measures the score error for every D against it.
constexpr float KAPPA = 1.0f / 131072.0f;  // 2^-17 = 7.6e-6 at sub_dim 8

// ------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// One probe, then sleep between probes.  try_wait returns after a few cycles on this part whatever suspend-time hint it is
// given (r02a profile: 23 % of all issued instructions were try_wait probes of waiting warps), so waiting warps back off
// with nanosleep; SLEEP_NS is chosen per hand-off from the slack the waiting role has.
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    if (SLEEP_NS == 0) { while (!mbar_try(bar, parity)) { } return; }
    do { __nanosleep(SLEEP_NS); } while (!mbar_try(bar, parity));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; LBO = step between the two 16-byte
// K chunks of one MMA (128 B here), SBO = step between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// K-major SWIZZLE_128B (the layout TMA writes): 8-row atoms of 1024 B, the K offset inside the 128-byte
// row is added to the start address and swizzled by the hardware (address bits 4-6 ^= bits 7-9).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}
// tcgen05.wait::ld tied to the destination registers, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait128(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]), "+r"(v[32]), "+r"(v[33]), "+r"(v[34]), "+r"(v[35]), "+r"(v[36]), "+r"(v[37]), "+r"(v[38]), "+r"(v[39]), "+r"(v[40]), "+r"(v[41]), "+r"(v[42]), "+r"(v[43]), "+r"(v[44]), "+r"(v[45]), "+r"(v[46]), "+r"(v[47]), "+r"(v[48]), "+r"(v[49]), "+r"(v[50]), "+r"(v[51]), "+r"(v[52]), "+r"(v[53]), "+r"(v[54]), "+r"(v[55]), "+r"(v[56]), "+r"(v[57]), "+r"(v[58]), "+r"(v[59]), "+r"(v[60]), "+r"(v[61]), "+r"(v[62]), "+r"(v[63]), "+r"(v[64]), "+r"(v[65]), "+r"(v[66]), "+r"(v[67]), "+r"(v[68]), "+r"(v[69]), "+r"(v[70]), "+r"(v[71]), "+r"(v[72]), "+r"(v[73]), "+r"(v[74]), "+r"(v[75]), "+r"(v[76]), "+r"(v[77]), "+r"(v[78]), "+r"(v[79]), "+r"(v[80]), "+r"(v[81]), "+r"(v[82]), "+r"(v[83]), "+r"(v[84]), "+r"(v[85]), "+r"(v[86]), "+r"(v[87]), "+r"(v[88]), "+r"(v[89]), "+r"(v[90]), "+r"(v[91]), "+r"(v[92]), "+r"(v[93]), "+r"(v[94]), "+r"(v[95]), "+r"(v[96]), "+r"(v[97]), "+r"(v[98]), "+r"(v[99]), "+r"(v[100]), "+r"(v[101]), "+r"(v[102]), "+r"(v[103]), "+r"(v[104]), "+r"(v[105]), "+r"(v[106]), "+r"(v[107]), "+r"(v[108]), "+r"(v[109]), "+r"(v[110]), "+r"(v[111]), "+r"(v[112]), "+r"(v[113]), "+r"(v[114]), "+r"(v[115]), "+r"(v[116]), "+r"(v[117]), "+r"(v[118]), "+r"(v[119]), "+r"(v[120]), "+r"(v[121]), "+r"(v[122]), "+r"(v[123]), "+r"(v[124]), "+r"(v[125]), "+r"(v[126]), "+r"(v[127])
                 :
                 : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {  // round-to-nearest tf32, low 13 bits zero
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fsat_ind(float g, float negH, float thH) {  // sat((th - g) * H): 1 iff g < th
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(g), "f"(negH), "f"(thH));
    return r;
}
__device__ __forceinline__ float tf32_lo(float v) {  // v - trunc_tf32(v): exact, what the tensor core drops when it reads v as tf32
    return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
}

// ------------------------------------------------------------------------------- prepare kernel
// One CTA per subspace, thread j = centroid slot j.  Builds the B operand tile
// [c_hi | c_lo | norm pieces] in its shared-memory image, a raw f32 copy and the cosine norms.
template <int MK, int TC_D>
__global__ void __launch_bounds__(TC_N) k_tc_prepare(const float* __restrict__ codebooks, int k, uint8_t* __restrict__ prep,
                                                      SubInfo* __restrict__ sinfo, const uint32_t* __restrict__ go) {
    __shared__ float red[TC_N];
    __shared__ uint32_t bad;
    if (go && *go == 0) return;
    using C = Cfg<TC_D>;
    constexpr uint32_t BP_SBO = C::BP_SBO, BP_BYTES = C::BP_BYTES, CB_BYTES = C::CB_BYTES;
    constexpr int NCH = C::NCH;
    const int s = blockIdx.x, j = threadIdx.x;
    if (j == 0) bad = 0;
    __syncthreads();
    uint8_t* img = prep + (size_t)s * C::PREP_BYTES;
    float c[TC_D];
    const bool real = j < k;
#pragma unroll
    for (int i = 0; i < TC_D; ++i) c[i] = real ? codebooks[((size_t)s * k + j) * TC_D + i] : 0.0f;
    double n2 = 0.0;
    bool finite = true;
#pragma unroll
    for (int i = 0; i < TC_D; ++i) {
        n2 += (double)c[i] * (double)c[i];
        finite = finite && !vqb_bad(c[i]) && fabsf(c[i]) < 1e15f;
    }
    if (!finite) atomicOr(&bad, 1u);
    float b[TC_D];
    float npiece[3] = {0.f, 0.f, 0.f};
    float rescore = 0.0f;   // cosine: -1/||c|| (score = x.c * rescore) ; L2 kinds: ||c||^2 (score = rescore - 2 x.c)
    if (MK == MK_COSINE) {
        // scores = -x.c/||c||;  ||c||^2 < FLT_MIN counts as a zero vector (cosine.c:38-45): score 0
        double inv = (finite && n2 >= (double)FLT_MIN) ? 1.0 / sqrt(n2) : 0.0;
        rescore = (float)(-inv);
#pragma unroll
        for (int i = 0; i < TC_D; ++i) b[i] = finite ? (float)(-(double)c[i] * inv) : 0.0f;
    } else {
#pragma unroll
        for (int i = 0; i < TC_D; ++i) b[i] = finite ? -2.0f * c[i] : 0.0f;
        double r = finite ? n2 : 0.0;
        rescore = (float)r;
        npiece[0] = to_tf32((float)r); r -= (double)npiece[0];
        npiece[1] = to_tf32((float)r); r -= (double)npiece[1];
        npiece[2] = to_tf32((float)r);
    }
    if (!real) { npiece[0] = 1.0e30f; npiece[1] = npiece[2] = 0.f; }  // padding slots can never be a minimum
    float hi[TC_D], lo[TC_D];
#pragma unroll
    for (int i = 0; i < TC_D; ++i) {
        hi[i] = to_tf32(b[i]);
        lo[i] = to_tf32(b[i] - hi[i]);
    }
    // B' image: row j, 16-byte k-chunk q at (j/8)*BP_SBO + q*128 + (j%8)*16
    float4* row = reinterpret_cast<float4*>(img + (j >> 3) * BP_SBO + (j & 7) * 16);
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
        row[q * 8] = make_float4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
        row[(NCH + q) * 8] = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
    }
    row[2 * NCH * 8] = make_float4(npiece[0], npiece[1], npiece[2], 0.f);
    float4* raw = reinterpret_cast<float4*>(img + BP_BYTES + j * (TC_D * 4));
#pragma unroll
    for (int q = 0; q < NCH; ++q) raw[q] = make_float4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
    {   // cosine norms exactly as hsd_sim_cosine_f32 accumulates them (cosine.c:163-198)
        bool tok;
        RegArr<TC_D> ca;
#pragma unroll
        for (int i = 0; i < TC_D; ++i) ca.v[i] = c[i];
        float nb = hsd_cosine_norm<TC_D>(ca, TC_D, tok);
        float2* aux = reinterpret_cast<float2*>(img + BP_BYTES + CB_BYTES + j * 8);
        *aux = make_float2(nb, __fsqrt_rn(nb));
        reinterpret_cast<float*>(img + BP_BYTES + CB_BYTES + AUX_BYTES)[j] = rescore;
    }
    red[j] = (real && finite) ? (float)n2 : 0.0f;
    __syncthreads();
    for (int st = TC_N / 2; st > 0; st >>= 1) {
        if (j < st) red[j] = fmaxf(red[j], red[j + st]);
        __syncthreads();
    }
    if (j == 0) {
        SubInfo si;
        si.cmax2 = red[0] * 1.0000005f;
        si.unsafe = bad;
        sinfo[s] = si;
    }
}

// ---------------------------------------------------------------------------- exact evaluation
// The reference's distance between the row's sub-vector (registers) and centroid j (shared memory).
template <int MK, int TC_D>
struct ExactEval {
    RegArr<TC_D> x;
    float na, sa;
    bool a_tail_ok;
    __device__ __forceinline__ void init() {
        na = sa = 0.f; a_tail_ok = true;
        if (MK == MK_COSINE) { na = hsd_cosine_norm<TC_D>(x, TC_D, a_tail_ok); sa = __fsqrt_rn(na); }
    }
    __device__ __forceinline__ float operator()(const float* cb, const float2* aux, int j) const {
        PtrAcc cp{cb + j * TC_D};
        if (MK == MK_TRAIN) return dist2_seq<TC_D>(x, cp, TC_D);
        if (MK == MK_COSINE) {
            float dot = hsd_cosine_dot<TC_D>(x, cp, TC_D);
            float2 a = aux[j];
            bool ok = a_tail_ok;  // codebook components are finite here (else the subspace is `unsafe`)
            float sim = 0.f;
            if (ok) sim = hsd_cosine_from_sums(dot, na, a.x, sa, a.y, ok);
            return ok ? __fsub_rn(1.0f, sim) : rust_cos<TC_D>(x, cp, TC_D);
        }
        return vq_distance<TC_D>(MK, x, cp, TC_D);
    }
};

// ROLE_HELP0/1: splitter of the even / odd row tiles and resolver of the even / odd units.  ROLE_SCAN0/1: scan of
// TMEM accumulator 0 / 1 (= units of that parity).
enum { ROLE_CTRL = 0, ROLE_HELP0 = 1, ROLE_HELP1 = 2, ROLE_SCAN0 = 3, ROLE_SCAN1 = 4 };

struct TcParams {
    uint32_t role_map;          // 4 bits per warpgroup: the ROLE_* it plays
    const uint32_t* go;         // training loop: the launch is a no-op when *go == 0 (speculatively enqueued iteration)
    const uint8_t* prep;        // [m] prepared images
    const SubInfo* sinfo;       // [m]
    const int* active;          // [m] 0/1 or nullptr (all active)
    void* codes;
    __half* recon;
    unsigned long long n;
    unsigned long long stride_row, stride_sub;
    int dim, m, k, n_groups, parts, num_tiles;
    uint32_t code_bytes;
    float* dbg_scores;            // DEBUG kernels only: [n][256] raw tensor-core scores of subspace dbg_sub
    unsigned long long* dbg_stats;  // DEBUG kernels only: [0] (row, subspace) pairs resolved by the full re-scan
    int dbg_sub;
    unsigned long long* dbg_ts;   // DEBUG kernels only: [dbg_ts_units][8] SM-clock stamps of CTA 0's hand-offs (vqb_debug_tc_timeline)
    int dbg_ts_units;
};

__device__ __forceinline__ void store_code(void* codes, uint32_t code_bytes, size_t off, uint32_t v) {
    if (code_bytes == 1) static_cast<uint8_t*>(codes)[off] = (uint8_t)v;
    else if (code_bytes == 2) static_cast<uint16_t*>(codes)[off] = (uint16_t)v;
    else static_cast<uint32_t*>(codes)[off] = v;
}

// ----------------------------------------------------------------------------------- main kernel
__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }

// Five warpgroups, roles assigned by p.role_map: control (TMA producer, MMA issuer) | splitter | resolve | scan columns 0-127 |
// scan columns 128-255.
// Both scan warpgroups drain EVERY accumulator, half the columns each, straight into registers (2 x tcgen05.ld.x64 per
// thread) and release it as soon as the loads have landed: the accumulator is busy for one TMEM read latency instead of
// a whole reduction, so the next MMA chain into it starts while its scores are still being reduced from registers.
template <int MK, int TC_D, bool DEBUG>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_assign(const __grid_constant__ CUtensorMap xmap, const TcParams p) {
    using C = Cfg<TC_D>;
    constexpr int TC_G = C::G, KC = C::KC, NCH = C::NCH;
    constexpr uint32_t BP_SBO = C::BP_SBO, BP_BYTES = C::BP_BYTES, CB_BYTES = C::CB_BYTES, MG_BYTES = C::MG_BYTES;
    constexpr uint32_t PREP_BYTES = C::PREP_BYTES;
    constexpr uint32_t OFF_RAW = C::OFF_RAW, OFF_AT = C::OFF_AT, OFF_BP = C::OFF_BP, OFF_CB = C::OFF_CB, OFF_AUX = C::OFF_AUX;
    constexpr uint32_t OFF_RINV = C::OFF_RINV, OFF_ONES = C::OFF_ONES, OFF_MG = C::OFF_MG, OFF_RES = C::OFF_RES;
    constexpr uint32_t OFF_SINFO = C::OFF_SINFO, OFF_BAR = C::OFF_BAR;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wgrp = warp >> 2, wq = warp & 3;
    const int role = (int)((p.role_map >> (4 * wgrp)) & 15u);   // which role this warpgroup plays (ROLE_*)
    const int grp = blockIdx.x % p.n_groups, part = blockIdx.x / p.n_groups;
    const int s0 = grp * TC_G;
    const int g_cnt = min(TC_G, p.m - s0);

    // barriers (arrival counts are per WARP: a warp's lanes synchronise with __syncwarp, lane 0 arrives)
    const uint32_t bar0 = sbase + OFF_BAR;
    constexpr int B_RAW_EMPTY = RAW_STAGES, B_AT_FULL = 2 * RAW_STAGES, B_AT_EMPTY = B_AT_FULL + AT_STAGES;
    constexpr int B_ACC_FULL = B_AT_EMPTY + AT_STAGES, B_ACC_EMPTY = B_ACC_FULL + 2;
    constexpr int B_MG_FULL = B_ACC_EMPTY + 2, B_MG_EMPTY = B_MG_FULL + MG_STAGES;
    constexpr int B_RES_FULL = B_MG_EMPTY + MG_STAGES, B_RES_EMPTY = B_RES_FULL + RES_STAGES;
    constexpr int B_COUNT = B_RES_EMPTY + RES_STAGES;
    static_assert(B_COUNT * 8 + 8 <= 512, "barrier area too small");
    auto RAW_FULL = [&](int i) { return bar0 + 8u * i; };
    auto RAW_EMPTY = [&](int i) { return bar0 + 8u * (B_RAW_EMPTY + i); };
    auto AT_FULL = [&](int i) { return bar0 + 8u * (B_AT_FULL + i); };
    auto AT_EMPTY = [&](int i) { return bar0 + 8u * (B_AT_EMPTY + i); };
    auto ACC_FULL = [&](int i) { return bar0 + 8u * (B_ACC_FULL + i); };
    auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (B_ACC_EMPTY + i); };
    auto MG_FULL = [&](int i) { return bar0 + 8u * (B_MG_FULL + i); };
    auto MG_EMPTY = [&](int i) { return bar0 + 8u * (B_MG_EMPTY + i); };
    auto RES_FULL = [&](int i) { return bar0 + 8u * (B_RES_FULL + i); };
    auto RES_EMPTY = [&](int i) { return bar0 + 8u * (B_RES_EMPTY + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 8 * B_COUNT);
    // the warp's lanes have all finished what the arrival stands for; one arrival per warp
    auto warp_arrive = [&](uint32_t bar) { __syncwarp(); if (lane == 0) mbar_arrive(bar); };

    if (p.go && *p.go == 0) return;
    // active subspaces of this group (same list for every role)
    uint32_t act_mask = 0;
    for (int i = 0; i < g_cnt; ++i)
        if (!p.active || p.active[s0 + i]) act_mask |= 1u << i;
    const int n_act = __popc(act_mask);
    if (n_act == 0) return;
    const int last_act = 31 - __clz(act_mask);

    // ---- one-time setup: prepared codebooks -> smem, constant ones tile, barriers, TMEM
    for (int i = 0; i < TC_G; ++i) {
        const int s = min(s0 + i, p.m - 1);
        const float4* src = reinterpret_cast<const float4*>(p.prep + (size_t)s * PREP_BYTES);
        float4* dbp = reinterpret_cast<float4*>(sm + OFF_BP + i * BP_BYTES);
        float4* dcb = reinterpret_cast<float4*>(sm + OFF_CB + i * CB_BYTES);
        float4* dax = reinterpret_cast<float4*>(sm + OFF_AUX + i * AUX_BYTES);
        float4* dri = reinterpret_cast<float4*>(sm + OFF_RINV + i * RINV_BYTES);
        for (int t = threadIdx.x; t < (int)(BP_BYTES / 16); t += TC_THREADS) dbp[t] = __ldg(src + t);
        for (int t = threadIdx.x; t < (int)(CB_BYTES / 16); t += TC_THREADS) dcb[t] = __ldg(src + BP_BYTES / 16 + t);
        for (int t = threadIdx.x; t < (int)(AUX_BYTES / 16); t += TC_THREADS) dax[t] = __ldg(src + (BP_BYTES + CB_BYTES) / 16 + t);
        for (int t = threadIdx.x; t < (int)(RINV_BYTES / 16); t += TC_THREADS)
            dri[t] = __ldg(src + (BP_BYTES + CB_BYTES + AUX_BYTES) / 16 + t);
    }
    if (threadIdx.x < BP_PAD / 16) reinterpret_cast<float4*>(sm + OFF_BP + TC_G * BP_BYTES)[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = threadIdx.x; t < (int)(ONES_BYTES / 16); t += TC_THREADS) {
        // [16 groups][2 chunks][8 rows][16 B]: chunk 0 = (1,1,1,0), chunk 1 = 0
        const bool chunk0 = ((t >> 3) & 1) == 0;
        reinterpret_cast<float4*>(sm + OFF_ONES)[t] = chunk0 ? make_float4(1.f, 1.f, 1.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x < TC_G) {  // per-subspace margin inputs: {sqrt(max ||c||^2), unsafe}
        const SubInfo si = p.sinfo[min(s0 + (int)threadIdx.x, p.m - 1)];
        reinterpret_cast<float2*>(sm + OFF_SINFO)[threadIdx.x] = make_float2(sqrtf(si.cmax2) * 1.0000005f, si.unsafe ? 1.0f : 0.0f);
    }
    if (threadIdx.x == 0) {
        // a raw tile is released by the splitter (4 warps), the resolve warpgroup once per unit (4 warps each) and the
        // last MMA that read it (1)
        for (int i = 0; i < RAW_STAGES; ++i) { mbar_init(RAW_FULL(i), 1); mbar_init(RAW_EMPTY(i), 4 + 4 * n_act + 1); }
        for (int i = 0; i < AT_STAGES; ++i) { mbar_init(AT_FULL(i), 4); mbar_init(AT_EMPTY(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), 4); }
        for (int i = 0; i < MG_STAGES; ++i) { mbar_init(MG_FULL(i), 4); mbar_init(MG_EMPTY(i), 4 * n_act); }
        for (int i = 0; i < RES_STAGES; ++i) { mbar_init(RES_FULL(i), 4); mbar_init(RES_EMPTY(i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();  // generic-proxy smem writes above -> visible to the tensor core / TMA
    if (role == ROLE_CTRL && wq == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_tiles = (p.num_tiles - part + p.parts - 1) / p.parts;  // tiles part, part+parts, ...

    if (role == ROLE_CTRL) {
        reg_dec<REGS_CTRL>();
        if (wq == 0) {
            // ================================ TMA producer ================================
            if (lane == 0) {
                for (int it = 0; it < my_tiles; ++it) {
                    const int st = it % RAW_STAGES, ph = (it / RAW_STAGES) & 1;
                    mbar_wait<256>(RAW_EMPTY(st), ph ^ 1);
                    mbar_expect_tx(RAW_FULL(st), RAW_BYTES);
                    const int tile = part + it * p.parts;
                    tma_load_2d(sbase + OFF_RAW + st * RAW_BYTES, &xmap, s0 * TC_D, tile * TC_ROWS, RAW_FULL(st));   // 32 columns from the group's first
                }
            }
        } else if (wq == 1) {
            // ================================ MMA issuer ==================================
            // The whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers, so each
            // tcgen05.mma is one instruction instead of a divergence loop); one elected lane issues.
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
            const uint64_t ones_desc = make_desc(sbase + OFF_ONES, 128, 256);
            // descriptor bases; the 14-bit address field (bytes >> 4) never carries: every operand lies below 256 KB
            const uint64_t raw_desc0 = make_desc_sw128(sbase + OFF_RAW), at_desc0 = make_desc_sw128(sbase + OFF_AT);
            const uint64_t b_desc0 = make_desc(sbase + OFF_BP, 128, BP_SBO);
            const bool use_norm = (MK != MK_COSINE) || (p.k < TC_N);
            uint32_t u = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int st = it % RAW_STAGES, ph = (it / RAW_STAGES) & 1;
                const int at = it % AT_STAGES, atph = (it / AT_STAGES) & 1;
                mbar_wait<64>(RAW_FULL(st), ph);   // TMA bytes have landed: the tensor core reads x_hi from the tile itself
                mbar_wait<64>(AT_FULL(at), atph);  // x_lo tile (the splitter runs a tile ahead)
                for (int i = 0; i < g_cnt; ++i) {
                    if (!(act_mask >> i & 1)) continue;
                    const int acc = u & 1, cph = (u >> 1) & 1;
                    mbar_wait<VQB_SPIN_EMPTY>(ACC_EMPTY(acc), cph ^ 1);
                    tc_fence_after();
                    // K = 8 step j of the sub-vector: 32 bytes further along the 128-byte row (A), two 16-byte chunks
                    // further in the B image
                    const uint64_t xhi = raw_desc0 + (uint64_t)((st * RAW_BYTES + i * (TC_D * 4)) >> 4);
                    const uint64_t xlo = at_desc0 + (uint64_t)((at * RAW_BYTES + i * (TC_D * 4)) >> 4);
                    const uint64_t bhi = b_desc0 + (uint64_t)((i * BP_BYTES) >> 4);
                    const uint32_t d = tmem_base + acc * TC_N;
                    if (elect_one()) {
                        if (DEBUG && p.dbg_ts && blockIdx.x == 0 && (int)u < p.dbg_ts_units) p.dbg_ts[u * 8 + 0] = clock64();
#pragma unroll
                        for (int j = 0; j < KC; ++j) umma_tf32(d, xhi + 2 * j, bhi + 16 * j, idesc, j > 0);          // x_hi . c_hi
#pragma unroll
                        for (int j = 0; j < KC; ++j) umma_tf32(d, xhi + 2 * j, bhi + 8 * NCH + 16 * j, idesc, 1);    // x_hi . c_lo
                        if (use_norm) umma_tf32(d, ones_desc, bhi + 16 * NCH, idesc, 1);                             // + ||c||^2
#pragma unroll
                        for (int j = 0; j < KC; ++j) umma_tf32(d, xlo + 2 * j, bhi + 16 * j, idesc, 1);              // x_lo . c_hi
                        umma_commit(ACC_FULL(acc));
                        if (i == last_act) { umma_commit(AT_EMPTY(at)); umma_commit(RAW_EMPTY(st)); }
                        if (DEBUG && p.dbg_ts && blockIdx.x == 0 && (int)u < p.dbg_ts_units) p.dbg_ts[u * 8 + 1] = clock64();
                    }
                    __syncwarp();
                    ++u;
                }
            }
        }
    } else if (role >= ROLE_SCAN0) {
        // ================================ scan ========================================
        reg_inc<REGS_SCAN>();
        const int acc = role - ROLE_SCAN0;          // TMEM accumulator (= unit parity) this warpgroup drains, whole rows
        const int quarter = wq;                     // TMEM lane quarter this warp may read
        const int r = quarter * 32 + lane;          // tile row = TMEM lane
        const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * TC_N;
        uint32_t u = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int mg = it % MG_STAGES, mph = (it / MG_STAGES) & 1;
            for (int i = 0; i < g_cnt; ++i) {
                if (!(act_mask >> i & 1)) continue;
                if ((int)(u & 1) != acc) { ++u; continue; }
                const int cph = (u >> 1) & 1;
                const int rs = u % RES_STAGES, rph = (u / RES_STAGES) & 1;
                ++u;

                mbar_wait<VQB_SPIN_FULL>(ACC_FULL(acc), cph);
                tc_fence_after();
                const bool stamp = DEBUG && p.dbg_ts && blockIdx.x == 0 && r == 0 && (int)(u - 1) < p.dbg_ts_units;
                if (stamp) p.dbg_ts[(u - 1) * 8 + 2] = clock64();
                // ---- 256 scores -> 64 minima of four columns; the next x32 load is in flight while one is reduced
                float gm[64];
                uint32_t va[32], vb[32];
                auto dump = [&](const uint32_t (&v)[32], int c) {   // DEBUG only: raw scores of one subspace
                    const unsigned long long row = (unsigned long long)(part + it * p.parts) * TC_ROWS + r;
                    if (p.dbg_scores && s0 + i == p.dbg_sub && row < p.n) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) p.dbg_scores[(size_t)row * TC_N + c * 32 + q] = __uint_as_float(v[q]);
                    }
                };
                auto gmin8 = [&](const uint32_t (&v)[32], float* g) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        g[q] = fminf(fmin3(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2])),
                                     __uint_as_float(v[4 * q + 3]));
                };
                tmem_ld32(tcol, va);
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    tmem_ld_wait32(va);
                    tmem_ld32(tcol + (c + 1) * 32, vb);
                    if (DEBUG) dump(va, c);
                    gmin8(va, &gm[8 * c]);
                    tmem_ld_wait32(vb);
                    if (c + 2 < 8) tmem_ld32(tcol + (c + 2) * 32, va);
                    if (DEBUG) dump(vb, c + 1);
                    gmin8(vb, &gm[8 * c + 8]);
                }
                tc_fence_before();
                warp_arrive(ACC_EMPTY(acc));  // TMEM accumulator may be overwritten by the next MMA chain
                if (stamp) p.dbg_ts[(u - 1) * 8 + 3] = clock64();
                if (DEBUG && p.dbg_ts && blockIdx.x == 0 && lane == 0 && (int)(u - 1) < p.dbg_ts_units)   // last warp's release
                    atomicMax(&p.dbg_ts[(u - 1) * 8 + 7], (unsigned long long)clock64());

                // ---- row minimum
                float t1[22];
#pragma unroll
                for (int q = 0; q < 21; ++q) t1[q] = fmin3(gm[3 * q], gm[3 * q + 1], gm[3 * q + 2]);
                t1[21] = gm[63];
                float t2[8];
#pragma unroll
                for (int q = 0; q < 7; ++q) t2[q] = fmin3(t1[3 * q], t1[3 * q + 1], t1[3 * q + 2]);
                t2[7] = t1[21];
                const float mall = fmin3(fmin3(t2[0], t2[1], t2[2]), fmin3(t2[3], t2[4], t2[5]), fminf(t2[6], t2[7]));

                mbar_wait<64>(MG_FULL(mg), mph);   // complete long before the first unit of the tile gets here
                const float2 hm = reinterpret_cast<const float2*>(sm + OFF_MG + mg * MG_BYTES)[i * TC_ROWS + r];
                warp_arrive(MG_EMPTY(mg));
                const float H = hm.x, M = hm.y;
                const float negH = -H;
                const float thH = fmaf(mall, H, M * H);  // (mall + M) * H

                // ---- groups within M of the minimum: indicator = sat((th - g) * H) in {0, 1}; the odd weights
                // 129 + 2t make the sum decode to t iff exactly one group is flagged (two flagged groups sum to
                // >= 260; a fractional indicator -- g within S * 2^-40 of the threshold -- cannot produce an odd
                // integer together with the always-full indicator of the minimum, and the resolve stage re-checks
                // the decoded group against the row minimum anyway)
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int t = 0; t < 64; ++t) {
                    const float w = (float)(129 + 2 * t), ind = fsat_ind(gm[t], negH, thH);
                    if ((t & 3) == 0) a0 = fmaf(ind, w, a0);
                    if ((t & 3) == 1) a1 = fmaf(ind, w, a1);
                    if ((t & 3) == 2) a2 = fmaf(ind, w, a2);
                    if ((t & 3) == 3) a3 = fmaf(ind, w, a3);
                }
                const float accw = (a0 + a1) + (a2 + a3);
                const int wi = (int)accw;
                const bool single = (M >= 0.f) && (accw >= 129.f) && (accw <= 255.f) && (accw == floorf(accw)) && (wi & 1);
                // result word: the margin with its low six mantissa bits replaced by the group index; sign bit = ambiguous
                const uint32_t word = single ? ((__float_as_uint(M) & ~63u) | (uint32_t)((wi - 129) >> 1)) : RES_AMBIGUOUS;
                mbar_wait<64>(RES_EMPTY(rs), rph ^ 1);
                reinterpret_cast<float2*>(sm + OFF_RES + rs * RES_BYTES)[r] = make_float2(mall, __uint_as_float(word));
                warp_arrive(RES_FULL(rs));
                if (stamp) p.dbg_ts[(u - 1) * 8 + 4] = clock64();
            }
        }
    } else {
        // ================================ helpers: split the row tiles of one parity, resolve the units of one parity ======
        reg_dec<REGS_HELP>();
        const int par = role - ROLE_HELP0;
        const int r = wq * 32 + lane;               // tile row
        const uint32_t r7 = (uint32_t)(r & 7);
        // lane-rotated gather order: lane-dependent rotations of the 16-byte chunk inside a candidate (rc) and of the
        // candidate inside its group of four (rot) make the eight lanes of a quarter-warp read eight different 16-byte
        // bank slots of their (arbitrary) candidate groups -> conflict-free LDS.128 gathers (D = 24: two-way at worst)
        constexpr int RCB = (NCH == 2) ? 1 : (NCH == 8 ? 3 : 2);       // lane bits that rotate the chunk
        const int rc = (lane & ((1 << RCB) - 1)) % NCH, rot = (lane >> RCB) & 3;
        const bool padded = p.k < TC_N;

        // x_lo tile + margins of row tile `ts` (runs one tile ahead of the MMAs)
        auto split_tile = [&](int ts) {
            const int st = ts % RAW_STAGES, ph = (ts / RAW_STAGES) & 1;
            const int at = ts % AT_STAGES, atph = (ts / AT_STAGES) & 1;
            const int mg = ts % MG_STAGES, mph = (ts / MG_STAGES) & 1;
            mbar_wait<128>(RAW_FULL(st), ph);
            const uint8_t* rawrow = sm + OFF_RAW + st * RAW_BYTES + r * 128;
            uint8_t* atrow = sm + OFF_AT + at * RAW_BYTES + r * 128;
            // SWIZZLE_128B: 16-byte chunk c of row r lives at chunk c ^ (r & 7); the x_lo tile keeps the same layout
            float4 v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(rawrow + ((c ^ r7) << 4));
            mbar_wait<128>(AT_EMPTY(at), atph ^ 1);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(atrow + ((c ^ r7) << 4)) = make_float4(tf32_lo(v[c].x), tf32_lo(v[c].y), tf32_lo(v[c].z), tf32_lo(v[c].w));
            fence_proxy_async();
            warp_arrive(AT_FULL(at));
            // every subspace's error margin M = KAPPA * S and indicator scale H = 2^(40 - floor(log2 S)):
            // (th - g) * H >= 1 for every representable g < th in the score range, th * H far from overflow
            float2 hm[TC_G];
#pragma unroll
            for (int i = 0; i < TC_G; ++i) {
                float nx2 = 0.f;
#pragma unroll
                for (int q = 0; q < NCH; ++q) {
                    const float4 vq = v[NCH * i + q];
                    nx2 = fmaf(vq.x, vq.x, nx2); nx2 = fmaf(vq.y, vq.y, nx2); nx2 = fmaf(vq.z, vq.z, nx2); nx2 = fmaf(vq.w, vq.w, nx2);
                }
                const float2 si = reinterpret_cast<const float2*>(sm + OFF_SINFO)[i];
                float S;
                if (MK == MK_COSINE) S = sqrt_approx(nx2) * 1.0000005f;
                else { const float t = sqrt_approx(nx2) * 1.0000005f + si.x; S = t * t; }
                // rows the pruning cannot be trusted on: NaN/Inf/huge/tiny magnitudes (negative M marks them);
                // cosine rows near hsdlib's zero-vector rule (||x||^2 < FLT_MIN, cosine.c:38-45) go the exact way too
                bool amb = (si.y != 0.f) || !(nx2 < 1e30f) || !(S > 1e-25f) || !(S < 1e30f);
                if (MK == MK_COSINE) amb = amb || !(nx2 > 4.0f * FLT_MIN);
                const uint32_t sexp = (__float_as_uint(S) >> 23) & 0xFFu;
                hm[i].x = amb ? 1.0f : __uint_as_float((294u - sexp) << 23);
                hm[i].y = amb ? -1.0f : (KAPPA * (float)(TC_D / 8)) * S;
            }
            mbar_wait<128>(MG_EMPTY(mg), mph ^ 1);
#pragma unroll
            for (int i = 0; i < TC_G; ++i) reinterpret_cast<float2*>(sm + OFF_MG + mg * MG_BYTES)[i * TC_ROWS + r] = hm[i];
            warp_arrive(MG_FULL(mg));
            warp_arrive(RAW_EMPTY(st));
        };

        if (par == 0) split_tile(0);
        uint32_t u = 0;
        for (int it = 0; it < my_tiles; ++it) {
            if (it + 1 < my_tiles && ((it + 1) & 1) == par) split_tile(it + 1);
            const int st = it % RAW_STAGES, ph = (it / RAW_STAGES) & 1;
            const unsigned long long row = (unsigned long long)(part + it * p.parts) * TC_ROWS + r;
            const bool live = row < p.n;
            bool raw_seen = false;
            const uint8_t* rawrow = sm + OFF_RAW + st * RAW_BYTES + r * 128;
            for (int i = 0; i < g_cnt; ++i) {
                if (!(act_mask >> i & 1)) continue;
                if ((int)(u & 1) != par) { ++u; continue; }
                const int rs = u % RES_STAGES, rph = (u / RES_STAGES) & 1;
                ++u;
                if (!raw_seen) { mbar_wait<128>(RAW_FULL(st), ph); raw_seen = true; }
                const int s = s0 + i;
                // this row's sub-vector, its 16-byte chunks in this lane's gather order: xr[q] = chunk (q + rc) mod NCH
                float4 xr[NCH];
#pragma unroll
                for (int q = 0; q < NCH; ++q) {
                    int ch = q + rc; if (ch >= NCH) ch -= NCH;
                    xr[q] = *reinterpret_cast<const float4*>(rawrow + (((NCH * i + ch) ^ r7) << 4));
                }
                mbar_wait<64>(RES_FULL(rs), rph);
                const float2 rv = reinterpret_cast<const float2*>(sm + OFF_RES + rs * RES_BYTES)[r];
                warp_arrive(RES_EMPTY(rs));
                const bool stamp = DEBUG && p.dbg_ts && blockIdx.x == 0 && r == 0 && (int)(u - 1) < p.dbg_ts_units;
                if (stamp) p.dbg_ts[(u - 1) * 8 + 5] = clock64();
                const float ref = rv.x;
                const uint32_t word = __float_as_uint(rv.y);
                const float* cb = reinterpret_cast<const float*>(sm + OFF_CB + i * CB_BYTES);
                const float2* aux = reinterpret_cast<const float2*>(sm + OFF_AUX + i * AUX_BYTES);
                uint32_t best = 0;
                bool amb = true;
                if (!(word & RES_AMBIGUOUS)) {
                    // fp32 re-score of the four candidates in the tensor core's own form: -x.c/||c|| or ||c||^2 - 2 x.c
                    const int t = (int)(word & 63u);
                    const float M = __uint_as_float(word & ~63u);
                    const uint8_t* gbase = reinterpret_cast<const uint8_t*>(cb) + t * (4 * TC_D * 4);
                    const float* rsc = reinterpret_cast<const float*>(sm + OFF_RINV + i * RINV_BYTES) + 4 * t;
                    float sc[4];
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        const int q = (ii + rot) & 3;
                        float dot = 0.f;
#pragma unroll
                        for (int qq = 0; qq < NCH; ++qq) {
                            int ch = qq + rc; if (ch >= NCH) ch -= NCH;
                            const float4 cv = *reinterpret_cast<const float4*>(gbase + q * (TC_D * 4) + ch * 16);
                            dot = (qq == 0) ? cv.x * xr[0].x : fmaf(cv.x, xr[qq].x, dot);
                            dot = fmaf(cv.y, xr[qq].y, dot); dot = fmaf(cv.z, xr[qq].z, dot); dot = fmaf(cv.w, xr[qq].w, dot);
                        }
                        sc[ii] = (MK == MK_COSINE) ? dot * rsc[q] : fmaf(dot, -2.0f, rsc[q]);
                        if (padded && 4 * t + q >= p.k) sc[ii] = __int_as_float(0x7f800000);  // padding slot
                    }
                    const float lo01 = fminf(sc[0], sc[1]), hi01 = fmaxf(sc[0], sc[1]);
                    const float lo23 = fminf(sc[2], sc[3]), hi23 = fmaxf(sc[2], sc[3]);
                    const float b1 = fminf(lo01, lo23);
                    const float b2 = fmin3(hi01, hi23, fmaxf(lo01, lo23));
                    const int ib = (lo01 <= lo23) ? ((sc[0] <= sc[1]) ? 0 : 1) : ((sc[2] <= sc[3]) ? 2 : 3);
                    // accept iff this group really holds the row minimum and its best beats its runner-up by more than M
                    const bool ok = (fabsf(b1 - ref) <= 0.5f * M) && (b2 - b1 > M);
                    amb = !ok;
                    best = (uint32_t)(4 * t + ((ib + rot) & 3));
                }
                // ---- everything else: the warp scans all k centroids of the row with the reference formula
                uint32_t todo = __ballot_sync(0xFFFFFFFFu, amb && live);
                if (DEBUG && p.dbg_stats && lane == 0 && todo) atomicAdd(p.dbg_stats, (unsigned long long)__popc(todo));
                if (todo) {
                    ExactEval<MK, TC_D> ev;  // this row's sub-vector in natural order (re-read: the rotation is per lane)
#pragma unroll
                    for (int q = 0; q < NCH; ++q) {
                        const float4 vq = *reinterpret_cast<const float4*>(rawrow + (((NCH * i + q) ^ r7) << 4));
                        ev.x.v[4 * q] = vq.x; ev.x.v[4 * q + 1] = vq.y; ev.x.v[4 * q + 2] = vq.z; ev.x.v[4 * q + 3] = vq.w;
                    }
                    while (todo) {
                        const int L = __ffs(todo) - 1;
                        todo &= todo - 1;
                        ExactEval<MK, TC_D> eo;
#pragma unroll
                        for (int q = 0; q < TC_D; ++q) eo.x.v[q] = __shfl_sync(0xFFFFFFFFu, ev.x.v[q], L);
                        eo.init();
                        float bd = __int_as_float(0x7f800000);
                        uint32_t bj = 0xFFFFFFFFu;
                        bool d0nan = false;
                        for (int j = lane; j < p.k; j += 32) {
                            float dd = eo(cb, aux, j);
                            if (j == 0) d0nan = isnan(dd);
                            if (isnan(dd)) dd = __int_as_float(0x7f800000);
                            if (bj == 0xFFFFFFFFu || dd < bd) { bd = dd; bj = (uint32_t)j; }
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            const float od = __shfl_xor_sync(0xFFFFFFFFu, bd, off);
                            const uint32_t oj = __shfl_xor_sync(0xFFFFFFFFu, bj, off);
                            if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
                        }
                        d0nan = __shfl_sync(0xFFFFFFFFu, d0nan ? 1 : 0, 0) != 0;
                        if (lane == L) best = d0nan ? 0u : bj;  // vector.rs:354-361: a NaN at index 0 is never replaced
                    }
                }
                if (live) {
                    if (p.codes) store_code(p.codes, p.code_bytes, (size_t)row * p.stride_row + (size_t)s * p.stride_sub, best);
                    if (p.recon) {  // pq.rs:193-195: f16::from_f32 of the chosen centroid
                        const float4* c = reinterpret_cast<const float4*>(cb + best * TC_D);
                        __half* dst = p.recon + (size_t)row * p.dim + (size_t)s * TC_D;
#pragma unroll
                        for (int q = 0; q < NCH; q += 2) {
                            const float4 c0 = c[q], c1 = c[q + 1];
                            __half2 h[4];
                            h[0] = __floats2half2_rn(c0.x, c0.y); h[1] = __floats2half2_rn(c0.z, c0.w);
                            h[2] = __floats2half2_rn(c1.x, c1.y); h[3] = __floats2half2_rn(c1.z, c1.w);
                            *reinterpret_cast<uint4*>(dst + 4 * q) = *reinterpret_cast<uint4*>(h);
                        }
                    }
                }
                warp_arrive(RAW_EMPTY(st));
                if (stamp) p.dbg_ts[(u - 1) * 8 + 6] = clock64();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (role == ROLE_CTRL && wq == 1) tmem_dealloc(tmem_base, 512);
}

// --------------------------------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(f);
    });
    return fn;
}

// timing variant selected by vqb_debug_tc_variant: the warpgroup -> role map (see vqb_tc_assign_launch)
int g_tc_variant = 1;

template <int MK, int D, bool DEBUG = false>
int launch_tc(vqb_ctx* ctx, const CUtensorMap& map, const TcParams& p, int grid) {
    auto kern = k_tc_assign<MK, D, DEBUG>;
    VQB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<D>::SMEM_BYTES));
    kern<<<grid, TC_THREADS, Cfg<D>::SMEM_BYTES, ctx->stream>>>(map, p);
    VQB_LAUNCHED(ctx);
    return VQB_SUCCESS;
}
template <int MK>
int launch_tc_d(vqb_ctx* ctx, const CUtensorMap& map, const TcParams& p, int grid, int d) {
    switch (d) {
        case 8: return launch_tc<MK, 8>(ctx, map, p, grid);
        case 16: return launch_tc<MK, 16>(ctx, map, p, grid);
        case 24: return launch_tc<MK, 24>(ctx, map, p, grid);
        case 32: return launch_tc<MK, 32>(ctx, map, p, grid);
    }
    return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "sub_dim %d has no tensor-core path", d);
}
inline int tc_group(int d) { return d == 24 ? 1 : 32 / d; }   // Cfg<D>::G

}  // namespace

int vqb_make_x_tensormap(vqb_ctx* ctx, const float* x, size_t n, size_t dim, CUtensorMap_st* out, int box_cols, int swizzle) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return vqb_fail(ctx, VQB_FAILURE, "cuTensorMapEncodeTiled is not available");
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)n};
    cuuint64_t gstr[1] = {(cuuint64_t)dim * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)TC_ROWS};
    cuuint32_t estr[2] = {1, 1};
    // 32 columns = one 128-byte line per row, swizzled (what the tensor core and the tile kernels expect); any other
    // width lands as plain rows of box_cols floats
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (swizzle < 0 ? box_cols == 32 : swizzle != 0) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return vqb_fail(ctx, VQB_FAILURE, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VQB_SUCCESS;
}

size_t vqb_tc_prep_bytes(size_t m, size_t d) {
    if (d != 8 && d != 16 && d != 24 && d != 32) d = 8;
    return m * prep_bytes_of(d) + m * sizeof(SubInfo) + 256;
}

bool vqb_tc_supported(int mk, const float* x, size_t n, size_t dim, size_t m, size_t k, size_t d) {
    (void)m;
    if (mk == MK_MANHATTAN || mk == MK_CHEBYSHEV) return false;   // not contractions: CUDA-core kernels by design
    if ((d != 8 && d != 16 && d != 24 && d != 32) || k == 0 || k > TC_N) return false;
    if (dim % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return false;  // TMA: 16-byte aligned rows
    if (n == 0 || n >= (size_t)1 << 31) return false;
    return get_encode_fn() != nullptr;
}

int vqb_tc_prepare(vqb_ctx* ctx, int mk, const float* codebooks, size_t m, size_t k, size_t d, void* prep, const uint32_t* go) {
    uint8_t* img = static_cast<uint8_t*>(prep);
    SubInfo* sinfo = reinterpret_cast<SubInfo*>(img + ((m * prep_bytes_of(d) + 255) & ~(size_t)255));
#define VQB_PREP_CASE(DD)                                                                                                  \
    case DD:                                                                                                               \
        if (mk == MK_COSINE) k_tc_prepare<MK_COSINE, DD><<<(unsigned)m, TC_N, 0, ctx->stream>>>(codebooks, (int)k, img, sinfo, go); \
        else k_tc_prepare<MK_TRAIN, DD><<<(unsigned)m, TC_N, 0, ctx->stream>>>(codebooks, (int)k, img, sinfo, go);          \
        break;
    switch (d) {
        VQB_PREP_CASE(8)
        VQB_PREP_CASE(16)
        VQB_PREP_CASE(24)
        VQB_PREP_CASE(32)
        default: return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "sub_dim %zu has no tensor-core path", d);
    }
#undef VQB_PREP_CASE
    VQB_LAUNCHED(ctx);
    return VQB_SUCCESS;
}

int vqb_tc_assign_launch(vqb_ctx* ctx, int mk, const float* x, size_t n, size_t dim, size_t m, size_t k, const void* prep,
                         const int* active_dev, void* codes, uint32_t code_bytes, size_t stride_row, size_t stride_sub,
                         __half* recon, float* dbg_scores, unsigned long long* dbg_stats, int dbg_sub,
                         unsigned long long* dbg_ts, int dbg_ts_units, const uint32_t* go) {
    if (n == 0) return VQB_SUCCESS;
    CUtensorMap map;
    VQB_TRY(vqb_make_x_tensormap(ctx, x, n, dim, &map));
    TcParams p;
    p.go = go;
    const uint8_t* img = static_cast<const uint8_t*>(prep);
    p.prep = img;
    const int d = (int)(dim / m);
    p.sinfo = reinterpret_cast<const SubInfo*>(img + ((m * prep_bytes_of((size_t)d) + 255) & ~(size_t)255));
    p.active = active_dev;
    p.codes = codes; p.recon = recon;
    p.n = n; p.stride_row = stride_row; p.stride_sub = stride_sub;
    p.dim = (int)dim; p.m = (int)m; p.k = (int)k;
    p.n_groups = (int)((m + tc_group(d) - 1) / tc_group(d));
    p.num_tiles = (int)((n + TC_ROWS - 1) / TC_ROWS);
    p.parts = std::max(1, std::min(p.num_tiles, ctx->sm_count / p.n_groups));
    p.code_bytes = code_bytes;
    p.dbg_scores = dbg_scores; p.dbg_stats = dbg_stats; p.dbg_sub = dbg_sub;
    p.dbg_ts = dbg_ts; p.dbg_ts_units = dbg_ts_units;
    // warpgroup -> role.  The SM's warp arbiter prefers higher warp ids: the latency-critical MMA issuer sits highest.
    static const uint32_t role_maps[4] = {
        ROLE_HELP0 | ROLE_HELP1 << 4 | ROLE_SCAN0 << 8 | ROLE_SCAN1 << 12 | ROLE_CTRL << 16,
        ROLE_CTRL | ROLE_HELP0 << 4 | ROLE_HELP1 << 8 | ROLE_SCAN0 << 12 | ROLE_SCAN1 << 16,
        ROLE_SCAN0 | ROLE_SCAN1 << 4 | ROLE_HELP0 << 8 | ROLE_HELP1 << 12 | ROLE_CTRL << 16,
        ROLE_HELP0 | ROLE_SCAN0 << 4 | ROLE_CTRL << 8 | ROLE_SCAN1 << 12 | ROLE_HELP1 << 16};
    p.role_map = role_maps[g_tc_variant & 3];
    const int grid = p.n_groups * p.parts;
    if (dbg_scores || dbg_stats || dbg_ts) {
        if (mk != MK_COSINE && mk != MK_TRAIN)
            return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "debug capture exists for the training and cosine kinds only");
        switch (d) {
            case 8: return mk == MK_COSINE ? launch_tc<MK_COSINE, 8, true>(ctx, map, p, grid) : launch_tc<MK_TRAIN, 8, true>(ctx, map, p, grid);
            case 16: return mk == MK_COSINE ? launch_tc<MK_COSINE, 16, true>(ctx, map, p, grid) : launch_tc<MK_TRAIN, 16, true>(ctx, map, p, grid);
            case 24: return mk == MK_COSINE ? launch_tc<MK_COSINE, 24, true>(ctx, map, p, grid) : launch_tc<MK_TRAIN, 24, true>(ctx, map, p, grid);
            case 32: return mk == MK_COSINE ? launch_tc<MK_COSINE, 32, true>(ctx, map, p, grid) : launch_tc<MK_TRAIN, 32, true>(ctx, map, p, grid);
        }
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "sub_dim %d has no tensor-core path", d);
    }
    switch (mk) {
        case MK_SQEUCLID: return launch_tc_d<MK_SQEUCLID>(ctx, map, p, grid, d);
        case MK_EUCLID: return launch_tc_d<MK_EUCLID>(ctx, map, p, grid, d);
        case MK_COSINE: return launch_tc_d<MK_COSINE>(ctx, map, p, grid, d);
        case MK_TRAIN: return launch_tc_d<MK_TRAIN>(ctx, map, p, grid, d);
    }
    return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "metric kind %d has no tensor-core path", mk);
}

// Diagnostics: selects the warp-role order of the tensor kernel (0: scan warpgroups 2-3, 1: scan warpgroups 4-5).
// Every variant produces the same codes; exists so that A/B timings need one build.
extern "C" int vqb_debug_tc_variant(int order) {
    g_tc_variant = order;
    return VQB_SUCCESS;
}

extern "C" int vqb_debug_tc_scores(vqb_ctx* ctx, int cosine, const float* x, size_t n, size_t dim, size_t m, size_t k,
                                   const float* codebooks, int sub, float* scores_out, uint64_t* rescans_out,
                                   uint32_t* codes_out) {
    if (!ctx || !x || !codebooks) return VQB_ERR_NULL_PTR;
    if (m == 0 || dim % m) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "dim must be divisible by m");
    const size_t d = dim / m;
    const int mk = cosine ? MK_COSINE : MK_TRAIN;
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    InputView xin, cin;
    OutputView so, co;
    VQB_TRY(xin.bind(ctx, x, n * dim * 4));
    VQB_TRY(cin.bind(ctx, codebooks, m * k * d * 4));
    VQB_TRY(so.bind(ctx, scores_out, scores_out ? n * TC_N * 4 : 0));
    VQB_TRY(co.bind(ctx, codes_out, codes_out ? m * n * 4 : 0));
    if (!vqb_tc_supported(mk, static_cast<const float*>(xin.dev), n, dim, m, k, d))
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "shape not covered by the tensor-core kernel");
    DevBuf prep, stats;
    VQB_CUDA(ctx, prep.alloc(vqb_tc_prep_bytes(m, d)));
    VQB_CUDA(ctx, stats.alloc(8));
    VQB_CUDA(ctx, cudaMemsetAsync(stats.p, 0, 8, ctx->stream));
    VQB_TRY(vqb_tc_prepare(ctx, mk, static_cast<const float*>(cin.dev), m, k, d, prep.p));
    VQB_TRY(vqb_tc_assign_launch(ctx, mk, static_cast<const float*>(xin.dev), n, dim, m, k, prep.p, nullptr, co.dev, 4,
                                 /*stride_row=*/1, /*stride_sub=*/n, nullptr, static_cast<float*>(so.dev),
                                 stats.as<unsigned long long>(), sub));
    VQB_TRY(so.finish(ctx));
    VQB_TRY(co.finish(ctx));
    if (rescans_out) VQB_CUDA(ctx, cudaMemcpyAsync(rescans_out, stats.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}

// Diagnostics: SM-clock stamps of CTA 0's per-unit hand-offs in one cosine-encode pass (ts_out[units][8], host):
// 0 issuer saw ACC_EMPTY, 1 issuer issued + committed, 2 scan saw ACC_FULL, 3 scan released the accumulator,
// 4 scan published its result, 5 resolve saw it, 6 resolve done, 7 the last scan warp released the accumulator.
extern "C" int vqb_debug_tc_timeline(vqb_ctx* ctx, const float* x, size_t n, size_t dim, size_t m, size_t k,
                                     const float* codebooks, uint64_t* ts_out, int units) {
    if (!ctx || !x || !codebooks || !ts_out) return VQB_ERR_NULL_PTR;
    if (m == 0 || dim % m) return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "dim must be divisible by m");
    const size_t d = dim / m;
    std::lock_guard<std::mutex> lk(ctx->mu);
    VQB_CUDA(ctx, cudaSetDevice(ctx->device));
    InputView xin, cin;
    VQB_TRY(xin.bind(ctx, x, n * dim * 4));
    VQB_TRY(cin.bind(ctx, codebooks, m * k * d * 4));
    if (!vqb_tc_supported(MK_COSINE, static_cast<const float*>(xin.dev), n, dim, m, k, d))
        return vqb_fail(ctx, VQB_ERR_INVALID_INPUT, "shape not covered by the tensor-core kernel");
    DevBuf prep, ts, codes;
    VQB_CUDA(ctx, prep.alloc(vqb_tc_prep_bytes(m, d)));
    VQB_CUDA(ctx, ts.alloc((size_t)units * 8 * 8));
    VQB_CUDA(ctx, codes.alloc(n * m));
    VQB_CUDA(ctx, cudaMemsetAsync(ts.p, 0, (size_t)units * 64, ctx->stream));
    VQB_TRY(vqb_tc_prepare(ctx, MK_COSINE, static_cast<const float*>(cin.dev), m, k, d, prep.p));
    VQB_TRY(vqb_tc_assign_launch(ctx, MK_COSINE, static_cast<const float*>(xin.dev), n, dim, m, k, prep.p, nullptr, codes.p, 1,
                                 m, 1, nullptr, nullptr, nullptr, 0, ts.as<unsigned long long>(), units));
    VQB_CUDA(ctx, cudaMemcpyAsync(ts_out, ts.p, (size_t)units * 64, cudaMemcpyDeviceToHost, ctx->stream));
    VQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VQB_SUCCESS;
}
