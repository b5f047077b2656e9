// sm_100a tensor-core plumbing shared by the tcgen05 kernels: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05.mma / commit / ld / alloc wrappers, shared
// memory and instruction descriptors, host-side tensor-map encoding.
// Inline PTX only -- no CUTLASS dependency.
#pragma once

#include <cuda.h>  // CUtensorMap types only; the driver entry point is fetched at run time

#include <type_traits>

#include "common.cuh"

namespace tt {
namespace tc {

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Device-side error word: a wait that never completes (a pipeline bug) traps instead of hanging the GPU.
__device__ int g_tc_timeout = 0;

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    atomicExch(&g_tc_timeout, 1);
    __trap();
}

// One lane of a fully converged warp (elect.sync).  tcgen05.mma / TMA take their operands from uniform
// registers: issued under `if (lane == 0)` ptxas cannot prove uniformity and wraps EVERY instruction in an
// elect/branch "waterfall" loop (~100 cycles per MMA, measured); issued under elect_one_sync() inside
// warp-uniform control flow they are single predicated instructions.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred)::"memory");
    return pred != 0;
}

// one non-blocking-ish probe of a phase (try_wait may suspend for a short, bounded time)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 2-D tiled load: coordinates (c0 = inner / contiguous dim, c1 = row)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// same, multicast to the CTAs of the cluster named in cta_mask: the box lands at the same CTA-relative offset in each
// of them and completes the transaction bytes on the mbarrier at the same offset in each
__device__ __forceinline__ void tma_load_2d_multicast(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1,
                                                      uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ tcgen05
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; kind::f16 (bf16/fp16 inputs, fp32 accumulate); one thread issues
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// immediate-predicate forms for the issue loops (one thread issues every MMA of a CTA: keep the loop short)
__device__ __forceinline__ void umma_f16_first(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_f16_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts_first(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts_acc(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc) : "memory");
}

// arrive on an mbarrier once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// same, arriving on the mbarrier at this CTA-relative offset in every CTA of the cluster named in cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint64_t *bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (lane_base + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// store 32 lanes x 32 consecutive 32-bit columns: thread i of the warp writes lane (lane_base + i), columns [col, col+32)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]; A = 128 lanes x K packed bf16 pairs (K/2 32-bit columns), kind::f16
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA pipe (Cody-Waite split + degree-4 minimax polynomial of 2^f on [0,1), relative error < 3e-6):
// the MUFU does 4 ex2 per clock per SM sub-partition and is THE bound of the softmax epilogues, the FMA pipe is mostly
// idle there, so a fixed fraction of the exponentials is computed here instead (same idea as FlashAttention-4).
// Valid for x >= -126 after the clamp; x = -inf (masked logits) gives 2^-126 ~ 1e-38 instead of 0.
__device__ __forceinline__ float ex2_poly(float x) {
    const float xc = fmaxf(x, -126.0f);
    float t;
    asm("add.rm.f32 %0, %1, 0f4B400000;" : "=f"(t) : "f"(xc));      // + 1.5 * 2^23, rounded down: floor(xc) in the low mantissa bits
    const float fl = t - 12582912.0f;
    const float f = xc - fl;                                       // in [0, 1)
    float p = 1.352060e-2f;                                        // weighted least-squares fit, max rel. error 2.7e-6
    p = fmaf(p, f, 5.203743e-2f);
    p = fmaf(p, f, 2.4142749e-1f);
    p = fmaf(p, f, 6.9300662e-1f);
    p = fmaf(p, f, 1.00000252f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));   // p * 2^floor(xc)
}

// compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(integral_constant<N-1>)
template <int N, int I = 0, typename F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<N, I + 1>(f);
    }
}

// element j of an unrolled row: which unit computes its exponential.  Bit (j & 7) of TT_POLY_MASK set = FMA pipe.
// Measured on B200 (C4, round 1): mask 0x52 (3 of 8 on the FMA pipe) is parity-clean but NOT faster (fwd 1.69 -> 1.81 ms,
// fwd+bwd 6.44 -> 6.84 ms): the epilogues are bound by their serial per-tile chain (barrier wake-ups, tcgen05.ld/st
// round trips, issue slots), not by raw MUFU throughput (ncu XU pipe 46-63 %).  Default: everything on the MUFU.
#ifndef TT_POLY_MASK
#define TT_POLY_MASK 0x00
#endif
template <int J>
__device__ __forceinline__ float ex2_mixed(float x) {
#ifdef TT_EXPERIMENT_NO_EXP   // developer experiment only (wrong results): how much do the exponentials cost the tile?
    return x * 0.5f;
#endif
    if constexpr (((TT_POLY_MASK >> (J & 7)) & 1) != 0) return ex2_poly(x);
    else return ex2_approx(x);
}

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, bf16: rows of 64 elements (128 B), groups of
// 8 rows 1024 B apart.  (bit layout: cute/arch/mma_sm100_desc.hpp UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);       // start address      [0,14)
    d |= static_cast<uint64_t>(1) << 16;                          // leading byte offset (unused for SW128 K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // stride byte offset  [32,46): 8 rows * 128 B
    d |= static_cast<uint64_t>(1) << 46;                          // descriptor version  [46,48) = 1 on sm_100
    d |= static_cast<uint64_t>(2) << 61;                          // layout type         [61,64) = SWIZZLE_128B
    return d;
}

// MN-major operand, SWIZZLE_128B, bf16 (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units,
// cute/atom/mma_traits_sm100.hpp): 64 MN-contiguous elements per 128-byte row, one row per k, 8-row swizzle
// atoms `sbo_bytes` apart along K, next block of 64 MN elements `lbo_bytes` away.  A [rows=k][64 cols] box
// that TMA wrote with SWIZZLE_128B is exactly this layout with sbo = 1024.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major, dense.
// (bit layout: UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int m, int n, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4)                                   // D format: F32
           | (1u << 7)                                 // A format: BF16
           | (1u << 10)                                // B format: BF16
           | (static_cast<uint32_t>(a_mn_major) << 15) // A major
           | (static_cast<uint32_t>(b_mn_major) << 16) // B major
           | (static_cast<uint32_t>(n >> 3) << 17)     // N / 8
           | (static_cast<uint32_t>(m >> 4) << 24);    // M / 16
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major bf16 [rows, dim] -> boxes of {64 columns (128 B), box_rows rows}, 128B swizzle, zero fill out of range
inline int make_tmap_bf16_rows(CUtensorMap *map, const void *base, int64_t rows, int dim, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point not available"); return TT_E_DEVICE; }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r)); return TT_E_BADARG; }
    return 0;
}

// ---------------------------------------------------------------- stream-K schedule
// Tiles g = m * n_tiles + n, CTA k owns [k * per_cta, (k + 1) * per_cta).  A row tile m is cut into
// segments seg = k - first_cta(m); every segment leaves one partial (two for FWD: one per softmax group).
struct Sched {
    int m_tiles, n_tiles;
    int64_t total, per_cta;
    int grid, max_seg;
};

inline Sched make_sched(int m_tiles, int n_tiles, int max_segments = 0) {
    Sched s;
    s.m_tiles = m_tiles; s.n_tiles = n_tiles;
    s.total = static_cast<int64_t>(m_tiles) * n_tiles;
    int64_t g = (s.total + 1) / 2;               // at least 2 tiles per CTA
    if (g > sm_count()) g = sm_count();
    if (g < 1) g = 1;
    s.per_cta = (s.total + g - 1) / g;
    // optional cap on how many CTAs share one row tile (each segment leaves a partial the epilogue must merge)
    if (max_segments > 0 && s.per_cta * max_segments < n_tiles) s.per_cta = (n_tiles + max_segments - 1) / max_segments;
    s.grid = static_cast<int>((s.total + s.per_cta - 1) / s.per_cta);
    s.max_seg = static_cast<int>((n_tiles + s.per_cta - 2) / s.per_cta) + 1;
    return s;
}

__host__ __device__ __forceinline__ int sched_first_cta(int m, int n_tiles, int64_t per_cta) {
    return static_cast<int>((static_cast<int64_t>(m) * n_tiles) / per_cta);
}
__host__ __device__ __forceinline__ int sched_last_cta(int m, int n_tiles, int64_t per_cta) {
    return static_cast<int>((static_cast<int64_t>(m + 1) * n_tiles - 1) / per_cta);
}

// position of a role inside its CTA's tile range
struct Cursor {
    int i;        // local tile index
    int m, n;     // row tile, W tile
    int r;        // local row counter (0 for the first row tile this CTA touches)
    int n_tiles;
    __device__ __forceinline__ void init(int64_t g0, int nt) {
        n_tiles = nt; i = 0; r = 0;
        m = static_cast<int>(g0 / nt);
        n = static_cast<int>(g0 - static_cast<int64_t>(m) * nt);
    }
    __device__ __forceinline__ void next() {
        ++i;
        if (++n == n_tiles) { n = 0; ++m; ++r; }
    }
};


}  // namespace tc
}  // namespace tt
