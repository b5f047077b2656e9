// Fused in-batch (+ shared-pool) softmax cross-entropy on the 5th-gen tensor
// cores: tcgen05.mma with TMEM accumulators, operands staged by TMA
// (kernel 3 of the hot path, bf16 path; ce_f32.cu is the exact fp32 path).
//
// Replaces mm / div / masked_fill / cat / log_softmax / nll_loss at
// TwoTowerModel.py:95-140 of the reference.  The B x (B+H) logits live only in
// TMEM: a CTA owns 128 user rows, streams 128-row item tiles through a 4-stage
// TMA ring, one elected thread issues tcgen05.mma (M=128, N=128, K=16 x D/16)
// into one of four 128-column TMEM accumulators, and four epilogue warps read
// them back with tcgen05.ld (thread = row) and fold them into an online
// (max, sum) in the exp2 domain.
//
// False-negative mask without per-element id compares: rows are processed in
// item-id-sorted order (U and I permuted together while they are converted to
// bf16; the loss is invariant to that permutation), so the columns that collide
// with row p are the contiguous run [lo_p, hi_p) around the diagonal and only
// tiles intersecting that run take the masked path.
//
// Roofline: tensor pipe.  Algorithmic flops fwd = 2*B*(B+H)*D.
#include <cub/cub.cuh>

#include "tc_common.cuh"

namespace tt {

using namespace tt::tc;

constexpr int TC_BM = 128;
constexpr int TC_BN = 128;
constexpr int TC_STAGES = 4;
constexpr int TC_ACC = 4;
constexpr int TC_THREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
constexpr float LOG2E = 1.4426950408889634f;

// ---------------------------------------------------------------- prep kernels
// out[p, :] = bf16(in[perm ? perm[p] : p, :]); one warp per row; NaN detection for the flag word
__global__ void __launch_bounds__(256)
tc_convert_rows(const float *__restrict__ in, const int32_t *__restrict__ perm, int64_t n, int dim,
                __nv_bfloat16 *__restrict__ out, int *__restrict__ nan_flags, int nan_bit) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t p = warp; p < n; p += n_warps) {
        const int64_t src = perm ? perm[p] : p;
        bool nan = false;
        for (int c = lane * 4; c < dim; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(in + src * dim + c));
            nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
            __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
            uint2 raw;
            raw.x = *reinterpret_cast<uint32_t *>(&a);
            raw.y = *reinterpret_cast<uint32_t *>(&b);
            *reinterpret_cast<uint2 *>(out + p * dim + c) = raw;
        }
        if (__any_sync(0xffffffffu, nan) && lane == 0) atomicOr(nan_flags, nan_bit);
    }
}

__global__ void tc_iota_keys(const int64_t *__restrict__ ids, int64_t n, int64_t *__restrict__ keys,
                             int32_t *__restrict__ vals) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        keys[i] = ids[i];
        vals[i] = static_cast<int32_t>(i);
    }
}

// [lo, hi) = run of equal ids around p in the sorted order; always contains p (identity perm when ids == NULL)
__global__ void tc_runs(const int64_t *__restrict__ sorted_ids, int64_t n, int32_t *__restrict__ lo,
                        int32_t *__restrict__ hi, int32_t *__restrict__ perm_identity) {
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n;
         p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (sorted_ids == nullptr) {  // no ids: the run is the diagonal element alone
            lo[p] = static_cast<int32_t>(p); hi[p] = static_cast<int32_t>(p + 1);
            perm_identity[p] = static_cast<int32_t>(p);
            continue;
        }
        const int64_t key = sorted_ids[p];
        int64_t a = 0, b = p;  // lower_bound in [0, p]
        while (a < b) { const int64_t mid = (a + b) >> 1; if (sorted_ids[mid] < key) a = mid + 1; else b = mid; }
        lo[p] = static_cast<int32_t>(a);
        a = p; b = n;  // upper_bound in [p, n)
        while (a < b) { const int64_t mid = (a + b) >> 1; if (sorted_ids[mid] <= key) a = mid + 1; else b = mid; }
        hi[p] = static_cast<int32_t>(a);
    }
}

// ---------------------------------------------------------------- main kernel
struct CeTcParams {
    int64_t batch;       // rows of U / I
    int64_t pool_rows;   // rows of the shared pool (0 if none)
    int tiles_item, tiles_total, splits;
    float scale2;        // inv_temp * log2(e)
    const int32_t *lo, *hi;
    float *part_m, *part_s;  // [splits, batch] in the exp2 domain
};

template <int D>
__global__ void __launch_bounds__(TC_THREADS, 1)
ce_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_i,
                 const __grid_constant__ CUtensorMap map_p, const CeTcParams prm) {
    constexpr int KB = D / 64;                       // 64-column (128-byte) K blocks
    constexpr int TILE_BYTES = TC_BN * D * 2;        // one operand tile
    constexpr int KBLOCK_BYTES = TC_BN * 128;        // one 64-column box of 128 rows
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *u_tile = smem;
    uint8_t *y_tiles = smem + TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(y_tiles + TC_STAGES * TILE_BYTES);
    uint64_t *full = bars;                     // [TC_STAGES]
    uint64_t *empty = full + TC_STAGES;        // [TC_STAGES]
    uint64_t *tfull = empty + TC_STAGES;       // [TC_ACC]
    uint64_t *tempty = tfull + TC_ACC;         // [TC_ACC]
    uint64_t *ufull = tempty + TC_ACC;         // [1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ufull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_tile = blockIdx.x, split = blockIdx.y;
    const int tiles_per = (prm.tiles_total + prm.splits - 1) / prm.splits;
    const int t0 = split * tiles_per;
    const int t1 = min(prm.tiles_total, t0 + tiles_per);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_u);
        prefetch_tensormap(&map_i);
        if (prm.pool_rows > 0) prefetch_tensormap(&map_p);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < TC_ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
        mbar_init(ufull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TC_ACC * TC_BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_arrive_expect_tx(ufull, TILE_BYTES);
            for (int kb = 0; kb < KB; ++kb)
                tma_load_2d(u_tile + kb * KBLOCK_BYTES, &map_u, ufull, kb * 64, row_tile * TC_BM);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], TILE_BYTES);
                const bool item = t < prm.tiles_item;
                const CUtensorMap *m = item ? &map_i : &map_p;
                const int row0 = (item ? t : t - prm.tiles_item) * TC_BN;
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(y_tiles + stage * TILE_BYTES + kb * KBLOCK_BYTES, m, &full[stage], kb * 64, row0);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16_f32(TC_BM, TC_BN);
            mbar_wait(ufull, 0);
            tc_fence_after();
            const uint32_t u_addr = smem_u32(u_tile);
            int stage = 0, acc = 0;
            uint32_t phase = 0, aphase = 0;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&tempty[acc], aphase ^ 1);
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t y_addr = smem_u32(y_tiles + stage * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t off = (k / 4) * KBLOCK_BYTES + (k % 4) * 32;  // 16 bf16 = 32 B inside the 128-B swizzle atom
                    umma_f16(tmem_base + acc * TC_BN, smem_desc_k_sw128(u_addr + off), smem_desc_k_sw128(y_addr + off),
                             idesc, k > 0 ? 1u : 0u);
                }
                umma_commit(&empty[stage]);   // smem slot reusable once these MMAs retire
                umma_commit(&tfull[acc]);     // accumulator ready for the epilogue
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                if (++acc == TC_ACC) { acc = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: 4 warps, thread = row =====
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int r_in_tile = quarter * 32 + lane;
        const int64_t p = static_cast<int64_t>(row_tile) * TC_BM + r_in_tile;
        const bool row_ok = p < prm.batch;
        const int lo = row_ok ? prm.lo[p] : 0, hi = row_ok ? prm.hi[p] : 0;
        float m = -INFINITY, s = 0.f;
        int acc = 0;
        uint32_t aphase = 0;
        for (int t = t0; t < t1; ++t) {
            const bool item = t < prm.tiles_item;
            const int64_t col0 = static_cast<int64_t>(item ? t : t - prm.tiles_item) * TC_BN;
            const int64_t ncol = item ? prm.batch : prm.pool_rows;
            const bool special = (col0 + TC_BN > ncol) || (item && hi > col0 && lo < col0 + TC_BN);
            mbar_wait(&tfull[acc], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < TC_BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * TC_BN + c * 32, r);
                tmem_ld_wait();
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]) * prm.scale2;
                if (special) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int64_t col = col0 + c * 32 + j;
                        const bool dead = (col >= ncol) || (item && col >= lo && col < hi && col != p);
                        if (dead) x[j] = -INFINITY;
                    }
                }
                float cmax = x[0];
#pragma unroll
                for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, x[j]);
                if (cmax > m) {
                    s *= ex2_approx(m - cmax);  // m = -inf -> 0 * s(=0)
                    m = cmax;
                }
                if (m > -INFINITY) {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        a0 += ex2_approx(x[j] - m);
                        a1 += ex2_approx(x[j + 1] - m);
                        a2 += ex2_approx(x[j + 2] - m);
                        a3 += ex2_approx(x[j + 3] - m);
                    }
                    s += (a0 + a1) + (a2 + a3);
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
            if (++acc == TC_ACC) { acc = 0; aphase ^= 1; }
        }
        if (row_ok) {
            prm.part_m[static_cast<int64_t>(split) * prm.batch + p] = m;
            prm.part_s[static_cast<int64_t>(split) * prm.batch + p] = s;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<TC_ACC * TC_BN>(tmem_base);
    }
}

// one warp per (sorted) row: merge splits (exp2 domain), per-row hard negatives, positive logit
__global__ void __launch_bounds__(256)
ce_tc_finalize(const __nv_bfloat16 *__restrict__ ub, const __nv_bfloat16 *__restrict__ ib,
               const float *__restrict__ user32, const float *__restrict__ hn_rows, int n_rowneg,
               const int32_t *__restrict__ perm, int64_t B, int dim, float inv_temp, int splits,
               const float *__restrict__ part_m, const float *__restrict__ part_s, float *__restrict__ row_lse,
               float *__restrict__ row_pos, float *__restrict__ row_loss, float *__restrict__ lse_sorted,
               int *__restrict__ nan_flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const float scale2 = inv_temp * LOG2E;
    for (int64_t p = warp; p < B; p += n_warps) {
        float m = -INFINITY, s = 0.f;  // exp2 domain
        for (int k = 0; k < splits; ++k) {
            const float mk = part_m[static_cast<int64_t>(k) * B + p];
            const float sk = part_s[static_cast<int64_t>(k) * B + p];
            if (mk == -INFINITY) continue;
            const float mn = fmaxf(m, mk);
            s = ((m == -INFINITY) ? 0.f : s * exp2f(m - mn)) + sk * exp2f(mk - mn);
            m = mn;
        }
        const int64_t orig = perm[p];
        bool nan = false;
        for (int n = 0; n < n_rowneg; ++n) {
            const float *h = hn_rows + (orig * n_rowneg + n) * dim;
            float d = 0.f;
            for (int k = lane; k < dim; k += 32) {
                const float hv = h[k];
                nan |= (hv != hv);
                d = fmaf(user32[orig * dim + k], hv, d);
            }
            d = warp_sum(d) * scale2;
            const float mn = fmaxf(m, d);
            s = ((m == -INFINITY) ? 0.f : s * exp2f(m - mn)) + exp2f(d - mn);
            m = mn;
        }
        if (__any_sync(0xffffffffu, nan) && lane == 0) atomicOr(nan_flags, 4);
        float pos = 0.f;
        for (int k = lane; k < dim; k += 32)
            pos = fmaf(__bfloat162float(ub[p * dim + k]), __bfloat162float(ib[p * dim + k]), pos);
        pos = warp_sum(pos) * inv_temp;
        if (lane == 0) {
            const float lse = (m + log2f(s)) * (1.0f / LOG2E);
            row_lse[orig] = lse;
            row_pos[orig] = pos;
            row_loss[p] = lse - pos;
            lse_sorted[p] = lse;
        }
    }
}

__global__ void __launch_bounds__(1024) ce_tc_mean(const float *__restrict__ x, int64_t n, float *__restrict__ out) {
    __shared__ float sh[1024];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0] / static_cast<float>(n);
}

struct CeTcPlan {
    int row_tiles, tiles_item, tiles_pool, tiles_total, splits;
    size_t sort_bytes;
};

static CeTcPlan ce_tc_plan(int64_t batch, int64_t pool_rows) {
    CeTcPlan p;
    p.row_tiles = static_cast<int>((batch + TC_BM - 1) / TC_BM);
    p.tiles_item = static_cast<int>((batch + TC_BN - 1) / TC_BN);
    p.tiles_pool = static_cast<int>((pool_rows + TC_BN - 1) / TC_BN);
    p.tiles_total = p.tiles_item + p.tiles_pool;
    int want = (4 * sm_count() + p.row_tiles - 1) / p.row_tiles;  // ~4 CTAs per SM over the launch
    int cap = p.tiles_total / 8;                                  // keep >= 8 tiles per CTA
    if (cap < 1) cap = 1;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    p.splits = want;
    p.sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, p.sort_bytes, static_cast<int64_t *>(nullptr),
                                    static_cast<int64_t *>(nullptr), static_cast<int32_t *>(nullptr),
                                    static_cast<int32_t *>(nullptr), static_cast<int>(batch));
    return p;
}

struct CeTcWs {
    __nv_bfloat16 *ub, *ib, *pb;
    int64_t *keys_in, *keys_out;
    int32_t *vals_in, *perm, *lo, *hi;
    float *part_m, *part_s, *row_loss, *lse_sorted;
    void *cub_tmp;
    bool ok;
    size_t used;
};

static CeTcWs ce_tc_carve(void *workspace, size_t bytes, int64_t batch, int64_t pool_rows, int dim, const CeTcPlan &pl) {
    Workspace ws(workspace, bytes);
    CeTcWs w;
    w.ub = ws.take<__nv_bfloat16>(batch * dim);
    w.ib = ws.take<__nv_bfloat16>(batch * dim);
    w.pb = ws.take<__nv_bfloat16>((pool_rows > 0 ? pool_rows : 1) * dim);
    w.keys_in = ws.take<int64_t>(batch);
    w.keys_out = ws.take<int64_t>(batch);
    w.vals_in = ws.take<int32_t>(batch);
    w.perm = ws.take<int32_t>(batch);
    w.lo = ws.take<int32_t>(batch);
    w.hi = ws.take<int32_t>(batch);
    w.part_m = ws.take<float>(static_cast<size_t>(pl.splits) * batch);
    w.part_s = ws.take<float>(static_cast<size_t>(pl.splits) * batch);
    w.row_loss = ws.take<float>(batch);
    w.lse_sorted = ws.take<float>(batch);
    w.cub_tmp = ws.take<char>(pl.sort_bytes);
    w.ok = ws.ok();
    w.used = ws.off;
    return w;
}

template <int D>
static int launch_ce_tc_fwd(const CUtensorMap &mu, const CUtensorMap &mi, const CUtensorMap &mp, const CeTcParams &prm,
                            int row_tiles, cudaStream_t st) {
    constexpr size_t smem = 1024 + static_cast<size_t>(TC_BN) * D * 2 * (1 + TC_STAGES) + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ce_tc_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(ce_tc_fwd_kernel)");
        attr_set = true;
    }
    dim3 grid(row_tiles, prm.splits);
    ce_tc_fwd_kernel<D><<<grid, TC_THREADS, smem, st>>>(mu, mi, mp, prm);
    TT_LAUNCH_CHECK("ce_tc_fwd_kernel");
    return 0;
}

static inline unsigned tc_grid(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b);
}


// ================================================================== backward
// dU = c * sum_j (P - delta)[b, j] * Y_j ,  dY = c * sum_b (P - delta)[b, j] * U_b ,  c = grad_loss / (B * T),
// P = exp(z - lse) recomputed tile by tile (flash-attention style) -- two passes of ONE kernel:
//   TRANS = false : X = U rows (CTA owns 128 of them), W = [item ; pool] tiles, row statistic lse[x]
//   TRANS = true  : X = item / pool rows,               W = U tiles,           column statistic lse[w]
// Per W tile:  S = X W^T (SS tcgen05.mma into TMEM)  ->  8 softmax warps read S (tcgen05.ld, thread = row),
// G = bf16(P - delta) written back to TMEM (tcgen05.st, packed pairs)  ->  Out += G W (TS tcgen05.mma: A from
// TMEM, B = the SAME smem tile read through an MN-major descriptor).  S and G are double buffered; the tensor
// pipe executes in issue order, which is what makes the two extra "buffer free" barriers unnecessary (see the
// MMA warp).  Neither S nor G ever leaves the SM.
constexpr int BW_THREADS = 320;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..9: softmax (2 per lane quarter)
constexpr int BW_STAGES = 4;
constexpr uint32_t BW_TMEM_S = 0, BW_TMEM_G = 256, BW_TMEM_OUT = 384;

struct CeBwdParams {
    int64_t batch, pool_rows;
    int tiles_item, tiles_pool;
    int n_tiles;   // W tiles in this pass
    int splits;
    float scale2;  // inv_temp * log2(e)
    const int32_t *lo, *hi;
    const float *lse2p;  // [tiles_item * 128] lse * log2(e) in sorted order, +inf past batch
    float *part;         // [splits][part_rows][D] raw fp32 accumulators
    int64_t part_rows;
};

__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <int D, bool TRANS>
__global__ void __launch_bounds__(BW_THREADS, 1)
ce_tc_bwd_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_i,
                 const __grid_constant__ CUtensorMap map_p, const CeBwdParams prm) {
    constexpr int KB = D / 64;
    constexpr int TILE_BYTES = TC_BN * D * 2;
    constexpr int KBLOCK_BYTES = TC_BN * 128;
    constexpr int LSE_BYTES = TC_BN * 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *x_tile = smem;
    uint8_t *w_tiles = smem + TILE_BYTES;
    float *lse_s = reinterpret_cast<float *>(w_tiles + BW_STAGES * TILE_BYTES);  // [BW_STAGES][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(lse_s + BW_STAGES * TC_BN);
    uint64_t *full = bars;                  // [BW_STAGES]  TMA -> MMA (+ softmax for lse_s)
    uint64_t *empty = full + BW_STAGES;     // [BW_STAGES]  MMA -> TMA
    uint64_t *sfull = empty + BW_STAGES;    // [2]          S tile ready
    uint64_t *gfull = sfull + 2;            // [2]          G tile written (256 arrivals)
    uint64_t *xfull = gfull + 2;            // [1]
    uint64_t *ofull = xfull + 1;            // [1]          all Out MMAs retired
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ofull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, split = blockIdx.y;
    const int tiles_per = (prm.n_tiles + prm.splits - 1) / prm.splits;
    const int t0 = split * tiles_per;
    const int t1 = min(prm.n_tiles, t0 + tiles_per);
    const int n_local = max(t1 - t0, 0);
    const bool x_is_item = !TRANS || m_tile < prm.tiles_item;   // TRANS: X rows are item rows (else pool rows)
    const int x_row0 = (TRANS && !x_is_item ? m_tile - prm.tiles_item : m_tile) * TC_BM;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_u);
        prefetch_tensormap(&map_i);
        if (prm.pool_rows > 0) prefetch_tensormap(&map_p);
        for (int s = 0; s < BW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&sfull[a], 1); mbar_init(&gfull[a], 256); }
        mbar_init(xfull, 1);
        mbar_init(ofull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && n_local > 0) {
            const CUtensorMap *mx = !TRANS ? &map_u : (x_is_item ? &map_i : &map_p);
            mbar_arrive_expect_tx(xfull, TILE_BYTES);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(x_tile + kb * KBLOCK_BYTES, mx, xfull, kb * 64, x_row0);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], TILE_BYTES + (TRANS ? LSE_BYTES : 0));
                const CUtensorMap *m;
                int row0;
                if (TRANS) { m = &map_u; row0 = t * TC_BN; }
                else if (t < prm.tiles_item) { m = &map_i; row0 = t * TC_BN; }
                else { m = &map_p; row0 = (t - prm.tiles_item) * TC_BN; }
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(w_tiles + stage * TILE_BYTES + kb * KBLOCK_BYTES, m, &full[stage], kb * 64, row0);
                if (TRANS) bulk_load_1d(lse_s + stage * TC_BN, prm.lse2p + static_cast<int64_t>(t) * TC_BN, LSE_BYTES, &full[stage]);
                if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && n_local > 0) {
            constexpr uint32_t idesc_s = idesc_bf16_f32(TC_BM, TC_BN, 0, 0);
            constexpr uint32_t idesc_o = idesc_bf16_f32(TC_BM, D, 0, 1);   // B operand MN-major
            mbar_wait(xfull, 0);
            tc_fence_after();
            const uint32_t x_addr = smem_u32(x_tile);
            auto issue_s = [&](int i) {
                const int stage = i % BW_STAGES;
                mbar_wait(&full[stage], (i / BW_STAGES) & 1);
                tc_fence_after();
                const uint32_t w_addr = smem_u32(w_tiles + stage * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t off = (k / 4) * KBLOCK_BYTES + (k % 4) * 32;
                    umma_f16(tmem_base + BW_TMEM_S + (i & 1) * TC_BN, smem_desc_k_sw128(x_addr + off),
                             smem_desc_k_sw128(w_addr + off), idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit(&sfull[i & 1]);
            };
            issue_s(0);
            for (int i = 0; i < n_local; ++i) {
                // S(i+1) goes into the buffer softmax(i-1) read; it finished before gfull(i-1) completed, which
                // this thread waited for in the previous iteration.
                if (i + 1 < n_local) issue_s(i + 1);
                mbar_wait(&gfull[i & 1], (i >> 1) & 1);
                tc_fence_after();
                const int stage = i % BW_STAGES;
                const uint32_t w_addr = smem_u32(w_tiles + stage * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BN / 16; ++k)
                    umma_f16_ts(tmem_base + BW_TMEM_OUT, tmem_base + BW_TMEM_G + (i & 1) * (TC_BN / 2) + k * 8,
                                smem_desc_mn_sw128(w_addr + k * 2048, KBLOCK_BYTES, 1024), idesc_o,
                                (i > 0 || k > 0) ? 1u : 0u);
                // retiring Out(i) frees the smem stage; it also precedes S(i+2) on the in-order tensor pipe, so
                // sfull(i+2) implies G(i) has been consumed and may be overwritten
                umma_commit(&empty[stage]);
            }
            umma_commit(ofull);
        }
    } else {
        // ===== softmax warps: thread = X row, each warp a 32-lane quarter x 64-column half =====
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r_in_tile = quarter * 32 + lane;
        const int64_t p = static_cast<int64_t>(x_row0) + r_in_tile;       // row index inside its matrix
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        int lo = 0, hi = 0;
        float row_stat = 0.f;
        if (x_is_item && p < prm.batch) { lo = prm.lo[p]; hi = prm.hi[p]; }
        if (!TRANS) row_stat = (p < prm.batch) ? prm.lse2p[p] : INFINITY;
        for (int i = 0; i < n_local; ++i) {
            const int t = t0 + i;
            const bool w_is_item = TRANS || t < prm.tiles_item;
            const int64_t col0 = static_cast<int64_t>(w_is_item ? t : t - prm.tiles_item) * TC_BN;
            const int64_t ncol = w_is_item ? prm.batch : prm.pool_rows;
            // columns that need the per-element path: past the end (zero-filled rows of W would count as logit 0;
            // TRANS handles them through lse = +inf) and the collision run / diagonal of this row
            const bool special = (!TRANS && col0 + TC_BN > ncol) ||
                                 (x_is_item && w_is_item && hi > col0 && lo < col0 + TC_BN);
            if (TRANS) mbar_wait(&full[i % BW_STAGES], (i / BW_STAGES) & 1);   // lse_s of this stage has landed
            mbar_wait(&sfull[i & 1], (i >> 1) & 1);
            tc_fence_after();
            uint32_t ra[32], rb[32];
            const uint32_t s_addr = lane_addr + BW_TMEM_S + (i & 1) * TC_BN + half * 64;
            tmem_ld_32x32(s_addr, ra);
            tmem_ld_32x32(s_addr + 32, rb);
            tmem_ld_wait();
            const float *ls = lse_s + (i % BW_STAGES) * TC_BN + half * 64;
            uint32_t g[32];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float e[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 st = make_float4(row_stat, row_stat, row_stat, row_stat);
                    if (TRANS) st = *reinterpret_cast<const float4 *>(ls + c * 32 + j);
                    const uint32_t *r = c == 0 ? ra : rb;
                    e[j] = ex2_approx(fmaf(__uint_as_float(r[j]), prm.scale2, -st.x));
                    e[j + 1] = ex2_approx(fmaf(__uint_as_float(r[j + 1]), prm.scale2, -st.y));
                    e[j + 2] = ex2_approx(fmaf(__uint_as_float(r[j + 2]), prm.scale2, -st.z));
                    e[j + 3] = ex2_approx(fmaf(__uint_as_float(r[j + 3]), prm.scale2, -st.w));
                }
                if (special) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int64_t col = col0 + half * 64 + c * 32 + j;
                        if (col >= ncol) e[j] = 0.f;
                        else if (x_is_item && w_is_item && col >= lo && col < hi) e[j] = (col == p) ? e[j] - 1.0f : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) g[c * 16 + j] = pack_bf16x2(e[2 * j], e[2 * j + 1]);
            }
            tmem_st_32x32(lane_addr + BW_TMEM_G + (i & 1) * (TC_BN / 2) + half * 32, g);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&gfull[i & 1]);
        }
        // final: raw accumulators -> this split's partial (scaled / un-permuted by ce_tc_reduce_rows)
        if (n_local > 0) {
            mbar_wait(ofull, 0);
            tc_fence_after();
        }
        float *dst = prm.part + (static_cast<int64_t>(split) * prm.part_rows + static_cast<int64_t>(m_tile) * TC_BM + r_in_tile) * D +
                     half * (D / 2);
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
            uint32_t r[32];
            if (n_local > 0) {
                tmem_ld_32x32(lane_addr + BW_TMEM_OUT + half * (D / 2) + c * 32, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4 *>(dst + c * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<512>(tmem_base);
    }
}

// lse2p[p] = lse[perm[p]] * log2(e) for p < B, +inf for the padding up to a multiple of 128
__global__ void ce_tc_bwd_prep(const float *__restrict__ row_lse, const int32_t *__restrict__ perm, int64_t B,
                               int64_t padded, float *__restrict__ lse2p) {
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < padded;
         p += static_cast<int64_t>(gridDim.x) * blockDim.x)
        lse2p[p] = (p < B) ? row_lse[perm[p]] * LOG2E : INFINITY;
}

// per-row hard negatives (fp32 SIMT, N is ~10): dHN[b,n,:] = g U_b / T ; extra[b,:] = sum_n g HN[b,n,:]
__global__ void __launch_bounds__(256)
ce_tc_bwd_hn_rows(const float *__restrict__ user, const float *__restrict__ hn_rows, int n_rowneg, int64_t B, int dim,
                  float inv_temp, const float *__restrict__ row_lse, const float *__restrict__ grad_loss,
                  float *__restrict__ d_hn, float *__restrict__ extra) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const float gscale = (*grad_loss) / static_cast<float>(B);
    for (int64_t b = warp; b < B; b += n_warps) {
        const float lse = row_lse[b];
        for (int k = lane; k < dim; k += 32) extra[b * dim + k] = 0.f;
        for (int n = 0; n < n_rowneg; ++n) {
            const float *h = hn_rows + (b * n_rowneg + n) * dim;
            float d = 0.f;
            for (int k = lane; k < dim; k += 32) d = fmaf(user[b * dim + k], h[k], d);
            d = warp_sum(d) * inv_temp;
            const float g = gscale * expf(d - lse);
            for (int k = lane; k < dim; k += 32) {
                if (d_hn) d_hn[(b * n_rowneg + n) * dim + k] = g * inv_temp * user[b * dim + k];
                extra[b * dim + k] += g * h[k];
            }
        }
    }
}

// out[dst(p), :] = (*grad_loss * scale) * sum_s part[s][row_begin + p, :] (+ extra_scale * extra[dst(p), :]);
// dst(p) = perm ? perm[p] : p.  Fixed summation order.
__global__ void __launch_bounds__(256)
ce_tc_reduce_rows(const float *__restrict__ part, int splits, int64_t part_rows, int64_t row_begin, int64_t n_rows,
                  int dim, const int32_t *__restrict__ perm, const float *__restrict__ grad_loss, float scale,
                  const float *__restrict__ extra, float extra_scale, float *__restrict__ out) {
    const int vpr = dim / 4;
    const float c = (*grad_loss) * scale;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_rows * vpr;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t p = i / vpr;
        const int c4 = static_cast<int>(i - p * vpr) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < splits; ++s) {
            const float4 v = *reinterpret_cast<const float4 *>(part + (static_cast<int64_t>(s) * part_rows + row_begin + p) * dim + c4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        const int64_t dst = perm ? perm[p] : p;
        float4 o = make_float4(a.x * c, a.y * c, a.z * c, a.w * c);
        if (extra) {
            const float4 x = *reinterpret_cast<const float4 *>(extra + dst * dim + c4);
            o.x = fmaf(x.x, extra_scale, o.x); o.y = fmaf(x.y, extra_scale, o.y);
            o.z = fmaf(x.z, extra_scale, o.z); o.w = fmaf(x.w, extra_scale, o.w);
        }
        *reinterpret_cast<float4 *>(out + dst * dim + c4) = o;
    }
}

static int bwd_splits(int m_tiles, int n_tiles) {
    int want = (4 * sm_count() + m_tiles - 1) / m_tiles;
    int cap = n_tiles / 8;
    if (cap < 1) cap = 1;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return want;
}

struct CeBwdWs {
    float *lse2p, *extra, *part_u, *part_y;
    int splits_u, splits_y;
    int64_t rows_u, rows_y;
    bool ok;
    size_t used;
};

static CeBwdWs ce_bwd_carve(void *workspace, size_t bytes, int64_t batch, int64_t pool_rows, int n_rowneg, int dim,
                            const CeTcPlan &pl) {
    Workspace ws(workspace, bytes);
    CeBwdWs w;
    w.splits_u = bwd_splits(pl.tiles_item, pl.tiles_total);
    w.splits_y = bwd_splits(pl.tiles_total, pl.tiles_item);
    w.rows_u = static_cast<int64_t>(pl.tiles_item) * TC_BM;
    w.rows_y = static_cast<int64_t>(pl.tiles_total) * TC_BM;
    w.lse2p = ws.take<float>(w.rows_u);
    w.extra = ws.take<float>(n_rowneg > 0 ? batch * dim : 1);
    w.part_u = ws.take<float>(static_cast<size_t>(w.splits_u) * w.rows_u * dim);
    w.part_y = ws.take<float>(static_cast<size_t>(w.splits_y) * w.rows_y * dim);
    (void)pool_rows;
    w.ok = ws.ok();
    w.used = ws.off;
    return w;
}

template <int D, bool TRANS>
static int launch_ce_tc_bwd(const CUtensorMap &mu, const CUtensorMap &mi, const CUtensorMap &mp, const CeBwdParams &prm,
                            int m_tiles, cudaStream_t st) {
    constexpr size_t smem = 1024 + static_cast<size_t>(TC_BN) * D * 2 * (1 + BW_STAGES) + BW_STAGES * TC_BN * 4 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ce_tc_bwd_kernel<D, TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(ce_tc_bwd_kernel)");
        attr_set = true;
    }
    dim3 grid(m_tiles, prm.splits);
    ce_tc_bwd_kernel<D, TRANS><<<grid, BW_THREADS, smem, st>>>(mu, mi, mp, prm);
    TT_LAUNCH_CHECK("ce_tc_bwd_kernel");
    return 0;
}

}  // namespace tt

extern "C" int tt_ce_tc_workspace(int64_t batch, int64_t pool, int n_rowneg, int dim, size_t *bytes_host) {
    using namespace tt;
    TT_CHECK_ARG(bytes_host && batch > 0 && pool >= 0 && n_rowneg >= 0 && dim > 0, "bad size");
    const CeTcPlan pl = ce_tc_plan(batch, pool);
    const CeTcWs w = ce_tc_carve(nullptr, ~size_t(0), batch, pool, dim, pl);
    *bytes_host = w.used + 1024;
    return 0;
}

extern "C" int tt_ce_fwd_tc(const float *user, const float *item, const int64_t *item_ids, const float *hn_rows,
                            int n_rowneg, const float *pool, int64_t pool_rows, int64_t batch, int dim, float inv_temp,
                            float *loss, float *row_lse, float *row_pos, int *nan_flags, void *workspace,
                            size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(user && item && loss && row_lse && row_pos && nan_flags && workspace, "null pointer");
    TT_CHECK_ARG((hn_rows != nullptr) == (n_rowneg > 0), "hn_rows / n_rowneg mismatch");
    TT_CHECK_ARG((pool != nullptr) == (pool_rows > 0), "pool / pool_rows mismatch");
    TT_CHECK_ARG(batch > 0 && batch < (int64_t(1) << 31), "bad batch");
    if (dim != 64 && dim != 128) { set_error("tensor-core CE path supports dim 64 or 128 (got %d)", dim); return TT_E_UNSUPPORTED; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CeTcPlan pl = ce_tc_plan(batch, pool_rows);
    CeTcWs w = ce_tc_carve(workspace, workspace_bytes, batch, pool_rows, dim, pl);
    if (!w.ok) { set_error("ce_tc workspace too small: need %zu have %zu", w.used, workspace_bytes); return TT_E_WORKSPACE; }
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) { set_error("ce_tc workspace must be 256-byte aligned"); return TT_E_BADARG; }

    // 1. item-id sorted order -> perm, collision runs
    if (item_ids != nullptr) {
        tc_iota_keys<<<tc_grid(batch, 256), 256, 0, st>>>(item_ids, batch, w.keys_in, w.vals_in);
        TT_LAUNCH_CHECK("tc_iota_keys");
        size_t tmp = pl.sort_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_in, w.keys_out, w.vals_in, w.perm,
                                                        static_cast<int>(batch), 0, 64, st);
        if (e != cudaSuccess) return cuda_status(e, "cub SortPairs(ce_tc)");
        tc_runs<<<tc_grid(batch, 256), 256, 0, st>>>(w.keys_out, batch, w.lo, w.hi, nullptr);
    } else {
        tc_runs<<<tc_grid(batch, 256), 256, 0, st>>>(nullptr, batch, w.lo, w.hi, w.perm);
    }
    TT_LAUNCH_CHECK("tc_runs");
    // 2. bf16 operands in sorted order
    tc_convert_rows<<<tc_grid(batch * 32, 256), 256, 0, st>>>(user, w.perm, batch, dim, w.ub, nan_flags, 1);
    tc_convert_rows<<<tc_grid(batch * 32, 256), 256, 0, st>>>(item, w.perm, batch, dim, w.ib, nan_flags, 2);
    if (pool) tc_convert_rows<<<tc_grid(pool_rows * 32, 256), 256, 0, st>>>(pool, nullptr, pool_rows, dim, w.pb, nan_flags, 4);
    TT_LAUNCH_CHECK("tc_convert_rows");
    // 3. tensor maps + main kernel
    CUtensorMap mu, mi, mp;
    int rc;
    if ((rc = make_tmap_bf16_rows(&mu, w.ub, batch, dim, TC_BM))) return rc;
    if ((rc = make_tmap_bf16_rows(&mi, w.ib, batch, dim, TC_BN))) return rc;
    if ((rc = make_tmap_bf16_rows(&mp, pool ? w.pb : w.ib, pool ? pool_rows : batch, dim, TC_BN))) return rc;
    CeTcParams prm;
    prm.batch = batch; prm.pool_rows = pool_rows;
    prm.tiles_item = pl.tiles_item; prm.tiles_total = pl.tiles_total; prm.splits = pl.splits;
    prm.scale2 = inv_temp * LOG2E;
    prm.lo = w.lo; prm.hi = w.hi; prm.part_m = w.part_m; prm.part_s = w.part_s;
    rc = (dim == 128) ? launch_ce_tc_fwd<128>(mu, mi, mp, prm, pl.row_tiles, st)
                      : launch_ce_tc_fwd<64>(mu, mi, mp, prm, pl.row_tiles, st);
    if (rc) return rc;
    // 4. finalize
    ce_tc_finalize<<<tc_grid(batch * 32, 256), 256, 0, st>>>(w.ub, w.ib, user, hn_rows, n_rowneg, w.perm, batch, dim,
                                                            inv_temp, pl.splits, w.part_m, w.part_s, row_lse, row_pos,
                                                            w.row_loss, w.lse_sorted, nan_flags);
    TT_LAUNCH_CHECK("ce_tc_finalize");
    ce_tc_mean<<<1, 1024, 0, st>>>(w.row_loss, batch, loss);
    TT_LAUNCH_CHECK("ce_tc_mean");
    return 0;
}

extern "C" int tt_ce_bwd_tc_workspace(int64_t batch, int64_t pool, int n_rowneg, int dim, size_t *bytes_host) {
    using namespace tt;
    TT_CHECK_ARG(bytes_host && batch > 0 && pool >= 0 && n_rowneg >= 0 && dim > 0, "bad size");
    const CeTcPlan pl = ce_tc_plan(batch, pool);
    const CeBwdWs w = ce_bwd_carve(nullptr, ~size_t(0), batch, pool, n_rowneg, dim, pl);
    *bytes_host = w.used + 1024;
    return 0;
}

extern "C" int tt_ce_bwd_tc(const float *user, const float *hn_rows, int n_rowneg, int64_t pool_rows, int64_t batch,
                            int dim, float inv_temp, const float *row_lse, const float *grad_loss, float *d_user,
                            float *d_item, float *d_hn_rows, float *d_pool, void *fwd_workspace,
                            size_t fwd_workspace_bytes, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(user && row_lse && grad_loss && d_user && d_item && fwd_workspace && workspace, "null pointer");
    TT_CHECK_ARG((hn_rows != nullptr) == (n_rowneg > 0), "hn_rows / n_rowneg mismatch");
    TT_CHECK_ARG(pool_rows == 0 || d_pool != nullptr, "d_pool required with a pool");
    TT_CHECK_ARG(batch > 0 && batch < (int64_t(1) << 31), "bad batch");
    if (dim != 64 && dim != 128) { set_error("tensor-core CE path supports dim 64 or 128 (got %d)", dim); return TT_E_UNSUPPORTED; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CeTcPlan pl = ce_tc_plan(batch, pool_rows);
    const CeTcWs f = ce_tc_carve(fwd_workspace, fwd_workspace_bytes, batch, pool_rows, dim, pl);
    if (!f.ok) { set_error("ce_tc forward workspace too small: need %zu have %zu", f.used, fwd_workspace_bytes); return TT_E_WORKSPACE; }
    const CeBwdWs w = ce_bwd_carve(workspace, workspace_bytes, batch, pool_rows, n_rowneg, dim, pl);
    if (!w.ok) { set_error("ce_tc backward workspace too small: need %zu have %zu", w.used, workspace_bytes); return TT_E_WORKSPACE; }
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) { set_error("ce_tc workspace must be 256-byte aligned"); return TT_E_BADARG; }

    ce_tc_bwd_prep<<<tc_grid(w.rows_u, 256), 256, 0, st>>>(row_lse, f.perm, batch, w.rows_u, w.lse2p);
    TT_LAUNCH_CHECK("ce_tc_bwd_prep");
    CUtensorMap mu, mi, mp;
    int rc;
    if ((rc = make_tmap_bf16_rows(&mu, f.ub, batch, dim, TC_BM))) return rc;
    if ((rc = make_tmap_bf16_rows(&mi, f.ib, batch, dim, TC_BN))) return rc;
    if ((rc = make_tmap_bf16_rows(&mp, pool_rows ? f.pb : f.ib, pool_rows ? pool_rows : batch, dim, TC_BN))) return rc;
    CeBwdParams prm;
    prm.batch = batch; prm.pool_rows = pool_rows;
    prm.tiles_item = pl.tiles_item; prm.tiles_pool = pl.tiles_pool;
    prm.scale2 = inv_temp * LOG2E;
    prm.lo = f.lo; prm.hi = f.hi; prm.lse2p = w.lse2p;
    // pass 1: dU
    prm.n_tiles = pl.tiles_total; prm.splits = w.splits_u; prm.part = w.part_u; prm.part_rows = w.rows_u;
    rc = (dim == 128) ? launch_ce_tc_bwd<128, false>(mu, mi, mp, prm, pl.tiles_item, st)
                      : launch_ce_tc_bwd<64, false>(mu, mi, mp, prm, pl.tiles_item, st);
    if (rc) return rc;
    // pass 2: dI, dPool
    prm.n_tiles = pl.tiles_item; prm.splits = w.splits_y; prm.part = w.part_y; prm.part_rows = w.rows_y;
    rc = (dim == 128) ? launch_ce_tc_bwd<128, true>(mu, mi, mp, prm, pl.tiles_total, st)
                      : launch_ce_tc_bwd<64, true>(mu, mi, mp, prm, pl.tiles_total, st);
    if (rc) return rc;
    if (hn_rows) {
        ce_tc_bwd_hn_rows<<<tc_grid(batch * 32, 256), 256, 0, st>>>(user, hn_rows, n_rowneg, batch, dim, inv_temp, row_lse,
                                                                   grad_loss, d_hn_rows, w.extra);
        TT_LAUNCH_CHECK("ce_tc_bwd_hn_rows");
    }
    const float scale = inv_temp / static_cast<float>(batch);
    const int64_t vec = dim / 4;
    ce_tc_reduce_rows<<<tc_grid(batch * vec, 256), 256, 0, st>>>(w.part_u, w.splits_u, w.rows_u, 0, batch, dim, f.perm,
                                                                grad_loss, scale, hn_rows ? w.extra : nullptr, inv_temp, d_user);
    ce_tc_reduce_rows<<<tc_grid(batch * vec, 256), 256, 0, st>>>(w.part_y, w.splits_y, w.rows_y, 0, batch, dim, f.perm,
                                                                grad_loss, scale, nullptr, 0.f, d_item);
    if (pool_rows > 0)
        ce_tc_reduce_rows<<<tc_grid(pool_rows * vec, 256), 256, 0, st>>>(w.part_y, w.splits_y, w.rows_y,
                                                                        static_cast<int64_t>(pl.tiles_item) * TC_BM, pool_rows,
                                                                        dim, nullptr, grad_loss, scale, nullptr, 0.f, d_pool);
    TT_LAUNCH_CHECK("ce_tc_reduce_rows");
    return 0;
}
