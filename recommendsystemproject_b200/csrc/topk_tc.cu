// Corpus scoring + top-K for retrieval evaluation on the 5th-gen tensor cores
// (kernel 4 of the hot path, bf16 scoring; topk_f32.cu is the fp32 SIMT path).
//
// Replaces `scores = U @ E^T` -> per-user -inf masking -> torch.topk at
// training_utils.py:220-258 of the reference.  The [Bq, Nc] score matrix
// lives only in TMEM, 128 x 256 fp32 at a time.
//
// Exactness (bit-exact row indices under the stated tie-break) with bf16 scoring:
//   stage 1 is a FILTER.  Every (query, CTA segment, column half) keeps a candidate list: a score enters when it
//     beats the list's threshold tau; a full list is pruned to its best ~K' = K + margin and tau rises to the
//     score of the last survivor, so "approx score <= tau" holds for everything that was ever left out.
//   stage 2 keeps the K' best candidates by approximate score, re-scores them in fp64 from the fp32 inputs and
//     sorts by (score desc, row asc).  The result is PROVABLY the exact top-K whenever
//         exact_score[K-th] > tau_max + eps_q,   eps_q = (1.02 * 2^-8 + 2^-16) * |q| * max|e|
//     (bf16 rounding of both operands is at most 2^-9 relative per element, fp32 accumulation ~1e-5): nothing left
//     out can reach the K-th exact score.  Queries that fail the test are flagged; the caller re-runs them on the
//     exact fp32 path (rare: the margin puts tau_max ~10 eps below the K-th score on the 10M-item config).
//
// Main kernel = the persistent stream-K / warp-specialised shape of ce_tc.cu:
//   warp 0 TMA (query tile + 3-stage ring of 256-row corpus tiles, SWIZZLE_128B), warp 1 one elected lane issues
//   tcgen05.mma M=128 N=256 K=16 (N=256 because a 128-wide MMA is issue-bound at ~100 cycles on this part:
//   tools/tc_selftest), warps 2..9 = two groups (column halves) x four lane quarters, thread = query row:
//   tcgen05.ld 128 scores, release the TMEM buffer, max-reduce 32 at a time against tau (one FMNMX per score),
//   rare append path, warp-cooperative prune by value bisection.
//
// Roofline: tensor pipe, 2*Q*N*D flops; one pass over the bf16 corpus per 128 queries per CTA (L2-resident ring).
#include <stdlib.h>

#include "tc_common.cuh"
#include "topk_common.cuh"

namespace tt {

using namespace tt::tc;

constexpr int KT_BM = 128;
constexpr int KT_BN = 256;
constexpr int KT_HALF = 128;     // columns per softmax group
constexpr int KT_THREADS = 320;
constexpr int KT_CAP = 512;      // candidate list capacity
constexpr int KT_SLACK = 32;     // a prune keeps between K' and K' + KT_SLACK entries
constexpr int KT_MAX_KP = 256;
template <int D> struct KtStages { static constexpr int value = D == 128 ? 3 : 6; };

// Tile order of one pass.  The corpus tiles are grouped into super-blocks (sbt tiles ~ 48 MB of bf16 rows); inside a
// super-block the tiles are numbered (query tile m, corpus tile nn) and cut into one contiguous slice per CTA, and every
// CTA walks the super-blocks in the same order.  All CTAs therefore sit in the same super-block at (roughly) the same
// time and share its corpus tiles through L2 -- without this every query tile re-streams the whole corpus from HBM
// (ncu r1: 16.3 GB of DRAM reads per launch for a 320 MB corpus shard).
struct KSched {
    int cl;                   // CTAs per cluster: the CTAs of a cluster take `cl` consecutive query tiles through the SAME
                              // corpus tiles, each loading 1/cl of every corpus tile and multicasting it to the others
    int m_tiles, n_tiles;     // query tile GROUPS (cl x 128 rows), corpus tiles (256 rows) visited by this pass
    int sbt, n_sb;            // corpus tiles per super-block, number of super-blocks
    int64_t per_cta;          // slice of a super-block's m_tiles * sbt tiles owned by one cluster
    int grid, max_seg;        // clusters launched; max_seg: clusters that can share one (super-block, query group) run
    int lists;                // candidate lists per query = n_sb * max_seg * 2
};
__host__ __device__ __forceinline__ int ks_cnt(const KSched &s, int sb) { return min(s.sbt, s.n_tiles - sb * s.sbt); }
__host__ __device__ __forceinline__ int ks_first_cta(const KSched &s, int sb, int m) {
    return static_cast<int>((static_cast<int64_t>(m) * ks_cnt(s, sb)) / s.per_cta);
}

// 512 tiles x 256 rows x 256 B = 33.5 MB of a D=128 bf16 corpus per super-block.  Round 1 used 768 (50 MB): ncu showed
// 1.83 GB of DRAM reads per launch for a 320 MB shard -- the clusters sit at DIFFERENT corpus tiles of the super-block,
// so the whole super-block has to stay resident, and 50 MB does not survive in B200's two-partition L2 next to the
// candidate-list traffic.  Measured (Q=16384, K=100): 1.25M-row shard 6.22 -> 5.20 ms (50 % -> 60 % of the bf16 peak),
// 10M rows 36.5 -> 36.2 ms; 384 tiles: 5.11 / 38.1 ms, 256: 5.23 / 38.7 ms (more runs => more lists for stage 2).
constexpr int KT_SB_TILES_DEFAULT = 512;
// developer knob (tools/kbench.py): TT_TOPK_SB_TILES overrides the super-block size
static int kt_sb_tiles() {
    static int v = 0;
    if (v == 0) {
        const char *e = getenv("TT_TOPK_SB_TILES");
        v = (e && atoi(e) >= 64) ? atoi(e) : KT_SB_TILES_DEFAULT;
    }
    return v;
}

static KSched make_ksched(int q_tiles, int n_tiles, bool super_blocks, int cl) {
    KSched s;
    const int m_tiles = (q_tiles + cl - 1) / cl;
    const int n_sm = sm_count() / cl;
    s.cl = cl;
    s.m_tiles = m_tiles; s.n_tiles = n_tiles;
    const int KT_SB_TILES = kt_sb_tiles();
    s.sbt = (super_blocks && m_tiles >= 8 && n_tiles >= 2 * KT_SB_TILES) ? KT_SB_TILES : n_tiles;
    s.n_sb = (n_tiles + s.sbt - 1) / s.sbt;
    const int64_t sb_total = static_cast<int64_t>(m_tiles) * s.sbt;
    int64_t g = (sb_total + 1) / 2;
    if (g > n_sm) g = n_sm;
    if (g < 1) g = 1;
    s.per_cta = (sb_total + g - 1) / g;
    // at most ~8 CTAs per run (each leaves 2 lists for stage 2 to merge) -- unless there are so few query tiles that
    // this would idle SMs (the repair pass for a handful of queries): then one run spreads over sm_count / m_tiles CTAs
    int seg_cap = (n_sm + m_tiles - 1) / m_tiles;
    if (seg_cap < 8) seg_cap = 8;
    if (s.per_cta * seg_cap < s.sbt) s.per_cta = (s.sbt + seg_cap - 1) / seg_cap;
    s.grid = static_cast<int>((sb_total + s.per_cta - 1) / s.per_cta);
    s.max_seg = static_cast<int>((s.sbt + s.per_cta - 2) / s.per_cta) + 1;
    s.lists = s.n_sb * s.max_seg * 2;
    return s;
}

// position of a warp role inside its CTA's tiles (one slice per super-block)
struct KCursor {
    int i;             // local tile counter over all slices (barrier phases)
    int sb, m, nn;     // super-block, query tile, corpus tile inside the super-block
    int cnt;           // corpus tiles of this super-block
    int left;          // tiles left in this super-block's slice (including the current one)
    int r;             // local run counter (a run = consecutive tiles of one (sb, m))
    bool first;        // current tile starts a run
    bool done;
    __device__ __forceinline__ void enter(const KSched &s, int cta) {   // find the next non-empty slice from `sb` on
        for (; sb < s.n_sb; ++sb) {
            cnt = ks_cnt(s, sb);
            const int64_t tiles = static_cast<int64_t>(s.m_tiles) * cnt;
            const int64_t a = static_cast<int64_t>(cta) * s.per_cta;
            if (a >= tiles) continue;
            const int64_t b = min(tiles, a + s.per_cta);
            m = static_cast<int>(a / cnt);
            nn = static_cast<int>(a - static_cast<int64_t>(m) * cnt);
            left = static_cast<int>(b - a);
            first = true;
            return;
        }
        done = true;
    }
    __device__ __forceinline__ void init(const KSched &s, int cta) {
        i = 0; r = 0; sb = 0; done = false; left = 0; m = nn = 0; cnt = 1; first = true;
        enter(s, cta);
    }
    __device__ __forceinline__ void next(const KSched &s, int cta) {
        ++i;
        ++r;                                   // provisional: undone below when the run continues
        if (--left == 0) { ++sb; enter(s, cta); return; }
        if (++nn == cnt) { nn = 0; ++m; first = true; return; }
        first = false;
        --r;
    }
    __device__ __forceinline__ bool last() const { return nn == cnt - 1 || left == 1; }
    __device__ __forceinline__ int tile(const KSched &s) const { return sb * s.sbt + nn; }
};

struct TopkTcParams {
    int64_t n_query, n_corpus;
    KSched sch;
    int kp;
    int tile_stride;         // corpus tile n covers rows [n * tile_stride * 256, +256): > 1 for the sampling pass
    const float *tau0;       // [n_query] initial thresholds from the sampling pass (NULL: start at the floor)
    float *raw_v;            // full pass: [grid][256][KT_RAW_CAP][8] raw score groups of the filter threads
    int32_t *raw_c;          //            [grid][256][KT_RAW_CAP]    corpus row of each group's first score
    float *blk_max;          // sampling pass only: [m_tiles][n_blk][128] block maxima (see topk_tc_kernel)
    int blk_tiles, n_blk;    // sampled tiles per block; blocks per query = ceil(sampled tiles / blk_tiles) * 2
    const int64_t *mask_offsets, *mask_rows;
    float *cand_v;       // [n_query][lists][KT_CAP] approximate scores
    int32_t *cand_i;     //                          corpus rows
    int32_t *cand_n;     // [n_query][lists] entries (-1: the list overflowed, 0: unused slot)
    float *cand_tau;     // [n_query][lists] everything left out scored <= tau (-inf: unused slot)
};

// ---------------------------------------------------------------- prep: fp32 -> bf16 rows (+ error bounds)
// bounds[0] = max over rows of |bf16(x)|, bounds[1] = max over rows of |x - bf16(x)| (both rounded up): the two
// corpus-side terms of stage 2's proof obligation.
__global__ void __launch_bounds__(256)
tk_convert_rows(const float *__restrict__ in, int64_t n, int dim, __nv_bfloat16 *__restrict__ out,
                unsigned int *__restrict__ bounds_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    float best = 0.f, best_err = 0.f;
    for (int64_t p = warp; p < n; p += n_warps) {
        float ss = 0.f, se = 0.f;
        for (int c = lane * 4; c < dim; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(in + p * dim + c));
            __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
            uint2 raw;
            raw.x = *reinterpret_cast<uint32_t *>(&a);
            raw.y = *reinterpret_cast<uint32_t *>(&b);
            *reinterpret_cast<uint2 *>(out + p * dim + c) = raw;
            const float r0 = __uint_as_float(raw.x << 16), r1 = __uint_as_float(raw.x & 0xffff0000u);
            const float r2 = __uint_as_float(raw.y << 16), r3 = __uint_as_float(raw.y & 0xffff0000u);
            ss += r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
            const float d0 = v.x - r0, d1 = v.y - r1, d2 = v.z - r2, d3 = v.w - r3;   // exact (Sterbenz)
            se += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
        best = fmaxf(best, warp_sum(ss));
        best_err = fmaxf(best_err, warp_sum(se));
    }
    if (bounds_bits != nullptr && lane == 0) {   // non-negative floats order like their bits; 1e-5: fp32 sums above
        if (best > 0.f) atomicMax(bounds_bits, __float_as_uint(sqrtf(best) * 1.00001f));
        if (best_err > 0.f) atomicMax(bounds_bits + 1, __float_as_uint(sqrtf(best_err) * 1.00001f));
    }
}

// ---------------------------------------------------------------- list prune (one warp, one list)
// Keeps the entries with score >= t where t is found by bisection so that between kp and kp + KT_SLACK entries
// survive (more only if many scores are equal).  Returns the new count; *tau_new = t if anything was dropped.
__device__ __forceinline__ int tk_prune_list(float *__restrict__ lv, int32_t *__restrict__ li, int n, int kp, int lane,
                                             float *tau_new) {
    float v[KT_CAP / 32];
    int32_t ix[KT_CAP / 32];
#pragma unroll
    for (int t = 0; t < KT_CAP / 32; ++t) {
        const int idx = t * 32 + lane;
        const bool ok = idx < n;
        v[t] = ok ? lv[idx] : -INFINITY;
        ix[t] = ok ? li[idx] : 0;
    }
    float hi = v[0], lo = (lane < n) ? v[0] : INFINITY;
#pragma unroll
    for (int t = 1; t < KT_CAP / 32; ++t) {
        hi = fmaxf(hi, v[t]);
        if (t * 32 + lane < n) lo = fminf(lo, v[t]);
    }
    hi = warp_max(hi);
    lo = -warp_max(-lo);
    int c_lo = n;
    if (n > kp + KT_SLACK) {
        for (int it = 0; it < 26; ++it) {
            const float mid = lo + 0.5f * (hi - lo);
            if (!(mid > lo && mid < hi)) break;
            int c = 0;
#pragma unroll
            for (int t = 0; t < KT_CAP / 32; ++t) c += (v[t] >= mid) ? 1 : 0;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= kp) {
                lo = mid;
                c_lo = c;
                if (c <= kp + KT_SLACK) break;
            } else {
                hi = mid;
            }
        }
    }
    *tau_new = (c_lo < n) ? lo : -INFINITY;
    if (c_lo == n) return n;
    __syncwarp();
    int mine = 0;
#pragma unroll
    for (int t = 0; t < KT_CAP / 32; ++t) mine += (v[t] >= lo) ? 1 : 0;
    int pos = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, pos, o);
        if (lane >= o) pos += x;
    }
    pos -= mine;
#pragma unroll
    for (int t = 0; t < KT_CAP / 32; ++t)
        if (v[t] >= lo) { lv[pos] = v[t]; li[pos] = ix[t]; ++pos; }
    __syncwarp();
    return c_lo;
}

// ---------------------------------------------------------------- main kernel
constexpr int KT_S2_CAP = 1024;               // stage 2 holds this many candidates at once (>= K' + one full list)
constexpr int KT_S2_LISTS = 512;              // stage 2 gathers up to this many lists per query in one step
constexpr int KT_RAW_CAP = 64;                // raw 8-score groups a filter thread can hold before they are drained
constexpr uint32_t KT_MASKED = 0xff7fffe0u;   // -3.4028e38: the score of a corpus row past the end (finite, below every threshold)
constexpr float KT_TAU_FLOOR = -3.0e38f;      // thresholds start here ("nothing left out yet"), above KT_MASKED

__device__ __forceinline__ void st_global_v4_pred(float *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
                 :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(static_cast<uint32_t>(pred)) : "memory");
}
__device__ __forceinline__ void st_global_u32_pred(int32_t *p, uint32_t v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.global.b32 [%0], %1;\n\t}"
                 :: "l"(p), "r"(v), "r"(static_cast<uint32_t>(pred)) : "memory");
}

// float max through integer atomics (the slot starts at -inf): non-negative floats order like ints, negative ones
// like unsigned ints reversed
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// SAMPLE = true: the sampling pass over every KT_STRIDE_B-th corpus tile.  No candidate lists, no append path at all:
// a filter thread only keeps the maximum of its row over a BLOCK of blk_tiles consecutive sampled half tiles (the
// 3-input-max chain the filter runs anyway) and folds it into blk_max[query tile][block][row].  tk_tau0_kernel turns
// the ~600 block maxima of a query into its starting threshold: with p = P(score > t), a block of b scores has
// P(max > t) = 1 - exp(-b p), so the r-th largest block maximum estimates the score above which C corpus items
// lie for r = n_blk * (1 - exp(-b C / N)).
template <int D, bool SAMPLE, int CL>
__global__ void __launch_bounds__(KT_THREADS, 1)
topk_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_e,
               const TopkTcParams prm) {
    constexpr int KB = D / 64;
    constexpr int ST = KtStages<D>::value;
    constexpr int X_BYTES = KT_BM * D * 2;
    constexpr int W_BYTES = KT_BN * D * 2;
    constexpr int XK_BYTES = KT_BM * 128;     // one 64-column block of the query tile
    constexpr int WK_BYTES = KT_BN * 128;     // one 64-column block of a corpus tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *x_tile = smem;
    uint8_t *w_tiles = smem + X_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(w_tiles + ST * W_BYTES);
    uint64_t *full = bars;              // [ST]
    uint64_t *empty = full + ST;        // [ST]
    uint64_t *sfull = empty + ST;       // [2]
    uint64_t *sfree = sfull + 2;        // [2]  256 arrivals
    uint64_t *xfull = sfree + 2;        // [1]
    uint64_t *xfree = xfull + 1;        // [1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(xfree + 1);
    // thresholds published to the other column-half group (same query rows): tau_sh[grp][row], tagged per warp with
    // the row tile they belong to
    float *tau_sh = reinterpret_cast<float *>(tmem_slot + 2);            // [2][128]
    volatile int *tau_tag = reinterpret_cast<volatile int *>(tau_sh + 2 * KT_BM);   // [8]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const KSched &sch = prm.sch;
    // cluster c = CTAs [c * CL, (c + 1) * CL): same tile sequence, query tile = group * CL + rank
    const int cta = static_cast<int>(blockIdx.x) / CL;
    const int rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
    constexpr uint16_t CL_MASK = static_cast<uint16_t>((1u << CL) - 1u);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_q);
        prefetch_tensormap(&map_e);
        // a stage is free again once EVERY CTA of the cluster has consumed it (the others multicast into it)
        for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
        for (int a = 0; a < 2; ++a) { mbar_init(&sfull[a], 1); mbar_init(&sfree[a], 256); }
        mbar_init(xfull, 1);
        mbar_init(xfree, 1);
        fence_barrier_init();
        for (int w = 0; w < 8; ++w) tau_tag[w] = -1;
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // the peers' barriers exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (whole warp in uniform control flow, one elected lane issues) =====
        KCursor c;
        c.init(sch, cta);
        for (; !c.done; c.next(sch, cta)) {
            if (c.first) {
                if (c.r >= 1) mbar_wait(xfree, (c.r - 1) & 1);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(xfull, X_BYTES);
                    for (int kb = 0; kb < KB; ++kb)
                        tma_load_2d(x_tile + kb * XK_BYTES, &map_q, xfull, kb * 64, (c.m * CL + rank) * KT_BM);
                }
                __syncwarp();
            }
            const int stage = c.i % ST;
            mbar_wait(&empty[stage], ((c.i / ST) & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&full[stage], W_BYTES);
                const int row0 = c.tile(sch) * prm.tile_stride * KT_BN;
                for (int kb = 0; kb < KB; ++kb) {
                    if (CL == 1) {
                        tma_load_2d(w_tiles + stage * W_BYTES + kb * WK_BYTES, &map_e, &full[stage], kb * 64, row0);
                    } else {   // this CTA's 1/CL of the rows, into every CTA of the cluster
                        constexpr int PART = KT_BN / CL;
                        tma_load_2d_multicast(w_tiles + stage * W_BYTES + kb * WK_BYTES + rank * PART * 128, &map_e,
                                              &full[stage], kb * 64, row0 + rank * PART, CL_MASK);
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = idesc_bf16_f32(KT_BM, KT_BN, 0, 0);
        const uint64_t xdesc = smem_desc_k_sw128(smem_u32(x_tile));
        const uint64_t wdesc = smem_desc_k_sw128(smem_u32(w_tiles));
        KCursor c;
        c.init(sch, cta);
        for (; !c.done; c.next(sch, cta)) {
            const int i = c.i, b = i & 1, stage = i % ST;
            if (i >= 2) mbar_wait(&sfree[b], ((i >> 1) - 1) & 1);
            if (c.first) mbar_wait(xfull, c.r & 1);
            mbar_wait(&full[stage], (i / ST) & 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t wd = wdesc + static_cast<uint64_t>(stage * (W_BYTES >> 4));
                const uint32_t acc = tmem_base + b * KT_BN;
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint32_t xo = ((k / 4) * XK_BYTES + (k % 4) * 32) >> 4;
                    const uint32_t wo = ((k / 4) * WK_BYTES + (k % 4) * 32) >> 4;
                    if (k == 0) umma_f16_first(acc, xdesc + xo, wd + wo, idesc);
                    else umma_f16_acc(acc, xdesc + xo, wd + wo, idesc);
                }
                if (CL == 1) umma_commit(&empty[stage]); else umma_commit_multicast(&empty[stage], CL_MASK);
                umma_commit(&sfull[b]);
                if (c.last()) umma_commit(xfree);
            }
            __syncwarp();
        }
    } else {
        // ===== filter warps: group = column half, thread = query row =====
        const int quarter = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int r_in_tile = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        KCursor c;
        c.init(sch, cta);
        float tau = INFINITY;
        int cnt = 0;        // entries of this thread's candidate list
        int rcnt = 0;       // 8-score groups waiting in this thread's raw buffer
        bool ovf = false, row_ok = false;
        float run_max = -INFINITY;   // SAMPLE: maximum of the current block
        int64_t q = 0, list = 0, m_lo = 0, m_hi = 0;
        // raw buffer of this thread: KT_RAW_CAP groups of 8 scores + the corpus row of each group's first score
        const int64_t raw_slot = static_cast<int64_t>(blockIdx.x) * (KT_THREADS - 64) + (threadIdx.x - 64);
        float *raw_v = SAMPLE ? nullptr : prm.raw_v + raw_slot * (KT_RAW_CAP * 8);
        int32_t *raw_c = SAMPLE ? nullptr : prm.raw_c + raw_slot * KT_RAW_CAP;

        // Every lane moves its own raw groups into its own candidate list, keeping the scores above its threshold that
        // its history mask does not exclude (all 32 lanes at once: the trip counts are similar, ~Poisson around the
        // same mean).  A lane whose list runs full stops; those lists are pruned by the whole warp (threshold rises)
        // and the lanes carry on against the new threshold.
        auto drain = [&]() {
            float *lv = prm.cand_v + list * KT_CAP;
            int32_t *li = prm.cand_i + list * KT_CAP;
            int g = 0;
            const int n_g = row_ok ? rcnt : 0;
            while (true) {
                float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na;
                int32_t nc = 0;
                if (g < n_g) {
                    na = *reinterpret_cast<const float4 *>(raw_v + g * 8);
                    nb = *reinterpret_cast<const float4 *>(raw_v + g * 8 + 4);
                    nc = raw_c[g];
                }
                while (g < n_g && cnt <= KT_CAP - 8) {
                    const float xs[8] = {na.x, na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w};
                    const int32_t col = nc;
                    ++g;
                    if (g < n_g) {   // next group's loads fly while this one is appended
                        na = *reinterpret_cast<const float4 *>(raw_v + g * 8);
                        nb = *reinterpret_cast<const float4 *>(raw_v + g * 8 + 4);
                        nc = raw_c[g];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        bool keep = xs[j] > tau;
                        if (keep && m_hi > m_lo) keep = !tk_masked(prm.mask_rows, m_lo, m_hi, static_cast<int64_t>(col + j));
                        if (keep) { lv[cnt] = xs[j]; li[cnt] = col + j; ++cnt; }
                    }
                }
                unsigned full = __ballot_sync(0xffffffffu, g < n_g);
                if (!full) break;
                while (full) {
                    const int src = __ffs(full) - 1;
                    full &= full - 1;
                    const int64_t l2 = __shfl_sync(0xffffffffu, list, src);
                    const int n2 = __shfl_sync(0xffffffffu, cnt, src);
                    float t_new;
                    __syncwarp();
                    const int kept = tk_prune_list(prm.cand_v + l2 * KT_CAP, prm.cand_i + l2 * KT_CAP, n2, prm.kp, lane, &t_new);
                    if (lane == src) {
                        cnt = kept;
                        if (t_new > tau) { tau = t_new; tau_sh[grp * KT_BM + r_in_tile] = tau; }
                        if (kept > KT_CAP - 8) { ovf = true; g = n_g; }   // only if > 500 scores tie
                    }
                }
            }
            rcnt = 0;
        };

        for (; !c.done; c.next(sch, cta)) {
            const int i = c.i, b = i & 1;
            if (c.first) {
                q = static_cast<int64_t>(c.m * CL + rank) * KT_BM + r_in_tile;
                row_ok = q < prm.n_query;
                const int part = cta - ks_first_cta(sch, c.sb, c.m);
                list = (q * sch.lists) + (c.sb * sch.max_seg + part) * 2 + grp;
                // never below KT_TAU_FLOOR: corpus rows past the end carry KT_MASKED, a finite value under the floor
                tau = row_ok ? fmaxf(prm.tau0 != nullptr ? prm.tau0[q] : KT_TAU_FLOOR, KT_TAU_FLOOR) : INFINITY;
                cnt = 0;
                rcnt = 0;
                ovf = false;
                run_max = -INFINITY;
                m_lo = m_hi = 0;
                if (row_ok && prm.mask_offsets != nullptr) { m_lo = prm.mask_offsets[q]; m_hi = prm.mask_offsets[q + 1]; }
                tau_sh[grp * KT_BM + r_in_tile] = -INFINITY;
                __syncwarp();
                __threadfence_block();
                if (lane == 0) tau_tag[grp * 4 + quarter] = c.sb * sch.m_tiles + c.m;
            }
            // the other group's K'-th best so far bounds the row's K'-th best from below just as well as ours does
            if (!SAMPLE && tau_tag[(grp ^ 1) * 4 + quarter] == c.sb * sch.m_tiles + c.m) {
                const float other = *reinterpret_cast<volatile float *>(tau_sh + (grp ^ 1) * KT_BM + r_in_tile);
                if (row_ok) tau = fmaxf(tau, other);
            }
            mbar_wait(&sfull[b], (i >> 1) & 1);
            tc_fence_after();
            uint32_t r[4][32];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) tmem_ld_32x32(lane_addr + b * KT_BN + grp * KT_HALF + qq * 32, r[qq]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&sfree[b]);
            const int64_t col0 = static_cast<int64_t>(c.tile(sch)) * prm.tile_stride * KT_BN + grp * KT_HALF;
            if (col0 + KT_HALF > prm.n_corpus) {   // last tile: rows past the corpus were zero-filled by TMA
#pragma unroll
                for (int qq = 0; qq < 4; ++qq)
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + qq * 32 + j >= prm.n_corpus) r[qq][j] = KT_MASKED;
            }
            float tile_max = -INFINITY;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                // four 8-score maxima (3-input max chains) and ONE compare + branch per 32-column chunk: two chunks in
                // three have no score above the threshold
                float sub[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t *x = &r[qq][g * 8];
                    float m = fmaxf(fmaxf(__uint_as_float(x[0]), __uint_as_float(x[1])), __uint_as_float(x[2]));
                    m = fmaxf(fmaxf(m, __uint_as_float(x[3])), __uint_as_float(x[4]));
                    m = fmaxf(fmaxf(m, __uint_as_float(x[5])), __uint_as_float(x[6]));
                    sub[g] = fmaxf(m, __uint_as_float(x[7]));
                }
                const float chunk_max = fmaxf(fmaxf(sub[0], sub[1]), fmaxf(sub[2], sub[3]));
                if (SAMPLE) {
                    tile_max = fmaxf(tile_max, chunk_max);
                } else if (chunk_max > tau) {
                    // A group that beats the threshold is copied RAW into the thread's buffer by predicated 16-byte
                    // stores: no search for "which score", no loop, no dependent chain.  (Finding and appending the
                    // individual scores right here cost the warp ~350 cycles of latency per hit -- 2.3 ms of a 6.6 ms
                    // pass -- whichever way it was coded: tag + max tree + loop, or 8 predicated appends.)
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t *x = &r[qq][g * 8];
                        const bool hit = sub[g] > tau;
                        st_global_v4_pred(raw_v + rcnt * 8, x[0], x[1], x[2], x[3], hit);
                        st_global_v4_pred(raw_v + rcnt * 8 + 4, x[4], x[5], x[6], x[7], hit);
                        st_global_u32_pred(raw_c + rcnt, static_cast<uint32_t>(col0 + qq * 32 + g * 8), hit);
                        rcnt += hit ? 1 : 0;
                    }
                }
            }
            const bool seg_end = c.last();
            if (SAMPLE) {
                run_max = fmaxf(run_max, tile_max);
                const int st = c.tile(sch);
                if ((st + 1) % prm.blk_tiles == 0 || seg_end) {
                    if (row_ok)
                        atomic_max_float(prm.blk_max + (static_cast<int64_t>(c.m * CL + rank) * prm.n_blk + (st / prm.blk_tiles) * 2 + grp) * KT_BM + r_in_tile,
                                         run_max);
                    run_max = -INFINITY;
                }
            } else {
                // the raw buffer must have room for the 16 groups of one more tile
                if (__any_sync(0xffffffffu, row_ok && (seg_end ? rcnt > 0 : rcnt > KT_RAW_CAP - 16))) drain();
                if (seg_end) {
                    // the list is cut to ~K' before it is handed to stage 2
                    unsigned cut = __ballot_sync(0xffffffffu, row_ok && cnt > prm.kp + KT_SLACK);
                    while (cut) {
                        const int src = __ffs(cut) - 1;
                        cut &= cut - 1;
                        const int64_t l2 = __shfl_sync(0xffffffffu, list, src);
                        const int n2 = __shfl_sync(0xffffffffu, cnt, src);
                        float t_new;
                        __syncwarp();
                        const int kept = tk_prune_list(prm.cand_v + l2 * KT_CAP, prm.cand_i + l2 * KT_CAP, n2, prm.kp, lane, &t_new);
                        if (lane == src) { cnt = kept; tau = fmaxf(tau, t_new); }
                    }
                    if (row_ok) {
                        prm.cand_n[list] = ovf ? -1 : cnt;
                        prm.cand_tau[list] = tau;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into it or arrive on its barriers
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------- stage 2: K' best by approximate score -> exact
// one CTA (128 threads) per query
__global__ void __launch_bounds__(128)
topk_tc_stage2(const float *__restrict__ query, const float *__restrict__ corpus, int dim, int k, int kp, int n_lists,
               int64_t row_offset, const float *__restrict__ cand_v,
               const int32_t *__restrict__ cand_i, const int32_t *__restrict__ cand_n,
               const float *__restrict__ cand_tau, const unsigned int *__restrict__ ebounds_bits,
               double *__restrict__ out_scores, int64_t *__restrict__ out_idx, int32_t *__restrict__ unverified) {
    __shared__ uint64_t ak[KT_S2_CAP];   // (approximate score, row) keys, see tk_pack_key
    __shared__ double ev[KT_MAX_KP];
    __shared__ int32_t ei[KT_MAX_KP];
    __shared__ float s_tau;
    __shared__ int s_bad;
    __shared__ double s_qn, s_qd;
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_tau = -INFINITY; s_bad = 0; }
    __syncthreads();
    int tot = 0;   // entries in av/ai (uniform across the CTA)
    // keep the kp largest keys (= best approximate scores, ties by row); what falls out bounds tau from below.
    // MSB-first radix SELECT over the 64-bit keys, 8 bits per pass, stopping as soon as the boundary bin is taken
    // whole (4 passes for distinct scores): ~10x fewer instructions than sorting the 512 / 1024 keys, which was 40 %
    // of this kernel.
    __shared__ int s_hist[256];
    __shared__ int s_sel[3];                  // boundary bin, keys above it, keys in it
    __shared__ unsigned long long s_wmax[4];
    __shared__ int s_wcnt[4];
    auto reduce_to_kp = [&]() {
        if (tot <= kp) return;                // uniform
        constexpr int PER_T = KT_S2_CAP / 128;
        uint64_t mine[PER_T];
#pragma unroll
        for (int j = 0; j < PER_T; ++j) mine[j] = (tid + 128 * j < tot) ? ak[tid + 128 * j] : 0ull;   // 0 = below every key
        uint64_t prefix = 0;
        int fixed_bits = 0, remaining = kp;
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            for (int t = tid; t < 256; t += blockDim.x) s_hist[t] = 0;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < PER_T; ++j) {
                if (tid + 128 * j < tot && (fixed_bits == 0 || (mine[j] >> (shift + 8)) == prefix))
                    atomicAdd(&s_hist[static_cast<int>((mine[j] >> shift) & 255u)], 1);
            }
            __syncthreads();
            if (warp == 0) {
                // bins from high to low: lane l owns bins 255 - 8l .. 248 - 8l
                int c[8], lane_sum = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { c[k] = s_hist[255 - (lane * 8 + k)]; lane_sum += c[k]; }
                int incl = lane_sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int x = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += x;
                }
                int above = incl - lane_sum;              // keys in bins higher than this lane's
                if (above < remaining && incl >= remaining) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (above < remaining && above + c[k] >= remaining) {
                            s_sel[0] = 255 - (lane * 8 + k); s_sel[1] = above; s_sel[2] = c[k];
                            above = remaining;            // stop
                        } else if (above < remaining) above += c[k];
                    }
                }
            }
            __syncthreads();
            prefix = (prefix << 8) | static_cast<uint64_t>(s_sel[0]);
            fixed_bits += 8;
            remaining -= s_sel[1];
            const int in_bin = s_sel[2];
            __syncthreads();
            if (in_bin == remaining) break;               // the whole boundary bin is kept
        }
        // keep keys whose fixed high bits are >= prefix; the largest dropped key bounds tau
        const int fs = 64 - fixed_bits;
        uint64_t dropped_max = 0;
        int n_keep = 0;
        bool keep[PER_T];
#pragma unroll
        for (int j = 0; j < PER_T; ++j) {
            const bool valid = tid + 128 * j < tot;
            const uint64_t hi = (fs >= 64) ? 0ull : (mine[j] >> fs);
            keep[j] = valid && hi >= prefix;
            if (valid && !keep[j] && mine[j] > dropped_max) dropped_max = mine[j];
            n_keep += keep[j] ? 1 : 0;
        }
        int incl = n_keep;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t x = __shfl_xor_sync(0xffffffffu, dropped_max, o);
            if (x > dropped_max) dropped_max = x;
        }
        if (lane == 31) s_wcnt[warp] = incl;
        if (lane == 0) s_wmax[warp] = dropped_max;
        __syncthreads();
        int pos = incl - n_keep;
        for (int w2 = 0; w2 < warp; ++w2) pos += s_wcnt[w2];
#pragma unroll
        for (int j = 0; j < PER_T; ++j)
            if (keep[j]) ak[pos++] = mine[j];
        if (tid == 0) {
            uint64_t dm = s_wmax[0];
            for (int w2 = 1; w2 < 4; ++w2) if (s_wmax[w2] > dm) dm = s_wmax[w2];
            if (dm != 0ull) s_tau = fmaxf(s_tau, tk_key_score(dm));
        }
        tot = s_wcnt[0] + s_wcnt[1] + s_wcnt[2] + s_wcnt[3];
        __syncthreads();
    };
    // gather the query's lists.  Fast path (the usual case: a few hundred candidates in a few dozen lists): counts and
    // thresholds of all lists at once, prefix sums in shared memory, then one flat copy with every load independent.
    __shared__ int s_off[KT_S2_LISTS + 1];
    __shared__ int s_wsum[4];
    __shared__ float s_wtau[4];
    bool gathered = false;
    if (n_lists <= KT_S2_LISTS) {
        constexpr int PER = KT_S2_LISTS / 128;
        int take[PER];
        int mine = 0, bad = 0;
        float tmax = -INFINITY;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int l = tid * PER + u;
            take[u] = 0;
            if (l < n_lists) {
                const int n = cand_n[q * n_lists + l];      // unused slots: n = 0, tau = -inf (tk_init_lists)
                tmax = fmaxf(tmax, cand_tau[q * n_lists + l]);
                bad |= n < 0 ? 1 : 0;
                take[u] = n < 0 ? KT_CAP : n;
            }
            mine += take[u];
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        tmax = warp_max(tmax);
        bad = __any_sync(0xffffffffu, bad != 0) ? 1 : 0;
        if (lane == 31) s_wsum[warp] = incl;
        if (lane == 0) { s_wtau[warp] = tmax; if (bad) s_bad = 1; }
        __syncthreads();
        int base = incl - mine;
        for (int w = 0; w < warp; ++w) base += s_wsum[w];
        const int total = s_wsum[0] + s_wsum[1] + s_wsum[2] + s_wsum[3];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int l = tid * PER + u;
            if (l < n_lists) s_off[l] = base;
            base += take[u];
        }
        if (tid == 0) {
            s_off[n_lists] = total;
            s_tau = fmaxf(fmaxf(s_wtau[0], s_wtau[1]), fmaxf(s_wtau[2], s_wtau[3]));
        }
        __syncthreads();
        if (total <= KT_S2_CAP) {
            for (int pI = tid; pI < total; pI += blockDim.x) {
                int lo = 0, hi = n_lists;            // last list with s_off[l] <= pI
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_off[mid] <= pI) lo = mid; else hi = mid;
                }
                const int64_t src = (q * n_lists + lo) * KT_CAP + (pI - s_off[lo]);
                ak[pI] = tk_pack_key(cand_v[src], cand_i[src]);
            }
            tot = total;
            gathered = true;
            __syncthreads();
        }
    }
    if (!gathered) {
        for (int l = 0; l < n_lists; ++l) {
            const int64_t list = q * n_lists + l;
            const int n = cand_n[list];
            if (tid == 0) {
                s_tau = fmaxf(s_tau, cand_tau[list]);
                if (n < 0) s_bad = 1;
            }
            const int take = n < 0 ? KT_CAP : n;
            if (tot + take > KT_S2_CAP) reduce_to_kp();
            for (int t = tid; t < take; t += blockDim.x) ak[tot + t] = tk_pack_key(cand_v[list * KT_CAP + t], cand_i[list * KT_CAP + t]);
            tot += take;
            __syncthreads();
        }
    }
    reduce_to_kp();
    // exact fp64 re-score of the survivors (fixed summation order: lane-strided, then xor tree); four candidates per
    // warp step so that their row loads are in flight together
    const float *qv = query + q * dim;
    for (int c0 = warp * 4; c0 < tot; c0 += 16) {
        double d[4] = {0.0, 0.0, 0.0, 0.0};
        const float *e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = corpus + static_cast<int64_t>(tk_key_row(ak[min(c0 + u, tot - 1)])) * dim;
        for (int t = lane; t < dim; t += 32) {
            const double qt = static_cast<double>(qv[t]);
            float x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = __ldg(e[u] + t);
#pragma unroll
            for (int u = 0; u < 4; ++u) d[u] = fma(qt, static_cast<double>(x[u]), d[u]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < 4; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + u < tot) { ev[c0 + u] = d[u]; ei[c0 + u] = tk_key_row(ak[c0 + u]); }
        }
    }
    if (warp == 0) {   // |q| and |q - bf16(q)| for the proof obligation
        double n2 = 0.0, d2 = 0.0;
        for (int t = lane; t < dim; t += 32) {
            const double x = static_cast<double>(qv[t]);
            const double dx = x - static_cast<double>(__bfloat162float(__float2bfloat16_rn(qv[t])));
            n2 = fma(x, x, n2);
            d2 = fma(dx, dx, d2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        }
        if (lane == 0) { s_qn = sqrt(n2); s_qd = sqrt(d2); }
    }
    int np2 = 32;
    while (np2 < tot) np2 <<= 1;
    __syncthreads();
    for (int t = tid; t < np2; t += blockDim.x)
        if (t >= tot) { ev[t] = -INFINITY; ei[t] = 0x7fffffff; }
    __syncthreads();
    tk_bitonic(ev, ei, np2, tid, static_cast<int>(blockDim.x), [] { __syncthreads(); });
    for (int t = tid; t < k; t += blockDim.x) {
        const bool ok = t < tot;
        out_scores[q * k + t] = ok ? ev[t] : -INFINITY;
        out_idx[q * k + t] = ok ? static_cast<int64_t>(ei[t]) + row_offset : -1;
    }
    if (tid == 0) {
        // nothing was ever left out (tau still at its floor): exact by construction.  Otherwise the K-th exact score
        // must clear the best exact score anything left out could have.  For a left-out row e with bf16 copies q~, e~:
        //   q.e - q~.e~ = (q - q~).e~ + q.(e - e~)  =>  q.e <= q~.e~ + |q - q~| max|e~| + |q| max|e - e~|
        // (Cauchy-Schwarz; |q - q~| is this query's own rounding error, the two corpus bounds come from
        // tk_convert_rows), and the filter's value of q~.e~ is below tau + 2e-5 |q~| |e~|: fp32 accumulation of 128
        // products in the tensor core (<= 1.6e-5, with slack).
        bool good = s_bad == 0;
        if (good && s_tau > KT_TAU_FLOOR) {
            const double e_norm = static_cast<double>(__uint_as_float(ebounds_bits[0]));
            const double e_err = static_cast<double>(__uint_as_float(ebounds_bits[1]));
            const double eps = s_qd * e_norm + s_qn * e_err + 2.0e-5 * (s_qn + s_qd) * e_norm + 1e-30;
            good = (tot >= k) && (ev[k - 1] > static_cast<double>(s_tau) + eps);
        }
        unverified[q] = good ? 0 : 1;
    }
}

// tau0[q] = the rank-th largest of the query's block maxima (one warp per query, values in registers, bisection on
// the value; the result may sit a little below the exact order statistic, which only admits more candidates).
// Rows the query masks out can be among those maxima: the rank is pushed down by the number of its masked rows that
// lie in sampled tiles.
constexpr int KT_MAX_BLK = 640;
__global__ void __launch_bounds__(256)
tk_tau0_kernel(const float *__restrict__ blk_max, int64_t n_query, int n_blk, int rank, int tile_stride,
               const int64_t *__restrict__ mask_offsets, const int64_t *__restrict__ mask_rows,
               float *__restrict__ tau0) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t q = warp; q < n_query; q += n_warps) {
        const float *src = blk_max + (q / KT_BM) * static_cast<int64_t>(n_blk) * KT_BM + (q % KT_BM);
        float v[KT_MAX_BLK / 32];
        float lo = INFINITY, hi = -INFINITY;
#pragma unroll
        for (int t = 0; t < KT_MAX_BLK / 32; ++t) {
            const int bI = t * 32 + lane;
            v[t] = bI < n_blk ? src[static_cast<int64_t>(bI) * KT_BM] : -INFINITY;
            if (bI < n_blk) lo = fminf(lo, v[t]);
            hi = fmaxf(hi, v[t]);
        }
        lo = -warp_max(-lo);
        hi = warp_max(hi);
        int want = rank;
        if (mask_offsets != nullptr) {
            int hidden = 0;
            for (int64_t t = mask_offsets[q] + lane; t < mask_offsets[q + 1]; t += 32)
                hidden += ((mask_rows[t] / KT_BN) % tile_stride == 0) ? 1 : 0;
            want += __reduce_add_sync(0xffffffffu, hidden);
        }
        float t0 = -INFINITY;
        if (want <= n_blk && hi > -INFINITY) {
            // invariant: count(v >= lo) >= want
            for (int it = 0; it < 30; ++it) {
                const float mid = lo + 0.5f * (hi - lo);
                if (!(mid > lo && mid < hi)) break;
                int c = 0;
#pragma unroll
                for (int t = 0; t < KT_MAX_BLK / 32; ++t) c += (v[t] >= mid) ? 1 : 0;
                c = __reduce_add_sync(0xffffffffu, c);
                if (c >= want) { lo = mid; if (c <= want + 1) break; } else hi = mid;
            }
            t0 = lo;
        }
        if (lane == 0) tau0[q] = t0;
    }
}

__global__ void tk_fill(float *__restrict__ p, int64_t n, float v) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        p[i] = v;
}

// every pass starts from empty lists: n = 0, tau = -inf (slots no CTA writes stay that way)
__global__ void tk_init_lists(int32_t *__restrict__ cand_n, float *__restrict__ cand_tau, int64_t n) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        cand_n[i] = 0;
        cand_tau[i] = -INFINITY;
    }
}

// The sampling pass gives every query a starting threshold (the full pass then stays on the filter's fast path): the
// main kernel in SAMPLE mode over every 16th corpus tile leaves ~600 block maxima per query, tk_tau0_kernel picks the
// rank-th largest.  The number of corpus items above that threshold is ~ stride * Gamma(rank): the rank is the smallest
// one for which fewer than 1.25 K' items (what the proof obligation needs with the usual crowd of near-ties around the
// K-th score) has probability < 1e-4 -- rank 26, ~420 items, for K = 100.  A threshold that still turns out too high
// makes the proof obligation of stage 2 fail and the query goes to the repair pass.
constexpr int KT_STRIDE_B = 16;
constexpr int64_t KT_SAMPLE_MIN_ROWS = 1 << 17;

struct KtPlan {
    KSched sched, sample_b;
    int kp;
    bool use_sample;
    int rank, blk_tiles, n_blk;   // sampling pass
    int cl;                       // CTAs per cluster (both passes)
    int lists;
};

// smallest r with P(Gamma(r) < x) = P(Poisson(x) >= r) < 1e-4
static int kt_rank_for(double x) {
    double term = exp(-x), below = 0.0;   // below = P(Poisson(x) < r)
    for (int r = 0; r < 4096; ++r) {
        if (1.0 - below < 1e-4) return r < 4 ? 4 : r;
        below += term;
        term *= x / (r + 1);
    }
    return 4096;
}

static KtPlan kt_plan(int64_t n_query, int64_t n_corpus, int k, bool sampling = true, bool wide = false) {
    KtPlan p;
    const int m_tiles = static_cast<int>((n_query + KT_BM - 1) / KT_BM);
    const int n_tiles = static_cast<int>((n_corpus + KT_BN - 1) / KT_BN);
    // pairs of CTAs share every corpus tile (each loads half and multicasts it): the pass is bound by the L2 -> SM
    // stream of corpus tiles (64 KB per 128 x 256 scores: 148 SMs at the tensor rate would need 17 TB/s), so halving
    // that stream is what buys speed.  With only a few query tiles the second CTA of a pair would mostly idle.
    p.cl = m_tiles >= 8 ? 2 : 1;
    p.sched = make_ksched(m_tiles, n_tiles, true, p.cl);
    int margin = k / 2;
    if (margin < 32) margin = 32;
    p.kp = k + margin;
    if (wide || p.kp > KT_MAX_KP) p.kp = KT_MAX_KP;   // wide: the repair pass for queries whose proof failed
    const int s_tiles = (n_tiles + KT_STRIDE_B - 1) / KT_STRIDE_B;
    p.sample_b = make_ksched(m_tiles, s_tiles, false, p.cl);
    p.blk_tiles = (s_tiles + KT_MAX_BLK / 2 - 1) / (KT_MAX_BLK / 2);
    p.n_blk = (s_tiles + p.blk_tiles - 1) / p.blk_tiles * 2;
    // items wanted above the threshold: C = stride * rank; as a rank among block maxima: n_blk * (1 - exp(-b C / N))
    const double want = static_cast<double>(KT_STRIDE_B) * kt_rank_for(1.25 * p.kp / KT_STRIDE_B);
    const double b = static_cast<double>(p.blk_tiles) * KT_HALF;
    p.rank = static_cast<int>(ceil(p.n_blk * (1.0 - exp(-b * want / static_cast<double>(n_corpus)))));
    if (p.rank < 4) p.rank = 4;
    // a meaningful order statistic needs the rank well inside the blocks (small corpora: scan without thresholds)
    p.use_sample = sampling && n_corpus >= KT_SAMPLE_MIN_ROWS && 4 * p.rank <= p.n_blk;
    p.lists = p.sched.lists;
    return p;
}

struct KtWs {
    __nv_bfloat16 *qb, *eb;
    unsigned int *emax;
    float *cand_v, *cand_tau, *tau0, *blk_max, *raw_v;
    int32_t *raw_c;
    int32_t *cand_i, *cand_n;
    bool ok;
    size_t used;
};

static KtWs kt_carve(void *workspace, size_t bytes, int64_t n_query, int64_t n_corpus, int dim, bool own_corpus,
                     const KtPlan &pl) {
    Workspace ws(workspace, bytes);
    KtWs w;
    const size_t lists = static_cast<size_t>(n_query) * pl.lists;
    w.qb = ws.take<__nv_bfloat16>(n_query * dim);
    w.eb = ws.take<__nv_bfloat16>(own_corpus ? n_corpus * dim : 1);
    w.emax = ws.take<unsigned int>(2);
    w.cand_v = ws.take<float>(lists * KT_CAP);
    w.cand_i = ws.take<int32_t>(lists * KT_CAP);
    w.cand_n = ws.take<int32_t>(lists);
    w.cand_tau = ws.take<float>(lists);
    w.tau0 = ws.take<float>(n_query);
    w.raw_v = ws.take<float>(static_cast<size_t>(pl.sched.grid) * pl.cl * (KT_THREADS - 64) * KT_RAW_CAP * 8);
    w.raw_c = ws.take<int32_t>(static_cast<size_t>(pl.sched.grid) * pl.cl * (KT_THREADS - 64) * KT_RAW_CAP);
    w.blk_max = ws.take<float>(static_cast<size_t>((n_query + KT_BM - 1) / KT_BM) * pl.n_blk * KT_BM);
    w.ok = ws.ok();
    w.used = ws.off;
    return w;
}

static inline unsigned kt_grid(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b);
}

// row-major bf16 [rows, dim] -> boxes of {64 columns, box_rows rows}, 128B swizzle (box_rows up to 256)
template <int D, bool SAMPLE, int CL>
static int launch_topk_tc_cl(const CUtensorMap &mq, const CUtensorMap &me, const TopkTcParams &prm, cudaStream_t st) {
    constexpr int ST = KtStages<D>::value;
    constexpr size_t smem = 1024 + static_cast<size_t>(KT_BM) * D * 2 + static_cast<size_t>(ST) * KT_BN * D * 2 + 256 + 2 * KT_BM * 4 + 64;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(topk_tc_kernel<D, SAMPLE, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(topk_tc_kernel)");
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(prm.sch.grid * CL));
    cfg.blockDim = dim3(KT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, topk_tc_kernel<D, SAMPLE, CL>, mq, me, prm);
    if (e != cudaSuccess) return cuda_status(e, "cudaLaunchKernelEx(topk_tc_kernel)");
    TT_LAUNCH_CHECK("topk_tc_kernel");
    return 0;
}

// `me` must have been built with box rows = KT_BN / prm.sch.cl
template <bool SAMPLE>
static int launch_topk_tc(int dim, const CUtensorMap &mq, const CUtensorMap &me, const TopkTcParams &prm, cudaStream_t st) {
    if (prm.sch.cl == 2)
        return dim == 128 ? launch_topk_tc_cl<128, SAMPLE, 2>(mq, me, prm, st) : launch_topk_tc_cl<64, SAMPLE, 2>(mq, me, prm, st);
    return dim == 128 ? launch_topk_tc_cl<128, SAMPLE, 1>(mq, me, prm, st) : launch_topk_tc_cl<64, SAMPLE, 1>(mq, me, prm, st);
}

}  // namespace tt

extern "C" int tt_topk_tc_prepare_corpus(const float *corpus, int64_t n_corpus, int dim, void *corpus_bf16,
                                         float *bounds, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(corpus && corpus_bf16 && bounds && n_corpus > 0, "null pointer / empty corpus");
    if (dim != 64 && dim != 128) { set_error("tensor-core top-K supports dim 64 or 128 (got %d)", dim); return TT_E_UNSUPPORTED; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(bounds, 0, 2 * sizeof(float), st);
    if (e != cudaSuccess) return cuda_status(e, "cudaMemsetAsync(bounds)");
    tk_convert_rows<<<kt_grid(n_corpus * 32, 256), 256, 0, st>>>(corpus, n_corpus, dim, static_cast<__nv_bfloat16 *>(corpus_bf16),
                                                                reinterpret_cast<unsigned int *>(bounds));
    TT_LAUNCH_CHECK("tk_convert_rows");
    return 0;
}

extern "C" int tt_score_topk_tc_workspace(int64_t n_query, int64_t n_corpus, int dim, int k, int own_corpus,
                                          size_t *bytes_host) {
    using namespace tt;
    TT_CHECK_ARG(bytes_host && n_query > 0 && n_corpus > 0 && dim > 0 && k > 0, "bad size");
    const KtPlan pl = kt_plan(n_query, n_corpus, k);
    const KtWs w = kt_carve(nullptr, ~size_t(0), n_query, n_corpus, dim, own_corpus != 0, pl);
    *bytes_host = w.used + 1024;
    return 0;
}

extern "C" int tt_score_topk_tc(const float *query, int64_t n_query, const float *corpus, const void *corpus_bf16,
                                const float *corpus_bounds, int64_t n_corpus, int dim, int k, int64_t row_offset,
                                const int64_t *mask_offsets, const int64_t *mask_rows, double *out_scores,
                                int64_t *out_idx, int32_t *unverified, int flags, void *workspace,
                                size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(query && corpus && out_scores && out_idx && unverified && workspace, "null pointer");
    TT_CHECK_ARG(n_query > 0 && n_corpus > 0 && k > 0, "non-positive size");
    TT_CHECK_ARG((mask_offsets == nullptr) == (mask_rows == nullptr), "mask_offsets / mask_rows mismatch");
    TT_CHECK_ARG((corpus_bf16 == nullptr) == (corpus_bounds == nullptr), "corpus_bf16 / corpus_bounds mismatch");
    if (dim != 64 && dim != 128) { set_error("tensor-core top-K supports dim 64 or 128 (got %d)", dim); return TT_E_UNSUPPORTED; }
    if (k > KT_MAX_KP - 32) { set_error("tensor-core top-K supports k <= %d (got %d)", KT_MAX_KP - 32, k); return TT_E_UNSUPPORTED; }
    if (n_corpus >= (int64_t(1) << 31)) { set_error("corpus shard must have < 2^31 rows"); return TT_E_UNSUPPORTED; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const KtPlan pl = kt_plan(n_query, n_corpus, k, (flags & TT_TOPK_SAMPLING) != 0, (flags & TT_TOPK_WIDE) != 0);
    const bool own = corpus_bf16 == nullptr;
    const KtWs w = kt_carve(workspace, workspace_bytes, n_query, n_corpus, dim, own, pl);
    if (!w.ok) { set_error("top-K (tensor core) workspace too small: need %zu have %zu", w.used, workspace_bytes); return TT_E_WORKSPACE; }
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) { set_error("top-K workspace must be 256-byte aligned"); return TT_E_BADARG; }

    tk_convert_rows<<<kt_grid(n_query * 32, 256), 256, 0, st>>>(query, n_query, dim, w.qb, nullptr);
    const __nv_bfloat16 *eb = static_cast<const __nv_bfloat16 *>(corpus_bf16);
    const unsigned int *emax = reinterpret_cast<const unsigned int *>(corpus_bounds);
    if (own) {
        cudaError_t e = cudaMemsetAsync(w.emax, 0, 2 * sizeof(unsigned int), st);
        if (e != cudaSuccess) return cuda_status(e, "cudaMemsetAsync(emax)");
        tk_convert_rows<<<kt_grid(n_corpus * 32, 256), 256, 0, st>>>(corpus, n_corpus, dim, w.eb, w.emax);
        eb = w.eb;
        emax = w.emax;
    }
    TT_LAUNCH_CHECK("tk_convert_rows");
    CUtensorMap mq, me;
    int rc;
    if ((rc = make_tmap_bf16_rows(&mq, w.qb, n_query, dim, KT_BM))) return rc;
    if ((rc = make_tmap_bf16_rows(&me, eb, n_corpus, dim, KT_BN / pl.cl))) return rc;
    TopkTcParams prm{};
    prm.n_query = n_query; prm.n_corpus = n_corpus;
    prm.mask_offsets = mask_offsets; prm.mask_rows = mask_rows;
    prm.cand_v = w.cand_v; prm.cand_i = w.cand_i; prm.cand_n = w.cand_n; prm.cand_tau = w.cand_tau;
    const float *tau_start = nullptr;
    if (pl.use_sample) {
        const int64_t n_max = (n_query + KT_BM - 1) / KT_BM * pl.n_blk * KT_BM;
        tk_fill<<<kt_grid(n_max, 256), 256, 0, st>>>(w.blk_max, n_max, -INFINITY);
        TT_LAUNCH_CHECK("tk_fill");
        prm.sch = pl.sample_b; prm.kp = pl.kp; prm.tile_stride = KT_STRIDE_B; prm.tau0 = nullptr;
        prm.blk_max = w.blk_max; prm.blk_tiles = pl.blk_tiles; prm.n_blk = pl.n_blk;
        if ((rc = launch_topk_tc<true>(dim, mq, me, prm, st))) return rc;
        tk_tau0_kernel<<<kt_grid(n_query * 32, 256), 256, 0, st>>>(w.blk_max, n_query, pl.n_blk, pl.rank, KT_STRIDE_B,
                                                                    mask_offsets, mask_rows, w.tau0);
        TT_LAUNCH_CHECK("tk_tau0_kernel");
        tau_start = w.tau0;
    }
    {
        const int64_t n_lists = static_cast<int64_t>(n_query) * pl.sched.lists;
        tk_init_lists<<<kt_grid(n_lists, 256), 256, 0, st>>>(w.cand_n, w.cand_tau, n_lists);
        TT_LAUNCH_CHECK("tk_init_lists");
        prm.sch = pl.sched; prm.kp = pl.kp; prm.tile_stride = 1; prm.tau0 = tau_start;
        prm.blk_max = nullptr; prm.blk_tiles = 1; prm.n_blk = 0;
        prm.raw_v = w.raw_v; prm.raw_c = w.raw_c;
        if ((rc = launch_topk_tc<false>(dim, mq, me, prm, st))) return rc;
    }
    topk_tc_stage2<<<static_cast<unsigned>(n_query), 128, 0, st>>>(query, corpus, dim, k, pl.kp, pl.sched.lists, row_offset,
                                                                   w.cand_v, w.cand_i, w.cand_n, w.cand_tau, emax, out_scores,
                                                                   out_idx, unverified);
    TT_LAUNCH_CHECK("topk_tc_stage2");
    return 0;
}
