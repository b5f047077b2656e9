// Corpus scoring + top-K for retrieval evaluation, fp32 SIMT scoring path
// (kernel 4 of the hot path; the bf16 tcgen05 scoring path lives in topk_tc.cu).
//
// Replaces `scores = U @ E^T` -> per-user -inf masking -> torch.topk at
// training_utils.py:220-258 of the reference.  The [Bq, Nc] score matrix is
// never written to HBM.
//
// Stage 1 (score + filter): CTA = 64 queries x a slice of the corpus, 64x64
// score tiles as in ce_f32.cu.  Each query row keeps a running threshold tau
// (its K'-th best score so far, K' = K + margin).  Only scores above tau are
// appended to the row's candidate list (expected K' ln(N/K') appends per
// row); a warp bitonic-sorts a list back down to K' when it fills up.
// Stage 2 (exact re-rank): one CTA per query gathers the <= splits*K'
// candidates, re-scores them in fp64 and sorts by (score desc, row asc) --
// the stated tie-break -- and emits the first K.
#include "topk_common.cuh"

namespace tt {

constexpr int TK_T = 64;
constexpr int TK_THREADS = 256;
constexpr int TK_CAP = 512;     // per (query, split) candidate list capacity
constexpr int TK_MARGIN = 8;    // extra candidates kept beyond K for the fp64 re-rank
constexpr int TK_MAX_K = 256;

__device__ __forceinline__ void tk_load_tile(float *__restrict__ dst, const float *__restrict__ src, int64_t row0,
                                             int64_t n, int dim, int stride) {
    const int vpr = dim / 4;
    for (int i = threadIdx.x; i < TK_T * vpr; i += TK_THREADS) {
        const int r = i / vpr, c = i - r * vpr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < n) v = __ldg(reinterpret_cast<const float4 *>(src + (row0 + r) * dim) + c);
        *reinterpret_cast<float4 *>(dst + r * stride + c * 4) = v;
    }
}

// one warp: shrink row r's list to its best kp entries (sorted), update cnt/tau
__device__ __forceinline__ void tk_prune_row(int r, int kp, float *__restrict__ gv, int32_t *__restrict__ gi,
                                             int *cnt, float *tau, float *sv, int32_t *si, int lane) {
    const int n = min(cnt[r], TK_CAP);
    int np2 = 32;
    while (np2 < n) np2 <<= 1;
    for (int t = lane; t < np2; t += 32) {
        sv[t] = (t < n) ? gv[t] : -INFINITY;
        si[t] = (t < n) ? gi[t] : 0x7fffffff;
    }
    __syncwarp();
    tk_bitonic(sv, si, np2, lane, 32, [] { __syncwarp(); });
    const int keep = min(n, kp);
    for (int t = lane; t < keep; t += 32) { gv[t] = sv[t]; gi[t] = si[t]; }
    __syncwarp();
    if (lane == 0) {
        cnt[r] = keep;
        if (n >= kp) tau[r] = sv[kp - 1];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(TK_THREADS)
topk_stage1(const float *__restrict__ query, int64_t Bq, const float *__restrict__ corpus, int64_t Nc, int dim, int kp,
            int tiles_total, const int64_t *__restrict__ mask_offsets, const int64_t *__restrict__ mask_rows,
            float *__restrict__ cand_v, int32_t *__restrict__ cand_i, int32_t *__restrict__ cand_n) {
    extern __shared__ __align__(16) float smem[];
    const int stride = dim + 4;
    float *Xs = smem;
    float *Ys = Xs + TK_T * stride;
    float *sv = Ys + TK_T * stride;                               // [8][TK_CAP]
    int32_t *si = reinterpret_cast<int32_t *>(sv + 8 * TK_CAP);   // [8][TK_CAP]
    __shared__ int cnt[TK_T];
    __shared__ float tau[TK_T];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t x0 = static_cast<int64_t>(blockIdx.x) * TK_T;
    const int splits = gridDim.y;
    const int split = blockIdx.y;

    tk_load_tile(Xs, query, x0, Bq, dim, stride);
    if (threadIdx.x < TK_T) { cnt[threadIdx.x] = 0; tau[threadIdx.x] = -INFINITY; }

    for (int t = split; t < tiles_total; t += splits) {
        const int64_t y0 = static_cast<int64_t>(t) * TK_T;
        __syncthreads();
        tk_load_tile(Ys, corpus, y0, Nc, dim, stride);
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k = 0; k < dim; k += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(Xs + (ty + 16 * i) * stride + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4 *>(Ys + (tx + 16 * j) * stride + k);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty + 16 * i;
            const int64_t gx = x0 + r;
            if (gx >= Bq) continue;
            const float th = tau[r];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t gy = y0 + tx + 16 * j;
                const float v = acc[i][j];
                if (gy < Nc && v > th) {
                    if (mask_offsets != nullptr && tk_masked(mask_rows, mask_offsets[gx], mask_offsets[gx + 1], gy)) continue;
                    const int slot = atomicAdd(&cnt[r], 1);
                    if (slot < TK_CAP) {
                        const int64_t base = (gx * splits + split) * TK_CAP;
                        cand_v[base + slot] = v;
                        cand_i[base + slot] = static_cast<int32_t>(gy);
                    }
                }
            }
        }
        __syncthreads();
        for (int r = warp; r < TK_T; r += TK_THREADS / 32) {
            if (cnt[r] > TK_CAP - TK_T) {
                const int64_t base = ((x0 + r) * splits + split) * TK_CAP;
                tk_prune_row(r, kp, cand_v + base, cand_i + base, cnt, tau, sv + warp * TK_CAP, si + warp * TK_CAP, lane);
            }
        }
    }
    __syncthreads();
    for (int r = warp; r < TK_T; r += TK_THREADS / 32) {
        if (x0 + r >= Bq) continue;
        const int64_t base = ((x0 + r) * splits + split) * TK_CAP;
        tk_prune_row(r, kp, cand_v + base, cand_i + base, cnt, tau, sv + warp * TK_CAP, si + warp * TK_CAP, lane);
        if (lane == 0) cand_n[(x0 + r) * splits + split] = cnt[r];
    }
}

// one CTA (128 threads) per query: fp64 re-score + final ordering
__global__ void __launch_bounds__(128)
topk_stage2(const float *__restrict__ query, const float *__restrict__ corpus, int dim, int k, int splits,
            int64_t row_offset, const float *__restrict__ cand_v, const int32_t *__restrict__ cand_i,
            const int32_t *__restrict__ cand_n, double *__restrict__ out_scores, int64_t *__restrict__ out_idx) {
    __shared__ double sv[TK_STAGE2_MAX];
    __shared__ int32_t si[TK_STAGE2_MAX];
    __shared__ int off[17];
    const int64_t q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int s = 0; s < splits; ++s) { off[s] = tot; tot += cand_n[q * splits + s]; }
        off[splits] = tot;
    }
    __syncthreads();
    const int total = off[splits];
    // compact the per-split lists into si (fixed order: split, then slot)
    for (int s = 0; s < splits; ++s) {
        const int n = off[s + 1] - off[s];
        for (int t = threadIdx.x; t < n; t += blockDim.x) si[off[s] + t] = cand_i[(q * splits + s) * TK_CAP + t];
    }
    __syncthreads();
    (void)cand_v;
    const float *qv = query + q * dim;
    for (int c = warp; c < total; c += 4) {
        const float *e = corpus + static_cast<int64_t>(si[c]) * dim;
        double d = 0.0;
        for (int t = lane; t < dim; t += 32) d = fma(static_cast<double>(qv[t]), static_cast<double>(e[t]), d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) sv[c] = d;
    }
    int np2 = 32;
    while (np2 < total) np2 <<= 1;
    __syncthreads();
    for (int t = threadIdx.x; t < np2; t += blockDim.x)
        if (t >= total) { sv[t] = -INFINITY; si[t] = 0x7fffffff; }
    __syncthreads();
    tk_bitonic(sv, si, np2, static_cast<int>(threadIdx.x), static_cast<int>(blockDim.x), [] { __syncthreads(); });
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const bool ok = t < total;
        out_scores[q * k + t] = ok ? sv[t] : -INFINITY;
        out_idx[q * k + t] = ok ? static_cast<int64_t>(si[t]) + row_offset : -1;
    }
}

// global merge of W per-shard top-k lists (scores fp64, global rows int64)
__global__ void __launch_bounds__(128)
topk_merge_kernel(const double *__restrict__ scores, const int64_t *__restrict__ idx, int W, int64_t Bq, int k,
                  double *__restrict__ out_scores, int64_t *__restrict__ out_idx) {
    __shared__ double sv[TK_STAGE2_MAX];
    __shared__ int64_t si[TK_STAGE2_MAX];
    const int64_t q = blockIdx.x;
    const int total = W * k;
    int np2 = 32;
    while (np2 < total) np2 <<= 1;
    for (int t = threadIdx.x; t < np2; t += blockDim.x) {
        if (t < total) {
            const int w = t / k, j = t - w * k;
            const int64_t src = (static_cast<int64_t>(w) * Bq + q) * k + j;
            const int64_t id = idx[src];
            sv[t] = (id < 0) ? -INFINITY : scores[src];
            si[t] = (id < 0) ? INT64_MAX : id;
        } else {
            sv[t] = -INFINITY;
            si[t] = INT64_MAX;
        }
    }
    __syncthreads();
    tk_bitonic(sv, si, np2, static_cast<int>(threadIdx.x), static_cast<int>(blockDim.x), [] { __syncthreads(); });
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const bool ok = si[t] != INT64_MAX;
        out_scores[q * k + t] = ok ? sv[t] : -INFINITY;
        out_idx[q * k + t] = ok ? si[t] : -1;
    }
}

struct TkPlan {
    int q_blocks, tiles, splits, kp;
};

static TkPlan tk_plan(int64_t Bq, int64_t Nc, int k) {
    TkPlan p;
    p.q_blocks = static_cast<int>((Bq + TK_T - 1) / TK_T);
    p.tiles = static_cast<int>((Nc + TK_T - 1) / TK_T);
    p.kp = k + TK_MARGIN;
    int want = (2 * 148 + p.q_blocks - 1) / p.q_blocks;
    const int cap = TK_STAGE2_MAX / p.kp;
    if (want > cap) want = cap;
    if (want > p.tiles) want = p.tiles;
    if (want > 16) want = 16;
    if (want < 1) want = 1;
    p.splits = want;
    return p;
}

}  // namespace tt

extern "C" int tt_score_topk_workspace(int64_t n_query, int64_t n_corpus, int dim, int k, size_t *bytes_host) {
    using namespace tt;
    TT_CHECK_ARG(bytes_host && n_query > 0 && n_corpus > 0 && dim > 0 && k > 0, "bad size");
    const TkPlan p = tk_plan(n_query, n_corpus, k);
    const size_t lists = static_cast<size_t>(n_query) * p.splits;
    *bytes_host = lists * TK_CAP * (sizeof(float) + sizeof(int32_t)) + lists * sizeof(int32_t) + 4096;
    return 0;
}

extern "C" int tt_score_topk_f32(const float *query, int64_t n_query, const float *corpus, int64_t n_corpus, int dim,
                                 int k, int64_t row_offset, const int64_t *mask_offsets, const int64_t *mask_rows,
                                 double *out_scores, int64_t *out_idx, void *workspace, size_t workspace_bytes,
                                 void *stream) {
    using namespace tt;
    TT_CHECK_ARG(query && corpus && out_scores && out_idx && workspace, "null pointer");
    TT_CHECK_ARG(n_query > 0 && n_corpus > 0 && dim > 0 && k > 0, "non-positive size");
    TT_CHECK_ARG((mask_offsets == nullptr) == (mask_rows == nullptr), "mask_offsets / mask_rows mismatch");
    if (k > TK_MAX_K) { set_error("top-K supports k <= %d (got %d)", TK_MAX_K, k); return TT_E_UNSUPPORTED; }
    if (dim % 4 != 0 || dim > 256) { set_error("top-K fp32 path supports dim %% 4 == 0 and dim <= 256 (got %d)", dim); return TT_E_UNSUPPORTED; }
    if (n_corpus >= (int64_t(1) << 31)) { set_error("corpus shard must have < 2^31 rows"); return TT_E_UNSUPPORTED; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const TkPlan p = tk_plan(n_query, n_corpus, k);
    const size_t lists = static_cast<size_t>(n_query) * p.splits;
    Workspace ws(workspace, workspace_bytes);
    float *cand_v = ws.take<float>(lists * TK_CAP);
    int32_t *cand_i = ws.take<int32_t>(lists * TK_CAP);
    int32_t *cand_n = ws.take<int32_t>(lists);
    if (!ws.ok()) { set_error("top-K workspace too small: need %zu have %zu", ws.off, workspace_bytes); return TT_E_WORKSPACE; }
    const size_t smem = sizeof(float) * 2 * TK_T * (dim + 4) + 8 * TK_CAP * (sizeof(float) + sizeof(int32_t));
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(topk_stage1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(float) * 2 * TK_T * (256 + 4) +
                                                              8 * TK_CAP * (sizeof(float) + sizeof(int32_t))));
        if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(topk_stage1)");
        attr_set = true;
    }
    dim3 grid(p.q_blocks, p.splits);
    topk_stage1<<<grid, TK_THREADS, smem, st>>>(query, n_query, corpus, n_corpus, dim, p.kp, p.tiles, mask_offsets,
                                               mask_rows, cand_v, cand_i, cand_n);
    TT_LAUNCH_CHECK("topk_stage1");
    topk_stage2<<<static_cast<unsigned>(n_query), 128, 0, st>>>(query, corpus, dim, k, p.splits, row_offset, cand_v,
                                                                cand_i, cand_n, out_scores, out_idx);
    TT_LAUNCH_CHECK("topk_stage2");
    return 0;
}

extern "C" int tt_topk_merge(const double *scores, const int64_t *idx, int n_shards, int64_t n_query, int k,
                             double *out_scores, int64_t *out_idx, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(scores && idx && out_scores && out_idx, "null pointer");
    TT_CHECK_ARG(n_shards > 0 && n_query > 0 && k > 0, "non-positive size");
    if (static_cast<int64_t>(n_shards) * k > TK_STAGE2_MAX) {
        set_error("top-K merge supports n_shards * k <= %d", TK_STAGE2_MAX);
        return TT_E_UNSUPPORTED;
    }
    topk_merge_kernel<<<static_cast<unsigned>(n_query), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        scores, idx, n_shards, n_query, k, out_scores, out_idx);
    TT_LAUNCH_CHECK("topk_merge_kernel");
    return 0;
}
