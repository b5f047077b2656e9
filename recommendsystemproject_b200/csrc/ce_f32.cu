// Fused in-batch (+ hard-negative) softmax cross-entropy, exact fp32 SIMT path
// (kernel 3 of the hot path; the bf16 tcgen05/TMA path lives in ce_tc.cu).
//
// Replaces mm / div / eq / masked_fill / bmm / cat / log_softmax / nll_loss at
// TwoTowerModel.py:95-140 of the reference and their autograd.  No B x (B+H)
// tensor is ever written to HBM: the forward keeps an online (max, sum) per
// row; the backward recomputes 64x64 logit tiles from U and I and feeds them
// straight into the dU / dI products (flash-attention style).
//
// Tiling: CTA = 64 "x" rows x 64 "y" rows per step, 256 threads as 16x16,
// each thread a 4x4 micro-tile with rows ty+16i and columns tx+16j so that
// shared-memory reads with row stride D+4 floats are conflict-free.  The y
// range is split across blockIdx.y so small batches still fill 148 SMs; the
// splits are combined in a fixed order (deterministic).
#include "common.cuh"

namespace tt {

constexpr int CE_T = 64;        // tile edge
constexpr int CE_THREADS = 256;
constexpr float CE_MASK = -1e9f;  // TwoTowerModel.py:114

struct CeSplits {
    int x_blocks, y_tiles, splits;
};

static CeSplits ce_splits(int64_t nx, int64_t ny) {
    CeSplits s;
    s.x_blocks = static_cast<int>((nx + CE_T - 1) / CE_T);
    s.y_tiles = static_cast<int>((ny + CE_T - 1) / CE_T);
    int want = (2 * 148 + s.x_blocks - 1) / s.x_blocks;
    if (want < 1) want = 1;
    if (want > s.y_tiles) want = s.y_tiles;
    if (want < 1) want = 1;
    s.splits = want;
    return s;
}

// load 64 rows x D floats (zero-filled past n) into smem with stride D+4; returns NaN seen
__device__ __forceinline__ bool ce_load_tile(float *__restrict__ dst, const float *__restrict__ src, int64_t row0,
                                             int64_t n, int dim, int stride) {
    const int vpr = dim / 4;
    bool nan = false;
    for (int i = threadIdx.x; i < CE_T * vpr; i += CE_THREADS) {
        const int r = i / vpr, c = i - r * vpr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < n) v = __ldg(reinterpret_cast<const float4 *>(src + (row0 + r) * dim) + c);
        nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
        *reinterpret_cast<float4 *>(dst + r * stride + c * 4) = v;
    }
    return nan;
}

// acc[i][j] = <Xs[ty+16i], Ys[tx+16j]>
__device__ __forceinline__ void ce_tile_dot(const float *__restrict__ Xs, const float *__restrict__ Ys, int dim,
                                            int stride, int tx, int ty, float (&acc)[4][4]) {
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
}

// ---------------------------------------------------------------- forward
// grid (x_blocks, splits).  y = [item rows (diag block, masked) | pool rows].
__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_tiles(const float *__restrict__ user, const float *__restrict__ item, const int64_t *__restrict__ item_ids,
             const float *__restrict__ pool, int64_t B, int64_t H, int dim, float inv_temp, int tiles_item,
             int tiles_total, float *__restrict__ part_m, float *__restrict__ part_s, float *__restrict__ row_pos,
             int *__restrict__ nan_flags) {
    extern __shared__ __align__(16) float smem[];
    const int stride = dim + 4;
    float *Xs = smem;
    float *Ys = smem + CE_T * stride;
    __shared__ int64_t xid[CE_T], yid[CE_T];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t x0 = static_cast<int64_t>(blockIdx.x) * CE_T;

    bool nan_x = ce_load_tile(Xs, user, x0, B, dim, stride);
    if (item_ids != nullptr && threadIdx.x < CE_T) xid[threadIdx.x] = (x0 + threadIdx.x < B) ? item_ids[x0 + threadIdx.x] : -1;
    if (nan_x && blockIdx.y == 0) atomicOr(nan_flags, 1);

    float m[4], s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; s[i] = 0.f; }

    for (int t = blockIdx.y; t < tiles_total; t += gridDim.y) {
        const bool diag = t < tiles_item;
        const int64_t y0 = static_cast<int64_t>(diag ? t : t - tiles_item) * CE_T;
        const int64_t ny = diag ? B : H;
        __syncthreads();  // previous tile fully consumed
        const bool nan_y = ce_load_tile(Ys, diag ? item : pool, y0, ny, dim, stride);
        if (diag && item_ids != nullptr && threadIdx.x < CE_T)
            yid[threadIdx.x] = (y0 + threadIdx.x < B) ? item_ids[y0 + threadIdx.x] : -2;
        if (nan_y && blockIdx.x == 0) atomicOr(nan_flags, diag ? 2 : 4);
        __syncthreads();
        float acc[4][4];
        ce_tile_dot(Xs, Ys, dim, stride, tx, ty, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty + 16 * i;
            const int64_t gx = x0 + r;
            float z[4];
            float tmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = tx + 16 * j;
                const int64_t gy = y0 + c;
                float v = acc[i][j] * inv_temp;
                if (gy >= ny) v = -INFINITY;
                else if (diag) {
                    if (gx == gy) { if (gx < B) row_pos[gx] = v; }
                    else if (item_ids != nullptr && xid[r] == yid[c]) v = CE_MASK;
                }
                z[j] = v;
                tmax = fmaxf(tmax, v);
            }
            if (tmax > -INFINITY) {
                const float mn = fmaxf(m[i], tmax);
                float add = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) add += expf(z[j] - mn);  // exp(-inf) = 0 for out-of-range columns
                s[i] = s[i] * expf(m[i] - mn) + add;
                m[i] = mn;
            }
        }
    }
    // combine the 16 threads (tx) that share each row: lanes differing in the low 4 bits
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const float mo = __shfl_xor_sync(0xffffffffu, m[i], o);
            const float so = __shfl_xor_sync(0xffffffffu, s[i], o);
            const float mn = fmaxf(m[i], mo);
            const float a = (m[i] == -INFINITY) ? 0.f : s[i] * expf(m[i] - mn);
            const float b = (mo == -INFINITY) ? 0.f : so * expf(mo - mn);
            s[i] = a + b;
            m[i] = mn;
        }
        const int64_t gx = x0 + ty + 16 * i;
        if (tx == 0 && gx < B) {
            part_m[static_cast<int64_t>(blockIdx.y) * B + gx] = m[i];
            part_s[static_cast<int64_t>(blockIdx.y) * B + gx] = s[i];
        }
    }
}

// one warp per row: merge splits + per-row hard negatives -> lse, row loss
__global__ void __launch_bounds__(256)
ce_fwd_finalize(const float *__restrict__ user, const float *__restrict__ hn_rows, int n_rowneg, int64_t B, int dim,
                float inv_temp, int splits, const float *__restrict__ part_m, const float *__restrict__ part_s,
                const float *__restrict__ row_pos, float *__restrict__ row_lse, float *__restrict__ row_loss,
                int *__restrict__ nan_flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t b = warp; b < B; b += n_warps) {
        float m = -INFINITY, s = 0.f;
        for (int k = 0; k < splits; ++k) {  // fixed order
            const float mk = part_m[static_cast<int64_t>(k) * B + b];
            const float sk = part_s[static_cast<int64_t>(k) * B + b];
            if (mk == -INFINITY) continue;
            const float mn = fmaxf(m, mk);
            s = ((m == -INFINITY) ? 0.f : s * expf(m - mn)) + sk * expf(mk - mn);
            m = mn;
        }
        bool nan = false;
        for (int n = 0; n < n_rowneg; ++n) {
            const float *h = hn_rows + (b * n_rowneg + n) * dim;
            float d = 0.f;
            for (int k = lane; k < dim; k += 32) {
                const float hv = h[k];
                nan |= (hv != hv);
                d = fmaf(user[b * dim + k], hv, d);
            }
            d = warp_sum(d) * inv_temp;
            const float mn = fmaxf(m, d);
            s = ((m == -INFINITY) ? 0.f : s * expf(m - mn)) + expf(d - mn);
            m = mn;
        }
        if (__any_sync(0xffffffffu, nan) && lane == 0) atomicOr(nan_flags, 4);
        if (lane == 0) {
            const float lse = m + logf(s);
            row_lse[b] = lse;
            row_loss[b] = lse - row_pos[b];
        }
    }
}

__global__ void __launch_bounds__(1024)
ce_mean_fixed_order(const float *__restrict__ x, int64_t n, float *__restrict__ out) {
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

// ---------------------------------------------------------------- backward
// One pass kernel, used three ways:
//   dU    : X = user, Y = item (diag) ; lse indexed by x
//   dU    : X = user, Y = pool        ; lse indexed by x
//   dI    : X = item, Y = user (diag) ; lse indexed by y   (TRANS)
//   dPool : X = pool, Y = user        ; lse indexed by y   (TRANS)
// G[x][y] = gscale * (exp(z - lse) - [diag && x == y]);  out[split][x][:] = sum_y G[x][y] * Y[y][:]
template <int NJ, bool TRANS>
__global__ void __launch_bounds__(CE_THREADS)
ce_bwd_pass(const float *__restrict__ X, const float *__restrict__ Y, int64_t nx, int64_t ny, int dim,
            const int64_t *__restrict__ ids, int diag, float inv_temp, const float *__restrict__ row_lse,
            const float *__restrict__ grad_loss, int64_t batch, int tiles_y, float *__restrict__ out_part) {
    extern __shared__ __align__(16) float smem[];
    const int stride = dim + 4;
    float *Xs = smem;
    float *Ys = smem + CE_T * stride;
    float *Gs = Ys + CE_T * stride;  // [64][68]
    constexpr int GSTR = CE_T + 4;
    __shared__ int64_t xid[CE_T], yid[CE_T];
    __shared__ float xl[CE_T], yl[CE_T];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t x0 = static_cast<int64_t>(blockIdx.x) * CE_T;
    const float gscale = (*grad_loss) / static_cast<float>(batch);

    ce_load_tile(Xs, X, x0, nx, dim, stride);
    if (threadIdx.x < CE_T) {
        const int64_t gx = x0 + threadIdx.x;
        if (diag && ids != nullptr) xid[threadIdx.x] = (gx < nx) ? ids[gx] : -1;
        if (!TRANS) xl[threadIdx.x] = (gx < nx) ? row_lse[gx] : 0.f;
    }
    float acc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

    for (int t = blockIdx.y; t < tiles_y; t += gridDim.y) {
        const int64_t y0 = static_cast<int64_t>(t) * CE_T;
        __syncthreads();
        ce_load_tile(Ys, Y, y0, ny, dim, stride);
        if (threadIdx.x < CE_T) {
            const int64_t gy = y0 + threadIdx.x;
            if (diag && ids != nullptr) yid[threadIdx.x] = (gy < ny) ? ids[gy] : -2;
            if (TRANS) yl[threadIdx.x] = (gy < ny) ? row_lse[gy] : 0.f;
        }
        __syncthreads();
        float sacc[4][4];
        ce_tile_dot(Xs, Ys, dim, stride, tx, ty, sacc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty + 16 * i;
            const int64_t gx = x0 + r;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = tx + 16 * j;
                const int64_t gy = y0 + c;
                float g = 0.f;
                if (gx < nx && gy < ny) {
                    const float z = sacc[i][j] * inv_temp;
                    const float lse = TRANS ? yl[c] : xl[r];
                    const bool masked = diag && ids != nullptr && gx != gy && xid[r] == yid[c];
                    float p = masked ? 0.f : expf(z - lse);
                    if (diag && gx == gy) p -= 1.0f;
                    g = gscale * p;
                }
                Gs[r * GSTR + c] = g;
            }
        }
        __syncthreads();
        // acc[i][jj] += sum_y Gs[ty+16i][y] * Ys[y][tx+16jj]
        for (int y = 0; y < CE_T; y += 4) {
            float4 gv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) gv[i] = *reinterpret_cast<const float4 *>(Gs + (ty + 16 * i) * GSTR + y);
#pragma unroll
            for (int yy = 0; yy < 4; ++yy) {
                float yv[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int d = tx + 16 * j;
                    yv[j] = (d < dim) ? Ys[(y + yy) * stride + d] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float gg = (yy == 0) ? gv[i].x : (yy == 1) ? gv[i].y : (yy == 2) ? gv[i].z : gv[i].w;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(gg, yv[j], acc[i][j]);
                }
            }
        }
    }
    float *out = out_part + static_cast<int64_t>(blockIdx.y) * nx * dim;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gx = x0 + ty + 16 * i;
        if (gx >= nx) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int d = tx + 16 * j;
            if (d < dim) out[gx * dim + d] = acc[i][j];
        }
    }
}

// per-row hard negatives: dHN[b,n,:] = g*U_b/T ; extra[b,:] = sum_n g*HN[b,n,:]  (scaled by 1/T later)
__global__ void __launch_bounds__(256)
ce_bwd_hn_rows(const float *__restrict__ user, const float *__restrict__ hn_rows, int n_rowneg, int64_t B, int dim,
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

// out[i] = scale * sum_{k < slabs} part[k][i]   (fixed order)
__global__ void ce_reduce_slabs(const float *__restrict__ part, int slabs, int64_t n, float scale,
                                float *__restrict__ out) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float a = 0.f;
        for (int k = 0; k < slabs; ++k) a += part[static_cast<int64_t>(k) * n + i];
        out[i] = a * scale;
    }
}

static inline unsigned ce_grid(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b);
}

static int ce_check_dims(int64_t batch, int dim) {
    if (batch <= 0 || dim <= 0) { set_error("bad argument: non-positive batch/dim"); return TT_E_BADARG; }
    if (dim % 4 != 0 || dim > 256) {
        set_error("fused CE fp32 path supports dim %% 4 == 0 and dim <= 256 (got %d)", dim);
        return TT_E_UNSUPPORTED;
    }
    return 0;
}

template <bool TRANS>
static int launch_bwd_pass(const float *X, const float *Y, int64_t nx, int64_t ny, int dim, const int64_t *ids,
                           int diag, float inv_temp, const float *row_lse, const float *grad_loss, int64_t batch,
                           const CeSplits &sp, float *out_part, cudaStream_t st) {
    const size_t smem = sizeof(float) * (2 * CE_T * (dim + 4) + CE_T * (CE_T + 4));
    dim3 grid(sp.x_blocks, sp.splits);
    const int nj = (dim + 15) / 16;
#define TT_BWD(NJ)                                                                                               \
    do {                                                                                                         \
        static bool attr_set = false;                                                                            \
        if (!attr_set) {                                                                                         \
            cudaError_t e = cudaFuncSetAttribute(ce_bwd_pass<NJ, TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 static_cast<int>(sizeof(float) * (2 * CE_T * (NJ * 16 + 4) + CE_T * (CE_T + 4)))); \
            if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(ce_bwd_pass)");                    \
            attr_set = true;                                                                                     \
        }                                                                                                        \
        ce_bwd_pass<NJ, TRANS><<<grid, CE_THREADS, smem, st>>>(X, Y, nx, ny, dim, ids, diag, inv_temp, row_lse,  \
                                                              grad_loss, batch, sp.y_tiles, out_part);          \
    } while (0)
    if (nj <= 1) TT_BWD(1);
    else if (nj <= 2) TT_BWD(2);
    else if (nj <= 4) TT_BWD(4);
    else if (nj <= 8) TT_BWD(8);
    else TT_BWD(16);
#undef TT_BWD
    TT_LAUNCH_CHECK("ce_bwd_pass");
    return 0;
}

}  // namespace tt

extern "C" int tt_ce_workspace(int64_t batch, int64_t pool, int n_rowneg, int dim, size_t *bytes_host) {
    using namespace tt;
    TT_CHECK_ARG(bytes_host && batch > 0 && pool >= 0 && n_rowneg >= 0 && dim > 0, "bad size");
    const CeSplits fwd = ce_splits(batch, batch + pool + CE_T);
    const CeSplits s_ui = ce_splits(batch, batch);
    const CeSplits s_up = ce_splits(batch, pool > 0 ? pool : 1);
    const CeSplits s_pu = ce_splits(pool > 0 ? pool : 1, batch);
    size_t fwd_b = sizeof(float) * (2 * static_cast<size_t>(fwd.splits + 1) * batch + batch) + 1024;
    const size_t du_slabs = s_ui.splits + (pool > 0 ? s_up.splits : 0) + 1;
    size_t bwd_b = sizeof(float) * (du_slabs * batch * dim + static_cast<size_t>(s_ui.splits) * batch * dim +
                                    (pool > 0 ? static_cast<size_t>(s_pu.splits) * pool * dim : 0)) + 4096;
    *bytes_host = (fwd_b > bwd_b ? fwd_b : bwd_b) + 1024;
    return 0;
}

extern "C" int tt_ce_fwd_f32(const float *user, const float *item, const int64_t *item_ids, const float *hn_rows,
                             int n_rowneg, const float *pool, int64_t pool_rows, int64_t batch, int dim,
                             float inv_temp, float *loss, float *row_lse, float *row_pos, int *nan_flags,
                             void *workspace, size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(user && item && loss && row_lse && row_pos && nan_flags && workspace, "null pointer");
    TT_CHECK_ARG((hn_rows != nullptr) == (n_rowneg > 0), "hn_rows / n_rowneg mismatch");
    TT_CHECK_ARG((pool != nullptr) == (pool_rows > 0), "pool / pool_rows mismatch");
    if (int rc = ce_check_dims(batch, dim)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tiles_item = static_cast<int>((batch + CE_T - 1) / CE_T);
    const int tiles_pool = static_cast<int>((pool_rows + CE_T - 1) / CE_T);
    const int tiles_total = tiles_item + tiles_pool;
    CeSplits sp = ce_splits(batch, static_cast<int64_t>(tiles_total) * CE_T);
    Workspace ws(workspace, workspace_bytes);
    float *part_m = ws.take<float>(static_cast<size_t>(sp.splits) * batch);
    float *part_s = ws.take<float>(static_cast<size_t>(sp.splits) * batch);
    float *row_loss = ws.take<float>(batch);
    if (!ws.ok()) { set_error("ce_fwd workspace too small: need %zu have %zu", ws.off, workspace_bytes); return TT_E_WORKSPACE; }
    const size_t smem = sizeof(float) * 2 * CE_T * (dim + 4);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ce_fwd_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(float) * 2 * CE_T * (256 + 4)));
        if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(ce_fwd_tiles)");
        attr_set = true;
    }
    dim3 grid(sp.x_blocks, sp.splits);
    ce_fwd_tiles<<<grid, CE_THREADS, smem, st>>>(user, item, item_ids, pool, batch, pool_rows, dim, inv_temp,
                                                tiles_item, tiles_total, part_m, part_s, row_pos, nan_flags);
    TT_LAUNCH_CHECK("ce_fwd_tiles");
    ce_fwd_finalize<<<ce_grid(batch * 32, 256), 256, 0, st>>>(user, hn_rows, n_rowneg, batch, dim, inv_temp, sp.splits,
                                                             part_m, part_s, row_pos, row_lse, row_loss, nan_flags);
    TT_LAUNCH_CHECK("ce_fwd_finalize");
    ce_mean_fixed_order<<<1, 1024, 0, st>>>(row_loss, batch, loss);
    TT_LAUNCH_CHECK("ce_mean_fixed_order");
    return 0;
}

extern "C" int tt_ce_bwd_f32(const float *user, const float *item, const int64_t *item_ids, const float *hn_rows,
                             int n_rowneg, const float *pool, int64_t pool_rows, int64_t batch, int dim,
                             float inv_temp, const float *row_lse, const float *grad_loss, float *d_user,
                             float *d_item, float *d_hn_rows, float *d_pool, void *workspace,
                             size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(user && item && row_lse && grad_loss && d_user && d_item && workspace, "null pointer");
    TT_CHECK_ARG((hn_rows != nullptr) == (n_rowneg > 0), "hn_rows / n_rowneg mismatch");
    TT_CHECK_ARG((pool != nullptr) == (pool_rows > 0), "pool / pool_rows mismatch");
    TT_CHECK_ARG(pool == nullptr || d_pool != nullptr, "d_pool required with pool");
    if (int rc = ce_check_dims(batch, dim)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CeSplits s_ui = ce_splits(batch, batch);
    const CeSplits s_up = ce_splits(batch, pool_rows > 0 ? pool_rows : 1);
    const CeSplits s_pu = ce_splits(pool_rows > 0 ? pool_rows : 1, batch);
    const int du_slabs = s_ui.splits + (pool ? s_up.splits : 0) + (hn_rows ? 1 : 0);
    Workspace ws(workspace, workspace_bytes);
    float *du_part = ws.take<float>(static_cast<size_t>(du_slabs) * batch * dim);
    float *di_part = ws.take<float>(static_cast<size_t>(s_ui.splits) * batch * dim);
    float *dp_part = pool ? ws.take<float>(static_cast<size_t>(s_pu.splits) * pool_rows * dim) : nullptr;
    if (!ws.ok()) { set_error("ce_bwd workspace too small: need %zu have %zu", ws.off, workspace_bytes); return TT_E_WORKSPACE; }
    const int64_t bd = batch * dim;
    int rc;
    // dU: in-batch block, pool block, per-row negatives
    if ((rc = launch_bwd_pass<false>(user, item, batch, batch, dim, item_ids, 1, inv_temp, row_lse, grad_loss, batch,
                                     s_ui, du_part, st))) return rc;
    int slab = s_ui.splits;
    if (pool) {
        if ((rc = launch_bwd_pass<false>(user, pool, batch, pool_rows, dim, nullptr, 0, inv_temp, row_lse, grad_loss,
                                         batch, s_up, du_part + static_cast<size_t>(slab) * bd, st))) return rc;
        slab += s_up.splits;
    }
    if (hn_rows) {
        ce_bwd_hn_rows<<<ce_grid(batch * 32, 256), 256, 0, st>>>(user, hn_rows, n_rowneg, batch, dim, inv_temp, row_lse,
                                                                grad_loss, d_hn_rows, du_part + static_cast<size_t>(slab) * bd);
        TT_LAUNCH_CHECK("ce_bwd_hn_rows");
        slab += 1;
    }
    ce_reduce_slabs<<<ce_grid(bd, 256), 256, 0, st>>>(du_part, slab, bd, inv_temp, d_user);
    TT_LAUNCH_CHECK("ce_reduce_slabs(dU)");
    // dI
    if ((rc = launch_bwd_pass<true>(item, user, batch, batch, dim, item_ids, 1, inv_temp, row_lse, grad_loss, batch,
                                    s_ui, di_part, st))) return rc;
    ce_reduce_slabs<<<ce_grid(bd, 256), 256, 0, st>>>(di_part, s_ui.splits, bd, inv_temp, d_item);
    TT_LAUNCH_CHECK("ce_reduce_slabs(dI)");
    if (pool) {
        if ((rc = launch_bwd_pass<true>(pool, user, pool_rows, batch, dim, nullptr, 0, inv_temp, row_lse, grad_loss,
                                        batch, s_pu, dp_part, st))) return rc;
        ce_reduce_slabs<<<ce_grid(pool_rows * dim, 256), 256, 0, st>>>(dp_part, s_pu.splits, pool_rows * dim, inv_temp, d_pool);
        TT_LAUNCH_CHECK("ce_reduce_slabs(dPool)");
    }
    return 0;
}
