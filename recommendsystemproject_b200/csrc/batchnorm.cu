// BatchNorm1d (training mode) with the MLP block's ReLU and Dropout fused in, and with statistics that can span
// several ranks (data-parallel towers: the reference normalises over the WHOLE batch, GenericTower.py:234 and
// Tower.py:18, so ranks exchange per-channel (mean, M2, n) between the two forward kernels and the two per-channel
// gradient sums between the two backward kernels).
//
// Replaces native_batch_norm (+ relu + dropout and their backward) at GenericTower.py:234 / Tower.py:16-21.
//   y = dropout(relu(gamma[c % P] * (x - mean_c) * rstd_c + beta[c % P]))        (relu / dropout optional)
// P = param_period: rows of x may hold G slabs side by side ([B, G*C] view of the positive + hard-negative item pass,
// TwoTowerModel.py:54-60: each slab normalises with its OWN batch statistics); P = cols for a plain layer.
//
// Numerics: per 128-column block and row chunk the kernel accumulates sum(x - K) and sum((x - K)^2) with the shift
// K = first row of the chunk (no cancellation), turns them into (mean, M2) and merges chunks -- and ranks -- in a fixed
// order with Chan's parallel formula: deterministic, and as accurate as torch's Welford pass.
// Dropout masks: the counter-based hash of (seed, call id, element) used by encoder_small.cu, rebuilt in the backward.
// HBM-bound: forward reads x twice and writes y once, backward reads x and dy twice and writes dx once.
#include "common.cuh"

namespace tt {

constexpr int BN_TX = 32;    // float4 column groups per block (128 columns)
constexpr int BN_TY = 8;     // row lanes per block
// at or below: statistics in one launch (one block per 128 columns).  Measured on the B = 512 step (torch profiler,
// graph replay): 9 us (forward) / 38 us (backward) per launch against ~3 + 3 us for the chunked pair -- a [512, 2816]
// slab view gives the single launch only 22 blocks -- so the one-launch path is kept for tiny inputs only.
constexpr int BN_SMALL_ROWS = 64;

__device__ __forceinline__ uint32_t bn_hash(uint64_t seed, uint64_t call_id, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (call_id + 1) + idx * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return static_cast<uint32_t>(z >> 32);
}
// keep-probability test for element idx (one hash per element)
__device__ __forceinline__ bool bn_keep(uint64_t seed, uint64_t call_id, uint64_t idx, uint32_t thresh) {
    return bn_hash(seed, call_id, idx) >= thresh;
}

// ---- forward statistics -------------------------------------------------------------------------------------------
// partial[chunk][3][cols]: shift K, sum(x - K), sum((x - K)^2) over the chunk's rows
__global__ void __launch_bounds__(BN_TX *BN_TY)
bn_stats_partial(const float *__restrict__ x, int64_t rows, int cols, int64_t stride, int rows_per_chunk,
                 float *__restrict__ partial) {
    __shared__ float4 sh[2][BN_TY][BN_TX];
    const int c4 = blockIdx.x * BN_TX + threadIdx.x;     // float4 column index
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk;
    const int64_t r1 = min(rows, r0 + rows_per_chunk);
    const bool ok = c4 * 4 < cols;
    float4 K = make_float4(0.f, 0.f, 0.f, 0.f), s = K, q = K;
    if (ok) {
        K = __ldg(reinterpret_cast<const float4 *>(x + r0 * stride) + c4);
        for (int64_t r = r0 + threadIdx.y; r < r1; r += BN_TY) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * stride) + c4);
            const float dx = v.x - K.x, dy = v.y - K.y, dz = v.z - K.z, dw = v.w - K.w;
            s.x += dx; s.y += dy; s.z += dz; s.w += dw;
            q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
        }
    }
    sh[0][threadIdx.y][threadIdx.x] = s;
    sh[1][threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && ok) {
        for (int k = 1; k < BN_TY; ++k) {   // fixed order
            const float4 a = sh[0][k][threadIdx.x], b = sh[1][k][threadIdx.x];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
        }
        float *p = partial + static_cast<int64_t>(blockIdx.y) * 3 * cols;
        *(reinterpret_cast<float4 *>(p) + c4) = K;
        *(reinterpret_cast<float4 *>(p + cols) + c4) = s;
        *(reinterpret_cast<float4 *>(p + 2 * cols) + c4) = q;
    }
}

// The affine form of the normalisation, spelled with explicit roundings: the backward RECOMPUTES the ReLU mask from x, so
// forward and backward must evaluate bit-identical expressions (left to the compiler, one side contracted
// beta - a * mean into an FMA and the other did not: one element in ~10^6 sat within an ulp of zero and flipped).
__device__ __forceinline__ float bn_scale(float gamma, float rstd) { return __fmul_rn(gamma, rstd); }
__device__ __forceinline__ float bn_shift(float a, float mean, float beta) { return __fsub_rn(beta, __fmul_rn(a, mean)); }
__device__ __forceinline__ float bn_affine(float a, float b, float x) { return __fmaf_rn(a, x, b); }

__device__ __forceinline__ void chan_merge(float &n, float &mean, float &m2, float nk, float mk, float m2k) {
    if (nk <= 0.f) return;
    const float tot = n + nk;
    const float d = mk - mean;
    mean += d * (nk / tot);
    m2 += m2k + d * d * (n * nk / tot);
    n = tot;
}

// stats[0..cols) = mean, [cols..2cols) = M2, stats[2*cols] = n  (this rank's rows)
// 32 columns x 32 chunk lanes per block: lane j merges chunks j, j+32, ... (Chan), the 32 results are merged in lane
// order.  The loads of four chunks are issued before their (dependent, division-heavy) merges: with 8 lanes and one
// chunk in flight this kernel was a 28 us chain of L2 latencies (ncu launch list, 512 chunks).
constexpr int BN_FL = 32;   // chunk lanes of the final kernels
__global__ void __launch_bounds__(32 * BN_FL)
bn_stats_final(const float *__restrict__ partial, int n_chunks, int64_t rows, int rows_per_chunk, int cols,
               float *__restrict__ stats) {
    __shared__ float sn[BN_FL][32], sm[BN_FL][32], s2[BN_FL][32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float n = 0.f, mean = 0.f, m2 = 0.f;
    if (c < cols) {
        for (int k0 = threadIdx.y; k0 < n_chunks; k0 += 4 * BN_FL) {
            float kk[4], sd[4], sq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * BN_FL;
                if (k < n_chunks) {
                    const float *p = partial + static_cast<int64_t>(k) * 3 * cols;
                    kk[u] = p[c]; sd[u] = p[cols + c]; sq[u] = p[2 * cols + c];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * BN_FL;
                if (k < n_chunks) {
                    const float nk = static_cast<float>(min(static_cast<int64_t>(rows_per_chunk), rows - static_cast<int64_t>(k) * rows_per_chunk));
                    chan_merge(n, mean, m2, nk, kk[u] + sd[u] / nk, fmaxf(sq[u] - sd[u] * sd[u] / nk, 0.f));
                }
            }
        }
    }
    sn[threadIdx.y][threadIdx.x] = n; sm[threadIdx.y][threadIdx.x] = mean; s2[threadIdx.y][threadIdx.x] = m2;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        for (int j = 1; j < BN_FL; ++j) chan_merge(n, mean, m2, sn[j][threadIdx.x], sm[j][threadIdx.x], s2[j][threadIdx.x]);
        stats[c] = mean;
        stats[cols + c] = m2;
        if (c == 0) stats[2 * cols] = n;
    }
}

// few rows (the B = 512 step of the shipped config): partial + final in ONE launch, one block per 128 columns
__global__ void __launch_bounds__(BN_TX *BN_TY)
bn_stats_small(const float *__restrict__ x, int64_t rows, int cols, int64_t stride, float *__restrict__ stats) {
    __shared__ float4 sh[2][BN_TY][BN_TX];
    const int c4 = blockIdx.x * BN_TX + threadIdx.x;
    const bool ok = c4 * 4 < cols;
    float4 K = make_float4(0.f, 0.f, 0.f, 0.f), s = K, q = K;
    if (ok) {
        K = __ldg(reinterpret_cast<const float4 *>(x) + c4);
        for (int64_t r = threadIdx.y; r < rows; r += BN_TY) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * stride) + c4);
            const float dx = v.x - K.x, dy = v.y - K.y, dz = v.z - K.z, dw = v.w - K.w;
            s.x += dx; s.y += dy; s.z += dz; s.w += dw;
            q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
        }
    }
    sh[0][threadIdx.y][threadIdx.x] = s;
    sh[1][threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && ok) {
        for (int k = 1; k < BN_TY; ++k) {
            const float4 a = sh[0][k][threadIdx.x], b = sh[1][k][threadIdx.x];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
        }
        const float n = static_cast<float>(rows);
        const float sv[4] = {s.x, s.y, s.z, s.w}, qv[4] = {q.x, q.y, q.z, q.w}, kv[4] = {K.x, K.y, K.z, K.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            stats[c4 * 4 + e] = kv[e] + sv[e] / n;
            stats[cols + c4 * 4 + e] = fmaxf(qv[e] - sv[e] * sv[e] / n, 0.f);
        }
        if (c4 == 0) stats[2 * cols] = n;
    }
}

// ---- forward apply -----------------------------------------------------------------------------------------------
// prologue: merge the n_ranks stats blocks (rank order) for this block's columns, publish mean / rstd (block row 0),
// update the running statistics; then normalise this block's rows
__global__ void __launch_bounds__(BN_TX *BN_TY)
bn_apply_kernel(const float *__restrict__ x, int64_t rows, int cols, int64_t stride, int rows_per_chunk,
                const float *__restrict__ stats_all, int n_ranks, const float *__restrict__ gamma,
                const float *__restrict__ beta, int period, float eps, int relu, float dropout_p,
                const int64_t *__restrict__ seed_dev, int64_t call_id, float *__restrict__ y, int64_t y_stride,
                float *__restrict__ save_mean, float *__restrict__ save_rstd, float *__restrict__ batch_var_unbiased,
                float *__restrict__ running_mean, float *__restrict__ running_var, float momentum,
                int64_t *__restrict__ num_batches) {
    __shared__ float s_mean[BN_TX * 4], s_rstd[BN_TX * 4];
    const int tid = threadIdx.y * BN_TX + threadIdx.x;
    const int col0 = blockIdx.x * BN_TX * 4;
    if (tid < BN_TX * 4) {
        const int c = col0 + tid;
        if (c < cols) {
            float n = 0.f, mean = 0.f, m2 = 0.f;
            const int sstride = 2 * cols + 1;
            for (int r = 0; r < n_ranks; ++r) {
                const float *st = stats_all + static_cast<int64_t>(r) * sstride;
                chan_merge(n, mean, m2, st[2 * cols], st[c], st[cols + c]);
            }
            const float var = m2 / n;
            const float rstd = rsqrtf(var + eps);
            s_mean[tid] = mean;
            s_rstd[tid] = rstd;
            if (blockIdx.y == 0) {
                save_mean[c] = mean;
                save_rstd[c] = rstd;
                const float unb = n > 1.f ? m2 / (n - 1.f) : var;
                if (batch_var_unbiased) batch_var_unbiased[c] = unb;
                if (running_mean) {   // plain layer (period == cols): r <- (1 - m) r + m stat, unbiased variance
                    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
                }
                if (c == 0 && num_batches) *num_batches += 1;
            }
        }
    }
    __syncthreads();
    const int c4 = blockIdx.x * BN_TX + threadIdx.x;
    if (c4 * 4 >= cols) return;
    const int cl = threadIdx.x * 4;
    const int pc = (c4 * 4) % period;   // period % 4 == 0 (checked on the host): a float4 never straddles slabs
    const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + pc));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(beta + pc));
    const float a0 = bn_scale(g.x, s_rstd[cl]), a1 = bn_scale(g.y, s_rstd[cl + 1]), a2 = bn_scale(g.z, s_rstd[cl + 2]),
                a3 = bn_scale(g.w, s_rstd[cl + 3]);
    const float b0 = bn_shift(a0, s_mean[cl], b.x), b1 = bn_shift(a1, s_mean[cl + 1], b.y), b2 = bn_shift(a2, s_mean[cl + 2], b.z),
                b3 = bn_shift(a3, s_mean[cl + 3], b.w);
    const bool drop = dropout_p > 0.f && seed_dev != nullptr;
    const uint64_t seed = drop ? static_cast<uint64_t>(*seed_dev) : 0;
    const uint32_t thresh = drop ? static_cast<uint32_t>(fminf(dropout_p, 0.999999f) * 4294967296.0f) : 0u;
    const float keep_scale = drop ? 1.0f / (1.0f - dropout_p) : 1.0f;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk;
    const int64_t r1 = min(rows, r0 + rows_per_chunk);
    for (int64_t r = r0 + threadIdx.y; r < r1; r += BN_TY) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * stride) + c4);
        float o[4] = {bn_affine(a0, b0, v.x), bn_affine(a1, b1, v.y), bn_affine(a2, b2, v.z), bn_affine(a3, b3, v.w)};
        if (relu) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
        }
        if (drop) {
            const uint64_t base = static_cast<uint64_t>(r) * cols + c4 * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = bn_keep(seed, call_id, base + e, thresh) ? o[e] * keep_scale : 0.f;
        }
        *(reinterpret_cast<float4 *>(y + r * y_stride) + c4) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---- backward ----------------------------------------------------------------------------------------------------
// g = dy * dropout mask / (1-p) * [bn(x) > 0];  partial[chunk][2][cols] = sum g, sum g * xhat
template <bool APPLY>
__global__ void __launch_bounds__(BN_TX *BN_TY)
bn_bwd_kernel(const float *__restrict__ dy, int64_t dy_stride, const float *__restrict__ x, int64_t rows, int cols,
              int64_t stride, int rows_per_chunk, const float *__restrict__ save_mean, const float *__restrict__ save_rstd,
              const float *__restrict__ gamma, const float *__restrict__ beta, int period, int relu, float dropout_p,
              const int64_t *__restrict__ seed_dev, int64_t call_id, float *__restrict__ partial,
              const float *__restrict__ sums_global, int n_ranks, float inv_total_rows, float *__restrict__ dx,
              int64_t dx_stride) {
    __shared__ float4 sh[2][BN_TY][BN_TX];
    const int c4 = blockIdx.x * BN_TX + threadIdx.x;
    const bool ok = c4 * 4 < cols;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk;
    const int64_t r1 = min(rows, r0 + rows_per_chunk);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (ok) {
        const int pc = (c4 * 4) % period;
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + pc));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(beta + pc));
        const float4 mu = __ldg(reinterpret_cast<const float4 *>(save_mean) + c4);
        const float4 rs = __ldg(reinterpret_cast<const float4 *>(save_rstd) + c4);
        const float gm[4] = {g.x, g.y, g.z, g.w}, bt[4] = {b.x, b.y, b.z, b.w};
        const float m[4] = {mu.x, mu.y, mu.z, mu.w}, rr[4] = {rs.x, rs.y, rs.z, rs.w};
        float k1[4] = {0.f, 0.f, 0.f, 0.f}, k2[4] = {0.f, 0.f, 0.f, 0.f};
        if (APPLY) {
            // the per-rank sums blocks [n_ranks][2 * cols] added in rank order (one block: the caller's global sums)
            float4 t1 = make_float4(0.f, 0.f, 0.f, 0.f), t2 = t1;
            for (int rk = 0; rk < n_ranks; ++rk) {
                const float *sg = sums_global + static_cast<int64_t>(rk) * 2 * cols;
                const float4 a = __ldg(reinterpret_cast<const float4 *>(sg) + c4);
                const float4 b2 = __ldg(reinterpret_cast<const float4 *>(sg + cols) + c4);
                t1.x += a.x; t1.y += a.y; t1.z += a.z; t1.w += a.w;
                t2.x += b2.x; t2.y += b2.y; t2.z += b2.z; t2.w += b2.w;
            }
            k1[0] = t1.x * inv_total_rows; k1[1] = t1.y * inv_total_rows; k1[2] = t1.z * inv_total_rows; k1[3] = t1.w * inv_total_rows;
            k2[0] = t2.x * inv_total_rows; k2[1] = t2.y * inv_total_rows; k2[2] = t2.z * inv_total_rows; k2[3] = t2.w * inv_total_rows;
        }
        const bool drop = dropout_p > 0.f && seed_dev != nullptr;
        const uint64_t seed = drop ? static_cast<uint64_t>(*seed_dev) : 0;
        const uint32_t thresh = drop ? static_cast<uint32_t>(fminf(dropout_p, 0.999999f) * 4294967296.0f) : 0u;
        const float keep_scale = drop ? 1.0f / (1.0f - dropout_p) : 1.0f;
        float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
        for (int64_t r = r0 + threadIdx.y; r < r1; r += BN_TY) {
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + r * stride) + c4);
            const float4 dv = __ldg(reinterpret_cast<const float4 *>(dy + r * dy_stride) + c4);
            const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
            float ge[4] = {dv.x, dv.y, dv.z, dv.w};
            float o[4];
            const uint64_t base = static_cast<uint64_t>(r) * cols + c4 * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xh = (xe[e] - m[e]) * rr[e];
                if (drop) ge[e] = bn_keep(seed, call_id, base + e, thresh) ? ge[e] * keep_scale : 0.f;
                if (relu) {
                    // bit-identical to the forward's value (bn_scale / bn_shift / bn_affine)
                    const float a = bn_scale(gm[e], rr[e]);
                    if (!(bn_affine(a, bn_shift(a, m[e], bt[e]), xe[e]) > 0.f)) ge[e] = 0.f;
                }
                if (APPLY) o[e] = gm[e] * rr[e] * (ge[e] - k1[e] - xh * k2[e]);
                else { a1[e] += ge[e]; a2[e] = fmaf(ge[e], xh, a2[e]); }
            }
            if (APPLY) *(reinterpret_cast<float4 *>(dx + r * dx_stride) + c4) = make_float4(o[0], o[1], o[2], o[3]);
        }
        s1 = make_float4(a1[0], a1[1], a1[2], a1[3]);
        s2 = make_float4(a2[0], a2[1], a2[2], a2[3]);
    }
    if (APPLY) return;
    sh[0][threadIdx.y][threadIdx.x] = s1;
    sh[1][threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && ok) {
        for (int k = 1; k < BN_TY; ++k) {
            const float4 a = sh[0][k][threadIdx.x], b = sh[1][k][threadIdx.x];
            s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
            s2.x += b.x; s2.y += b.y; s2.z += b.z; s2.w += b.w;
        }
        float *p = partial + static_cast<int64_t>(blockIdx.y) * 2 * cols;
        *(reinterpret_cast<float4 *>(p) + c4) = s1;
        *(reinterpret_cast<float4 *>(p + cols) + c4) = s2;
    }
}

// sums[0..cols) = sum g, sums[cols..2cols) = sum g xhat: 32 columns x 32 chunk lanes per block, fixed order
__global__ void __launch_bounds__(32 * BN_FL)
bn_bwd_final(const float *__restrict__ partial, int n_chunks, int cols, float *__restrict__ sums) {
    __shared__ float sa[BN_FL][32], sb[BN_FL][32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float a = 0.f, b = 0.f;
    if (c < cols) {
        for (int k0 = threadIdx.y; k0 < n_chunks; k0 += 4 * BN_FL) {
            float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * BN_FL;
                if (k < n_chunks) {
                    va[u] = partial[static_cast<int64_t>(k) * 2 * cols + c];
                    vb[u] = partial[static_cast<int64_t>(k) * 2 * cols + cols + c];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { a += va[u]; b += vb[u]; }
        }
    }
    sa[threadIdx.y][threadIdx.x] = a; sb[threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        for (int j = 1; j < BN_FL; ++j) { a += sa[j][threadIdx.x]; b += sb[j][threadIdx.x]; }
        sums[c] = a;
        sums[cols + c] = b;
    }
}

// parameter gradients from the LOCAL sums (a data-parallel caller all-reduces them with the other dense gradients):
// dgamma[c % P] (+)= sum g xhat, dbeta (+)= sum g
__global__ void bn_param_grads(const float *__restrict__ sums, int cols, int period, float *__restrict__ dgamma,
                               float *__restrict__ dbeta, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= period) return;
    float a = 0.f, b = 0.f;
    for (int g = c; g < cols; g += period) { a += sums[g]; b += sums[cols + g]; }   // slabs in order
    if (accumulate) { dbeta[c] += a; dgamma[c] += b; }
    else { dbeta[c] = a; dgamma[c] = b; }
}

static inline int bn_rows_per_chunk(int64_t rows, int cols) {
    // enough blocks for ~2 waves, chunks of at least 8 rows; depends on the shape only (reproducible order)
    const int col_blocks = (cols + BN_TX * 4 - 1) / (BN_TX * 4);
    int r = 4096;
    while (r > 8 && ((rows + r - 1) / r) * col_blocks < 2 * 148) r >>= 1;
    return r;
}

}  // namespace tt

extern "C" int tt_bn_workspace(int64_t rows, int cols, size_t *bytes_host) {
    TT_CHECK_ARG(bytes_host && rows > 0 && cols > 0, "bad size");
    const int rpc = tt::bn_rows_per_chunk(rows, cols);
    const int64_t chunks = (rows + rpc - 1) / rpc;
    *bytes_host = static_cast<size_t>(chunks) * 3 * cols * sizeof(float) + 256;
    return 0;
}

#define TT_BN_CHECKS()                                                                                         \
    TT_CHECK_ARG(rows > 0 && cols > 0 && cols % 4 == 0 && x_stride % 4 == 0, "bn: cols and strides must be multiples of 4"); \
    TT_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0, "bn: x must be 16-byte aligned")

extern "C" int tt_bn_stats(const float *x, int64_t rows, int cols, int64_t x_stride, float *stats, void *workspace,
                           size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(x && stats && workspace, "null pointer");
    TT_BN_CHECKS();
    const int rpc = bn_rows_per_chunk(rows, cols);
    const int chunks = static_cast<int>((rows + rpc - 1) / rpc);
    if (workspace_bytes < static_cast<size_t>(chunks) * 3 * cols * sizeof(float)) { set_error("bn workspace too small"); return TT_E_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace);
    dim3 grid((cols + BN_TX * 4 - 1) / (BN_TX * 4), chunks), block(BN_TX, BN_TY);
    if (rows <= BN_SMALL_ROWS) {
        bn_stats_small<<<grid.x, block, 0, st>>>(x, rows, cols, x_stride, stats);
        TT_LAUNCH_CHECK("bn_stats_small");
        return 0;
    }
    bn_stats_partial<<<grid, block, 0, st>>>(x, rows, cols, x_stride, rpc, partial);
    TT_LAUNCH_CHECK("bn_stats_partial");
    bn_stats_final<<<(cols + 31) / 32, dim3(32, BN_FL), 0, st>>>(partial, chunks, rows, rpc, cols, stats);
    TT_LAUNCH_CHECK("bn_stats_final");
    return 0;
}

extern "C" int tt_bn_apply(const float *x, int64_t rows, int cols, int64_t x_stride, const float *stats_all, int n_ranks,
                           const float *gamma, const float *beta, int param_period, float eps, int relu, float dropout_p,
                           const int64_t *seed_dev, int64_t call_id, float *y, int64_t y_stride, float *save_mean,
                           float *save_rstd, float *batch_var_unbiased, float *running_mean, float *running_var,
                           float momentum, int64_t *num_batches, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(x && stats_all && gamma && beta && y && save_mean && save_rstd, "null pointer");
    TT_BN_CHECKS();
    TT_CHECK_ARG(n_ranks >= 1 && param_period > 0 && param_period % 4 == 0 && cols % param_period == 0, "bn: bad period");
    TT_CHECK_ARG(y_stride % 4 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0, "bn: y alignment");
    TT_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "running stats: both or neither");
    TT_CHECK_ARG(running_mean == nullptr || param_period == cols, "running stats are updated for plain layers only");
    const int rpc = bn_rows_per_chunk(rows, cols);
    const int chunks = static_cast<int>((rows + rpc - 1) / rpc);
    dim3 grid((cols + BN_TX * 4 - 1) / (BN_TX * 4), chunks), block(BN_TX, BN_TY);
    bn_apply_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
        x, rows, cols, x_stride, rpc, stats_all, n_ranks, gamma, beta, param_period, eps, relu, dropout_p, seed_dev, call_id,
        y, y_stride, save_mean, save_rstd, batch_var_unbiased, running_mean, running_var, momentum, num_batches);
    TT_LAUNCH_CHECK("bn_apply_kernel");
    return 0;
}

extern "C" int tt_bn_bwd_stats(const float *dy, int64_t dy_stride, const float *x, int64_t rows, int cols, int64_t x_stride,
                               const float *save_mean, const float *save_rstd, const float *gamma, const float *beta,
                               int param_period, int relu, float dropout_p, const int64_t *seed_dev, int64_t call_id,
                               float *sums, float *dgamma, float *dbeta, int accumulate, void *workspace,
                               size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(dy && x && save_mean && save_rstd && gamma && beta && sums && workspace, "null pointer");
    TT_BN_CHECKS();
    TT_CHECK_ARG(dy_stride % 4 == 0 && reinterpret_cast<uintptr_t>(dy) % 16 == 0, "bn: dy alignment");
    TT_CHECK_ARG(param_period > 0 && param_period % 4 == 0 && cols % param_period == 0, "bn: bad period");
    const int rpc = bn_rows_per_chunk(rows, cols);
    const int chunks = static_cast<int>((rows + rpc - 1) / rpc);
    if (workspace_bytes < static_cast<size_t>(chunks) * 2 * cols * sizeof(float)) { set_error("bn workspace too small"); return TT_E_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace);
    dim3 grid((cols + BN_TX * 4 - 1) / (BN_TX * 4), chunks), block(BN_TX, BN_TY);
    if (rows <= BN_SMALL_ROWS) {
        // one chunk = all rows: the "partial" of chunk 0 IS the result, written straight into sums
        dim3 g1(grid.x, 1);
        bn_bwd_kernel<false><<<g1, block, 0, st>>>(dy, dy_stride, x, rows, cols, x_stride, static_cast<int>(rows), save_mean,
                                                    save_rstd, gamma, beta, param_period, relu, dropout_p, seed_dev, call_id,
                                                    sums, nullptr, 1, 0.f, nullptr, 0);
        TT_LAUNCH_CHECK("bn_bwd_kernel<stats, small>");
    } else {
        bn_bwd_kernel<false><<<grid, block, 0, st>>>(dy, dy_stride, x, rows, cols, x_stride, rpc, save_mean, save_rstd, gamma, beta,
                                                      param_period, relu, dropout_p, seed_dev, call_id, partial, nullptr, 1, 0.f,
                                                      nullptr, 0);
        TT_LAUNCH_CHECK("bn_bwd_kernel<stats>");
        bn_bwd_final<<<(cols + 31) / 32, dim3(32, BN_FL), 0, st>>>(partial, chunks, cols, sums);
        TT_LAUNCH_CHECK("bn_bwd_final");
    }
    if (dgamma && dbeta) {
        bn_param_grads<<<(param_period + 127) / 128, 128, 0, st>>>(sums, cols, param_period, dgamma, dbeta, accumulate);
        TT_LAUNCH_CHECK("bn_param_grads");
    }
    return 0;
}

extern "C" int tt_bn_bwd_apply(const float *dy, int64_t dy_stride, const float *x, int64_t rows, int cols, int64_t x_stride,
                               const float *save_mean, const float *save_rstd, const float *gamma, const float *beta,
                               int param_period, int relu, float dropout_p, const int64_t *seed_dev, int64_t call_id,
                               const float *sums_all, int n_ranks, double total_rows, float *dx, int64_t dx_stride,
                               void *stream) {
    using namespace tt;
    TT_CHECK_ARG(dy && x && save_mean && save_rstd && gamma && beta && sums_all && dx && n_ranks >= 1, "null pointer");
    TT_BN_CHECKS();
    TT_CHECK_ARG(dy_stride % 4 == 0 && dx_stride % 4 == 0 && reinterpret_cast<uintptr_t>(dy) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(dx) % 16 == 0, "bn: dy / dx alignment");
    TT_CHECK_ARG(total_rows >= 1.0, "total_rows");
    const int rpc = bn_rows_per_chunk(rows, cols);
    const int chunks = static_cast<int>((rows + rpc - 1) / rpc);
    dim3 grid((cols + BN_TX * 4 - 1) / (BN_TX * 4), chunks), block(BN_TX, BN_TY);
    bn_bwd_kernel<true><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
        dy, dy_stride, x, rows, cols, x_stride, rpc, save_mean, save_rstd, gamma, beta, param_period, relu, dropout_p, seed_dev,
        call_id, nullptr, sums_all, n_ranks, static_cast<float>(1.0 / total_rows), dx, dx_stride);
    TT_LAUNCH_CHECK("bn_bwd_kernel<apply>");
    return 0;
}
