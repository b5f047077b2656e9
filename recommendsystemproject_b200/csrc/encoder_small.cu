// Two fused pieces of the small-sequence Transformer behaviour encoder: what is left of
// nn.TransformerEncoderLayer (SequenceEncoder.py:13-21, batch_first, post-norm, ReLU) once its four GEMMs go to
// cuBLAS.  At the shipped shape (B=512 samples x L=20 positions, d=64, 4 heads) the torch layer is ~55 launches per
// layer and step, and the slow ones are exactly these: LayerNorm backward (gamma/beta reduction: 88 us each),
// memory-efficient attention forward / backward (56 / 130 us at L=20), plus the dropout / residual / transpose
// elementwise kernels between them.
//
//   attn_small_*        softmax(q k^T / sqrt(dh) + key padding mask) -> dropout -> . v for L <= 32, straight from
//                       the packed in_proj output [B, L, 3d] (no head transposes, no [B, H, L, L] tensor in HBM);
//                       one warp per (sample, head), lane = query row; the backward recomputes the probabilities.
//   add_dropout_ln_*    y = LayerNorm(x + dropout(z)); one warp per row; the backward returns dx, dz and per-CTA
//                       partial sums of dgamma / dbeta that a second kernel adds in fixed order (deterministic).
//
// Dropout masks are a counter-based hash of (seed read from device memory, call site, element index): the same
// mask is rebuilt in the backward, nothing is stored, and a CUDA-graph replay sees a new seed every step.
#include "common.cuh"

namespace tt {

__device__ __forceinline__ uint32_t drop_hash(uint64_t seed, uint64_t call, uint64_t elem) {
    uint64_t z = seed + call * 0x9E3779B97F4A7C15ull + elem * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return static_cast<uint32_t>(z);
}
// keep with probability 1 - p: hash >= p * 2^32
__device__ __forceinline__ bool drop_keep(uint64_t seed, uint64_t call, uint64_t elem, uint32_t thresh) {
    return drop_hash(seed, call, elem) >= thresh;
}
static inline uint32_t drop_threshold(float p) {
    double t = static_cast<double>(p) * 4294967296.0;
    if (t < 0.0) t = 0.0;
    if (t > 4294967295.0) t = 4294967295.0;
    return static_cast<uint32_t>(t);
}

constexpr int AT_MAXL = 32;

struct AttnArgs {
    const float *qkv;            // [B, L, 3d]: q | k | v, head h = columns [h*dh, (h+1)*dh) of each third
    const uint8_t *key_pad;      // [B, L], 1 = key is padding (ignored); may be NULL
    int64_t batch;
    int len, heads;
    float scale, keep_scale;     // 1/sqrt(dh), 1/(1-p)
    uint32_t thresh;             // 0: no dropout
    const int64_t *seed;
    int64_t call;
};

// probabilities of query row `i` (this lane) against all keys: p[j] = softmax_j((q*scale) . k_j), masked keys -> 0
template <int DH>
__device__ __forceinline__ void attn_row_probs(const float (&q)[DH], const float (*Ks)[DH], const uint8_t *pad, int L,
                                               float (&p)[AT_MAXL]) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < AT_MAXL; ++j) {
        float s = -INFINITY;
        if (j < L && !(pad != nullptr && pad[j])) {
            s = 0.f;
#pragma unroll
            for (int c = 0; c < DH; ++c) s = fmaf(q[c], Ks[j][c], s);
        }
        p[j] = s;
        mx = fmaxf(mx, s);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < AT_MAXL; ++j) {
        p[j] = (p[j] == -INFINITY) ? 0.f : __expf(p[j] - mx);
        sum += p[j];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < AT_MAXL; ++j) p[j] *= inv;
}

template <int DH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_small_fwd_kernel(AttnArgs a, float *__restrict__ out) {
    __shared__ float Ks[WARPS][AT_MAXL][DH];
    __shared__ float Vs[WARPS][AT_MAXL][DH];
    __shared__ uint8_t Pad[WARPS][AT_MAXL];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t w = static_cast<int64_t>(blockIdx.x) * WARPS + warp;
    if (w >= a.batch * a.heads) return;
    const int64_t b = w / a.heads;
    const int h = static_cast<int>(w % a.heads);
    const int L = a.len, d = a.heads * DH;
    const float *base = a.qkv + b * L * 3 * d + h * DH;
    for (int idx = lane; idx < L * DH; idx += 32) {
        const int j = idx / DH, c = idx % DH;
        Ks[warp][j][c] = base[static_cast<int64_t>(j) * 3 * d + d + c];
        Vs[warp][j][c] = base[static_cast<int64_t>(j) * 3 * d + 2 * d + c];
    }
    if (lane < L) Pad[warp][lane] = a.key_pad ? a.key_pad[b * L + lane] : 0;
    __syncwarp();
    if (lane >= L) return;
    const int i = lane;
    float q[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) q[c] = base[static_cast<int64_t>(i) * 3 * d + c] * a.scale;
    float p[AT_MAXL];
    attn_row_probs<DH>(q, Ks[warp], a.key_pad ? Pad[warp] : nullptr, L, p);
    const uint64_t seed = a.thresh ? static_cast<uint64_t>(*a.seed) : 0;
    float o[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] = 0.f;
#pragma unroll
    for (int j = 0; j < AT_MAXL; ++j) {
        if (j < L) {
            float pj = p[j];
            if (a.thresh) pj = drop_keep(seed, a.call, (w * L + i) * L + j, a.thresh) ? pj * a.keep_scale : 0.f;
#pragma unroll
            for (int c = 0; c < DH; ++c) o[c] = fmaf(pj, Vs[warp][j][c], o[c]);
        }
    }
    float *dst = out + (b * L + i) * d + h * DH;
#pragma unroll
    for (int c = 0; c < DH; ++c) dst[c] = o[c];
}

template <int DH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_small_bwd_kernel(AttnArgs a, const float *__restrict__ grad_out, float *__restrict__ grad_qkv) {
    extern __shared__ float smem_dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per warp: K, V, Q (scaled), dO: [L][DH] each; Pd, dS: [L][L + 1]  (sized by the actual L: more warps per SM)
    const int LS = a.len + 1;
    const int per_warp = 4 * a.len * DH + 2 * a.len * LS;
    float *mine = smem_dyn + warp * per_warp;
    float (*Ks)[DH] = reinterpret_cast<float (*)[DH]>(mine);
    float (*Vs)[DH] = reinterpret_cast<float (*)[DH]>(mine + a.len * DH);
    float (*Qs)[DH] = reinterpret_cast<float (*)[DH]>(mine + 2 * a.len * DH);
    float (*Gs)[DH] = reinterpret_cast<float (*)[DH]>(mine + 3 * a.len * DH);
    float *Pd = mine + 4 * a.len * DH;
    float *dS = Pd + a.len * LS;
    __shared__ uint8_t Pad[WARPS][AT_MAXL];
    const int64_t w = static_cast<int64_t>(blockIdx.x) * WARPS + warp;
    if (w >= a.batch * a.heads) return;
    const int64_t b = w / a.heads;
    const int h = static_cast<int>(w % a.heads);
    const int L = a.len, d = a.heads * DH;
    const float *base = a.qkv + b * L * 3 * d + h * DH;
    const float *gbase = grad_out + b * L * d + h * DH;
    for (int idx = lane; idx < L * DH; idx += 32) {
        const int j = idx / DH, c = idx % DH;
        Qs[j][c] = base[static_cast<int64_t>(j) * 3 * d + c] * a.scale;
        Ks[j][c] = base[static_cast<int64_t>(j) * 3 * d + d + c];
        Vs[j][c] = base[static_cast<int64_t>(j) * 3 * d + 2 * d + c];
        Gs[j][c] = gbase[static_cast<int64_t>(j) * d + c];
    }
    if (lane < L) Pad[warp][lane] = a.key_pad ? a.key_pad[b * L + lane] : 0;
    __syncwarp();
    const uint64_t seed = a.thresh ? static_cast<uint64_t>(*a.seed) : 0;
    float *gq = grad_qkv + b * L * 3 * d + h * DH;
    if (lane < L) {
        // ---- phase 1, lane = query row i: probabilities again, dS, dQ
        const int i = lane;
        float q[DH], go[DH];
#pragma unroll
        for (int c = 0; c < DH; ++c) { q[c] = Qs[i][c]; go[c] = Gs[i][c]; }
        float p[AT_MAXL];
        attn_row_probs<DH>(q, Ks, a.key_pad ? Pad[warp] : nullptr, L, p);
        float dp[AT_MAXL];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < AT_MAXL; ++j) {
            dp[j] = 0.f;
            if (j < L) {
                float x = 0.f;
#pragma unroll
                for (int c = 0; c < DH; ++c) x = fmaf(go[c], Vs[j][c], x);      // d(dropped prob)
                float keep = 1.0f;
                if (a.thresh) keep = drop_keep(seed, a.call, (w * L + i) * L + j, a.thresh) ? a.keep_scale : 0.f;
                Pd[i * LS + j] = p[j] * keep;
                dp[j] = x * keep;                                               // d(prob)
                dot = fmaf(p[j], dp[j], dot);
            }
        }
        float dq[DH];
#pragma unroll
        for (int c = 0; c < DH; ++c) dq[c] = 0.f;
#pragma unroll
        for (int j = 0; j < AT_MAXL; ++j) {
            if (j < L) {
                const float ds = p[j] * (dp[j] - dot);
                dS[i * LS + j] = ds;
#pragma unroll
                for (int c = 0; c < DH; ++c) dq[c] = fmaf(ds, Ks[j][c], dq[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < DH; ++c) gq[static_cast<int64_t>(i) * 3 * d + c] = dq[c] * a.scale;
    }
    __syncwarp();
    if (lane < L) {
        // ---- phase 2, lane = key row j: dK, dV
        const int j = lane;
        float dk[DH], dv[DH];
#pragma unroll
        for (int c = 0; c < DH; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
        for (int i = 0; i < L; ++i) {
            const float ds = dS[i * LS + j], pd = Pd[i * LS + j];
#pragma unroll
            for (int c = 0; c < DH; ++c) {
                dk[c] = fmaf(ds, Qs[i][c], dk[c]);
                dv[c] = fmaf(pd, Gs[i][c], dv[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < DH; ++c) {
            gq[static_cast<int64_t>(j) * 3 * d + d + c] = dk[c];
            gq[static_cast<int64_t>(j) * 3 * d + 2 * d + c] = dv[c];
        }
    }
}

// ---------------------------------------------------------------- y = LayerNorm(x + dropout(z))
struct LnArgs {
    int64_t rows;
    int dim;
    float eps, keep_scale;
    uint32_t thresh;
    const int64_t *seed;
    int64_t call;
};

template <int VPL>   // values per lane: dim = 32 * VPL
__global__ void __launch_bounds__(256)
add_dropout_ln_fwd_kernel(LnArgs a, const float *__restrict__ x, const float *__restrict__ z,
                          const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ y,
                          float *__restrict__ xhat, float *__restrict__ rstd) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint64_t seed = a.thresh ? static_cast<uint64_t>(*a.seed) : 0;
    float g[VPL], bt[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) { g[k] = gamma[lane + 32 * k]; bt[k] = beta[lane + 32 * k]; }
    for (int64_t r = warp; r < a.rows; r += n_warps) {
        float hval[VPL];
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int64_t e = r * a.dim + lane + 32 * k;
            float zz = z[e];
            if (a.thresh) zz = drop_keep(seed, a.call, e, a.thresh) ? zz * a.keep_scale : 0.f;
            hval[k] = x[e] + zz;
            sum += hval[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / a.dim;
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) { const float dlt = hval[k] - mean; var = fmaf(dlt, dlt, var); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float rs = rsqrtf(var / a.dim + a.eps);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int64_t e = r * a.dim + lane + 32 * k;
            const float xh = (hval[k] - mean) * rs;
            xhat[e] = xh;
            y[e] = fmaf(xh, g[k], bt[k]);
        }
        if (lane == 0) rstd[r] = rs;
    }
}

template <int VPL>
__global__ void __launch_bounds__(256)
add_dropout_ln_bwd_kernel(LnArgs a, const float *__restrict__ gy, const float *__restrict__ xhat,
                          const float *__restrict__ rstd, const float *__restrict__ gamma, float *__restrict__ gx,
                          float *__restrict__ gz, float *__restrict__ partial) {
    __shared__ float red[8][2][32 * VPL];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint64_t seed = a.thresh ? static_cast<uint64_t>(*a.seed) : 0;
    float g[VPL], dg[VPL], db[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) { g[k] = gamma[lane + 32 * k]; dg[k] = 0.f; db[k] = 0.f; }
    for (int64_t r = warp; r < a.rows; r += n_warps) {
        float av[VPL], xh[VPL];
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int64_t e = r * a.dim + lane + 32 * k;
            const float dy = gy[e];
            xh[k] = xhat[e];
            dg[k] = fmaf(dy, xh[k], dg[k]);
            db[k] += dy;
            av[k] = dy * g[k];
            m1 += av[k];
            m2 = fmaf(av[k], xh[k], m2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m1 += __shfl_xor_sync(0xffffffffu, m1, o);
            m2 += __shfl_xor_sync(0xffffffffu, m2, o);
        }
        m1 /= a.dim;
        m2 /= a.dim;
        const float rs = rstd[r];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int64_t e = r * a.dim + lane + 32 * k;
            const float dh = rs * (av[k] - m1 - xh[k] * m2);
            gx[e] = dh;
            float dz = dh;
            if (a.thresh) dz = drop_keep(seed, a.call, e, a.thresh) ? dh * a.keep_scale : 0.f;
            gz[e] = dz;
        }
    }
    // per-CTA partial sums of dgamma / dbeta (warps added in fixed order)
#pragma unroll
    for (int k = 0; k < VPL; ++k) { red[wib][0][lane + 32 * k] = dg[k]; red[wib][1][lane + 32 * k] = db[k]; }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * 32 * VPL; t += blockDim.x) {
        const int which = t / (32 * VPL), c = t % (32 * VPL);
        float s = 0.f;
        for (int w2 = 0; w2 < 8; ++w2) s += red[w2][which][c];
        partial[(static_cast<int64_t>(blockIdx.x) * 2 + which) * a.dim + c] = s;
    }
}

// one warp per output column (dgamma[c] or dbeta[c]): lanes stride over the CTAs' partials, fixed-order tree
__global__ void __launch_bounds__(256)
ln_reduce_partials(const float *__restrict__ partial, int n_cta, int dim, int accumulate, float *__restrict__ ggamma,
                   float *__restrict__ gbeta) {
    const int lane = threadIdx.x & 31;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= 2 * dim) return;
    const int which = t / dim, c = t % dim;
    float s = 0.f;
    for (int k = lane; k < n_cta; k += 32) s += partial[(static_cast<int64_t>(k) * 2 + which) * dim + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        float *dst = (which == 0 ? ggamma : gbeta) + c;
        *dst = accumulate ? *dst + s : s;
    }
}

static int ln_grid(int64_t rows) {
    int64_t b = (rows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 2;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

static AttnArgs make_attn(const float *qkv, const uint8_t *key_pad, int64_t batch, int len, int heads, int head_dim,
                          float dropout_p, const int64_t *seed, int64_t call) {
    AttnArgs a;
    a.qkv = qkv; a.key_pad = key_pad; a.batch = batch; a.len = len; a.heads = heads;
    a.scale = 1.0f / sqrtf(static_cast<float>(head_dim));
    const bool drop = dropout_p > 0.f && seed != nullptr;
    a.keep_scale = drop ? 1.0f / (1.0f - dropout_p) : 1.0f;
    a.thresh = drop ? drop_threshold(dropout_p) : 0u;
    a.seed = seed; a.call = call;
    return a;
}

}  // namespace tt

extern "C" int tt_attn_small_fwd(const float *qkv, const uint8_t *key_pad_mask, int64_t batch, int len, int heads,
                                 int head_dim, float dropout_p, const int64_t *seed_dev, int64_t call_id, float *out,
                                 void *stream) {
    using namespace tt;
    TT_CHECK_ARG(qkv && out && batch > 0 && len > 0 && heads > 0, "null pointer / empty input");
    TT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p must be in [0, 1)");
    if (len > AT_MAXL) { set_error("attn_small supports sequence length <= %d (got %d)", AT_MAXL, len); return TT_E_UNSUPPORTED; }
    const AttnArgs a = make_attn(qkv, key_pad_mask, batch, len, heads, head_dim, dropout_p, seed_dev, call_id);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    constexpr int WARPS = 8;
    const unsigned grid = static_cast<unsigned>((batch * heads + WARPS - 1) / WARPS);
    if (head_dim == 8) attn_small_fwd_kernel<8, WARPS><<<grid, WARPS * 32, 0, st>>>(a, out);
    else if (head_dim == 16) attn_small_fwd_kernel<16, WARPS><<<grid, WARPS * 32, 0, st>>>(a, out);
    else if (head_dim == 32) attn_small_fwd_kernel<32, 4><<<static_cast<unsigned>((batch * heads + 3) / 4), 128, 0, st>>>(a, out);
    else { set_error("attn_small supports head_dim 8, 16 or 32 (got %d)", head_dim); return TT_E_UNSUPPORTED; }
    TT_LAUNCH_CHECK("attn_small_fwd_kernel");
    return 0;
}

namespace tt {
template <int DH, int WARPS>
static int launch_attn_bwd(const AttnArgs &a, const float *grad_out, float *grad_qkv, cudaStream_t st) {
    constexpr size_t smem_max = static_cast<size_t>(WARPS) * (4 * AT_MAXL * DH + 2 * AT_MAXL * (AT_MAXL + 1)) * sizeof(float);
    const size_t smem = static_cast<size_t>(WARPS) * (4 * a.len * DH + 2 * a.len * (a.len + 1)) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_small_bwd_kernel<DH, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem_max));
        if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(attn_small_bwd_kernel)");
        attr_set = true;
    }
    const unsigned grid = static_cast<unsigned>((a.batch * a.heads + WARPS - 1) / WARPS);
    attn_small_bwd_kernel<DH, WARPS><<<grid, WARPS * 32, smem, st>>>(a, grad_out, grad_qkv);
    TT_LAUNCH_CHECK("attn_small_bwd_kernel");
    return 0;
}
}  // namespace tt

extern "C" int tt_attn_small_bwd(const float *qkv, const uint8_t *key_pad_mask, const float *grad_out, int64_t batch,
                                 int len, int heads, int head_dim, float dropout_p, const int64_t *seed_dev,
                                 int64_t call_id, float *grad_qkv, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(qkv && grad_out && grad_qkv && batch > 0 && len > 0 && heads > 0, "null pointer / empty input");
    TT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p must be in [0, 1)");
    if (len > AT_MAXL) { set_error("attn_small supports sequence length <= %d (got %d)", AT_MAXL, len); return TT_E_UNSUPPORTED; }
    const AttnArgs a = make_attn(qkv, key_pad_mask, batch, len, heads, head_dim, dropout_p, seed_dev, call_id);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (head_dim == 8) return launch_attn_bwd<8, 8>(a, grad_out, grad_qkv, st);
    if (head_dim == 16) return launch_attn_bwd<16, 8>(a, grad_out, grad_qkv, st);
    if (head_dim == 32) return launch_attn_bwd<32, 4>(a, grad_out, grad_qkv, st);
    set_error("attn_small supports head_dim 8, 16 or 32 (got %d)", head_dim);
    return TT_E_UNSUPPORTED;
}

extern "C" int tt_add_dropout_ln_fwd(const float *x, const float *z, int64_t rows, int dim, const float *gamma,
                                     const float *beta, float eps, float dropout_p, const int64_t *seed_dev,
                                     int64_t call_id, float *y, float *xhat, float *rstd, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(x && z && gamma && beta && y && xhat && rstd && rows > 0, "null pointer / empty input");
    TT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p must be in [0, 1)");
    if (dim % 32 != 0 || dim > 256) { set_error("add_dropout_ln supports dim = 32k <= 256 (got %d)", dim); return TT_E_UNSUPPORTED; }
    const bool drop = dropout_p > 0.f && seed_dev != nullptr;
    LnArgs a{rows, dim, eps, drop ? 1.0f / (1.0f - dropout_p) : 1.0f, drop ? drop_threshold(dropout_p) : 0u, seed_dev, call_id};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ln_grid(rows) * 8;   // forward: plenty of small CTAs
#define TT_LN_FWD(V) add_dropout_ln_fwd_kernel<V><<<grid, 256, 0, st>>>(a, x, z, gamma, beta, y, xhat, rstd)
    switch (dim / 32) {
        case 1: TT_LN_FWD(1); break; case 2: TT_LN_FWD(2); break; case 3: TT_LN_FWD(3); break; case 4: TT_LN_FWD(4); break;
        case 5: TT_LN_FWD(5); break; case 6: TT_LN_FWD(6); break; case 7: TT_LN_FWD(7); break; default: TT_LN_FWD(8); break;
    }
#undef TT_LN_FWD
    TT_LAUNCH_CHECK("add_dropout_ln_fwd_kernel");
    return 0;
}

extern "C" int tt_add_dropout_ln_bwd_workspace(int64_t rows, int dim, size_t *bytes_host) {
    TT_CHECK_ARG(bytes_host && rows > 0 && dim > 0, "bad size");
    *bytes_host = static_cast<size_t>(tt::ln_grid(rows)) * 2 * dim * sizeof(float) + 256;
    return 0;
}

extern "C" int tt_add_dropout_ln_bwd(const float *grad_y, const float *xhat, const float *rstd, const float *gamma,
                                     int64_t rows, int dim, float dropout_p, const int64_t *seed_dev, int64_t call_id,
                                     float *grad_x, float *grad_z, float *grad_gamma, float *grad_beta, int accumulate,
                                     void *workspace, size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(grad_y && xhat && rstd && gamma && grad_x && grad_z && grad_gamma && grad_beta && workspace && rows > 0,
                 "null pointer / empty input");
    TT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p must be in [0, 1)");
    if (dim % 32 != 0 || dim > 256) { set_error("add_dropout_ln supports dim = 32k <= 256 (got %d)", dim); return TT_E_UNSUPPORTED; }
    const int grid = ln_grid(rows);
    if (workspace_bytes < static_cast<size_t>(grid) * 2 * dim * sizeof(float)) {
        set_error("add_dropout_ln_bwd workspace too small");
        return TT_E_WORKSPACE;
    }
    const bool drop = dropout_p > 0.f && seed_dev != nullptr;
    LnArgs a{rows, dim, 0.f, drop ? 1.0f / (1.0f - dropout_p) : 1.0f, drop ? drop_threshold(dropout_p) : 0u, seed_dev, call_id};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace);
#define TT_LN_BWD(V) add_dropout_ln_bwd_kernel<V><<<grid, 256, 0, st>>>(a, grad_y, xhat, rstd, gamma, grad_x, grad_z, partial)
    switch (dim / 32) {
        case 1: TT_LN_BWD(1); break; case 2: TT_LN_BWD(2); break; case 3: TT_LN_BWD(3); break; case 4: TT_LN_BWD(4); break;
        case 5: TT_LN_BWD(5); break; case 6: TT_LN_BWD(6); break; case 7: TT_LN_BWD(7); break; default: TT_LN_BWD(8); break;
    }
#undef TT_LN_BWD
    TT_LAUNCH_CHECK("add_dropout_ln_bwd_kernel");
    ln_reduce_partials<<<(2 * dim * 32 + 255) / 256, 256, 0, st>>>(partial, grid, dim, accumulate, grad_gamma, grad_beta);
    TT_LAUNCH_CHECK("ln_reduce_partials");
    return 0;
}
