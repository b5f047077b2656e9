// Weight and bias gradient of a Linear layer in one pass:  dW[O, I] = dY^T X,  db[O] = column sums of dY.
//
// Replaces, in the backward of every nn.Linear of the towers and of the sequence encoder (Tower.py:17-24,
// SequenceFeatureProcessor.py:30, nn.TransformerEncoderLayer inside SequenceEncoder.py:13-21), the pair torch
// launches: an fp32 SIMT GEMM "nt" whose reduction dimension is the ROW count (5632 item rows, 10240 sequence
// positions) over a 64..256-wide output -- cuBLAS picks a 64x64 tile without split-K, i.e. 1..16 CTAs on 148 SMs,
// 27 us each -- and a separate reduce_kernel for the bias (15 us each): together a quarter of the kernel time of a
// C2 training step (profiles/r1_c2_step_launches_fused.md).  Here the rows are split over ~2 waves of CTAs, each
// writes a partial tile (and, for the first column tile, the partial column sums of dY it has in shared memory
// anyway), and a second kernel adds the partials in fixed order: deterministic, fp32 FMA.
#include "common.cuh"

namespace tt {

constexpr int LG_T = 64;     // tile of dW: 64 outputs x 64 inputs
constexpr int LG_K = 16;     // rows per shared-memory stage

__global__ void __launch_bounds__(256)
linear_wgrad_partial(const float *__restrict__ dy, const float *__restrict__ x, int64_t rows, int n_out, int n_in,
                     int tiles_i, int64_t rows_per, float *__restrict__ part_w, float *__restrict__ part_b) {
    __shared__ __align__(16) float ys[LG_K][LG_T];
    __shared__ __align__(16) float xs[LG_K][LG_T];
    const int tile_o = blockIdx.x / tiles_i, tile_i = blockIdx.x % tiles_i;
    const int chunk = blockIdx.y;
    const int64_t r_begin = chunk * rows_per, r_end = min(rows, r_begin + rows_per);
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    const int lr = t / 16, lc = (t % 16) * 4;            // loader: row lr of the stage, columns lc..lc+3
    const int o0 = tile_o * LG_T, i0 = tile_i * LG_T;
    float acc[4][4] = {};
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec_y = (n_out % 4 == 0), vec_x = (n_in % 4 == 0);
    auto load_stage = [&](int64_t r0, float4 &vy, float4 &vx) {
        const int64_t r = r0 + lr;
        vy = make_float4(0.f, 0.f, 0.f, 0.f);
        vx = vy;
        if (r < r_end) {
            if (vec_y && o0 + lc + 3 < n_out) vy = *reinterpret_cast<const float4 *>(dy + r * n_out + o0 + lc);
            else {
                float tmp[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = 0; k < 4; ++k) if (o0 + lc + k < n_out) tmp[k] = dy[r * n_out + o0 + lc + k];
                vy = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            }
            if (vec_x && i0 + lc + 3 < n_in) vx = *reinterpret_cast<const float4 *>(x + r * n_in + i0 + lc);
            else {
                float tmp[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = 0; k < 4; ++k) if (i0 + lc + k < n_in) tmp[k] = x[r * n_in + i0 + lc + k];
                vx = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            }
        }
    };
    float4 vy, vx;
    if (r_begin < r_end) load_stage(r_begin, vy, vx);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += LG_K) {
        __syncthreads();
        *reinterpret_cast<float4 *>(&ys[lr][lc]) = vy;
        *reinterpret_cast<float4 *>(&xs[lr][lc]) = vx;
        __syncthreads();
        if (r0 + LG_K < r_end) load_stage(r0 + LG_K, vy, vx);     // next stage's rows fly while this one is multiplied
#pragma unroll
        for (int k = 0; k < LG_K; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&ys[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&xs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int p = 0; p < 4; ++p) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
                bsum[p] += av[p];
            }
        }
    }
    float *pw = part_w + static_cast<int64_t>(chunk) * n_out * n_in;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int o = o0 + ty * 4 + p;
        if (o >= n_out) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + tx * 4 + q;
            if (i < n_in) pw[static_cast<int64_t>(o) * n_in + i] = acc[p][q];
        }
        if (tile_i == 0 && tx == 0 && part_b != nullptr) part_b[static_cast<int64_t>(chunk) * n_out + o] = bsum[p];
    }
}

// 64 outputs per CTA, 4 threads per output: thread g of an output adds chunks g, g+4, ... and the four sums are
// combined in fixed order (deterministic; 4x shorter dependent chains than one thread per output)
__global__ void __launch_bounds__(256)
linear_wgrad_reduce(const float *__restrict__ part_w, const float *__restrict__ part_b, int n_chunks, int64_t n_w,
                    int n_out, int accumulate, float *__restrict__ dw, float *__restrict__ db) {
    __shared__ float red[4][64];
    const int el = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int64_t e = static_cast<int64_t>(blockIdx.x) * 64 + el;
    const int64_t total = n_w + (db != nullptr ? n_out : 0);
    float s = 0.f;
    if (e < n_w) {
        for (int c = grp; c < n_chunks; c += 4) s += part_w[static_cast<int64_t>(c) * n_w + e];
    } else if (e < total) {
        const int o = static_cast<int>(e - n_w);
        for (int c = grp; c < n_chunks; c += 4) s += part_b[static_cast<int64_t>(c) * n_out + o];
    }
    red[grp][el] = s;
    __syncthreads();
    if (grp == 0 && e < total) {
        const float v = ((red[0][el] + red[1][el]) + red[2][el]) + red[3][el];
        float *dst = e < n_w ? dw + e : db + (e - n_w);
        *dst = accumulate ? *dst + v : v;
    }
}

struct LgPlan { int tiles_o, tiles_i, chunks; int64_t rows_per; };
static LgPlan lg_plan(int64_t rows, int n_out, int n_in) {
    LgPlan p;
    p.tiles_o = (n_out + LG_T - 1) / LG_T;
    p.tiles_i = (n_in + LG_T - 1) / LG_T;
    const int tiles = p.tiles_o * p.tiles_i;
    int64_t want = (static_cast<int64_t>(sm_count()) * 2 + tiles - 1) / tiles;    // ~2 waves of CTAs
    const int64_t most = (rows + 4 * LG_K - 1) / (4 * LG_K);     // at least 64 rows per chunk
    if (want > most) want = most;
    if (want < 1) want = 1;
    p.rows_per = ((rows + want - 1) / want + LG_K - 1) / LG_K * LG_K;
    p.chunks = static_cast<int>((rows + p.rows_per - 1) / p.rows_per);
    return p;
}

}  // namespace tt

extern "C" int tt_linear_wgrad_workspace(int64_t rows, int n_out, int n_in, size_t *bytes_host) {
    TT_CHECK_ARG(bytes_host && rows > 0 && n_out > 0 && n_in > 0, "bad size");
    const tt::LgPlan p = tt::lg_plan(rows, n_out, n_in);
    *bytes_host = static_cast<size_t>(p.chunks) * (static_cast<size_t>(n_out) * n_in + n_out) * sizeof(float) + 512;
    return 0;
}

extern "C" int tt_linear_wgrad(const float *grad_out, const float *input, int64_t rows, int n_out, int n_in,
                               float *grad_weight, float *grad_bias, int accumulate, void *workspace,
                               size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(grad_out && input && grad_weight && workspace && rows > 0 && n_out > 0 && n_in > 0, "null pointer / empty input");
    TT_CHECK_ARG(reinterpret_cast<uintptr_t>(grad_out) % 16 == 0 && reinterpret_cast<uintptr_t>(input) % 16 == 0,
                 "grad_out / input must be 16-byte aligned");
    const LgPlan p = lg_plan(rows, n_out, n_in);
    const size_t n_w = static_cast<size_t>(n_out) * n_in;
    Workspace ws(workspace, workspace_bytes);
    float *part_w = ws.take<float>(static_cast<size_t>(p.chunks) * n_w);
    float *part_b = ws.take<float>(static_cast<size_t>(p.chunks) * n_out);
    if (!ws.ok()) { set_error("linear_wgrad workspace too small: need %zu have %zu", ws.off, workspace_bytes); return TT_E_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(static_cast<unsigned>(p.tiles_o * p.tiles_i), static_cast<unsigned>(p.chunks));
    linear_wgrad_partial<<<grid, 256, 0, st>>>(grad_out, input, rows, n_out, n_in, p.tiles_i, p.rows_per, part_w,
                                               grad_bias ? part_b : nullptr);
    TT_LAUNCH_CHECK("linear_wgrad_partial");
    const int64_t total = static_cast<int64_t>(n_w) + (grad_bias ? n_out : 0);
    linear_wgrad_reduce<<<static_cast<unsigned>((total + 63) / 64), 256, 0, st>>>(part_w, part_b, p.chunks,
                                                                                    static_cast<int64_t>(n_w), n_out, accumulate, grad_weight, grad_bias);
    TT_LAUNCH_CHECK("linear_wgrad_reduce");
    return 0;
}

// ---------------------------------------------------------------- forward and input gradient
// y[R, N] = x[R, K] . W[N, K]^T + b   (B_TRANS = true,  B = W)          -- nn.Linear forward
// dx[R, K'] = dy[R, N'] . W[N', K']   (B_TRANS = false, B = W, no bias)  -- its input gradient
// One launch each.  cuBLASLt runs these skinny fp32 products (R = 512..10240 rows, 8..256 columns) as a split-K SIMT
// GEMM + a split-K reduce + a bias epilogue kernel: three launches per Linear and direction in a step that is
// launch-bound (profiles/r1_c2_step_launches_fused.md).  64 x 64 tile, 16-wide k stages, 4 x 4 outputs per thread,
// next stage prefetched into registers; the k loop runs in order => deterministic.
namespace tt {

template <bool B_TRANS>
__global__ void __launch_bounds__(256)
linear_gemm64(const float *__restrict__ A, const float *__restrict__ B, const float *__restrict__ bias,
              float *__restrict__ C, int64_t R, int N, int K, int relu) {
    __shared__ __align__(16) float As[LG_K][LG_T + 4];
    __shared__ __align__(16) float Bs[LG_K][LG_T + 4];
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * LG_T;
    const int n0 = blockIdx.x * LG_T;
    const bool vecA = (K % 4 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
    const bool vecB = (B_TRANS ? (K % 4 == 0) : (N % 4 == 0)) && (reinterpret_cast<uintptr_t>(B) % 16 == 0);
    float acc[4][4] = {};
    float ra[4], rb[4];
    auto load_stage = [&](int k0) {
        {   // A tile: row = t / 4, four consecutive k
            const int64_t r = r0 + t / 4;
            const int kq = k0 + (t % 4) * 4;
            ra[0] = ra[1] = ra[2] = ra[3] = 0.f;
            if (r < R) {
                if (vecA && kq + 3 < K) {
                    const float4 v = *reinterpret_cast<const float4 *>(A + r * K + kq);
                    ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
                } else {
                    for (int j = 0; j < 4; ++j) if (kq + j < K) ra[j] = A[r * K + kq + j];
                }
            }
        }
        rb[0] = rb[1] = rb[2] = rb[3] = 0.f;
        if (B_TRANS) {   // B[n, k]: row n = t / 4, four consecutive k
            const int n = n0 + t / 4;
            const int kq = k0 + (t % 4) * 4;
            if (n < N) {
                if (vecB && kq + 3 < K) {
                    const float4 v = *reinterpret_cast<const float4 *>(B + static_cast<int64_t>(n) * K + kq);
                    rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
                } else {
                    for (int j = 0; j < 4; ++j) if (kq + j < K) rb[j] = B[static_cast<int64_t>(n) * K + kq + j];
                }
            }
        } else {         // B[k, n]: k = t / 16, four consecutive n
            const int k = k0 + t / 16;
            const int nq = n0 + (t % 16) * 4;
            if (k < K) {
                if (vecB && nq + 3 < N) {
                    const float4 v = *reinterpret_cast<const float4 *>(B + static_cast<int64_t>(k) * N + nq);
                    rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
                } else {
                    for (int j = 0; j < 4; ++j) if (nq + j < N) rb[j] = B[static_cast<int64_t>(k) * N + nq + j];
                }
            }
        }
    };
    load_stage(0);
    for (int k0 = 0; k0 < K; k0 += LG_K) {
        __syncthreads();
        {
            const int row = t / 4, kq = (t % 4) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) As[kq + j][row] = ra[j];
            if (B_TRANS) {
#pragma unroll
                for (int j = 0; j < 4; ++j) Bs[kq + j][row] = rb[j];
            } else {
                *reinterpret_cast<float4 *>(&Bs[t / 16][(t % 16) * 4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
            }
        }
        __syncthreads();
        if (k0 + LG_K < K) load_stage(k0 + LG_K);
#pragma unroll
        for (int k = 0; k < LG_K; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int64_t r = r0 + ty * 4 + p;
        if (r >= R) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int n = n0 + tx * 4 + q;
            if (n >= N) continue;
            float v = acc[p][q] + (bias != nullptr ? bias[n] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            C[r * N + n] = v;
        }
    }
}

}  // namespace tt

extern "C" int tt_linear_fwd(const float *input, const float *weight, const float *bias, int64_t rows, int n_out, int n_in,
                             int relu, float *out, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(input && weight && out && rows > 0 && n_out > 0 && n_in > 0, "null pointer / empty input");
    dim3 grid(static_cast<unsigned>((n_out + LG_T - 1) / LG_T), static_cast<unsigned>((rows + LG_T - 1) / LG_T));
    linear_gemm64<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(input, weight, bias, out, rows, n_out, n_in, relu);
    TT_LAUNCH_CHECK("linear_gemm64<fwd>");
    return 0;
}

extern "C" int tt_linear_dgrad(const float *grad_out, const float *weight, int64_t rows, int n_out, int n_in,
                               float *grad_input, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(grad_out && weight && grad_input && rows > 0 && n_out > 0 && n_in > 0, "null pointer / empty input");
    dim3 grid(static_cast<unsigned>((n_in + LG_T - 1) / LG_T), static_cast<unsigned>((rows + LG_T - 1) / LG_T));
    // dx[R, n_in] = dy[R, n_out] . W[n_out, n_in]: A = dy (K = n_out), B = W as [K, N = n_in]
    linear_gemm64<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, weight, nullptr, grad_input, rows, n_in, n_out, 0);
    TT_LAUNCH_CHECK("linear_gemm64<dgrad>");
    return 0;
}
