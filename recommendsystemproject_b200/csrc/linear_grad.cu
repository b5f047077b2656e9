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
