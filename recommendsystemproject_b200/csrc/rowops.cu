// Row-wise L2 normalisation of the tower outputs, forward and backward, one kernel each.
// Replaces F.normalize(x, p=2, dim=1) at Tower.py:41 (y = x / max(||x||_2, 1e-12)) and its autograd (torch runs it
// as norm + clamp + div and ~8 elementwise kernels backward).  One warp per row, fp32, fixed shuffle order.
#include "common.cuh"

namespace tt {

__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float *__restrict__ x, int64_t rows, int dim, float eps, float *__restrict__ y,
                  float *__restrict__ inv_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int vpr = dim / 4;
    for (int64_t r = warp; r < rows; r += n_warps) {
        const float4 *xr = reinterpret_cast<const float4 *>(x + r * dim);
        float ss = 0.f;
        for (int c = lane; c < vpr; c += 32) {
            const float4 v = __ldg(xr + c);
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
        for (int c = lane; c < vpr; c += 32) {
            const float4 v = __ldg(xr + c);
            *(reinterpret_cast<float4 *>(y + r * dim) + c) = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
        }
        if (lane == 0) inv_norm[r] = inv;
    }
}

// dx = inv * (g - y * <g, y>)   (norm above eps);   dx = g / eps   (clamped row: d max(n, eps) / dn = 0)
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float *__restrict__ gy, const float *__restrict__ y, const float *__restrict__ inv_norm,
                  int64_t rows, int dim, float eps, float *__restrict__ gx) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int vpr = dim / 4;
    for (int64_t r = warp; r < rows; r += n_warps) {
        const float4 *gr = reinterpret_cast<const float4 *>(gy + r * dim);
        const float4 *yr = reinterpret_cast<const float4 *>(y + r * dim);
        const float inv = inv_norm[r];
        const bool clamped = inv >= 1.0f / eps;
        float dot = 0.f;
        if (!clamped) {
            for (int c = lane; c < vpr; c += 32) {
                const float4 g = __ldg(gr + c), v = __ldg(yr + c);
                dot += g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w;
            }
            dot = warp_sum(dot);
        }
        for (int c = lane; c < vpr; c += 32) {
            const float4 g = __ldg(gr + c), v = __ldg(yr + c);
            *(reinterpret_cast<float4 *>(gx + r * dim) + c) =
                make_float4(inv * (g.x - v.x * dot), inv * (g.y - v.y * dot), inv * (g.z - v.z * dot), inv * (g.w - v.w * dot));
        }
    }
}

static inline unsigned rowops_grid(int64_t rows) {
    int64_t b = (rows * 32 + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b);
}

}  // namespace tt

extern "C" int tt_l2_normalize_fwd(const float *x, int64_t rows, int dim, float eps, float *y, float *inv_norm, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(x && y && inv_norm && rows > 0 && dim > 0 && dim % 4 == 0, "l2_normalize: dim % 4 == 0");
    TT_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0, "l2_normalize: alignment");
    l2norm_fwd_kernel<<<rowops_grid(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, dim, eps, y, inv_norm);
    TT_LAUNCH_CHECK("l2norm_fwd_kernel");
    return 0;
}

extern "C" int tt_l2_normalize_bwd(const float *grad_y, const float *y, const float *inv_norm, int64_t rows, int dim, float eps,
                                   float *grad_x, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(grad_y && y && inv_norm && grad_x && rows > 0 && dim > 0 && dim % 4 == 0, "l2_normalize: dim % 4 == 0");
    TT_CHECK_ARG(reinterpret_cast<uintptr_t>(grad_y) % 16 == 0 && reinterpret_cast<uintptr_t>(grad_x) % 16 == 0, "l2_normalize: alignment");
    l2norm_bwd_kernel<<<rowops_grid(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_y, y, inv_norm, rows, dim, eps, grad_x);
    TT_LAUNCH_CHECK("l2norm_bwd_kernel");
    return 0;
}
