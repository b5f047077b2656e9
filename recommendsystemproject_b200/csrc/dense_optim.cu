// Dense-parameter side of clip_grad_norm_ + Adam (training_utils.py:53-56,
// train_twotower.py:111 of the reference): squared-norm reduction, the clip
// coefficient, and a flat fused Adam.  All scalars stay on the device so the
// whole optimizer step can sit inside one CUDA graph.
#include "common.cuh"

namespace tt {

constexpr int SQ_BLOCKS = 592;  // 4 x 148 SMs; fixed so the reduction order is reproducible

__global__ void __launch_bounds__(256)
sq_norm_partial(const float *__restrict__ x, int64_t n, float *__restrict__ partial) {
    __shared__ float sh[8];
    float acc = 0.f;
    const int64_t n4 = n / 4;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const bool aligned = (reinterpret_cast<uintptr_t>(x) % 16) == 0;
    if (aligned) {
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
             i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
            const float4 v = __ldg(x4 + i);
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        for (int64_t i = n4 * 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
             i += static_cast<int64_t>(gridDim.x) * blockDim.x)
            acc += x[i] * x[i];
    } else {
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
             i += static_cast<int64_t>(gridDim.x) * blockDim.x)
            acc += x[i] * x[i];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sh[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) sq_norm_final(const float *__restrict__ partial, int n, float *__restrict__ out) {
    __shared__ float sh[1024];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out += sh[0];
}

__global__ void clip_coef_kernel(const float *__restrict__ sq_terms, int n_terms, float max_norm,
                                 float *__restrict__ coef, float *__restrict__ total_norm) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n_terms; ++i) s += static_cast<double>(sq_terms[i]);
        const float tn = static_cast<float>(sqrt(s));
        const float c = max_norm / (tn + 1e-6f);  // torch.nn.utils.clip_grad_norm_
        *coef = c < 1.0f ? c : 1.0f;
        if (total_norm) *total_norm = tn;
    }
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                 int64_t n, const float *__restrict__ clip_coef, AdamHyper h, const int64_t *__restrict__ step_dev) {
    const float coef = clip_coef ? *clip_coef : 1.0f;
    float step_size, bc2_sqrt;
    adam_step_consts(h, step_dev, step_size, bc2_sqrt);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_elem(h, step_size, bc2_sqrt, g[i] * coef, pp, mm, vv);
        m[i] = mm;
        v[i] = vv;
        p[i] = pp;
    }
}

}  // namespace tt

extern "C" int tt_sq_norm_accum(const float *x, int64_t n, float *out, void *workspace, size_t workspace_bytes,
                                void *stream) {
    using namespace tt;
    TT_CHECK_ARG(x && out && workspace && n >= 0, "bad argument");
    if (workspace_bytes < sizeof(float) * SQ_BLOCKS) { set_error("sq_norm workspace needs %zu bytes", sizeof(float) * SQ_BLOCKS); return TT_E_WORKSPACE; }
    if (n == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace);
    sq_norm_partial<<<SQ_BLOCKS, 256, 0, st>>>(x, n, partial);
    TT_LAUNCH_CHECK("sq_norm_partial");
    sq_norm_final<<<1, 1024, 0, st>>>(partial, SQ_BLOCKS, out);
    TT_LAUNCH_CHECK("sq_norm_final");
    return 0;
}

extern "C" int tt_clip_coef(const float *sq_terms, int n_terms, float max_norm, float *coef, float *total_norm,
                            void *stream) {
    TT_CHECK_ARG(sq_terms && coef && n_terms > 0, "bad argument");
    tt::clip_coef_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(sq_terms, n_terms, max_norm, coef, total_norm);
    TT_LAUNCH_CHECK("clip_coef_kernel");
    return 0;
}

extern "C" int tt_adam_flat(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                            const float *clip_coef, double lr, double beta1, double beta2, double eps,
                            const int64_t *step_dev, const double *lr_dev, void *stream) {
    TT_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step_dev && n >= 0, "bad argument");
    if (n == 0) return 0;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = static_cast<int64_t>(tt::sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    tt::adam_flat_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        param, grad, exp_avg, exp_avg_sq, n, clip_coef, tt::make_adam(lr, beta1, beta2, eps, lr_dev), step_dev);
    TT_LAUNCH_CHECK("adam_flat_kernel");
    return 0;
}
