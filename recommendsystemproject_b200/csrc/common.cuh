// Shared helpers for the tt_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tt_b200.h"

namespace tt {

void set_error(const char *fmt, ...);

inline int cuda_status(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
}

#define TT_CHECK_ARG(cond, msg)                       \
    do {                                              \
        if (!(cond)) {                                \
            tt::set_error("bad argument: %s", msg);   \
            return TT_E_BADARG;                       \
        }                                             \
    } while (0)

#define TT_LAUNCH_CHECK(name)                                              \
    do {                                                                   \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) return tt::cuda_status(e__, name);        \
    } while (0)

int sm_count();

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct Workspace {
    char *base;
    size_t size;
    size_t off;
    Workspace(void *p, size_t n) : base(static_cast<char *>(p)), size(n), off(0) {}
    template <typename T>
    T *take(size_t count) {
        off = align_up(off, 256);
        T *p = reinterpret_cast<T *>(base + off);
        off += count * sizeof(T);
        return p;
    }
    bool ok() const { return off <= size; }
};

// Adam hyper-parameters: doubles as the host gave them (torch computes 1-beta and the bias
// corrections in double) plus the fp32 constants the element loop uses.
struct AdamHyper {
    double lr, beta1, beta2;
    const double *lr_dev;   // when not NULL the learning rate is read from device memory (schedulers + CUDA graphs)
    float b1, omb1, b2, omb2, eps;
};
inline AdamHyper make_adam(double lr, double beta1, double beta2, double eps, const double *lr_dev = nullptr) {
    AdamHyper h;
    h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.lr_dev = lr_dev;
    h.b1 = static_cast<float>(beta1); h.omb1 = static_cast<float>(1.0 - beta1);
    h.b2 = static_cast<float>(beta2); h.omb2 = static_cast<float>(1.0 - beta2);
    h.eps = static_cast<float>(eps);
    return h;
}
__device__ __forceinline__ void adam_step_consts(const AdamHyper &h, const int64_t *step_dev, float &step_size,
                                                 float &bc2_sqrt) {
    const double t = static_cast<double>(*step_dev);
    const double lr = h.lr_dev ? *h.lr_dev : h.lr;
    step_size = static_cast<float>(lr / (1.0 - pow(h.beta1, t)));
    bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(h.beta2, t)));
}
__device__ __forceinline__ void adam_elem(const AdamHyper &h, float step_size, float bc2_sqrt, float g, float &p,
                                          float &m, float &v) {
    m = h.b1 * m + h.omb1 * g;
    v = h.b2 * v + h.omb2 * g * g;
    p -= step_size * (m / (sqrtf(v) / bc2_sqrt + h.eps));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 16-byte vector of table elements -> floats
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void load(const float *p, float (&f)[4]) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    __device__ static __forceinline__ void store(float *p, const float (&f)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
};
template <>
struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void load(const __nv_bfloat16 *p, float (&f)[8]) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static __forceinline__ void store(__nv_bfloat16 *p, const float (&f)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

}  // namespace tt
