// MUFU ex2 throughput per SM: f32 vs packed f16x2 / bf16x2 (run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/mufu_probe tools/mufu_probe.cu && /tmp/mufu_probe)
// Question behind it: the fused-CE backward rounds its probabilities to bf16 anyway; does the packed form give two
// exponentials per XU slot?
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(int reps, long long *out, uint32_t *sink, float seed) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(-(seed + i * 0.01f + threadIdx.x * 1e-4f));
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(v[i]));
            if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
            if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= v[i];
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    long long *dclk, h;
    uint32_t *sink;
    cudaMalloc(&dclk, 8);
    cudaMalloc(&sink, 4);
    const int reps = 4000;
    const char *names[3] = {"ex2.approx.ftz.f32   ", "ex2.approx.ftz.bf16x2", "ex2.approx.f16x2     "};
    for (int mode = 0; mode < 3; ++mode) {
        if (mode == 0) probe<0><<<1, 256>>>(reps, dclk, sink, 1.0f);
        if (mode == 1) probe<1><<<1, 256>>>(reps, dclk, sink, 1.0f);
        if (mode == 2) probe<2><<<1, 256>>>(reps, dclk, sink, 1.0f);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h, dclk, 8, cudaMemcpyDeviceToHost);
        const double instr = 256.0 * reps * 8;
        printf("%s: %lld cycles, %.2f thread-instructions / clk / SM (%.2f results / clk)\n", names[mode], h, instr / h,
               instr / h * (mode == 0 ? 1 : 2));
    }
    return 0;
}
