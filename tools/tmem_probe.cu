// TMEM -> register read throughput of one SM (run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tmem_probe tools/tmem_probe.cu -lcuda && /tmp/tmem_probe).
// W warps (warp w reads the 32 lanes of quarter w % 4) each issue `reps` x 4 tcgen05.ld.32x32b.x32 (128 columns,
// 16 KB per warp and repetition) with one tcgen05.wait::ld per repetition.  Prints bytes per cycle per SM: this is
// the roofline of every epilogue that has to see each fp32 accumulator once (fused CE, top-K filter).
#include <stdio.h>
#include <stdlib.h>

#include "../recommendsystemproject_b200/csrc/tc_common.cuh"

namespace tt {
void set_error(const char *, ...) {}
int sm_count() { return 148; }
}  // namespace tt

using namespace tt::tc;

template <int COLS_PER_LD>
__global__ void __launch_bounds__(512, 1) ldtm_probe(int reps, long long *out, uint32_t *sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        uint32_t v[4][32];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) tmem_ld_32x32(base + qq * 32, v[qq]);
        tmem_ld_wait();
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
#pragma unroll
            for (int j = 0; j < 32; j += 8) acc ^= v[qq][j];
    }
    const long long t1 = clock64();
    __syncthreads();
    if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_slot);
}

int main() {
    long long *dclk;
    uint32_t *sink;
    cudaMalloc(&dclk, 16 * 8);
    cudaMalloc(&sink, 4);
    const int reps = 2000;
    for (int warps : {1, 2, 4, 8, 16}) {
        ldtm_probe<32><<<1, warps * 32>>>(reps, dclk, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[16];
        cudaMemcpy(h, dclk, warps * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
        const double bytes = static_cast<double>(warps) * reps * 128 * 32 * 4;
        printf("%2d warps (%d per TMEM quarter): %lld cycles for %d x 128 columns per warp -> %.1f B/clk per SM, %.1f B/clk per warp\n",
               warps, warps >= 4 ? warps / 4 : 1, mx, reps, bytes / mx, bytes / mx / warps);
    }
    return 0;
}
