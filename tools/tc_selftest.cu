// Known-answer self-test of the tcgen05 operand conventions used by the fused kernels
// (run on a B200:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tc_selftest tools/tc_selftest.cu -lcuda
//  && tools/bin/tc_selftest).  One 128x128x128 bf16 product per mode:
//   mode 0  SS, B K-major              D = A * Y^T   (the forward kernel's convention)
//   mode 1  SS, B MN-major             D = A * Y     (same smem bytes, MN-major descriptor)
//   mode 2  TS (A in TMEM), B K-major  D = A * Y^T
//   mode 3  TS (A in TMEM), B MN-major D = A * Y
// Prints max |err| against a host fp32 reference for each (mode, lbo, sbo) tried.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../recommendsystemproject_b200/csrc/tc_common.cuh"

namespace tt {
void set_error(const char *, ...) {}
int sm_count() { return 148; }
}  // namespace tt

using namespace tt::tc;

constexpr int T = 128;
constexpr int KBLOCK_BYTES = T * 128;

__global__ void __launch_bounds__(128, 1)
selftest_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_y,
                const __nv_bfloat16 *__restrict__ a_gmem, int mode, uint32_t lbo, uint32_t sbo, uint32_t kadv,
                float *__restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *a_tile = smem;
    uint8_t *y_tile = smem + 2 * KBLOCK_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(y_tile + 2 * KBLOCK_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars[0], 4 * KBLOCK_BYTES);
        for (int kb = 0; kb < 2; ++kb) {
            tma_load_2d(a_tile + kb * KBLOCK_BYTES, &map_a, &bars[0], kb * 64, 0);
            tma_load_2d(y_tile + kb * KBLOCK_BYTES, &map_y, &bars[0], kb * 64, 0);
        }
    }
    if (mode >= 2) {
        // thread = row: pack the row's 128 bf16 into 64 words and store them to TMEM columns [128, 192)
        const int row = warp * 32 + lane;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a_gmem + row * T);
        uint32_t r[32];
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = src[h * 32 + j];
            tmem_st_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + 128 + h * 32, r);
        }
        tmem_st_wait();
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const bool mn = (mode & 1) != 0;
        const uint32_t idesc = idesc_bf16_f32(T, T, 0, mn ? 1 : 0);
        const uint32_t a_addr = smem_u32(a_tile), y_addr = smem_u32(y_tile);
        for (int k = 0; k < T / 16; ++k) {
            const uint32_t koff = (k / 4) * KBLOCK_BYTES + (k % 4) * 32;
            const uint64_t bdesc = mn ? smem_desc_mn_sw128(y_addr + k * kadv, lbo, sbo) : smem_desc_k_sw128(y_addr + koff);
            if (mode >= 2)
                umma_f16_ts(tmem_base, tmem_base + 128 + k * 8, bdesc, idesc, k > 0 ? 1u : 0u);
            else
                umma_f16(tmem_base, smem_desc_k_sw128(a_addr + koff), bdesc, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c = 0; c < T / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[row * T + c * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
}


// ---- throughput probe: `reps` x 8 back-to-back MMAs (one 128x128x128 product each), cycles per MMA instruction
__global__ void __launch_bounds__(128, 1)
perf_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_y, int mode, int reps,
            int n_cols, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *a_tile = smem;
    uint8_t *y_tile = smem + 2 * KBLOCK_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(y_tile + 2 * KBLOCK_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars[0], 4 * KBLOCK_BYTES);
        for (int kb = 0; kb < 2; ++kb) {
            tma_load_2d(a_tile + kb * KBLOCK_BYTES, &map_a, &bars[0], kb * 64, 0);
            tma_load_2d(y_tile + kb * KBLOCK_BYTES, &map_y, &bars[0], kb * 64, 0);
        }
    }
    __syncwarp();
    if (warp == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
    }
    if (warp == 0 && (mode >= 4 ? elect_one_sync() : threadIdx.x == 0)) {
        const bool mn = (mode & 1) != 0;
        const uint32_t idesc = idesc_bf16_f32(T, n_cols, 0, mn ? 1 : 0);
        const uint32_t a_addr = smem_u32(a_tile), y_addr = smem_u32(y_tile);
        const long long t0 = clock64();
        if (mode == 6 || mode == 7) {   // two independent accumulation chains interleaved (6: SS+SS, 7: SS+TS MN-major)
            const uint64_t ad = smem_desc_k_sw128(a_addr), bdk = smem_desc_k_sw128(y_addr);
            const uint64_t bdm = smem_desc_mn_sw128(y_addr, KBLOCK_BYTES, 1024);
            const uint32_t idesc_mn = idesc_bf16_f32(T, n_cols, 0, 1), idesc_k = idesc_bf16_f32(T, n_cols, 0, 0);
            for (int r = 0; r < reps; r += 2) {
#pragma unroll
                for (int k = 0; k < T / 16; ++k) {
                    const uint32_t koff = ((k / 4) * KBLOCK_BYTES + (k % 4) * 32) >> 4;
                    if (k == 0) umma_f16_first(tmem_base, ad + koff, bdk + koff, idesc_k); else umma_f16_acc(tmem_base, ad + koff, bdk + koff, idesc_k);
                    if (mode == 6) { if (k == 0) umma_f16_first(tmem_base + 128, ad + koff, bdk + koff, idesc_k); else umma_f16_acc(tmem_base + 128, ad + koff, bdk + koff, idesc_k); }
                    else { if (k == 0) umma_f16_ts_first(tmem_base + 128, tmem_base + 256, bdm, idesc_mn); else umma_f16_ts_acc(tmem_base + 128, tmem_base + 256 + k * 8, bdm + k * 128, idesc_mn); }
                }
            }
        } else if (mode == 8) {   // same accumulator chain but M=128 x N=256 per instruction (B = the tile twice: LBO trick not needed, N=256 reads 256 rows)
            const uint64_t ad = smem_desc_k_sw128(a_addr), bdk = smem_desc_k_sw128(a_addr);   // 256 "n" rows = a_tile + y_tile (contiguous 64-col blocks differ, timing only)
            const uint32_t idesc256 = idesc_bf16_f32(T, 256, 0, 0);
            for (int r = 0; r < reps; r += 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // K within the first 64-col block only (timing probe; results unused)
                    const uint32_t koff = (k * 32) >> 4;
                    if (k == 0) umma_f16_first(tmem_base, ad + koff, bdk + koff, idesc256); else umma_f16_acc(tmem_base, ad + koff, bdk + koff, idesc256);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t koff = (k * 32) >> 4;
                    umma_f16_acc(tmem_base, ad + koff, bdk + koff, idesc256);
                }
            }
        } else if (mode >= 4) {   // hoisted descriptors + immediate predicates (the production issue loop)
            const uint64_t ad = smem_desc_k_sw128(a_addr), bdk = smem_desc_k_sw128(y_addr);
            const uint64_t bdm = smem_desc_mn_sw128(y_addr, KBLOCK_BYTES, 1024);
            for (int r = 0; r < reps; ++r) {
                const uint32_t acc = tmem_base + (r & 1) * 128;
#pragma unroll
                for (int k = 0; k < T / 16; ++k) {
                    const uint32_t koff = ((k / 4) * KBLOCK_BYTES + (k % 4) * 32) >> 4;
                    if (mode == 4) { if (k == 0) umma_f16_first(acc, ad + koff, bdk + koff, idesc); else umma_f16_acc(acc, ad + koff, bdk + koff, idesc); }
                    else { if (k == 0) umma_f16_ts_first(acc, tmem_base + 256, bdm, idesc); else umma_f16_ts_acc(acc, tmem_base + 256 + k * 8, bdm + k * 128, idesc); }
                }
            }
        } else
        for (int r = 0; r < reps; ++r) {
            const uint32_t acc = tmem_base + (r & 1) * 128;   // alternate two accumulators like the real kernels
#pragma unroll
            for (int k = 0; k < T / 16; ++k) {
                const uint32_t koff = (k / 4) * KBLOCK_BYTES + (k % 4) * 32;
                const uint64_t bdesc = mn ? smem_desc_mn_sw128(y_addr + k * 2048, KBLOCK_BYTES, 1024) : smem_desc_k_sw128(y_addr + koff);
                if (mode >= 2) umma_f16_ts(acc, tmem_base + 256 + k * 8, bdesc, idesc, k > 0 ? 1u : 0u);
                else umma_f16(acc, smem_desc_k_sw128(a_addr + koff), bdesc, idesc, k > 0 ? 1u : 0u);
            }
        }
        const long long t1 = clock64();
        umma_commit(&bars[1]);
        mbar_wait(&bars[1], 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    std::vector<float> a(T * T), y(T * T);
    std::vector<__nv_bfloat16> ab(T * T), yb(T * T);
    srand(7);
    for (int i = 0; i < T * T; ++i) {
        a[i] = bf16_round((rand() % 2001 - 1000) / 1000.0f);
        y[i] = bf16_round((rand() % 2001 - 1000) / 1000.0f);
        ab[i] = __float2bfloat16(a[i]);
        yb[i] = __float2bfloat16(y[i]);
    }
    std::vector<float> ref_t(T * T), ref_n(T * T);  // A*Y^T and A*Y
    for (int i = 0; i < T; ++i)
        for (int j = 0; j < T; ++j) {
            double st = 0, sn = 0;
            for (int k = 0; k < T; ++k) {
                st += static_cast<double>(a[i * T + k]) * y[j * T + k];
                sn += static_cast<double>(a[i * T + k]) * y[k * T + j];
            }
            ref_t[i * T + j] = static_cast<float>(st);
            ref_n[i * T + j] = static_cast<float>(sn);
        }
    __nv_bfloat16 *da, *dy;
    float *dout;
    cudaMalloc(&da, T * T * 2);
    cudaMalloc(&dy, T * T * 2);
    cudaMalloc(&dout, T * T * 4);
    cudaMemcpy(da, ab.data(), T * T * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, yb.data(), T * T * 2, cudaMemcpyHostToDevice);
    CUtensorMap ma, my;
    if (make_tmap_bf16_rows(&ma, da, T, T, T) || make_tmap_bf16_rows(&my, dy, T, T, T)) { printf("tmap failed\n"); return 1; }
    const size_t smem = 1024 + 4 * KBLOCK_BYTES + 256;
    cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    struct Case { int mode; uint32_t lbo, sbo, kadv; };
    const Case cases[] = {
        {0, 0, 0, 0},
        {1, KBLOCK_BYTES, 1024, 2048},   // derived from the canonical MN-major SW128 layout
        {1, 1024, KBLOCK_BYTES, 2048},   // LBO/SBO swapped, in case the convention is the other way round
        {2, 0, 0, 0},
        {3, KBLOCK_BYTES, 1024, 2048},
        {3, 1024, KBLOCK_BYTES, 2048},
    };
    std::vector<float> out(T * T);
    int bad = 0;
    for (const Case &c : cases) {
        cudaMemset(dout, 0, T * T * 4);
        selftest_kernel<<<1, 128, smem>>>(ma, my, da, c.mode, c.lbo, c.sbo, c.kadv, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d lbo %u sbo %u: CUDA error %s\n", c.mode, c.lbo, c.sbo, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(out.data(), dout, T * T * 4, cudaMemcpyDeviceToHost);
        const std::vector<float> &ref = (c.mode & 1) ? ref_n : ref_t;
        double err = 0;
        for (int i = 0; i < T * T; ++i) err = fmax(err, fabs(static_cast<double>(out[i]) - ref[i]));
        printf("mode %d lbo %5u sbo %5u kadv %4u: max|err| = %.3e %s\n", c.mode, c.lbo, c.sbo, c.kadv, err,
               err < 1e-3 ? "OK" : "MISMATCH");
        if (err >= 1e-3) ++bad;
    }
    printf("selftest done, %d mismatching case(s)\n", bad);
    // throughput probe
    long long *dclk;
    cudaMalloc(&dclk, 16);
    cudaFuncSetAttribute(perf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const char *names[9] = {"SS  B K-major ", "SS  B MN-major", "TS  B K-major ", "TS  B MN-major", "SS  K-major  hoisted", "TS  MN-major hoisted", "2 chains SS+SS       ", "2 chains SS+TS(MN)   ", "1 chain  N=256 (x8)  "};
    for (int n_cols : {128, 64}) {
        for (int mode = 0; mode < 9; ++mode) {
            if (mode == 8 && n_cols != 128) continue;
            for (int reps : {4, 64}) {
                perf_kernel<<<1, 128, smem>>>(ma, my, mode, reps, n_cols, dclk);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("perf mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 3; }
                long long h[2];
                cudaMemcpy(h, dclk, 16, cudaMemcpyDeviceToHost);
                printf("perf %s M=128 N=%3d K=16: %3d x 8 MMAs  issue %7lld cyc  complete %7lld cyc  -> %.1f cyc/MMA\n", names[mode],
                       n_cols, reps, h[0], h[1], static_cast<double>(h[1]) / (reps * 8));
            }
        }
    }
    return 0;
}
