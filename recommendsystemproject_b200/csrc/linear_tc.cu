// nn.Linear forward / input gradient / weight gradient on the 5th-gen tensor cores in TF32: tcgen05.mma kind::tf32
// straight from the fp32 activations (no conversion pass, fp32 accumulation in TMEM), operands staged by TMA.
//
// Replaces addmm and its two backward GEMMs for every Linear of the towers (Tower.py:17-24, GenericTower.py:221,
// SequenceFeatureProcessor.py:33) when the caller allows TF32 (torch.backends.cuda.matmul.allow_tf32); the exact fp32
// kernels of linear_grad.cu stay the parity path.
//
//   FWD    Y[R, O]  = X[R, I] . W[O, I]^T + b (+ ReLU)     A = X  K-major,   B = W  K-major
//   DGRAD  dX[R, I] = dY[R, O] . W[O, I]                  A = dY K-major,   B = W  MN-major (no transposed copy)
//   WGRAD  dW[O, I] = dY[R, O]^T . X[R, I]                A = dY MN-major,  B = X  MN-major, split over R across the
//                                                         chip; partials added in a fixed order (deterministic)
// MN-major TF32 operands have ONE legal shared-memory layout (CUTLASS sm100_common.inl: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout"): 128-byte rows of 32 fp32, one row per k, swizzled in 32-byte chunks
// with a 4-row period (UMMA layout type SWIZZLE_128B_BASE32B, canonical ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte
// units) -- exactly what a TMA box of [32 k-rows][32 fp32] with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes: 4-row atoms
// 512 B apart along K (SBO), 32-element blocks one box = 4096 B apart along M/N (LBO).
//
// One persistent kernel, 6 warps per CTA: warp 0 TMA producer (ring of 4 stages of 32 k), warp 1 issues the MMAs
// (M = 128, N = BN in {128, 256}, K = 8 per instruction) into one of two TMEM accumulators, warps 2..5 drain the other
// accumulator (thread = output row): bias, ReLU, 128-byte row segments straight to global memory.
// Roofline: HBM for the tower shapes (R >> I, O): X is read once, Y written once.
#include "tc_common.cuh"

namespace tt {

using namespace tt::tc;

constexpr int LT_BM = 128;
constexpr int LT_BK = 32;           // fp32 elements per k block = one 128-byte swizzle row
constexpr int LT_STAGES = 4;
constexpr int LT_THREADS = 192;
constexpr int LT_A_BYTES = LT_BM * 128;

enum { LT_FWD = 0, LT_DGRAD = 1, LT_WGRAD = 2 };

__host__ __device__ constexpr uint32_t idesc_tf32_f32(int m, int n, int a_mn_major, int b_mn_major) {
    return (1u << 4)                                   // D format: F32
           | (2u << 7)                                 // A format: TF32
           | (2u << 10)                                // B format: TF32
           | (static_cast<uint32_t>(a_mn_major) << 15)
           | (static_cast<uint32_t>(b_mn_major) << 16)
           | (static_cast<uint32_t>(n >> 3) << 17)
           | (static_cast<uint32_t>(m >> 4) << 24);
}

// MN-major TF32 operand: layout type 1 = SWIZZLE_128B_BASE32B (bit layout: UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(1) << 61;
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct LinTcParams {
    int64_t M, N;             // output rows / columns of the GEMM
    int m_tiles, n_tiles;     // 128 x BN output tiles
    int k_blocks;             // ceil(K / 32)
    int splits, kb_per_split; // split-K (WGRAD); 1 / k_blocks otherwise
    const float *bias;        // [N] or null (FWD)
    int relu;
    float *out;               // [M, ldo] final output (splits == 1) ...
    int64_t ldo;
    float *partial;           // ... or [splits][m_tiles*128][n_tiles*BN] raw partials (splits > 1)
};

template <int MODE, int BN>
__global__ void __launch_bounds__(LT_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const LinTcParams prm) {
    constexpr bool A_MN = MODE == LT_WGRAD;
    constexpr bool B_MN = MODE != LT_FWD;
    constexpr int B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = LT_A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + LT_STAGES * STAGE_BYTES);
    uint64_t *full = bars;                      // [ST] TMA -> MMA
    uint64_t *empty = full + LT_STAGES;         // [ST] MMA -> TMA
    uint64_t *acc_full = empty + LT_STAGES;     // [2]  accumulator complete
    uint64_t *acc_empty = acc_full + 2;         // [2]  accumulator drained (128 arrivals)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total = static_cast<int64_t>(prm.splits) * prm.m_tiles * prm.n_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
        for (int s = 0; s < LT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int64_t t, int &s, int &m, int &n) {
        n = static_cast<int>(t % prm.n_tiles);
        const int64_t r = t / prm.n_tiles;
        m = static_cast<int>(r % prm.m_tiles);
        s = static_cast<int>(r / prm.m_tiles);
    };

    if (warp == 0) {
        // ===== TMA producer =====
        int it = 0;   // running k-block counter over all my tiles (ring position)
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
            int s, m, n;
            decode(t, s, m, n);
            const int kb0 = s * prm.kb_per_split, kb1 = min(prm.k_blocks, kb0 + prm.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int stage = it % LT_STAGES;
                mbar_wait(&empty[stage], ((it / LT_STAGES) & 1) ^ 1);
                if (elect_one_sync()) {
                    uint8_t *a_dst = smem + stage * STAGE_BYTES;
                    uint8_t *b_dst = a_dst + LT_A_BYTES;
                    mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                    if (!A_MN) {
                        tma_load_2d(a_dst, &map_a, &full[stage], kb * LT_BK, m * LT_BM);
                    } else {
#pragma unroll
                        for (int j = 0; j < LT_BM / 32; ++j)
                            tma_load_2d(a_dst + j * 4096, &map_a, &full[stage], m * LT_BM + j * 32, kb * LT_BK);
                    }
                    if (!B_MN) {
                        tma_load_2d(b_dst, &map_b, &full[stage], kb * LT_BK, n * BN);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 32; ++j)
                            tma_load_2d(b_dst + j * 4096, &map_b, &full[stage], n * BN + j * 32, kb * LT_BK);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = idesc_tf32_f32(LT_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
        int it = 0, tl = 0;
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x, ++tl) {
            int s, m, n;
            decode(t, s, m, n);
            const int kb0 = s * prm.kb_per_split, kb1 = min(prm.k_blocks, kb0 + prm.kb_per_split);
            const int b = tl & 1;
            if (tl >= 2) mbar_wait(&acc_empty[b], ((tl >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + b * BN;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int stage = it % LT_STAGES;
                mbar_wait(&full[stage], (it / LT_STAGES) & 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t b_addr = a_addr + LT_A_BYTES;
                    const uint64_t adesc = A_MN ? smem_desc_mn_sw128_32b(a_addr, 4096, 512) : smem_desc_k_sw128(a_addr);
                    const uint64_t bdesc = B_MN ? smem_desc_mn_sw128_32b(b_addr, 4096, 512) : smem_desc_k_sw128(b_addr);
#pragma unroll
                    for (int j = 0; j < LT_BK / 8; ++j) {
                        // K-major: 8 tf32 = 32 B inside the 128-byte swizzle row; MN-major: 8 k-rows = two 512-byte atoms
                        const uint64_t ao = A_MN ? static_cast<uint64_t>((j * 1024) >> 4) : static_cast<uint64_t>((j * 32) >> 4);
                        const uint64_t bo = B_MN ? static_cast<uint64_t>((j * 1024) >> 4) : static_cast<uint64_t>((j * 32) >> 4);
                        umma_tf32(acc, adesc + ao, bdesc + bo, idesc, (kb > kb0 || j > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                    if (kb == kb1 - 1) umma_commit(&acc_full[b]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: 4 warps = 128 accumulator lanes, thread = output row =====
        const int quarter = warp & 3;
        const int r_in_tile = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        int tl = 0;
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x, ++tl) {
            int s, m, n;
            decode(t, s, m, n);
            const int b = tl & 1;
            mbar_wait(&acc_full[b], (tl >> 1) & 1);
            tc_fence_after();
            const int64_t row = static_cast<int64_t>(m) * LT_BM + r_in_tile;
            const int col0 = n * BN;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(lane_addr + b * BN + c, r);
                tmem_ld_wait();
                if (prm.splits > 1) {
                    float *dst = prm.partial + (static_cast<int64_t>(s) * prm.m_tiles * LT_BM + row) * (static_cast<int64_t>(prm.n_tiles) * BN) + col0 + c;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4 *>(dst + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                } else if (row < prm.M) {
                    float *dst = prm.out + row * prm.ldo + col0 + c;
                    if (col0 + c + 32 <= prm.N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float v[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                v[e] = __uint_as_float(r[j + e]);
                                if (prm.bias) v[e] += __ldg(prm.bias + col0 + c + j + e);
                                if (prm.relu) v[e] = fmaxf(v[e], 0.f);
                            }
                            *reinterpret_cast<float4 *>(dst + j) = make_float4(v[0], v[1], v[2], v[3]);
                        }
                    } else {
                        for (int j = 0; j < 32; ++j) {
                            if (col0 + c + j >= prm.N) break;
                            float v = __uint_as_float(r[j]);
                            if (prm.bias) v += __ldg(prm.bias + col0 + c + j);
                            if (prm.relu) v = fmaxf(v, 0.f);
                            dst[j] = v;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[b]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<512>(tmem_base);
    }
}

// out[m, n] (+)= sum_s partial[s][m][n]   (fixed order over s)
__global__ void __launch_bounds__(256)
linear_tc_reduce(const float *__restrict__ partial, int splits, int64_t m_pad, int64_t n_pad, int64_t M, int64_t N,
                 float *__restrict__ out, int64_t ldo, int accumulate) {
    const int64_t n4 = N / 4;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < M * n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t m = i / n4, c = (i - m * n4) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *src = reinterpret_cast<const float4 *>(partial + m * n_pad + c);
        const int64_t step = m_pad * n_pad / 4;   // float4s between two splits (n_pad % 4 == 0)
        int s = 0;
        for (; s + 8 <= splits; s += 8) {         // eight loads in flight, added in split order
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (s + j) * step);
#pragma unroll
            for (int j = 0; j < 8; ++j) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
        }
        for (; s < splits; ++s) {
            const float4 v = __ldg(src + s * step);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        float *o = out + m * ldo + c;
        if (accumulate) {
            const float4 p = *reinterpret_cast<const float4 *>(o);
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        *reinterpret_cast<float4 *>(o) = a;
    }
}

// bias gradient: column sums of dY[R, O]; 128 columns x row chunks per block, partials added in fixed order
__global__ void __launch_bounds__(256)
colsum_partial(const float *__restrict__ x, int64_t rows, int cols, int rows_per_chunk, float *__restrict__ partial) {
    __shared__ float4 sh[8][32];
    const int c4 = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 * 4 < cols)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * cols) + c4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c4 * 4 < cols) {
        for (int k = 1; k < 8; ++k) { const float4 a = sh[k][threadIdx.x]; s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
        *(reinterpret_cast<float4 *>(partial + static_cast<int64_t>(blockIdx.y) * cols) + c4) = s;
    }
}
// 32 columns x 32 chunk lanes per block; lane j adds chunks j, j+32, ... (four loads in flight), lanes added in order
__global__ void __launch_bounds__(1024)
colsum_final(const float *__restrict__ partial, int n_chunks, int cols, float *__restrict__ out, int accumulate) {
    __shared__ float sh[32][32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float a = 0.f;
    if (c < cols) {
        for (int k0 = threadIdx.y; k0 < n_chunks; k0 += 128) {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k0 + u * 32 < n_chunks) v[u] = partial[static_cast<int64_t>(k0 + u * 32) * cols + c];
            a += (v[0] + v[1]) + (v[2] + v[3]);
        }
    }
    sh[threadIdx.y][threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        for (int j = 1; j < 32; ++j) a += sh[j][threadIdx.x];
        out[c] = accumulate ? out[c] + a : a;
    }
}

static inline int colsum_rows_per_chunk(int64_t rows, int cols) {
    // ~2 waves of 128-column blocks, at least 32 rows per chunk; a function of the shape only (reproducible order)
    const int col_blocks = (cols / 4 + 31) / 32;
    int r = 1024;
    while (r > 32 && ((rows + r - 1) / r) * col_blocks < 2 * 148) r >>= 1;
    return r;
}

// row-major fp32 [rows, cols] -> boxes of {32 columns (128 B), box_rows rows}, 128B swizzle, zero fill out of range
static int make_tmap_f32(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int box_rows, bool atom32 = false) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point not available"); return TT_E_DEVICE; }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 4};
    cuuint32_t box[2] = {32, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(f32) failed (CUresult %d)", static_cast<int>(r)); return TT_E_BADARG; }
    return 0;
}

template <int MODE, int BN>
static int launch_linear_tc(const CUtensorMap &ma, const CUtensorMap &mb, const LinTcParams &prm, cudaStream_t st) {
    constexpr size_t smem = LT_STAGES * (LT_A_BYTES + BN * 128) + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(linear_tc_kernel)");
        attr_set = true;
    }
    const int64_t total = static_cast<int64_t>(prm.splits) * prm.m_tiles * prm.n_tiles;
    const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
    linear_tc_kernel<MODE, BN><<<grid, LT_THREADS, smem, st>>>(ma, mb, prm);
    TT_LAUNCH_CHECK("linear_tc_kernel");
    return 0;
}

static inline bool lt_ok(const void *p, int64_t cols) { return reinterpret_cast<uintptr_t>(p) % 16 == 0 && cols % 4 == 0; }

}  // namespace tt

extern "C" int tt_linear_tc_supported(int64_t rows, int n_out, int n_in) {
    return (rows >= 1 && n_out % 4 == 0 && n_in % 4 == 0 && n_out >= 8 && n_in >= 8) ? 1 : 0;
}

extern "C" int tt_linear_fwd_tc(const float *input, const float *weight, const float *bias, int64_t rows, int n_out, int n_in,
                                int relu, float *out, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(input && weight && out && rows > 0, "bad argument");
    TT_CHECK_ARG(lt_ok(input, n_in) && lt_ok(weight, n_in) && lt_ok(out, n_out), "tf32 linear: 16-byte aligned rows (dims % 4 == 0)");
    const int bn = n_out > 128 ? 256 : 128;
    CUtensorMap ma, mb;
    int rc;
    if ((rc = make_tmap_f32(&ma, input, rows, n_in, LT_BM))) return rc;
    if ((rc = make_tmap_f32(&mb, weight, n_out, n_in, bn))) return rc;
    LinTcParams p{};
    p.M = rows; p.N = n_out;
    p.m_tiles = static_cast<int>((rows + LT_BM - 1) / LT_BM);
    p.n_tiles = (n_out + bn - 1) / bn;
    p.k_blocks = (n_in + LT_BK - 1) / LT_BK;
    p.splits = 1; p.kb_per_split = p.k_blocks;
    p.bias = bias; p.relu = relu; p.out = out; p.ldo = n_out;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return bn == 256 ? launch_linear_tc<LT_FWD, 256>(ma, mb, p, st) : launch_linear_tc<LT_FWD, 128>(ma, mb, p, st);
}

extern "C" int tt_linear_dgrad_tc(const float *grad_out, const float *weight, int64_t rows, int n_out, int n_in,
                                  float *grad_input, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(grad_out && weight && grad_input && rows > 0, "bad argument");
    TT_CHECK_ARG(lt_ok(grad_out, n_out) && lt_ok(weight, n_in) && lt_ok(grad_input, n_in), "tf32 linear: 16-byte aligned rows (dims % 4 == 0)");
    const int bn = n_in > 128 ? 256 : 128;
    CUtensorMap ma, mb;
    int rc;
    if ((rc = make_tmap_f32(&ma, grad_out, rows, n_out, LT_BM))) return rc;
    if ((rc = make_tmap_f32(&mb, weight, n_out, n_in, 32, true))) return rc;       // [32 k = out rows][32 n = in cols] boxes
    LinTcParams p{};
    p.M = rows; p.N = n_in;
    p.m_tiles = static_cast<int>((rows + LT_BM - 1) / LT_BM);
    p.n_tiles = (n_in + bn - 1) / bn;
    p.k_blocks = (n_out + LT_BK - 1) / LT_BK;
    p.splits = 1; p.kb_per_split = p.k_blocks;
    p.out = grad_input; p.ldo = n_in;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return bn == 256 ? launch_linear_tc<LT_DGRAD, 256>(ma, mb, p, st) : launch_linear_tc<LT_DGRAD, 128>(ma, mb, p, st);
}

extern "C" int tt_linear_wgrad_tc_workspace(int64_t rows, int n_out, int n_in, size_t *bytes_host) {
    using namespace tt;
    TT_CHECK_ARG(bytes_host && rows > 0 && n_out > 0 && n_in > 0, "bad size");
    const int bn = n_in > 128 ? 256 : 128;
    const int64_t m_tiles = (n_out + LT_BM - 1) / LT_BM, n_tiles = (n_in + bn - 1) / bn;
    const int64_t k_blocks = (rows + LT_BK - 1) / LT_BK;
    int64_t splits = sm_count() / (m_tiles * n_tiles);
    if (splits < 1) splits = 1;
    if (splits > k_blocks) splits = k_blocks;
    const int64_t chunks = (rows + 31) / 32;     // upper bound of colsum_rows_per_chunk's chunk count
    *bytes_host = static_cast<size_t>(splits * m_tiles * LT_BM * n_tiles * bn + chunks * n_out) * sizeof(float) + 512;
    return 0;
}

extern "C" int tt_linear_wgrad_tc(const float *grad_out, const float *input, int64_t rows, int n_out, int n_in,
                                  float *grad_weight, float *grad_bias, int accumulate, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(grad_out && input && grad_weight && workspace && rows > 0, "bad argument");
    TT_CHECK_ARG(lt_ok(grad_out, n_out) && lt_ok(input, n_in) && lt_ok(grad_weight, n_in), "tf32 linear: 16-byte aligned rows (dims % 4 == 0)");
    size_t need = 0;
    tt_linear_wgrad_tc_workspace(rows, n_out, n_in, &need);
    if (workspace_bytes < need) { set_error("linear_wgrad_tc workspace too small: need %zu have %zu", need, workspace_bytes); return TT_E_WORKSPACE; }
    TT_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "workspace alignment");
    const int bn = n_in > 128 ? 256 : 128;
    CUtensorMap ma, mb;
    int rc;
    if ((rc = make_tmap_f32(&ma, grad_out, rows, n_out, 32, true))) return rc;    // [32 k = rows][32 m = out cols] boxes
    if ((rc = make_tmap_f32(&mb, input, rows, n_in, 32, true))) return rc;       // [32 k = rows][32 n = in cols] boxes
    LinTcParams p{};
    p.M = n_out; p.N = n_in;
    p.m_tiles = (n_out + LT_BM - 1) / LT_BM;
    p.n_tiles = (n_in + bn - 1) / bn;
    p.k_blocks = static_cast<int>((rows + LT_BK - 1) / LT_BK);
    int splits = sm_count() / (p.m_tiles * p.n_tiles);
    if (splits < 1) splits = 1;
    if (splits > p.k_blocks) splits = p.k_blocks;
    p.kb_per_split = (p.k_blocks + splits - 1) / splits;
    p.splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    float *partial = static_cast<float *>(workspace);
    const int64_t m_pad = static_cast<int64_t>(p.m_tiles) * LT_BM, n_pad = static_cast<int64_t>(p.n_tiles) * bn;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // always through the partial buffer (even for one split): the reduce kernel also implements `accumulate`
    p.partial = partial;
    const int real_splits = p.splits;
    if (real_splits == 1) { p.out = partial; p.ldo = n_pad; p.M = m_pad; p.N = n_pad; }   // plain stores of the padded tile grid
    rc = bn == 256 ? launch_linear_tc<LT_WGRAD, 256>(ma, mb, p, st) : launch_linear_tc<LT_WGRAD, 128>(ma, mb, p, st);
    if (rc) return rc;
    int64_t blocks = (static_cast<int64_t>(n_out) * (n_in / 4) + 255) / 256;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    linear_tc_reduce<<<static_cast<unsigned>(blocks), 256, 0, st>>>(partial, real_splits, m_pad, n_pad, n_out, n_in, grad_weight, n_in, accumulate);
    TT_LAUNCH_CHECK("linear_tc_reduce");
    if (grad_bias) {
        float *cs = partial + static_cast<int64_t>(real_splits) * m_pad * n_pad;
        const int rpc = colsum_rows_per_chunk(rows, n_out);
        const int chunks = static_cast<int>((rows + rpc - 1) / rpc);
        dim3 grid((n_out / 4 + 31) / 32, chunks), block(32, 8);
        colsum_partial<<<grid, block, 0, st>>>(grad_out, rows, n_out, rpc, cs);
        colsum_final<<<(n_out + 31) / 32, dim3(32, 32), 0, st>>>(cs, chunks, n_out, grad_bias, accumulate);
        TT_LAUNCH_CHECK("colsum");
    }
    return 0;
}
