// Embedding gather + sum/mean/max pooling forward (kernel 1 of the hot path).
//
// Replaces aten::embedding followed by mean/sum/max(dim=1) at
// GenericTower.py:153-160,182 and SequenceFeatureProcessor.py:60-68 of the
// reference; the [B, L, D] intermediate is never materialised.
//
// Layout: a table row is D*sizeof(T) bytes, read as 16-byte vectors.  LPR
// ("lanes per row", a power of two <= 32) adjacent lanes cover one row, so a
// warp reads 32/LPR sample rows at once and every global load is a fully
// coalesced 16*LPR-byte segment.  The ids of a sample are staged LPR at a
// time into registers (one coalesced 8*LPR-byte load) and broadcast with
// shuffles; four row loads are kept in flight per lane.
// HBM-bound: algorithmic bytes = n_rows*len*(8 + D*sizeof(T)) + n_rows*D*4.
#include "common.cuh"

namespace tt {

template <typename T, int LPR, int MODE>
__global__ void __launch_bounds__(256)
gather_pool_kernel(const T *__restrict__ table, int64_t vocab, int dim, const int64_t *__restrict__ ids,
                   int64_t n_rows, int len, int64_t pad, float *__restrict__ out, int64_t out_stride,
                   int32_t *__restrict__ argmax, int *__restrict__ oob_flag, int out_vec_ok) {
    constexpr int VN = Vec16<T>::N;
    constexpr int UNROLL = 4;
    const int vpr = dim / VN;  // 16-byte vectors per table row
    const int sub = threadIdx.x % LPR;
    const int grp_in_block = threadIdx.x / LPR;
    constexpr int GROUPS_PER_WARP = 32 / LPR;
    const int groups_per_block = blockDim.x / LPR;
    const int lane = threadIdx.x & 31;
    const int grp_in_warp = lane / LPR;
    const float inv_len = 1.0f / static_cast<float>(len);

    // warp-uniform loop so that full-mask shuffles stay legal at the tail
    const int64_t warp_first = (static_cast<int64_t>(blockIdx.x) * groups_per_block + grp_in_block) - grp_in_warp;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * groups_per_block;
    for (int64_t base = warp_first; base < n_rows; base += stride) {
        const int64_t row = base + grp_in_warp;
        const bool row_ok = row < n_rows;
        (void)GROUPS_PER_WARP;
        for (int c0 = 0; c0 < vpr; c0 += LPR) {
            const int c = c0 + sub;
            const bool col_ok = row_ok && c < vpr;
            float acc[VN];
            int amax[VN];
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                acc[e] = (MODE == TT_POOL_MAX) ? -INFINITY : 0.0f;
                amax[e] = 0;
            }
            int n_pad = 0, first_pad = -1;
            for (int l0 = 0; l0 < len; l0 += LPR) {
                int64_t my_id = -1;
                if (row_ok && l0 + sub < len) my_id = __ldg(ids + row * len + l0 + sub);
                const int cnt = min(LPR, len - l0);
                for (int j = 0; j < cnt; j += UNROLL) {
                    float v[UNROLL][VN];
                    bool use[UNROLL];
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) {
                        const int src = (j + u < LPR) ? (j + u) : 0;
                        const int64_t id = __shfl_sync(0xffffffffu, my_id, src, LPR);
                        const bool in = (j + u) < cnt;
                        bool valid = in && id >= 0 && id < vocab;
                        if (in && row_ok && !valid && sub == 0) atomicOr(oob_flag, 1);
                        if (valid && id == pad) {
                            if (n_pad == 0) first_pad = l0 + j + u;
                            ++n_pad;
                            valid = false;
                        }
                        use[u] = valid && col_ok;
                        if (use[u]) Vec16<T>::load(table + id * dim + c * VN, v[u]);
                    }
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) {
                        if (!use[u]) continue;
#pragma unroll
                        for (int e = 0; e < VN; ++e) {
                            if (MODE == TT_POOL_MAX) {
                                if (v[u][e] > acc[e]) { acc[e] = v[u][e]; amax[e] = l0 + j + u; }
                            } else {
                                acc[e] += v[u][e];
                            }
                        }
                    }
                }
            }
            if (n_pad > 0 && col_ok) {  // pads pool like any id: add table[pad] once
                float pv[VN];
                Vec16<T>::load(table + pad * dim + c * VN, pv);
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                    if (MODE == TT_POOL_MAX) {
                        // first occurrence wins on ties, like a left-to-right scan
                        if (pv[e] > acc[e] || (pv[e] == acc[e] && first_pad < amax[e])) {
                            acc[e] = pv[e]; amax[e] = first_pad;
                        }
                    } else {
                        acc[e] += static_cast<float>(n_pad) * pv[e];
                    }
                }
            }
            if (col_ok) {
                if (MODE == TT_POOL_MEAN) {
#pragma unroll
                    for (int e = 0; e < VN; ++e) acc[e] *= inv_len;
                }
                float *o = out + row * out_stride + c * VN;
                if (out_vec_ok) {
#pragma unroll
                    for (int e = 0; e < VN; e += 4)
                        *reinterpret_cast<float4 *>(o + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < VN; ++e) o[e] = acc[e];
                }
                if (MODE == TT_POOL_MAX && argmax != nullptr) {
#pragma unroll
                    for (int e = 0; e < VN; ++e) argmax[row * dim + c * VN + e] = amax[e];
                }
            }
        }
    }
}

// any D / unaligned tables: one thread per (row, d)
template <typename T, int MODE>
__global__ void gather_pool_scalar_kernel(const T *__restrict__ table, int64_t vocab, int dim,
                                          const int64_t *__restrict__ ids, int64_t n_rows, int len,
                                          float *__restrict__ out, int64_t out_stride,
                                          int32_t *__restrict__ argmax, int *__restrict__ oob_flag) {
    const int64_t total = n_rows * dim;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = t / dim;
        const int d = static_cast<int>(t % dim);
        float acc = (MODE == TT_POOL_MAX) ? -INFINITY : 0.0f;
        int am = 0;
        for (int l = 0; l < len; ++l) {
            const int64_t id = ids[row * len + l];
            if (id < 0 || id >= vocab) { atomicOr(oob_flag, 1); continue; }
            const float v = static_cast<float>(table[id * dim + d]);
            if (MODE == TT_POOL_MAX) { if (v > acc) { acc = v; am = l; } }
            else acc += v;
        }
        if (MODE == TT_POOL_MEAN) acc /= static_cast<float>(len);
        out[row * out_stride + d] = acc;
        if (MODE == TT_POOL_MAX && argmax != nullptr) argmax[row * dim + d] = am;
    }
}

template <typename T, int LPR>
static int launch_vec(const T *table, int64_t vocab, int dim, const int64_t *ids, int64_t n_rows, int len, int mode,
                      int64_t pad, float *out, int64_t out_stride, int32_t *argmax, int *oob, cudaStream_t st) {
    const int threads = 256;
    const int groups_per_block = threads / LPR;
    int64_t blocks = (n_rows + groups_per_block - 1) / groups_per_block;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;  // 8 resident CTAs/SM x 2 waves, grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const int vec_ok = (reinterpret_cast<uintptr_t>(out) % 16 == 0 && out_stride % 4 == 0) ? 1 : 0;
#define TT_GP_LAUNCH(M)                                                                                      \
    gather_pool_kernel<T, LPR, M><<<static_cast<unsigned>(blocks), threads, 0, st>>>(                        \
        table, vocab, dim, ids, n_rows, len, pad, out, out_stride, argmax, oob, vec_ok)
    switch (mode) {
        case TT_POOL_NONE:
        case TT_POOL_SUM: TT_GP_LAUNCH(TT_POOL_SUM); break;
        case TT_POOL_MEAN: TT_GP_LAUNCH(TT_POOL_MEAN); break;
        default: TT_GP_LAUNCH(TT_POOL_MAX); break;
    }
#undef TT_GP_LAUNCH
    TT_LAUNCH_CHECK("gather_pool_kernel");
    return 0;
}

template <typename T>
static int dispatch(const T *table, int64_t vocab, int dim, const int64_t *ids, int64_t n_rows, int len, int mode,
                    int64_t pad, float *out, int64_t out_stride, int32_t *argmax, int *oob, cudaStream_t st) {
    constexpr int VN = Vec16<T>::N;
    const bool vec = (dim % VN == 0) && (reinterpret_cast<uintptr_t>(table) % 16 == 0);
    if (!vec) {
        const int64_t total = n_rows * dim;
        int64_t blocks = (total + 255) / 256;
        const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
#define TT_GS_LAUNCH(M)                                                                                 \
    gather_pool_scalar_kernel<T, M><<<static_cast<unsigned>(blocks), 256, 0, st>>>(                      \
        table, vocab, dim, ids, n_rows, len, out, out_stride, argmax, oob)
        switch (mode) {
            case TT_POOL_NONE:
            case TT_POOL_SUM: TT_GS_LAUNCH(TT_POOL_SUM); break;
            case TT_POOL_MEAN: TT_GS_LAUNCH(TT_POOL_MEAN); break;
            default: TT_GS_LAUNCH(TT_POOL_MAX); break;
        }
#undef TT_GS_LAUNCH
        TT_LAUNCH_CHECK("gather_pool_scalar_kernel");
        return 0;
    }
    const int vpr = dim / VN;
    if (vpr <= 1) return launch_vec<T, 1>(table, vocab, dim, ids, n_rows, len, mode, pad, out, out_stride, argmax, oob, st);
    if (vpr <= 2) return launch_vec<T, 2>(table, vocab, dim, ids, n_rows, len, mode, pad, out, out_stride, argmax, oob, st);
    if (vpr <= 4) return launch_vec<T, 4>(table, vocab, dim, ids, n_rows, len, mode, pad, out, out_stride, argmax, oob, st);
    if (vpr <= 8) return launch_vec<T, 8>(table, vocab, dim, ids, n_rows, len, mode, pad, out, out_stride, argmax, oob, st);
    if (vpr <= 16) return launch_vec<T, 16>(table, vocab, dim, ids, n_rows, len, mode, pad, out, out_stride, argmax, oob, st);
    return launch_vec<T, 32>(table, vocab, dim, ids, n_rows, len, mode, pad, out, out_stride, argmax, oob, st);
}

}  // namespace tt

extern "C" int tt_emb_gather_pool_fwd(const void *table, int table_dtype, int64_t vocab, int dim, const int64_t *ids,
                                      int64_t n_rows, int len, int mode, int64_t padding_idx, float *out,
                                      int64_t out_stride, int32_t *argmax, int *oob_flag, void *stream) {
    TT_CHECK_ARG(table && ids && out && oob_flag, "null pointer");
    TT_CHECK_ARG(vocab > 0 && dim > 0 && n_rows >= 0 && len > 0, "non-positive size");
    TT_CHECK_ARG(mode >= TT_POOL_NONE && mode <= TT_POOL_MAX, "unknown pooling mode");
    TT_CHECK_ARG(mode != TT_POOL_NONE || len == 1, "TT_POOL_NONE needs len == 1");
    TT_CHECK_ARG(out_stride >= dim, "out_stride < dim");
    TT_CHECK_ARG(table_dtype == TT_F32 || table_dtype == TT_BF16, "unknown table dtype");
    if (n_rows == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (table_dtype == TT_F32)
        return tt::dispatch<float>(static_cast<const float *>(table), vocab, dim, ids, n_rows, len, mode,
                                   padding_idx, out, out_stride, argmax, oob_flag, st);
    return tt::dispatch<__nv_bfloat16>(static_cast<const __nv_bfloat16 *>(table), vocab, dim, ids, n_rows, len,
                                       mode, padding_idx, out, out_stride, argmax, oob_flag, st);
}
