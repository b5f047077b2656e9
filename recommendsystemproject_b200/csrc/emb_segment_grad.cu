// Sparse embedding gradient: deterministic sorted-segment scatter-add, plus
// the fused row-wise Adam (kernel 2 of the hot path).
//
// Replaces embedding_dense_backward (autograd of GenericTower.py:153-160,182,
// SequenceFeatureProcessor.py:60-68; training_utils.py:51) and, for the
// tables, clip_grad_norm_ + Adam.step (training_utils.py:53-56,
// train_twotower.py:111).  The dense [V, D] gradient is never formed.
//
// Pipeline (all on one stream, no host sync; U lives in device memory):
//   keys   : key[p] = id (pad -> V so it sorts last), val[p] = p
//   sort   : stable LSD radix sort over ceil(log2(V+1)) bits (cub::DeviceRadixSort)
//   heads  : head[i] = key[i] != key[i-1]; exclusive scan -> segment index
//   chunks : segments longer than CHUNK positions are cut into CHUNK-sized
//            pieces reduced by separate warps (heavy hitters under Zipf ids)
//   reduce : LPR lanes per segment, 16-byte loads of grad_out rows, fp32
//            accumulation in sorted (= ascending position) order -> bitwise
//            reproducible; one row_grad[U, D] write; per-segment sum of squares
//   norm   : fixed-order tree over the per-segment squares
//   adam   : one pass over the U touched rows of table / exp_avg / exp_avg_sq
// HBM-bound: algorithmic bytes = n_pos*(8 + D*4) + U*D*4 (segment grad) and
// U*D*(4 + 3*4 + 3*4) (row-wise Adam, fp32 state).
#include <cub/cub.cuh>

#include "common.cuh"

namespace tt {

constexpr int SEG_CHUNK_MAX = 128;

// Long segments are cut into `chunk` positions per lane group.  128 amortises the partial-row traffic on big
// inputs; small inputs (an ML-1M batch is 10^4 positions) get shorter chunks so that a 30-row vocabulary still
// spreads over the whole chip.  Depends on n only, so the summation order stays a function of the input.
static inline int seg_chunk_for(int64_t n) {
    int c = SEG_CHUNK_MAX;
    while (c > 8 && n / c < 4096) c >>= 1;
    return c;
}

__global__ void seg_build_keys(const int64_t *__restrict__ ids, int64_t n, int64_t pad, int64_t vocab,
                               uint32_t *__restrict__ keys, int32_t *__restrict__ vals) {
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n;
         p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t id = ids[p];
        const bool drop = (id == pad) || id < 0 || id >= vocab;
        keys[p] = drop ? static_cast<uint32_t>(vocab) : static_cast<uint32_t>(id);
        vals[p] = static_cast<int32_t>(p);
    }
}

__global__ void seg_heads(const uint32_t *__restrict__ keys, int64_t n, uint32_t sentinel,
                          int32_t *__restrict__ head) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t k = keys[i];
        head[i] = (k != sentinel && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
    }
}

// seg_start[s], unique_rows[s]; counters[0] = U, counters[1] = n_valid
__global__ void seg_starts(const uint32_t *__restrict__ keys, const int32_t *__restrict__ head,
                           const int32_t *__restrict__ seg_id, int64_t n, uint32_t sentinel,
                           int32_t *__restrict__ seg_start, int64_t *__restrict__ unique_rows,
                           int32_t *__restrict__ counters, const int32_t *__restrict__ vals = nullptr,
                           int32_t *__restrict__ seg_first = nullptr) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t k = keys[i];
        if (head[i]) {
            seg_start[seg_id[i]] = static_cast<int32_t>(i);
            unique_rows[seg_id[i]] = static_cast<int64_t>(k);
            // the segment's first sorted value, next to its start: seg_adam_rows_wide reaches the first gradient row
            // (most segments have one or two) without the dependent load through sorted_pos
            if (seg_first) seg_first[seg_id[i]] = vals[i];
        }
        if (k == sentinel && (i == 0 || keys[i - 1] != sentinel)) counters[1] = static_cast<int32_t>(i);
        if (i == n - 1) {
            counters[0] = seg_id[i] + head[i];
            if (k != sentinel) counters[1] = static_cast<int32_t>(n);
        }
    }
}

// number of CHUNK pieces of each long segment (0 for short ones); also closes seg_start[U] = n_valid
__global__ void seg_chunk_counts(int32_t *__restrict__ seg_start, const int32_t *__restrict__ counters,
                                 int64_t n, int SEG_CHUNK, int32_t *__restrict__ n_chunks) {
    const int32_t U = counters[0];
    for (int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < n;
         s += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int32_t c = 0;
        if (s < U) {
            const int32_t end = (s + 1 < U) ? seg_start[s + 1] : counters[1];
            const int32_t len = end - seg_start[s];
            if (len > SEG_CHUNK) c = (len + SEG_CHUNK - 1) / SEG_CHUNK;
        }
        n_chunks[s] = c;
    }
}

// ---- small problems (an ML-1M batch: 10^2..10^4 positions per table, vocabularies of 10..10^4 rows) ----
// keys -> sort -> heads -> scan -> starts -> chunk counts -> scan are nine launches of the general path; here ONE CTA
// does all of it in shared memory: (row << pos_bits | position) packed into one 32-bit key, cub::BlockRadixSort over
// the significant bits only, block scans for the segment numbers and the chunk bases.  Same outputs as the general
// path (sorted positions ascending inside a segment => same summation order), for n <= 1024 * ITEMS positions with
// bits(vocab) + bits(n) <= 32; used up to 8192 positions.
template <int ITEMS>
__global__ void __launch_bounds__(1024, 1)
seg_small_prep(const int64_t *__restrict__ ids, int n, int64_t pad, int64_t vocab, int pos_bits, int end_bit, int SEG_CHUNK,
               int32_t *__restrict__ sorted_pos, int32_t *__restrict__ seg_start, int64_t *__restrict__ unique_rows,
               int32_t *__restrict__ chunk_base, int32_t *__restrict__ counters) {
    using Sort = cub::BlockRadixSort<uint32_t, 1024, ITEMS>;
    using Scan = cub::BlockScan<int, 1024>;
    extern __shared__ unsigned char seg_small_smem[];
    auto &sort_tmp = *reinterpret_cast<typename Sort::TempStorage *>(seg_small_smem);
    auto &scan_tmp = *reinterpret_cast<typename Scan::TempStorage *>(seg_small_smem);
    __shared__ uint32_t last_key[1024];
    __shared__ int s_U, s_valid;
    const int t = threadIdx.x;
    const uint32_t SENT = 0xffffffffu;
    uint32_t key[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int p = t * ITEMS + k;
        key[k] = SENT;
        if (p < n) {
            const int64_t id = ids[p];
            if (!(id == pad || id < 0 || id >= vocab)) key[k] = (static_cast<uint32_t>(id) << pos_bits) | static_cast<uint32_t>(p);
        }
    }
    Sort(sort_tmp).Sort(key, 0, end_bit);     // dropped positions carry all-ones in the sorted bits: they end up last
    __syncthreads();
    last_key[t] = key[ITEMS - 1];
    __syncthreads();
    // heads and the number of valid positions
    int heads = 0, valid = 0;
    uint32_t prev = (t == 0) ? SENT : last_key[t - 1];
    uint32_t head_mask = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const bool ok = key[k] != SENT;
        const bool hd = ok && (prev == SENT || (prev >> pos_bits) != (key[k] >> pos_bits) || (t == 0 && k == 0));
        prev = key[k];
        head_mask |= (hd ? 1u : 0u) << k;
        valid += ok ? 1 : 0;
    }
    heads = __popc(head_mask);
    int seg_base, total_heads;
    Scan(scan_tmp).ExclusiveSum(heads, seg_base, total_heads);
    __syncthreads();
    int valid_base, total_valid;
    Scan(scan_tmp).ExclusiveSum(valid, valid_base, total_valid);
    __syncthreads();
    if (t == 0) { s_U = total_heads; s_valid = total_valid; counters[0] = total_heads; counters[1] = total_valid; }
    int seg = seg_base;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int idx = t * ITEMS + k;
        if (key[k] != SENT) {
            sorted_pos[idx] = static_cast<int32_t>(key[k] & ((1u << pos_bits) - 1u));
            if ((head_mask >> k) & 1u) {
                seg_start[seg] = idx;
                unique_rows[seg] = static_cast<int64_t>(key[k] >> pos_bits);
                ++seg;
            }
        }
    }
    __syncthreads();   // seg_start[] of this CTA is visible to this CTA (global writes + barrier)
    const int U = s_U, n_valid = s_valid;
    // chunk bases: exclusive sum over segments of ceil(len / CHUNK) for the long ones
    auto chunks_of = [&](int sg) {
        if (sg >= U) return 0;
        const int end = (sg + 1 < U) ? seg_start[sg + 1] : n_valid;
        const int len = end - seg_start[sg];
        return len > SEG_CHUNK ? (len + SEG_CHUNK - 1) / SEG_CHUNK : 0;
    };
    int mine = 0;
    for (int k = 0; k < ITEMS; ++k) mine += chunks_of(t * ITEMS + k);
    int cb;
    Scan(scan_tmp).ExclusiveSum(mine, cb);
    for (int k = 0; k < ITEMS; ++k) {
        const int sg = t * ITEMS + k;
        if (sg < U) chunk_base[sg] = cb;
        cb += chunks_of(sg);
    }
}

template <int ITEMS>
static int launch_seg_small(const int64_t *ids, int n, int64_t pad, int64_t vocab, int pos_bits, int end_bit, int chunk,
                            int32_t *sorted_pos, int32_t *seg_start, int64_t *unique_rows, int32_t *chunk_base,
                            int32_t *counters, cudaStream_t st) {
    using Sort = cub::BlockRadixSort<uint32_t, 1024, ITEMS>;
    using Scan = cub::BlockScan<int, 1024>;
    constexpr size_t smem = sizeof(typename Sort::TempStorage) > sizeof(typename Scan::TempStorage)
                                ? sizeof(typename Sort::TempStorage) : sizeof(typename Scan::TempStorage);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(seg_small_prep<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(seg_small_prep)");
        attr_set = true;
    }
    seg_small_prep<ITEMS><<<1, 1024, smem, st>>>(ids, n, pad, vocab, pos_bits, end_bit, chunk, sorted_pos, seg_start,
                                                  unique_rows, chunk_base, counters);
    TT_LAUNCH_CHECK("seg_small_prep");
    return 0;
}

struct GradSrc {
    const float *grad_out;
    int64_t grad_stride;
    const int32_t *argmax;
    int len;
    int mode;
    int dim;
    // list form (tt_emb_segment_grad_lists): position p reads the gradient row at float4 offset pos_src[p] from
    // grad_out; without pos_src, row p of a buffer cut into pieces of piece_rows rows, `piece_stride` floats apart (the
    // per-source blocks of the row-sharded exchange); piece_rows == 0: the flat [n, grad_stride] form above
    const int32_t *pos_src;
    int64_t piece_rows;
    int64_t piece_stride;
};

// gradient row (and pooling slot) of sorted position p
template <bool LISTS>
__device__ __forceinline__ const float *seg_grad_row(const GradSrc &g, int32_t p, int &slot, int64_t &src) {
    slot = 0;
    if (LISTS) {
        src = p;
        // pos_src given: the sort carried each position's float4 gradient offset as its VALUE (seg_build_keys_lists), so
        // the "sorted position" p already is that offset -- no second dependent load, no division
        if (g.pos_src) return g.grad_out + 4 * static_cast<int64_t>(p);
        const uint32_t q = static_cast<uint32_t>(p), pr = static_cast<uint32_t>(g.piece_rows);
        const uint32_t piece = q / pr;
        return g.grad_out + static_cast<int64_t>(piece) * g.piece_stride + static_cast<int64_t>(q - piece * pr) * g.grad_stride;
    }
    src = p;
    if (g.len > 1) { src = p / g.len; slot = p - static_cast<int32_t>(src) * g.len; }
    return g.grad_out + src * g.grad_stride;
}

// accumulate positions [p0, p1) of the sorted order into acc (LPR lanes per row, 4 floats per lane)
template <int LPR, bool LISTS = false>
__device__ __forceinline__ void seg_accumulate(const GradSrc &g, const int32_t *__restrict__ sorted_pos, int p0,
                                               int p1, int c, bool col_ok, float (&acc)[4]) {
    constexpr int UNROLL = 4;
    for (int i = p0; i < p1; i += UNROLL) {
        float4 v[UNROLL];
        int slot[UNROLL];
        bool use[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            use[u] = (i + u) < p1 && col_ok;
            slot[u] = 0;
            if (use[u]) {
                const int32_t p = __ldg(sorted_pos + i + u);
                int64_t src;
                const float *grow = seg_grad_row<LISTS>(g, p, slot[u], src);
                v[u] = __ldg(reinterpret_cast<const float4 *>(grow + c * 4));
                if (!LISTS && g.mode == TT_POOL_MAX) {
                    const int4 am = __ldg(reinterpret_cast<const int4 *>(g.argmax + src * g.dim + c * 4));
                    if (am.x != slot[u]) v[u].x = 0.f;
                    if (am.y != slot[u]) v[u].y = 0.f;
                    if (am.z != slot[u]) v[u].z = 0.f;
                    if (am.w != slot[u]) v[u].w = 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (use[u]) { acc[0] += v[u].x; acc[1] += v[u].y; acc[2] += v[u].z; acc[3] += v[u].w; }
        }
    }
}

// one LPR-lane group per CHUNK piece of a long segment -> partial[c, D]
template <int LPR, bool LISTS = false>
__global__ void __launch_bounds__(256)
seg_reduce_chunks(GradSrc g, const int32_t *__restrict__ sorted_pos, const int32_t *__restrict__ seg_start,
                  const int32_t *__restrict__ chunk_base, const int32_t *__restrict__ counters,
                  int32_t SEG_CHUNK, float *__restrict__ partial) {
    const int U = counters[0];
    if (U == 0) return;
    const int n_valid = counters[1];
    const int total = chunk_base[U - 1] + 0;  // exclusive scan value at U-1 ...
    const int vpr = g.dim / 4;
    const int sub = threadIdx.x % LPR;
    const int64_t group0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
    const int64_t n_groups = static_cast<int64_t>(gridDim.x) * blockDim.x / LPR;
    // total chunks = chunk_base[U-1] + chunks of the last segment
    const int last_end = n_valid;
    const int last_len = last_end - seg_start[U - 1];
    const int n_total = total + (last_len > SEG_CHUNK ? (last_len + SEG_CHUNK - 1) / SEG_CHUNK : 0);
    for (int64_t ck = group0; ck < n_total; ck += n_groups) {
        // upper_bound(chunk_base, ck) - 1 over segments [0, U)
        int lo = 0, hi = U;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (chunk_base[mid] <= ck) lo = mid; else hi = mid;
        }
        // skip zero-chunk segments that share the same base: lo is the LAST segment with base <= ck,
        // and only a long segment advances the base, so lo is the owner.
        const int s = lo;
        const int send = (s + 1 < U) ? seg_start[s + 1] : n_valid;
        const int p0 = seg_start[s] + static_cast<int>(ck - chunk_base[s]) * SEG_CHUNK;
        const int p1 = min(p0 + SEG_CHUNK, send);
        for (int c0 = 0; c0 < vpr; c0 += LPR) {
            const int c = c0 + sub;
            const bool col_ok = c < vpr;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            seg_accumulate<LPR, LISTS>(g, sorted_pos, p0, p1, c, col_ok, acc);
            if (col_ok)
                *reinterpret_cast<float4 *>(partial + ck * g.dim + c * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
    }
}

// one LPR-lane group per segment -> row_grad[s, D], sq[s]
template <int LPR, bool LISTS = false>
__global__ void __launch_bounds__(256)
seg_reduce_rows(GradSrc g, const int32_t *__restrict__ sorted_pos, const int32_t *__restrict__ seg_start,
                const int32_t *__restrict__ chunk_base, const int32_t *__restrict__ counters,
                const float *__restrict__ partial, float scale, int SEG_CHUNK, float *__restrict__ row_grad,
                float *__restrict__ seg_sq) {
    const int U = counters[0];
    const int n_valid = counters[1];
    const int vpr = g.dim / 4;
    const int sub = threadIdx.x % LPR;
    const int lane = threadIdx.x & 31;
    const int grp_in_warp = lane / LPR;
    const int groups_per_block = blockDim.x / LPR;
    const int64_t warp_first = static_cast<int64_t>(blockIdx.x) * groups_per_block + threadIdx.x / LPR - grp_in_warp;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * groups_per_block;
    for (int64_t base = warp_first; base < U; base += stride) {
        const int64_t s = base + grp_in_warp;
        const bool ok = s < U;
        float sq = 0.f;
        if (ok) {
            const int p0 = seg_start[s];
            const int p1 = (s + 1 < U) ? seg_start[s + 1] : n_valid;
            const int len = p1 - p0;
            for (int c0 = 0; c0 < vpr; c0 += LPR) {
                const int c = c0 + sub;
                const bool col_ok = c < vpr;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                if (len <= SEG_CHUNK) {
                    seg_accumulate<LPR, LISTS>(g, sorted_pos, p0, p1, c, col_ok, acc);
                } else if (col_ok) {
                    const int nck = (len + SEG_CHUNK - 1) / SEG_CHUNK;
                    const int64_t cb = chunk_base[s];
                    for (int k = 0; k < nck; ++k) {
                        const float4 v = *reinterpret_cast<const float4 *>(partial + (cb + k) * g.dim + c * 4);
                        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
                    }
                }
                if (col_ok) {
                    acc[0] *= scale; acc[1] *= scale; acc[2] *= scale; acc[3] *= scale;
                    *reinterpret_cast<float4 *>(row_grad + s * g.dim + c * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    sq += acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2] + acc[3] * acc[3];
                }
            }
        }
        // fixed-order reduction over the LPR lanes of the group
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (ok && sub == 0) seg_sq[s] = sq;
    }
}

// sum of one segment's gradient rows for the "wide" layout (8 lanes per segment, lane l owns float4 columns l, l+8, ...),
// positions in sorted (= ascending position) order; a long segment adds its chunk partials in chunk order.  ONE function
// for seg_reduce_rows_wide and seg_adam_rows_wide: the deferred (fused-Adam) form must add the same numbers in the same
// order as the form that stores row_grad.
template <int CPL, bool LISTS, bool FIRST = false>
__device__ __forceinline__ void seg_wide_accumulate(const GradSrc &g, const int32_t *__restrict__ sorted_pos, int p0, int p1,
                                                    int SEG_CHUNK, const int32_t *__restrict__ chunk_base,
                                                    const float *__restrict__ partial, int64_t s, int sub,
                                                    float (&acc)[CPL][4], int32_t first = 0) {   // FIRST: sorted_pos[p0], preloaded
    constexpr int LPR = 8;
    const int len = p1 - p0;
#pragma unroll
    for (int k = 0; k < CPL; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }
    if (len <= SEG_CHUNK) {
        constexpr int UN = (CPL <= 2) ? 2 : 1;   // positions in flight per lane (register budget: 64)
        for (int i = p0; i < p1; i += UN) {
            float4 v[UN][CPL];
            bool use[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                use[u] = (i + u) < p1;
                if (use[u]) {
                    const int32_t p = (FIRST && i + u == p0) ? first : __ldg(sorted_pos + i + u);
                    int64_t src;
                    int slot;
                    const float *grow = seg_grad_row<LISTS>(g, p, slot, src);
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const int c = sub + LPR * k;
                        v[u][k] = __ldg(reinterpret_cast<const float4 *>(grow + c * 4));
                        if (!LISTS && g.mode == TT_POOL_MAX) {
                            const int4 am = __ldg(reinterpret_cast<const int4 *>(g.argmax + src * g.dim + c * 4));
                            if (am.x != slot) v[u][k].x = 0.f;
                            if (am.y != slot) v[u][k].y = 0.f;
                            if (am.z != slot) v[u][k].z = 0.f;
                            if (am.w != slot) v[u][k].w = 0.f;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u)
                if (use[u]) {
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        acc[k][0] += v[u][k].x; acc[k][1] += v[u][k].y; acc[k][2] += v[u][k].z; acc[k][3] += v[u][k].w;
                    }
                }
        }
    } else {
        const int nck = (len + SEG_CHUNK - 1) / SEG_CHUNK;
        const int64_t cb = chunk_base[s];
        for (int c2 = 0; c2 < nck; ++c2) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const float4 v = *reinterpret_cast<const float4 *>(partial + (cb + c2) * g.dim + (sub + LPR * k) * 4);
                acc[k][0] += v.x; acc[k][1] += v.y; acc[k][2] += v.z; acc[k][3] += v.w;
            }
        }
    }
}

// D >= 64: 8 lanes per segment, CPL float4 columns per lane (lane l owns columns l, l+8, ...: every load instruction
// of a segment is one contiguous 128-byte line), 4 segments per warp and CPL * 2 independent 16-byte loads in flight
// per lane.  Most segments of a 10M-row table hold one or two positions, so the kernel lives on memory-level
// parallelism, not on the length of the inner loop.
// <= 64 registers at D <= 128: 4 blocks = 32 warps per SM (ncu: long_scoreboard-bound at the 2 blocks per SM that 86
// registers allowed)
// row_grad == NULL: the deferred form's first phase -- per-segment squares (the gradient norm) only; the update
// phase (seg_adam_rows_wide) forms the sums again instead of reading them back from HBM.
template <int CPL, bool LISTS = false>
__global__ void __launch_bounds__(256, (CPL <= 4) ? 4 : 2)
seg_reduce_rows_wide(GradSrc g, const int32_t *__restrict__ sorted_pos, const int32_t *__restrict__ seg_start,
                     const int32_t *__restrict__ chunk_base, const int32_t *__restrict__ counters,
                     const float *__restrict__ partial, float scale, int SEG_CHUNK, float *__restrict__ row_grad,
                     float *__restrict__ seg_sq) {
    constexpr int LPR = 8;
    const int U = counters[0];
    const int n_valid = counters[1];
    const int sub = threadIdx.x % LPR;
    const int64_t group0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
    const int64_t n_groups = static_cast<int64_t>(gridDim.x) * blockDim.x / LPR;
    const int64_t rounds = (U + n_groups - 1) / n_groups;
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t s = group0 + it * n_groups;
        const bool ok = s < U;
        float sq = 0.f;
        if (ok) {
            const int p0 = seg_start[s];
            const int p1 = (s + 1 < U) ? seg_start[s + 1] : n_valid;
            float acc[CPL][4];
            seg_wide_accumulate<CPL, LISTS>(g, sorted_pos, p0, p1, SEG_CHUNK, chunk_base, partial, s, sub, acc);
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                acc[k][0] *= scale; acc[k][1] *= scale; acc[k][2] *= scale; acc[k][3] *= scale;
                if (row_grad)
                    *reinterpret_cast<float4 *>(row_grad + s * g.dim + (sub + LPR * k) * 4) =
                        make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
                sq += acc[k][0] * acc[k][0] + acc[k][1] * acc[k][1] + acc[k][2] * acc[k][2] + acc[k][3] * acc[k][3];
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (ok && sub == 0) seg_sq[s] = sq;
    }
}

// Deferred form, second phase: segment sum (same order as above) -> Adam on the row, in one kernel.  The sum is formed
// from the upstream gradient rows again ([B, D] pooled gradients: L2 hits) instead of a [U, D] row_grad buffer that
// would be written once and read once through HBM (2 x U x D x 4 bytes of the 9 x U x D x 4 the two-kernel form moves).
// The row's parameter / moment loads are issued BEFORE the gather chain (they depend on the row number only).
template <int CPL, bool LISTS>
__global__ void __launch_bounds__(256, 2)
seg_adam_rows_wide(GradSrc g, const int32_t *__restrict__ sorted_pos, const int32_t *__restrict__ seg_start,
                   const int32_t *__restrict__ chunk_base, const int32_t *__restrict__ counters,
                   const float *__restrict__ partial, float scale, int SEG_CHUNK, const int64_t *__restrict__ rows,
                   const int32_t *__restrict__ seg_first, float *__restrict__ table, float *__restrict__ m,
                   float *__restrict__ v, const float *__restrict__ clip_coef, AdamHyper h,
                   const int64_t *__restrict__ step_dev) {
    constexpr int LPR = 8;
    const int U = counters[0];
    const int n_valid = counters[1];
    const float coef = clip_coef ? *clip_coef : 1.0f;
    float step_size, bc2_sqrt;
    adam_step_consts(h, step_dev, step_size, bc2_sqrt);
    const int sub = threadIdx.x % LPR;
    const int64_t group0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
    const int64_t n_groups = static_cast<int64_t>(gridDim.x) * blockDim.x / LPR;
    // (row, start, end, first sorted value) of a segment are four independent loads; the NEXT segment's are issued at
    // the top of an iteration, so that inside an iteration both dependent chains are one load deep: first gradient row
    // (<- first) and the row's state (<- row).  Measured without this: 3.19 ms against 2.87 ms for the plain Adam kernel
    // (every 8-lane group sat through seg_start -> sorted_pos -> gradient row before its stores).
    int64_t r_n = 0;
    int p0_n = 0, p1_n = 0;
    int32_t f_n = 0;
    if (group0 < U) {
        r_n = rows[group0]; p0_n = seg_start[group0]; f_n = seg_first[group0];
        p1_n = (group0 + 1 < U) ? seg_start[group0 + 1] : n_valid;
    }
    for (int64_t s = group0; s < U; s += n_groups) {
        const int64_t r = r_n;
        const int p0 = p0_n, p1 = p1_n;
        const int32_t first = f_n;
        const int64_t sn = s + n_groups;
        if (sn < U) {
            r_n = rows[sn]; p0_n = seg_start[sn]; f_n = seg_first[sn];
            p1_n = (sn + 1 < U) ? seg_start[sn + 1] : n_valid;
        }
        constexpr bool PREFETCH = CPL <= 4;      // D <= 128: 12 float4 of row state fit next to the accumulators (128 registers)
        float4 p4[PREFETCH ? CPL : 1], m4[PREFETCH ? CPL : 1], v4[PREFETCH ? CPL : 1];
        if (PREFETCH) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int64_t o = r * g.dim + (sub + LPR * k) * 4;
                // the row state is touched exactly once per step: streaming (evict-first) loads and stores, so that the
                // 15 GB that pass through do not push the [B, D] gradient rows (re-read by every position) out of L2
                // (ncu without the hints: 9.0 GB of DRAM reads for 7.4 GB of state)
                p4[k] = __ldcs(reinterpret_cast<const float4 *>(table + o));
                m4[k] = __ldcs(reinterpret_cast<const float4 *>(m + o));
                v4[k] = __ldcs(reinterpret_cast<const float4 *>(v + o));
            }
        }
        float acc[CPL][4];
        seg_wide_accumulate<CPL, LISTS, true>(g, sorted_pos, p0, p1, SEG_CHUNK, chunk_base, partial, s, sub, acc, first);
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            constexpr int KK = 0;
            const int kk = PREFETCH ? k : KK;
            if (!PREFETCH) {
                const int64_t o = r * g.dim + (sub + LPR * k) * 4;
                p4[0] = __ldcs(reinterpret_cast<const float4 *>(table + o));
                m4[0] = __ldcs(reinterpret_cast<const float4 *>(m + o));
                v4[0] = __ldcs(reinterpret_cast<const float4 *>(v + o));
            }
            // (acc * scale) is what the two-kernel form stores as row_grad, (row_grad * coef) what its Adam kernel applies
            const float gg[4] = {(acc[k][0] * scale) * coef, (acc[k][1] * scale) * coef, (acc[k][2] * scale) * coef,
                                 (acc[k][3] * scale) * coef};
            float pp[4] = {p4[kk].x, p4[kk].y, p4[kk].z, p4[kk].w};
            float mm[4] = {m4[kk].x, m4[kk].y, m4[kk].z, m4[kk].w};
            float vv[4] = {v4[kk].x, v4[kk].y, v4[kk].z, v4[kk].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) adam_elem(h, step_size, bc2_sqrt, gg[e], pp[e], mm[e], vv[e]);
            const int64_t o = r * g.dim + (sub + LPR * k) * 4;
            __stcs(reinterpret_cast<float4 *>(m + o), make_float4(mm[0], mm[1], mm[2], mm[3]));
            __stcs(reinterpret_cast<float4 *>(v + o), make_float4(vv[0], vv[1], vv[2], vv[3]));
            __stcs(reinterpret_cast<float4 *>(table + o), make_float4(pp[0], pp[1], pp[2], pp[3]));
        }
    }
}

// deterministic two-level sum of x[0..n) (n on the device): SUM_BLOCKS fixed slices, then one block over the partials
constexpr int SUM_BLOCKS = 592;
__global__ void __launch_bounds__(256)
seg_sum_partial(const float *__restrict__ x, const int32_t *__restrict__ n_dev, float *__restrict__ part) {
    __shared__ float sh[256];
    const int64_t n = *n_dev;
    const int64_t per = (n + SUM_BLOCKS - 1) / SUM_BLOCKS;
    const int64_t a = per * blockIdx.x, b = min(n, a + per);
    float acc = 0.f;
    for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) acc += x[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

// generic-D fallback (D % 4 != 0 or unaligned grad rows): one thread per (segment, d)
__global__ void seg_reduce_rows_scalar(GradSrc g, const int32_t *__restrict__ sorted_pos,
                                       const int32_t *__restrict__ seg_start, const int32_t *__restrict__ counters,
                                       float scale, float *__restrict__ row_grad) {
    const int U = counters[0];
    const int n_valid = counters[1];
    const int64_t total = static_cast<int64_t>(U) * g.dim;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int s = static_cast<int>(t / g.dim);
        const int d = static_cast<int>(t % g.dim);
        const int p1 = (s + 1 < U) ? seg_start[s + 1] : n_valid;
        float acc = 0.f;
        for (int i = seg_start[s]; i < p1; ++i) {
            const int32_t p = sorted_pos[i];
            int64_t src;
            int slot;
            const float *grow = g.piece_rows > 0 ? seg_grad_row<true>(g, p, slot, src) : seg_grad_row<false>(g, p, slot, src);
            float v = grow[d];
            if (g.piece_rows == 0 && g.mode == TT_POOL_MAX && g.argmax[src * g.dim + d] != slot) v = 0.f;
            acc += v;
        }
        row_grad[t] = acc * scale;
    }
}

__global__ void seg_sq_scalar(const float *__restrict__ row_grad, const int32_t *__restrict__ counters, int dim,
                              float *__restrict__ seg_sq) {
    const int U = counters[0];
    for (int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < U;
         s += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float sq = 0.f;
        for (int d = 0; d < dim; ++d) { const float v = row_grad[s * dim + d]; sq += v * v; }
        seg_sq[s] = sq;
    }
}

// *out += sum(x[0..n)) with n read from device; single block, fixed order
__global__ void __launch_bounds__(1024)
sum_fixed_order(const float *__restrict__ x, const int32_t *__restrict__ n_dev, int64_t n_host,
                float *__restrict__ out, int32_t *__restrict__ n_unique_out) {
    __shared__ float sh[1024];
    const int64_t n = n_dev ? static_cast<int64_t>(*n_dev) : n_host;
    float acc = 0.f;
    // strided ownership with a fixed block size -> the summation order depends only on n
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *out += sh[0];
        if (n_unique_out) *n_unique_out = static_cast<int32_t>(n);
    }
}

__global__ void seg_store_count(const int32_t *__restrict__ counters, int32_t *__restrict__ n_unique) {
    if (threadIdx.x == 0) *n_unique = counters[0];
}

static inline unsigned grid_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b);
}

struct SegPlan {
    size_t cub_sort, cub_scan, cub_bytes;
    int64_t max_chunks;
    int chunk;
};

static SegPlan seg_plan(int64_t n) {
    SegPlan p{};
    cub::DeviceRadixSort::SortPairs(nullptr, p.cub_sort, static_cast<uint32_t *>(nullptr),
                                    static_cast<uint32_t *>(nullptr), static_cast<int32_t *>(nullptr),
                                    static_cast<int32_t *>(nullptr), static_cast<int>(n), 0, 32);
    cub::DeviceScan::ExclusiveSum(nullptr, p.cub_scan, static_cast<int32_t *>(nullptr),
                                  static_cast<int32_t *>(nullptr), static_cast<int>(n));
    p.cub_bytes = p.cub_sort > p.cub_scan ? p.cub_sort : p.cub_scan;
    p.chunk = seg_chunk_for(n);
    p.max_chunks = 2 * (n / p.chunk) + 2;  // each long segment wastes at most one partial chunk
    return p;
}

template <int LPR, bool LISTS = false>
static int launch_reduce(const GradSrc &g, const int32_t *sorted_pos, const int32_t *seg_start,
                         const int32_t *chunk_base, const int32_t *counters, float *partial, int64_t max_chunks,
                         float scale, float *row_grad, float *seg_sq, int64_t n, int SEG_CHUNK, cudaStream_t st) {
    const int threads = 256;
    const int gpb = threads / LPR;
    if (n > SEG_CHUNK) {
        seg_reduce_chunks<LPR, LISTS><<<grid_for(max_chunks * LPR, threads), threads, 0, st>>>(
            g, sorted_pos, seg_start, chunk_base, counters, SEG_CHUNK, partial);
        TT_LAUNCH_CHECK("seg_reduce_chunks");
    }
    int64_t blocks = (n + gpb - 1) / gpb;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    seg_reduce_rows<LPR, LISTS><<<static_cast<unsigned>(blocks), threads, 0, st>>>(g, sorted_pos, seg_start, chunk_base,
                                                                           counters, partial, scale, SEG_CHUNK, row_grad, seg_sq);
    TT_LAUNCH_CHECK("seg_reduce_rows");
    return 0;
}

}  // namespace tt

extern "C" int tt_emb_segment_grad_workspace(int64_t n_pos, int dim, size_t *bytes_host) {
    TT_CHECK_ARG(bytes_host && n_pos > 0 && dim > 0, "bad size");
    TT_CHECK_ARG(n_pos < (int64_t(1) << 31) - 2, "more than 2^31 positions per call");
    const tt::SegPlan p = tt::seg_plan(n_pos);
    size_t b = 0;
    auto add = [&](size_t x) { b = tt::align_up(b, 256) + x; };
    add(sizeof(uint32_t) * n_pos);      // keys in
    add(sizeof(uint32_t) * n_pos);      // keys out
    add(sizeof(int32_t) * n_pos);       // vals in
    add(sizeof(int32_t) * n_pos);       // vals out (sorted positions)
    add(sizeof(int32_t) * n_pos);       // head flags / chunk counts
    add(sizeof(int32_t) * n_pos);       // seg ids / chunk bases
    add(sizeof(int32_t) * (n_pos + 1)); // seg starts
    add(sizeof(float) * n_pos);         // per-segment squares
    add(sizeof(int32_t) * 4);           // counters
    add(sizeof(float) * p.max_chunks * dim);
    add(p.cub_bytes);
    add(sizeof(int32_t) * n_pos);       // first sorted value of every segment
    *bytes_host = tt::align_up(b, 256) + 256;
    return 0;
}

namespace tt {

// keys of the list form: key[p] = rows[(p / piece_len) * piece_stride + p % piece_len] (negative / >= vocab: dropped)
// value carried through the (stable) sort: the position p, or -- when pos_src is given -- the float4 offset of p's
// gradient row (what the reduce kernels would otherwise fetch through a second dependent load per position)
__global__ void seg_build_keys_lists(const int32_t *__restrict__ rows, int64_t n, int64_t piece_len, int64_t piece_stride,
                                     int64_t vocab, const int32_t *__restrict__ pos_src, uint32_t *__restrict__ keys,
                                     int32_t *__restrict__ vals) {
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n;
         p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t piece = p / piece_len;
        const int32_t r = __ldg(rows + piece * piece_stride + (p - piece * piece_len));
        const bool drop = r < 0 || r >= vocab;
        keys[p] = drop ? static_cast<uint32_t>(vocab) : static_cast<uint32_t>(r);
        vals[p] = (pos_src && !drop) ? __ldg(pos_src + p) : static_cast<int32_t>(p);
    }
}

struct KeySrc {
    const int64_t *ids;       // flat form: ids[n], pad dropped
    int64_t pad;
    const int32_t *rows;      // list form (ids == nullptr)
    int64_t piece_len, piece_stride;
    const int32_t *pos_src;   // list form: per-position float4 gradient offsets (become the sort values)
};

// the workspace of one segment-gradient call (same carve-up in tt_emb_segment_grad_workspace)
struct SegWs {
    uint32_t *keys_in, *keys_out;
    int32_t *vals_in, *vals_out, *head, *seg_id, *seg_start;
    float *seg_sq;
    int32_t *counters;
    float *partial;
    void *cub_tmp;
    int32_t *seg_first;     // first sorted value of every segment (general path)
    bool ok;
    size_t used;
};
static SegWs seg_carve(void *workspace, size_t workspace_bytes, int64_t n, int dim, const SegPlan &plan) {
    Workspace ws(workspace, workspace_bytes);
    SegWs w;
    w.keys_in = ws.take<uint32_t>(n);
    w.keys_out = ws.take<uint32_t>(n);
    w.vals_in = ws.take<int32_t>(n);
    w.vals_out = ws.take<int32_t>(n);
    w.head = ws.take<int32_t>(n);
    w.seg_id = ws.take<int32_t>(n);
    w.seg_start = ws.take<int32_t>(n + 1);
    w.seg_sq = ws.take<float>(n);
    w.counters = ws.take<int32_t>(4);
    w.partial = ws.take<float>(plan.max_chunks * dim);
    w.cub_tmp = ws.take<char>(plan.cub_bytes);
    w.seg_first = ws.take<int32_t>(n);
    w.ok = ws.ok();
    w.used = ws.off;
    return w;
}

// dims the 8-lane "wide" kernels cover: D / 32 float4 columns per lane
static inline int seg_wide_cpl(int dim) {
    const int vpr = dim / 4;
    const int cpl = (dim % 4 == 0 && vpr % 8 == 0) ? vpr / 8 : 0;
    return (cpl == 2 || cpl == 3 || cpl == 4 || cpl == 6 || cpl == 8) ? cpl : 0;
}

template <bool LISTS>
static int segment_grad_run(const KeySrc &ks, const GradSrc &g, int64_t n, int64_t vocab, float scale,
                            int64_t *unique_rows, float *row_grad, int32_t *n_unique, float *sq_norm, void *workspace,
                            size_t workspace_bytes, cudaStream_t st) {
    const int dim = g.dim;
    const SegPlan plan = seg_plan(n);

    const SegWs sw = seg_carve(workspace, workspace_bytes, n, dim, plan);
    if (!sw.ok) { set_error("segment_grad workspace too small: need %zu have %zu", sw.used, workspace_bytes); return TT_E_WORKSPACE; }
    uint32_t *keys_in = sw.keys_in, *keys_out = sw.keys_out;
    int32_t *vals_in = sw.vals_in, *vals_out = sw.vals_out, *head = sw.head, *seg_id = sw.seg_id, *seg_start = sw.seg_start;
    float *seg_sq = sw.seg_sq;
    int32_t *counters = sw.counters;
    float *partial = sw.partial;
    void *cub_tmp = sw.cub_tmp;

    const int threads = 256;
    const unsigned g1 = grid_for(n, threads);
    int32_t *n_chunks = head;       // the general path reuses head -> chunk counts, seg_id -> chunk bases
    int32_t *chunk_base = seg_id;
    // small problem: one CTA prepares everything (see seg_small_prep)
    int pos_bits = 1, id_bits = 1;
    while ((int64_t(1) << pos_bits) < n) ++pos_bits;
    while ((int64_t(1) << id_bits) < vocab) ++id_bits;
    // (measured on the C2 step: 17 us for <= 4096 positions against ~35 us of launches; at 16 / 32 keys per thread the
    // single CTA is no faster than the general path -- 41 / 90 us -- so those sizes stay there)
    const bool small = !LISTS && n <= 8192 && pos_bits + id_bits <= 32;
    if (small) {
        const int nn = static_cast<int>(n);
        int rc2;
        if (nn <= 4096) rc2 = launch_seg_small<4>(ks.ids, nn, ks.pad, vocab, pos_bits, pos_bits + id_bits, plan.chunk, vals_out, seg_start, unique_rows, chunk_base, counters, st);
        else rc2 = launch_seg_small<8>(ks.ids, nn, ks.pad, vocab, pos_bits, pos_bits + id_bits, plan.chunk, vals_out, seg_start, unique_rows, chunk_base, counters, st);
        if (rc2) return rc2;
    } else {
        if (LISTS) seg_build_keys_lists<<<g1, threads, 0, st>>>(ks.rows, n, ks.piece_len, ks.piece_stride, vocab, ks.pos_src, keys_in, vals_in);
        else seg_build_keys<<<g1, threads, 0, st>>>(ks.ids, n, ks.pad, vocab, keys_in, vals_in);
        TT_LAUNCH_CHECK("seg_build_keys");
        int end_bit = 1;
        while ((int64_t(1) << end_bit) <= vocab) ++end_bit;  // keys are in [0, vocab]
        size_t tmp = plan.cub_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_tmp, tmp, keys_in, keys_out, vals_in, vals_out,
                                                        static_cast<int>(n), 0, end_bit, st);
        if (e != cudaSuccess) return cuda_status(e, "cub SortPairs");
        const uint32_t sentinel = static_cast<uint32_t>(vocab);
        seg_heads<<<g1, threads, 0, st>>>(keys_out, n, sentinel, head);
        TT_LAUNCH_CHECK("seg_heads");
        tmp = plan.cub_bytes;
        e = cub::DeviceScan::ExclusiveSum(cub_tmp, tmp, head, seg_id, static_cast<int>(n), st);
        if (e != cudaSuccess) return cuda_status(e, "cub ExclusiveSum");
        seg_starts<<<g1, threads, 0, st>>>(keys_out, head, seg_id, n, sentinel, seg_start, unique_rows, counters, vals_out,
                                           sw.seg_first);
        TT_LAUNCH_CHECK("seg_starts");
        seg_chunk_counts<<<g1, threads, 0, st>>>(seg_start, counters, n, plan.chunk, n_chunks);
        TT_LAUNCH_CHECK("seg_chunk_counts");
        tmp = plan.cub_bytes;
        e = cub::DeviceScan::ExclusiveSum(cub_tmp, tmp, n_chunks, chunk_base, static_cast<int>(n), st);
        if (e != cudaSuccess) return cuda_status(e, "cub ExclusiveSum(chunks)");
    }   // general path

    const bool vec = (dim % 4 == 0) && (g.grad_stride % 4 == 0) && (reinterpret_cast<uintptr_t>(g.grad_out) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(row_grad) % 16 == 0) && (g.piece_stride % 4 == 0);
    int rc = 0;
    if (row_grad == nullptr && !(vec && seg_wide_cpl(dim) != 0)) {
        set_error("segment_grad without row_grad (deferred form) needs dim in {64, 96, 128, 192, 256} and 16-byte aligned gradient rows");
        return TT_E_UNSUPPORTED;
    }
    if (!vec) {
        seg_reduce_rows_scalar<<<grid_for(n * dim, threads), threads, 0, st>>>(g, vals_out, seg_start, counters, scale,
                                                                                row_grad);
        TT_LAUNCH_CHECK("seg_reduce_rows_scalar");
        seg_sq_scalar<<<g1, threads, 0, st>>>(row_grad, counters, dim, seg_sq);
        TT_LAUNCH_CHECK("seg_sq_scalar");
    } else {
        const int vpr = dim / 4;
        const int cpl = seg_wide_cpl(dim);
        if (cpl != 0) {
            if (n > plan.chunk) {
                seg_reduce_chunks<8, LISTS><<<grid_for(plan.max_chunks * 8, threads), threads, 0, st>>>(
                    g, vals_out, seg_start, chunk_base, counters, plan.chunk, partial);
                TT_LAUNCH_CHECK("seg_reduce_chunks");
            }
            int64_t blocks = (n + 31) / 32;
            const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
            if (blocks > cap) blocks = cap;
#define TT_WIDE(C) seg_reduce_rows_wide<C, LISTS><<<static_cast<unsigned>(blocks), threads, 0, st>>>(               \
        g, vals_out, seg_start, chunk_base, counters, partial, scale, plan.chunk, row_grad, seg_sq)
            if (cpl == 2) TT_WIDE(2); else if (cpl == 3) TT_WIDE(3); else if (cpl == 4) TT_WIDE(4);
            else if (cpl == 6) TT_WIDE(6); else TT_WIDE(8);
#undef TT_WIDE
            TT_LAUNCH_CHECK("seg_reduce_rows_wide");
        }
        else if (vpr <= 1) rc = launch_reduce<1, LISTS>(g, vals_out, seg_start, chunk_base, counters, partial, plan.max_chunks, scale, row_grad, seg_sq, n, plan.chunk, st);
        else if (vpr <= 2) rc = launch_reduce<2, LISTS>(g, vals_out, seg_start, chunk_base, counters, partial, plan.max_chunks, scale, row_grad, seg_sq, n, plan.chunk, st);
        else if (vpr <= 4) rc = launch_reduce<4, LISTS>(g, vals_out, seg_start, chunk_base, counters, partial, plan.max_chunks, scale, row_grad, seg_sq, n, plan.chunk, st);
        else if (vpr <= 8) rc = launch_reduce<8, LISTS>(g, vals_out, seg_start, chunk_base, counters, partial, plan.max_chunks, scale, row_grad, seg_sq, n, plan.chunk, st);
        else if (vpr <= 16) rc = launch_reduce<16, LISTS>(g, vals_out, seg_start, chunk_base, counters, partial, plan.max_chunks, scale, row_grad, seg_sq, n, plan.chunk, st);
        else rc = launch_reduce<32, LISTS>(g, vals_out, seg_start, chunk_base, counters, partial, plan.max_chunks, scale, row_grad, seg_sq, n, plan.chunk, st);
        if (rc) return rc;
    }
    // *sq_norm += sum(seg_sq[0..U)), *n_unique = U
    float *sq_target = sq_norm ? sq_norm : seg_sq + (n - 1);  // dummy target when the caller does not want the norm
    if (n > 65536) {
        float *sum_part = reinterpret_cast<float *>(keys_in);   // the sort input is dead by now
        seg_sum_partial<<<SUM_BLOCKS, 256, 0, st>>>(seg_sq, counters, sum_part);
        TT_LAUNCH_CHECK("seg_sum_partial");
        sum_fixed_order<<<1, 1024, 0, st>>>(sum_part, nullptr, SUM_BLOCKS, sq_target, nullptr);
        seg_store_count<<<1, 32, 0, st>>>(counters, n_unique);
    } else {
        sum_fixed_order<<<1, 1024, 0, st>>>(seg_sq, counters, 0, sq_target, n_unique);
    }
    TT_LAUNCH_CHECK("sum_fixed_order");
    return 0;
}

}  // namespace tt

extern "C" int tt_emb_segment_grad(const int64_t *ids, int64_t n_rows, int len, int mode, int64_t padding_idx,
                                   int64_t vocab, const float *grad_out, int64_t grad_stride, const int32_t *argmax,
                                   int dim, int64_t *unique_rows, float *row_grad, int32_t *n_unique,
                                   float *sq_norm, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(ids && grad_out && unique_rows && row_grad && n_unique && workspace, "null pointer");
    TT_CHECK_ARG(n_rows > 0 && len > 0 && dim > 0 && vocab > 0, "non-positive size");
    TT_CHECK_ARG(mode >= TT_POOL_NONE && mode <= TT_POOL_MAX, "unknown pooling mode");
    TT_CHECK_ARG(mode != TT_POOL_MAX || argmax, "TT_POOL_MAX needs argmax");
    TT_CHECK_ARG(mode != TT_POOL_NONE || len == 1, "TT_POOL_NONE needs len == 1");
    TT_CHECK_ARG(vocab < (int64_t(1) << 31) - 1, "vocab must fit 31 bits");
    const int64_t n = n_rows * len;
    TT_CHECK_ARG(n < (int64_t(1) << 31) - 2, "more than 2^31 positions per call");
    const KeySrc ks{ids, padding_idx, nullptr, 0, 0, nullptr};
    const GradSrc g{grad_out, grad_stride, argmax, len, mode, dim, nullptr, 0, 0};
    const float scale = (mode == TT_POOL_MEAN) ? 1.0f / static_cast<float>(len) : 1.0f;
    return segment_grad_run<false>(ks, g, n, vocab, scale, unique_rows, row_grad, n_unique, sq_norm, workspace,
                                   workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int tt_emb_segment_grad_lists(const int32_t *rows, int64_t n_pieces, int64_t piece_len, int64_t piece_stride,
                                         const int32_t *pos_src, int64_t vocab, const float *grad, int64_t grad_piece_rows,
                                         int64_t grad_piece_stride, int dim, int64_t *unique_rows, float *row_grad,
                                         int32_t *n_unique, float *sq_norm, void *workspace, size_t workspace_bytes,
                                         void *stream) {
    using namespace tt;
    TT_CHECK_ARG(rows && grad && unique_rows && n_unique && workspace, "null pointer");   // row_grad NULL = deferred form
    TT_CHECK_ARG(n_pieces > 0 && piece_len > 0 && piece_stride >= piece_len && dim > 0 && vocab > 0, "bad size");
    TT_CHECK_ARG(grad_piece_rows > 0 && grad_piece_stride >= grad_piece_rows * dim, "bad gradient piece layout");
    TT_CHECK_ARG(vocab < (int64_t(1) << 31) - 1, "vocab must fit 31 bits");
    const int64_t n = n_pieces * piece_len;
    TT_CHECK_ARG(n < (int64_t(1) << 31) - 2, "more than 2^31 positions per call");
    const KeySrc ks{nullptr, -1, rows, piece_len, piece_stride, pos_src};
    const GradSrc g{grad, dim, nullptr, 1, TT_POOL_SUM, dim, pos_src, grad_piece_rows, grad_piece_stride};
    return segment_grad_run<true>(ks, g, n, vocab, 1.0f, unique_rows, row_grad, n_unique, sq_norm, workspace,
                                  workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int tt_emb_segment_adam_lists(int64_t n_positions, const int32_t *pos_src, const float *grad, int64_t grad_piece_rows,
                                         int64_t grad_piece_stride, int dim, const int64_t *unique_rows, int64_t max_rows,
                                         void *workspace, size_t workspace_bytes, float *table, float *exp_avg,
                                         float *exp_avg_sq, const float *clip_coef, double lr, double beta1, double beta2,
                                         double eps, const int64_t *step_dev, const double *lr_dev, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(grad && unique_rows && workspace && table && exp_avg && exp_avg_sq && step_dev, "null pointer");
    TT_CHECK_ARG(n_positions > 0 && n_positions < (int64_t(1) << 31) - 2 && dim > 0, "bad size");
    TT_CHECK_ARG(grad_piece_rows > 0 && grad_piece_stride >= grad_piece_rows * dim, "bad gradient piece layout");
    const int cpl = seg_wide_cpl(dim);
    if (cpl == 0 || reinterpret_cast<uintptr_t>(grad) % 16 != 0 || grad_piece_stride % 4 != 0) {
        set_error("tt_emb_segment_adam_lists needs dim in {64, 96, 128, 192, 256} and 16-byte aligned gradient rows");
        return TT_E_UNSUPPORTED;
    }
    if (max_rows <= 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const SegPlan plan = seg_plan(n_positions);
    const SegWs sw = seg_carve(workspace, workspace_bytes, n_positions, dim, plan);
    if (!sw.ok) { set_error("segment_grad workspace too small: need %zu have %zu", sw.used, workspace_bytes); return TT_E_WORKSPACE; }
    const GradSrc g{grad, dim, nullptr, 1, TT_POOL_SUM, dim, pos_src, grad_piece_rows, grad_piece_stride};
    const AdamHyper hyper = make_adam(lr, beta1, beta2, eps, lr_dev);
    int64_t blocks = (max_rows + 31) / 32;                      // 32 segments per 256-thread block
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    // seg_id holds the chunk bases after the first phase (segment_grad_run: chunk_base = seg_id)
#define TT_WIDE_ADAM(C) seg_adam_rows_wide<C, true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(                            \
        g, sw.vals_out, sw.seg_start, sw.seg_id, sw.counters, sw.partial, 1.0f, plan.chunk, unique_rows, sw.seg_first,     \
        table, exp_avg, exp_avg_sq, clip_coef, hyper, step_dev)
    if (cpl == 2) TT_WIDE_ADAM(2); else if (cpl == 3) TT_WIDE_ADAM(3); else if (cpl == 4) TT_WIDE_ADAM(4);
    else if (cpl == 6) TT_WIDE_ADAM(6); else TT_WIDE_ADAM(8);
#undef TT_WIDE_ADAM
    TT_LAUNCH_CHECK("seg_adam_rows_wide");
    return 0;
}

namespace tt {

template <typename T, int LPR>
__global__ void __launch_bounds__(256)
rowwise_adam_kernel(T *__restrict__ table, float *__restrict__ m, float *__restrict__ v, int dim,
                    const int64_t *__restrict__ rows, const float *__restrict__ row_grad,
                    const int32_t *__restrict__ n_unique, const float *__restrict__ clip_coef, AdamHyper h,
                    const int64_t *__restrict__ step_dev) {
    const int U = *n_unique;
    const float coef = clip_coef ? *clip_coef : 1.0f;
    float step_size, bc2_sqrt;
    adam_step_consts(h, step_dev, step_size, bc2_sqrt);
    const int vpr = dim / 4;
    const int sub = threadIdx.x % LPR;
    const int64_t g0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
    const int64_t ng = static_cast<int64_t>(gridDim.x) * blockDim.x / LPR;
    for (int64_t s = g0; s < U; s += ng) {
        const int64_t r = rows[s];
        for (int c = sub; c < vpr; c += LPR) {
            const float4 g4 = *reinterpret_cast<const float4 *>(row_grad + s * dim + c * 4);
            float4 m4 = *reinterpret_cast<float4 *>(m + r * dim + c * 4);
            float4 v4 = *reinterpret_cast<float4 *>(v + r * dim + c * 4);
            float p[4];
            if (sizeof(T) == 4) {
                const float4 p4 = *reinterpret_cast<const float4 *>(reinterpret_cast<float *>(table) + r * dim + c * 4);
                p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w;
            } else {
                const uint2 raw = *reinterpret_cast<const uint2 *>(reinterpret_cast<__nv_bfloat16 *>(table) + r * dim + c * 4);
                p[0] = __uint_as_float(raw.x << 16); p[1] = __uint_as_float(raw.x & 0xffff0000u);
                p[2] = __uint_as_float(raw.y << 16); p[3] = __uint_as_float(raw.y & 0xffff0000u);
            }
            const float gg[4] = {g4.x * coef, g4.y * coef, g4.z * coef, g4.w * coef};
            float mm[4] = {m4.x, m4.y, m4.z, m4.w};
            float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) adam_elem(h, step_size, bc2_sqrt, gg[e], p[e], mm[e], vv[e]);
            *reinterpret_cast<float4 *>(m + r * dim + c * 4) = make_float4(mm[0], mm[1], mm[2], mm[3]);
            *reinterpret_cast<float4 *>(v + r * dim + c * 4) = make_float4(vv[0], vv[1], vv[2], vv[3]);
            if (sizeof(T) == 4) {
                *reinterpret_cast<float4 *>(reinterpret_cast<float *>(table) + r * dim + c * 4) = make_float4(p[0], p[1], p[2], p[3]);
            } else {
                __nv_bfloat162 a = __floats2bfloat162_rn(p[0], p[1]);
                __nv_bfloat162 b = __floats2bfloat162_rn(p[2], p[3]);
                uint2 raw;
                raw.x = *reinterpret_cast<uint32_t *>(&a);
                raw.y = *reinterpret_cast<uint32_t *>(&b);
                *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(table) + r * dim + c * 4) = raw;
            }
        }
    }
}

template <int LPR>
__global__ void __launch_bounds__(256)
scatter_rows_kernel(float *__restrict__ dense, int dim, const int64_t *__restrict__ rows,
                    const float *__restrict__ row_grad, const int32_t *__restrict__ n_unique) {
    const int U = *n_unique;
    const int sub = threadIdx.x % LPR;
    const int64_t g0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
    const int64_t ng = static_cast<int64_t>(gridDim.x) * blockDim.x / LPR;
    for (int64_t s = g0; s < U; s += ng) {
        const int64_t r = rows[s];
        for (int d = sub; d < dim; d += LPR) dense[r * dim + d] += row_grad[s * dim + d];  // rows are unique: no race
    }
}

}  // namespace tt

extern "C" int tt_emb_rowwise_adam(void *table, int table_dtype, float *exp_avg, float *exp_avg_sq, int dim,
                                   const int64_t *unique_rows, const float *row_grad, const int32_t *n_unique,
                                   int64_t max_rows, const float *clip_coef, double lr, double beta1, double beta2,
                                   double eps, const int64_t *step_dev, const double *lr_dev, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(table && exp_avg && exp_avg_sq && unique_rows && row_grad && n_unique && step_dev, "null pointer");
    TT_CHECK_ARG(dim > 0 && dim % 4 == 0, "row-wise Adam needs dim % 4 == 0");
    TT_CHECK_ARG(table_dtype == TT_F32 || table_dtype == TT_BF16, "unknown table dtype");
    if (max_rows <= 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int vpr = dim / 4;
    const int lpr = vpr >= 32 ? 32 : (vpr >= 16 ? 16 : (vpr >= 8 ? 8 : (vpr >= 4 ? 4 : (vpr >= 2 ? 2 : 1))));
    const unsigned grid = grid_for(max_rows * lpr, 256);
    const AdamHyper hyper = make_adam(lr, beta1, beta2, eps, lr_dev);
#define TT_ADAM(T, L)                                                                                              \
    rowwise_adam_kernel<T, L><<<grid, 256, 0, st>>>(static_cast<T *>(table), exp_avg, exp_avg_sq, dim, unique_rows, \
                                                   row_grad, n_unique, clip_coef, hyper, step_dev)
#define TT_ADAM_T(T)                                   \
    switch (lpr) {                                     \
        case 1: TT_ADAM(T, 1); break;                  \
        case 2: TT_ADAM(T, 2); break;                  \
        case 4: TT_ADAM(T, 4); break;                  \
        case 8: TT_ADAM(T, 8); break;                  \
        case 16: TT_ADAM(T, 16); break;                \
        default: TT_ADAM(T, 32); break;                \
    }
    if (table_dtype == TT_F32) { TT_ADAM_T(float) } else { TT_ADAM_T(__nv_bfloat16) }
#undef TT_ADAM_T
#undef TT_ADAM
    TT_LAUNCH_CHECK("rowwise_adam_kernel");
    return 0;
}

extern "C" int tt_emb_scatter_rows(float *dense, int dim, const int64_t *unique_rows, const float *row_grad,
                                   const int32_t *n_unique, int64_t max_rows, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(dense && unique_rows && row_grad && n_unique && dim > 0, "bad argument");
    if (max_rows <= 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int lpr = dim >= 32 ? 32 : (dim >= 16 ? 16 : (dim >= 8 ? 8 : 4));
    const unsigned grid = grid_for(max_rows * lpr, 256);
    switch (lpr) {
        case 4: scatter_rows_kernel<4><<<grid, 256, 0, st>>>(dense, dim, unique_rows, row_grad, n_unique); break;
        case 8: scatter_rows_kernel<8><<<grid, 256, 0, st>>>(dense, dim, unique_rows, row_grad, n_unique); break;
        case 16: scatter_rows_kernel<16><<<grid, 256, 0, st>>>(dense, dim, unique_rows, row_grad, n_unique); break;
        default: scatter_rows_kernel<32><<<grid, 256, 0, st>>>(dense, dim, unique_rows, row_grad, n_unique); break;
    }
    TT_LAUNCH_CHECK("scatter_rows_kernel");
    return 0;
}
