// Helpers shared by the fp32 (topk_f32.cu) and tcgen05 (topk_tc.cu) top-K paths.
#pragma once

#include "common.cuh"

namespace tt {

constexpr int TK_STAGE2_MAX = 2048;

template <typename V, typename I>
__device__ __forceinline__ bool tk_before(V va, I ia, V vb, I ib) {
    return va > vb || (va == vb && ia < ib);
}

// bitonic sort of n (power of two) pairs so that "before" elements come first; `nthreads` cooperating
// threads with id `tid`; SYNC() separates stages.
template <typename V, typename I, typename SyncFn>
__device__ __forceinline__ void tk_bitonic(V *v, I *ix, int n, int tid, int nthreads, SyncFn sync) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < n / 2; t += nthreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // j is a power of two: 2*j*(t/j) + t%j
                const int p = i + j;
                const bool asc = (i & k) == 0;
                const V va = v[i], vb = v[p];
                const I ia = ix[i], ib = ix[p];
                const bool swap = asc ? tk_before(vb, ib, va, ia) : tk_before(va, ia, vb, ib);
                // unconditional stores: no divergent branch around them
                v[i] = swap ? vb : va;
                v[p] = swap ? va : vb;
                ix[i] = swap ? ib : ia;
                ix[p] = swap ? ia : ib;
            }
            sync();
        }
    }
}

// (fp32 score, non-negative int32 row) as ONE 64-bit key whose unsigned order is "score ascending, then row descending":
// sorting keys in DESCENDING order yields (score desc, row asc) with a single compare per pair and no branches.
__device__ __forceinline__ uint64_t tk_pack_key(float score, int32_t row) {
    uint32_t u = __float_as_uint(score);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (static_cast<uint64_t>(u) << 32) | static_cast<uint64_t>(0x7fffffffu - static_cast<uint32_t>(row));
}
__device__ __forceinline__ float tk_key_score(uint64_t key) {
    uint32_t u = static_cast<uint32_t>(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ int32_t tk_key_row(uint64_t key) {
    return static_cast<int32_t>(0x7fffffffu - static_cast<uint32_t>(key & 0xffffffffu));
}

// bitonic sort of n (power of two) keys, largest first
template <typename SyncFn>
__device__ __forceinline__ void tk_bitonic_keys_desc(uint64_t *key, int n, int tid, int nthreads, SyncFn sync) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < n / 2; t += nthreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i + j;
                const uint64_t a = key[i], b = key[p];
                const bool first_big = (i & k) == 0;          // this pair's direction
                const bool swap = first_big ? (b > a) : (a > b);
                key[i] = swap ? b : a;
                key[p] = swap ? a : b;
            }
            sync();
        }
    }
}

__device__ __forceinline__ bool tk_masked(const int64_t *__restrict__ mask_rows, int64_t lo, int64_t hi, int64_t row) {
    while (lo < hi) {  // sorted ascending
        const int64_t mid = (lo + hi) >> 1;
        const int64_t v = mask_rows[mid];
        if (v == row) return true;
        if (v < row) lo = mid + 1; else hi = mid;
    }
    return false;
}


}  // namespace tt
