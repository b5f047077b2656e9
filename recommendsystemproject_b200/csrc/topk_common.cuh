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
                const int i = 2 * j * (t / j) + (t % j);
                const int p = i + j;
                const bool asc = (i & k) == 0;
                const V va = v[i], vb = v[p];
                const I ia = ix[i], ib = ix[p];
                const bool swap = asc ? tk_before(vb, ib, va, ia) : tk_before(va, ia, vb, ib);
                if (swap) { v[i] = vb; v[p] = va; ix[i] = ib; ix[p] = ia; }
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
