// Row-sharded embedding tables (owner = row % W, local row = row / W): the device side of the exchange.
//
// The reference is single-process (no collective anywhere); this is the SURVEY 8(e) design for its
// embedding layer (GenericTower.py:141-183) at the 100M-user / 10M-item scale of BASELINE configs[2].
//
// Wire format.  Every rank sends every owner ONE int32 block and gets ONE fp32 block back, both of a fixed size
// (capacities, not counts: no sizes cross the host, the whole step is CUDA-graph capturable).  Per table t a block holds
//   off_t [n_rows + 1]   exclusive offsets: the entries of sample b for this owner are [off[b], off[b+1])
//   rows_t[cap_t]        the owner's LOCAL rows, samples in order, positions in order inside a sample; unused = -1
// and the float block
//   pooled tables (len > 1): vec_t[n_rows][dim]  one partial sum per sample
//   single tables (len == 1): vec_t[cap_t][dim]  one row per entry
// The same float layout carries the gradients back in the backward pass.
//
// Kernels (all HBM / latency bound, no arithmetic to speak of):
//   shard_count / shard_scan / shard_fill   source side: bucket the ids of a batch by owner (ballots, no atomics:
//                                           order inside a sample is preserved, so the owner's sums are deterministic)
//   shard_owner_gather                      owner side: the gather + pool kernel over CSR segments of int32 rows
//   shard_combine                           source side: add the W partials in rank order (+ pads x pad row, / L)
//   shard_grad_pack                         source side of the backward: gradient rows into the float block
#include "common.cuh"

namespace tt {

constexpr int SHARD_MAX_WORLD = 32;

__device__ __forceinline__ void split_id(int64_t id, int world, bool narrow, int &owner, int32_t &local) {
    if (narrow) {   // vocab < 2^32: 32-bit division
        const uint32_t u = static_cast<uint32_t>(id);
        const uint32_t q = u / static_cast<uint32_t>(world);
        owner = static_cast<int>(u - q * static_cast<uint32_t>(world));
        local = static_cast<int32_t>(q);
    } else {
        const int64_t q = id / world;
        owner = static_cast<int>(id - q * world);
        local = static_cast<int32_t>(q);
    }
}

// ---- source side, len > 1: one warp per sample -------------------------------------------------------------------
// FILL = false: counts[w][row] (into the offsets slot, scanned afterwards) and n_pad[row]
// FILL = true : rows[w][off[w][row] + rank] = local row
template <bool FILL>
__global__ void __launch_bounds__(256)
shard_rows_kernel(const int64_t *__restrict__ ids, int64_t n_rows, int len, int64_t pad, int64_t vocab, int world,
                  int narrow, int32_t *__restrict__ send, int64_t block_ints, int64_t off_base, int64_t rows_base,
                  int64_t cap, int32_t *__restrict__ n_pad, int *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = warp0; row < n_rows; row += n_warps) {
        int32_t mine = 0;   // lane w < world: running count (FILL: running offset) of owner w
        if (FILL && lane < world) mine = send[lane * block_ints + off_base + row];
        int pads = 0;
        for (int l0 = 0; l0 < len; l0 += 32) {
            const int l = l0 + lane;
            int owner = -1;
            int32_t local = 0;
            bool is_pad = false;
            if (l < len) {
                const int64_t id = __ldg(ids + row * len + l);
                if (id == pad) is_pad = true;
                else if (id < 0 || id >= vocab) { if (!FILL) atomicOr(flags, 1); }
                else split_id(id, world, narrow != 0, owner, local);
            }
            if (!FILL) pads += __popc(__ballot_sync(0xffffffffu, is_pad));
            for (int w = 0; w < world; ++w) {
                const uint32_t b = __ballot_sync(0xffffffffu, owner == w);
                if (FILL) {
                    const int32_t base = __shfl_sync(0xffffffffu, mine, w);
                    if (owner == w) {
                        const int64_t pos = static_cast<int64_t>(base) + __popc(b & lt);
                        if (pos < cap) send[w * block_ints + rows_base + pos] = local;
                    }
                }
                if (lane == w) mine += __popc(b);
            }
        }
        if (!FILL) {
            if (lane < world) send[lane * block_ints + off_base + row] = mine;
            if (lane == 0 && n_pad) n_pad[row] = pads;
        }
    }
}

// ---- source side, len == 1: one thread per sample ------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256)
shard_single_kernel(const int64_t *__restrict__ ids, int64_t n_rows, int64_t pad, int64_t vocab, int world, int narrow,
                    int32_t *__restrict__ send, int64_t block_ints, int64_t off_base, int64_t rows_base, int64_t cap,
                    int32_t *__restrict__ n_pad, int *__restrict__ flags) {
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; row < n_rows;
         row += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t id = __ldg(ids + row);
        int owner = -1;
        int32_t local = 0;
        const bool is_pad = (id == pad);
        if (!is_pad) {
            if (id < 0 || id >= vocab) { if (!FILL) atomicOr(flags, 1); }
            else split_id(id, world, narrow != 0, owner, local);
        }
        if (!FILL) {
            for (int w = 0; w < world; ++w) send[w * block_ints + off_base + row] = (w == owner) ? 1 : 0;
            if (n_pad) n_pad[row] = is_pad ? 1 : 0;
        } else if (owner >= 0) {
            const int64_t pos = send[owner * block_ints + off_base + row];
            if (pos < cap) send[owner * block_ints + rows_base + pos] = local;
        }
    }
}

// in-place exclusive scan of the n counts of every owner block, two kernels, no single-CTA pass over the batch:
//   shard_scan_tiles   grid (tiles, world): a tile of SCAN_TILE counts becomes its LOCAL exclusive scan, tile total aside
//   shard_scan_bases   same grid: base = sum of the totals of the tiles before mine (at most a few dozen), added to the
//                      tile; the last tile also writes off[n] = grand total and raises the overflow flag
constexpr int SCAN_TILE = 4096;   // 1024 threads x 4

__global__ void __launch_bounds__(1024)
shard_scan_tiles(int32_t *__restrict__ send, int64_t block_ints, int64_t off_base, int64_t n, int32_t *__restrict__ totals) {
    __shared__ int32_t warp_sums[32];
    int32_t *off = send + static_cast<int64_t>(blockIdx.y) * block_ints + off_base;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + t * 4;
    int32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? off[i0 + k] : 0;
    const int32_t mine = v[0] + v[1] + v[2] + v[3];
    int32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        warp_sums[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    int32_t run = incl - mine + (warp > 0 ? warp_sums[warp - 1] : 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < n) off[i0 + k] = run;
        run += v[k];
    }
    if (t == 1023) totals[blockIdx.y * gridDim.x + blockIdx.x] = warp_sums[31];
}

__global__ void __launch_bounds__(1024)
shard_scan_bases(int32_t *__restrict__ send, int64_t block_ints, int64_t off_base, int64_t n, int64_t cap,
                 const int32_t *__restrict__ totals, int *__restrict__ flags) {
    int32_t *off = send + static_cast<int64_t>(blockIdx.y) * block_ints + off_base;
    const int32_t *tot = totals + blockIdx.y * gridDim.x;
    int32_t base = 0;
    for (unsigned k = 0; k < blockIdx.x; ++k) base += tot[k];     // uniform across the block, L2 hits
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * 4;
    if (base != 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k < n) off[i0 + k] += base;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        const int32_t total = base + tot[blockIdx.x];
        off[n] = total;
        if (total > cap) atomicOr(flags, 2);   // capacity overflow: entries beyond cap were dropped
    }
}

// ---- owner side ------------------------------------------------------------------------------------------------------
// pooled: one LPR-lane group per (source rank, sample): partial[s][b] = sum of table rows of that sample's entries
template <typename T, int LPR>
__global__ void __launch_bounds__(256)
shard_owner_pool_kernel(const T *__restrict__ table, int64_t local_rows, int dim, int world,
                        const int32_t *__restrict__ recv, int64_t block_ints, int64_t off_base, int64_t rows_base,
                        int64_t cap, int64_t n_rows, float *__restrict__ out, int64_t block_floats, int64_t vec_base,
                        int32_t *__restrict__ pos_src) {
    constexpr int VN = Vec16<T>::N;
    constexpr int UNROLL = 4;
    const int vpr = dim / VN;
    const int sub = threadIdx.x % LPR;
    const int lane = threadIdx.x & 31;
    const int grp_in_warp = lane / LPR;
    const int groups_per_block = blockDim.x / LPR;
    const int64_t total = static_cast<int64_t>(world) * n_rows;
    const int64_t warp_first = (static_cast<int64_t>(blockIdx.x) * groups_per_block + threadIdx.x / LPR) - grp_in_warp;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * groups_per_block;
    for (int64_t base = warp_first; base < total; base += stride) {
        const int64_t item = base + grp_in_warp;
        const bool ok = item < total;
        const int64_t s = ok ? item / n_rows : 0;
        const int64_t b = ok ? item - s * n_rows : 0;
        const int32_t *blk = recv + s * block_ints;
        int32_t e0 = 0, e1 = 0;
        if (ok) {
            e0 = __ldg(blk + off_base + b);
            e1 = __ldg(blk + off_base + b + 1);
            if (e1 > cap) e1 = static_cast<int32_t>(cap);
            if (e0 > e1) e0 = e1;
        }
        // the longest segment of the warp decides the trip count (shuffles need every lane)
        int n_e = e1 - e0;
        int n_max = n_e;
#pragma unroll
        for (int o = 16; o >= LPR; o >>= 1) n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, o));
        for (int c0 = 0; c0 < vpr; c0 += LPR) {
            const int c = c0 + sub;
            const bool col_ok = ok && c < vpr;
            float acc[VN];
#pragma unroll
            for (int e = 0; e < VN; ++e) acc[e] = 0.f;
            for (int j0 = 0; j0 < n_max; j0 += LPR) {
                int32_t my_row = -1;
                if (j0 + sub < n_e) {
                    my_row = __ldg(blk + rows_base + e0 + j0 + sub);
                    // where this entry's gradient row will sit in the float block of the backward (float4 units)
                    if (pos_src && c0 == 0)
                        pos_src[s * cap + e0 + j0 + sub] = static_cast<int32_t>((s * block_floats + vec_base + b * dim) >> 2);
                }
                const int cnt = min(LPR, n_max - j0);
                for (int j = 0; j < cnt; j += UNROLL) {
                    float v[UNROLL][VN];
                    bool use[UNROLL];
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) {
                        const int src = (j + u < LPR) ? (j + u) : 0;
                        const int32_t r = __shfl_sync(0xffffffffu, my_row, src, LPR);
                        use[u] = (j + u) < cnt && r >= 0 && r < local_rows && col_ok;
                        if (use[u]) Vec16<T>::load(table + static_cast<int64_t>(r) * dim + c * VN, v[u]);
                    }
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) {
                        if (!use[u]) continue;
#pragma unroll
                        for (int e = 0; e < VN; ++e) acc[e] += v[u][e];
                    }
                }
            }
            if (col_ok) {
                float *o = out + s * block_floats + vec_base + b * dim + c * VN;
#pragma unroll
                for (int e = 0; e < VN; e += 4)
                    *reinterpret_cast<float4 *>(o + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
            }
        }
    }
}

// single: one LPR-lane group per (source rank, entry): vec[s][e] = table[rows[s][e]]
template <typename T, int LPR>
__global__ void __launch_bounds__(256)
shard_owner_rows_kernel(const T *__restrict__ table, int64_t local_rows, int dim, int world,
                        const int32_t *__restrict__ recv, int64_t block_ints, int64_t rows_base, int64_t cap,
                        float *__restrict__ out, int64_t block_floats, int64_t vec_base, int32_t *__restrict__ pos_src) {
    constexpr int VN = Vec16<T>::N;
    const int vpr = dim / VN;
    const int sub = threadIdx.x % LPR;
    const int64_t total = static_cast<int64_t>(world) * cap;
    const int64_t g0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
    const int64_t ng = static_cast<int64_t>(gridDim.x) * blockDim.x / LPR;
    for (int64_t item = g0; item < total; item += ng) {
        const int64_t s = item / cap;
        const int64_t e = item - s * cap;
        const int32_t r = __ldg(recv + s * block_ints + rows_base + e);
        if (pos_src && sub == 0) pos_src[item] = static_cast<int32_t>((s * block_floats + vec_base + e * dim) >> 2);
        if (r < 0 || r >= local_rows) continue;
        for (int c = sub; c < vpr; c += LPR) {
            float v[VN];
            Vec16<T>::load(table + static_cast<int64_t>(r) * dim + c * VN, v);
            float *o = out + s * block_floats + vec_base + e * dim + c * VN;
#pragma unroll
            for (int k = 0; k < VN; k += 4) *reinterpret_cast<float4 *>(o + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
        }
    }
}

// ---- source side: combine the returned vectors --------------------------------------------------------------------
// one thread per (sample, float4 column)
__global__ void __launch_bounds__(256)
shard_combine_kernel(const float *__restrict__ recv_vec, int64_t block_floats, int64_t vec_base, int world,
                     const int64_t *__restrict__ ids, int64_t n_rows, int len, int64_t pad, int64_t vocab, int narrow,
                     int mode, const int32_t *__restrict__ send, int64_t block_ints, int64_t off_base, int64_t cap,
                     const int32_t *__restrict__ n_pad, const float *__restrict__ pad_row, int dim,
                     float *__restrict__ out, int64_t out_stride) {
    const int vpr = dim / 4;
    const int64_t total = n_rows * vpr;
    const float inv_len = 1.0f / static_cast<float>(len);
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t b = t / vpr;
        const int c = static_cast<int>(t - b * vpr);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (len > 1) {
            for (int w = 0; w < world; ++w) {   // rank order: the sum is a function of the inputs only
                const float4 v = __ldg(reinterpret_cast<const float4 *>(recv_vec + w * block_floats + vec_base + b * dim) + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            const int np = n_pad ? n_pad[b] : 0;
            if (np > 0 && pad_row) {   // pads pool like any id (GenericTower.py:153-160): add the pad row np times
                const float4 p = __ldg(reinterpret_cast<const float4 *>(pad_row) + c);
                const float f = static_cast<float>(np);
                acc.x += f * p.x; acc.y += f * p.y; acc.z += f * p.z; acc.w += f * p.w;
            }
            if (mode == TT_POOL_MEAN) { acc.x *= inv_len; acc.y *= inv_len; acc.z *= inv_len; acc.w *= inv_len; }
        } else {
            const int64_t id = __ldg(ids + b);
            if (id == pad) {
                if (pad_row) acc = __ldg(reinterpret_cast<const float4 *>(pad_row) + c);
            } else if (id >= 0 && id < vocab) {
                int owner;
                int32_t local;
                split_id(id, world, narrow != 0, owner, local);
                const int64_t slot = send[owner * block_ints + off_base + b];
                if (slot < cap)
                    acc = __ldg(reinterpret_cast<const float4 *>(recv_vec + owner * block_floats + vec_base + slot * dim) + c);
            }
        }
        *reinterpret_cast<float4 *>(out + b * out_stride + c * 4) = acc;
    }
}

// ---- source side of the backward: gradient rows into the float block ----------------------------------------------
__global__ void __launch_bounds__(256)
shard_grad_pack_kernel(const float *__restrict__ grad_out, int64_t grad_stride, int64_t n_rows, int len, int mode,
                       int dim, int world, const int64_t *__restrict__ ids, int64_t pad, int64_t vocab, int narrow,
                       const int32_t *__restrict__ send, int64_t block_ints, int64_t off_base, int64_t cap,
                       float *__restrict__ send_vec, int64_t block_floats, int64_t vec_base) {
    const int vpr = dim / 4;
    const int64_t total = n_rows * vpr;
    const float scale = (mode == TT_POOL_MEAN && len > 1) ? 1.0f / static_cast<float>(len) : 1.0f;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t b = t / vpr;
        const int c = static_cast<int>(t - b * vpr);
        float4 g = __ldg(reinterpret_cast<const float4 *>(grad_out + b * grad_stride) + c);
        if (len > 1) {
            g.x *= scale; g.y *= scale; g.z *= scale; g.w *= scale;
            for (int w = 0; w < world; ++w)
                *(reinterpret_cast<float4 *>(send_vec + w * block_floats + vec_base + b * dim) + c) = g;
        } else {
            const int64_t id = __ldg(ids + b);
            if (id == pad || id < 0 || id >= vocab) continue;
            int owner;
            int32_t local;
            split_id(id, world, narrow != 0, owner, local);
            const int64_t slot = send[owner * block_ints + off_base + b];
            if (slot < cap) *(reinterpret_cast<float4 *>(send_vec + owner * block_floats + vec_base + slot * dim) + c) = g;
        }
    }
}

static inline unsigned shard_grid(int64_t threads_needed, int threads) {
    int64_t b = (threads_needed + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b);
}

}  // namespace tt

extern "C" int tt_shard_route(const int64_t *ids, int64_t n_rows, int len, int64_t padding_idx, int64_t vocab, int world,
                              int32_t *send, int64_t block_ints, int64_t off_base, int64_t rows_base, int64_t cap,
                              int32_t *n_pad, int *flags, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(ids && send && flags && workspace, "null pointer");
    TT_CHECK_ARG(n_rows > 0 && len > 0 && vocab > 0 && cap > 0, "non-positive size");
    TT_CHECK_ARG(world >= 1 && world <= SHARD_MAX_WORLD, "world must be in [1, 32]");
    TT_CHECK_ARG(off_base >= 0 && rows_base >= off_base + n_rows + 1 && block_ints >= rows_base + cap, "block layout");
    TT_CHECK_ARG((vocab + world - 1) / world < (int64_t(1) << 31), "local rows must fit 31 bits");
    TT_CHECK_ARG(n_rows < (int64_t(1) << 31) && static_cast<int64_t>(n_rows) * len < (int64_t(1) << 31), "too many positions");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int narrow = vocab < (int64_t(1) << 32) ? 1 : 0;
    const unsigned tiles = static_cast<unsigned>((n_rows + SCAN_TILE - 1) / SCAN_TILE);
    if (workspace_bytes < sizeof(int32_t) * tiles * world) { set_error("shard_route workspace needs %zu bytes", sizeof(int32_t) * tiles * world); return TT_E_WORKSPACE; }
    int32_t *totals = static_cast<int32_t *>(workspace);
    const dim3 scan_grid(tiles, world);
    {   // unused row slots read as -1 on the owner: one strided memset over the W owner blocks
        cudaError_t e = cudaMemset2DAsync(send + rows_base, sizeof(int32_t) * block_ints, 0xff, sizeof(int32_t) * cap, world, st);
        if (e != cudaSuccess) return cuda_status(e, "cudaMemset2DAsync(shard rows)");
    }
    if (len > 1) {
        const unsigned grid = shard_grid(n_rows * 32, 256);
        shard_rows_kernel<false><<<grid, 256, 0, st>>>(ids, n_rows, len, padding_idx, vocab, world, narrow, send, block_ints,
                                                        off_base, rows_base, cap, n_pad, flags);
        TT_LAUNCH_CHECK("shard_rows_kernel<count>");
        shard_scan_tiles<<<scan_grid, 1024, 0, st>>>(send, block_ints, off_base, n_rows, totals);
        shard_scan_bases<<<scan_grid, 1024, 0, st>>>(send, block_ints, off_base, n_rows, cap, totals, flags);
        TT_LAUNCH_CHECK("shard_scan");
        shard_rows_kernel<true><<<grid, 256, 0, st>>>(ids, n_rows, len, padding_idx, vocab, world, narrow, send, block_ints,
                                                       off_base, rows_base, cap, n_pad, flags);
        TT_LAUNCH_CHECK("shard_rows_kernel<fill>");
    } else {
        const unsigned grid = shard_grid(n_rows, 256);
        shard_single_kernel<false><<<grid, 256, 0, st>>>(ids, n_rows, padding_idx, vocab, world, narrow, send, block_ints,
                                                          off_base, rows_base, cap, n_pad, flags);
        TT_LAUNCH_CHECK("shard_single_kernel<count>");
        shard_scan_tiles<<<scan_grid, 1024, 0, st>>>(send, block_ints, off_base, n_rows, totals);
        shard_scan_bases<<<scan_grid, 1024, 0, st>>>(send, block_ints, off_base, n_rows, cap, totals, flags);
        TT_LAUNCH_CHECK("shard_scan");
        shard_single_kernel<true><<<grid, 256, 0, st>>>(ids, n_rows, padding_idx, vocab, world, narrow, send, block_ints,
                                                         off_base, rows_base, cap, n_pad, flags);
        TT_LAUNCH_CHECK("shard_single_kernel<fill>");
    }
    return 0;
}

namespace tt {

template <typename T>
static int owner_gather(const T *table, int64_t local_rows, int dim, int world, const int32_t *recv, int64_t block_ints,
                        int64_t off_base, int64_t rows_base, int64_t cap, int64_t n_rows, int pooled, float *out,
                        int64_t block_floats, int64_t vec_base, int32_t *pos_src, cudaStream_t st) {
    constexpr int VN = Vec16<T>::N;
    const int vpr = dim / VN;
    const int lpr = vpr >= 32 ? 32 : (vpr >= 16 ? 16 : (vpr >= 8 ? 8 : (vpr >= 4 ? 4 : (vpr >= 2 ? 2 : 1))));
    const int64_t groups = pooled ? world * n_rows : world * cap;
    const unsigned grid = shard_grid(groups * lpr, 256);
#define TT_OG(L)                                                                                                          \
    do {                                                                                                                  \
        if (pooled)                                                                                                       \
            shard_owner_pool_kernel<T, L><<<grid, 256, 0, st>>>(table, local_rows, dim, world, recv, block_ints, off_base, \
                                                                 rows_base, cap, n_rows, out, block_floats, vec_base, pos_src); \
        else                                                                                                              \
            shard_owner_rows_kernel<T, L><<<grid, 256, 0, st>>>(table, local_rows, dim, world, recv, block_ints, rows_base, \
                                                                 cap, out, block_floats, vec_base, pos_src);              \
    } while (0)
    switch (lpr) {
        case 1: TT_OG(1); break;
        case 2: TT_OG(2); break;
        case 4: TT_OG(4); break;
        case 8: TT_OG(8); break;
        case 16: TT_OG(16); break;
        default: TT_OG(32); break;
    }
#undef TT_OG
    TT_LAUNCH_CHECK("shard_owner_gather");
    return 0;
}

}  // namespace tt

extern "C" int tt_shard_owner_gather(const void *table, int table_dtype, int64_t local_rows, int dim, int world,
                                     const int32_t *recv, int64_t block_ints, int64_t off_base, int64_t rows_base,
                                     int64_t cap, int64_t n_rows, int pooled, float *out, int64_t block_floats,
                                     int64_t vec_base, int32_t *pos_src, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(table && recv && out, "null pointer");
    TT_CHECK_ARG(local_rows >= 0 && dim > 0 && world >= 1 && world <= SHARD_MAX_WORLD && cap > 0 && n_rows > 0, "bad size");
    TT_CHECK_ARG(table_dtype == TT_F32 || table_dtype == TT_BF16, "unknown table dtype");
    TT_CHECK_ARG(dim % (table_dtype == TT_F32 ? 4 : 8) == 0, "sharded tables need 16-byte rows (dim % 4 == 0 fp32, % 8 bf16)");
    TT_CHECK_ARG(block_floats % 4 == 0 && vec_base % 4 == 0, "float block layout must be 16-byte aligned");
    TT_CHECK_ARG(static_cast<int64_t>(world) * block_floats < (int64_t(1) << 33), "float blocks exceed the 31-bit float4 offsets of pos_src");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (table_dtype == TT_F32)
        return owner_gather<float>(static_cast<const float *>(table), local_rows, dim, world, recv, block_ints, off_base,
                                   rows_base, cap, n_rows, pooled, out, block_floats, vec_base, pos_src, st);
    return owner_gather<__nv_bfloat16>(static_cast<const __nv_bfloat16 *>(table), local_rows, dim, world, recv, block_ints,
                                       off_base, rows_base, cap, n_rows, pooled, out, block_floats, vec_base, pos_src, st);
}

extern "C" int tt_shard_combine(const float *recv_vec, int64_t block_floats, int64_t vec_base, int world,
                                const int64_t *ids, int64_t n_rows, int len, int64_t padding_idx, int64_t vocab, int mode,
                                const int32_t *send, int64_t block_ints, int64_t off_base, int64_t cap,
                                const int32_t *n_pad, const float *pad_row, int dim, float *out, int64_t out_stride,
                                void *stream) {
    using namespace tt;
    TT_CHECK_ARG(recv_vec && ids && send && out, "null pointer");
    TT_CHECK_ARG(n_rows > 0 && len > 0 && dim > 0 && dim % 4 == 0 && out_stride % 4 == 0, "bad size");
    TT_CHECK_ARG(world >= 1 && world <= SHARD_MAX_WORLD, "world must be in [1, 32]");
    TT_CHECK_ARG(mode == TT_POOL_NONE || mode == TT_POOL_SUM || mode == TT_POOL_MEAN, "sharded tables pool with sum / mean");
    TT_CHECK_ARG(reinterpret_cast<uintptr_t>(out) % 16 == 0, "out must be 16-byte aligned");
    const int narrow = vocab < (int64_t(1) << 32) ? 1 : 0;
    shard_combine_kernel<<<shard_grid(n_rows * (dim / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        recv_vec, block_floats, vec_base, world, ids, n_rows, len, padding_idx, vocab, narrow, mode, send, block_ints,
        off_base, cap, n_pad, pad_row, dim, out, out_stride);
    TT_LAUNCH_CHECK("shard_combine_kernel");
    return 0;
}

extern "C" int tt_shard_grad_pack(const float *grad_out, int64_t grad_stride, int64_t n_rows, int len, int mode, int dim,
                                  int world, const int64_t *ids, int64_t padding_idx, int64_t vocab, const int32_t *send,
                                  int64_t block_ints, int64_t off_base, int64_t cap, float *send_vec, int64_t block_floats,
                                  int64_t vec_base, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(grad_out && ids && send && send_vec, "null pointer");
    TT_CHECK_ARG(n_rows > 0 && len > 0 && dim > 0 && dim % 4 == 0 && grad_stride % 4 == 0, "bad size");
    TT_CHECK_ARG(world >= 1 && world <= SHARD_MAX_WORLD, "world must be in [1, 32]");
    TT_CHECK_ARG(reinterpret_cast<uintptr_t>(grad_out) % 16 == 0, "grad_out must be 16-byte aligned");
    const int narrow = vocab < (int64_t(1) << 32) ? 1 : 0;
    shard_grad_pack_kernel<<<shard_grid(n_rows * (dim / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        grad_out, grad_stride, n_rows, len, mode, dim, world, ids, padding_idx, vocab, narrow, send, block_ints, off_base,
        cap, send_vec, block_floats, vec_base);
    TT_LAUNCH_CHECK("shard_grad_pack_kernel");
    return 0;
}
