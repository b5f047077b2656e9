// One-shot all-gather of a small vector over NVLink peer memory (SyncBatchNorm statistics of the data-parallel towers).
//
// NCCL's fixed cost for a tiny collective inside the step's CUDA graph is ~25 us at 8 GPUs and the step has twelve of
// them (6 BatchNorm layers, forward statistics + backward sums); the payload is 1-5 KB.  Here every rank stores its
// vector straight into slot [rank] of EVERY peer's symmetric buffer (ld/st on peer pointers mapped through NVSwitch),
// publishes a release flag to each peer and spins on its own W flags: one kernel, one NVLink round trip.
// The consumer (tt_bn_apply / tt_bn_bwd_apply) then reads the W vectors from LOCAL memory and merges them in rank order
// -- deterministic, unlike an all-reduce whose summation order is the library's.
//
// Reuse: a (region, flag) pair is used once per step by one call site; between two uses every rank passes other
// gathers of the same step (each is a full barrier), so a writer can never overtake a reader of the previous step.
// The epoch counter lives in device memory and is bumped by the kernel: CUDA-graph replays need no host update.
#include "common.cuh"

namespace tt {

constexpr int P2P_MAX_WORLD = 16;
struct P2pPeers {
    float *buf[P2P_MAX_WORLD];        // peer r's symmetric buffer (as mapped in THIS process)
};

__global__ void __launch_bounds__(256)
p2p_allgather_small_kernel(const float *__restrict__ src, int n, int rank, int world, P2pPeers peers,
                           int64_t region_off, int64_t slot_floats, int64_t flag_off, unsigned int *__restrict__ epoch_dev) {
    __shared__ unsigned int s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = *epoch_dev + 1u;
        *epoch_dev = s_epoch;
    }
    __syncthreads();
    const unsigned int e = s_epoch;
    // 1. my vector into slot [rank] of every peer (including myself)
    for (int w = 0; w < world; ++w) {
        float *dst = peers.buf[w] + region_off + static_cast<int64_t>(rank) * slot_floats;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: flag [rank] of every peer = epoch (release, system scope)
    if (threadIdx.x < world) {
        unsigned int *flag = reinterpret_cast<unsigned int *>(peers.buf[threadIdx.x] + flag_off) + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(e) : "memory");
    }
    // 3. wait for everybody's flag in MY buffer
    if (threadIdx.x < world) {
        const unsigned int *flag = reinterpret_cast<const unsigned int *>(peers.buf[rank] + flag_off) + threadIdx.x;
        unsigned int v = 0;
        unsigned long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (++spins > (1ull << 31)) { __trap(); }
        } while (static_cast<int>(v - e) < 0);
    }
    __syncthreads();
}

}  // namespace tt

extern "C" int tt_p2p_allgather_small(const float *src, int n, int rank, int world, const void *const *peer_bufs_host,
                                      int64_t region_off_floats, int64_t slot_floats, int64_t flag_off_floats,
                                      unsigned int *epoch_dev, void *stream) {
    using namespace tt;
    TT_CHECK_ARG(src && peer_bufs_host && epoch_dev, "null pointer");
    TT_CHECK_ARG(world >= 1 && world <= P2P_MAX_WORLD && rank >= 0 && rank < world, "world must be in [1, 16]");
    TT_CHECK_ARG(n > 0 && n <= slot_floats, "vector longer than its slot");
    P2pPeers peers{};
    for (int w = 0; w < world; ++w) {
        TT_CHECK_ARG(peer_bufs_host[w] != nullptr, "null peer buffer");
        peers.buf[w] = static_cast<float *>(const_cast<void *>(peer_bufs_host[w]));
    }
    p2p_allgather_small_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, n, rank, world, peers, region_off_floats,
                                                                                 slot_floats, flag_off_floats, epoch_dev);
    TT_LAUNCH_CHECK("p2p_allgather_small_kernel");
    return 0;
}
