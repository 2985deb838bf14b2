// p2p.cu — data-parallel exchange over NVLink peer memory, without NCCL on the hot path.
//
// The weight gradient is the path's one exchange step.  For large layers the exchange is fused into
// the kernels on either side of it:
//   * reduce-scatter: the tcgen05 weight-gradient kernel (gemm_tc.cu, scatter mode) stores every finished output tile
//     straight into the OWNER rank's gradient arena through its peer mapping — rows [o*R, (o+1)*R) of a layer belong
//     to rank o, and the owner's arena for that layer is viewed as `world` partial slices of R rows, one per source
//     rank.  The NVLink transfer of tile i overlaps the MMAs of tile i+1 (double-buffered TMEM accumulators).
//   * reduce + update + all-gather: the owner sums the `world` partial slices, applies the momentum update to its rows
//     (cublasNN.cpp:558-566 semantics) and stores the new weights into EVERY rank's weight arena.
// Ordering between ranks is by monotonically increasing step counters in peer-visible flag words:
//   wg_done[layer][src]   written by `src` into every rank after its gradient tiles of `layer` are stored
//   w_done[layer][owner]  written by `owner` into every rank after its rows of `layer` are updated everywhere
// Only tiny one-warp gate kernels ever spin, and they spin on LOCAL memory (peers write the flags into it).
#include "engine.h"

namespace b2 {
namespace {

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// Spin until flags[i] >= value for i in [0, count).  One warp; lanes take flags round-robin.
__global__ void gate_kernel(const int* flags, int count, int value) {
    const long long t0 = clock64();
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        while (ld_acquire_sys(flags + i) < value) {
            __nanosleep(200);
            if (clock64() - t0 > 20000000000LL) __trap();      // ~10 s: a peer died; fail loudly instead of hanging
        }
    }
}

// Same for every (layer, rank) entry selected by `layer_mask` of a [P2P_MAX_LAYERS][P2P_MAX_WORLD] flag block.
__global__ void gate_layers_kernel(const int* flags, unsigned long long layer_mask, int world, int value) {
    const long long t0 = clock64();
    for (int i = threadIdx.x; i < P2P_MAX_LAYERS * P2P_MAX_WORLD; i += blockDim.x) {
        const int layer = i / P2P_MAX_WORLD, r = i - layer * P2P_MAX_WORLD;
        if (!((layer_mask >> layer) & 1ull) || r >= world) continue;
        while (ld_acquire_sys(flags + i) < value) {
            __nanosleep(200);
            if (clock64() - t0 > 20000000000LL) __trap();
        }
    }
}

// Small layer, source side: this rank's whole gradient into slot `rank` of every peer (NVLink stores).
__global__ void __launch_bounds__(256, 8) small_scatter_kernel(P2pPeers peers, const float4* __restrict__ src, int64_t slot_off4,
                                                            int64_t n4) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = src[i];
#pragma unroll 1
        for (int p = 0; p < peers.world; ++p) reinterpret_cast<float4*>(peers.S[p])[slot_off4 + i] = v;
    }
}

// Small layer, every rank: g = sum of the slots in rank order (same order everywhere => replicas stay bit-identical),
// then the momentum update of the local replica.
// <= 32 registers (8 blocks / SM by the launch bound): a block then fits into the register file NEXT TO a persistent
// tcgen05 GEMM CTA (576 threads x 96 registers leave 10 240), so the owner side of the exchange runs under the GEMMs
// instead of waiting for one of them to retire.
__global__ void __launch_bounds__(256, 8) small_reduce_update_kernel(const float4* __restrict__ slots, int64_t stride4, int world,
                                                                  float4* __restrict__ theta, float4* __restrict__ vel,
                                                                  int64_t n4, float mu, float alpha) {
    const int64_t step = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += step) {
        float4 g = slots[i];
        for (int r = 1; r < world; ++r) {
            const float4 q = slots[r * stride4 + i];
            g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
        }
        float4 v = vel[i];
        float4 t = theta[i];
        v.x = fmaf(mu, v.x, alpha * g.x); v.y = fmaf(mu, v.y, alpha * g.y);
        v.z = fmaf(mu, v.z, alpha * g.z); v.w = fmaf(mu, v.w, alpha * g.w);
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        vel[i] = v;
        theta[i] = t;
    }
}

// flags_of_peer[p][index] = value on every rank p (including this one), after everything this GPU stored before.
__global__ void signal_kernel(P2pPeers peers, int index, int value) {
    __threadfence_system();
    if (threadIdx.x < peers.world) st_release_sys(peers.flags[threadIdx.x] + index, value);
}

// Owner-side fused kernel for one layer: g = sum over sources of the partial slices, v <- mu v + alpha g,
// theta <- theta + v, new theta stored into every rank's weight arena.  `count` = R * ldw floats (multiple of 4).
__global__ void __launch_bounds__(256, 8) reduce_update_allgather_kernel(P2pPeers peers, const float4* __restrict__ partial,
                                                                      int64_t slice4, float4* __restrict__ vel,
                                                                      int64_t w_off4, int64_t count4, float mu, float alpha) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count4; i += stride) {
        float4 g = __ldcs(partial + i);
        for (int r = 1; r < peers.world; ++r) {
            const float4 q = __ldcs(partial + r * slice4 + i);
            g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
        }
        float4 v = vel[i];
        float4 t = reinterpret_cast<const float4*>(peers.W[peers.rank])[w_off4 + i];
        v.x = fmaf(mu, v.x, alpha * g.x); v.y = fmaf(mu, v.y, alpha * g.y);
        v.z = fmaf(mu, v.z, alpha * g.z); v.w = fmaf(mu, v.w, alpha * g.w);
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        vel[i] = v;
#pragma unroll 1
        for (int p = 0; p < peers.world; ++p) reinterpret_cast<float4*>(peers.W[p])[w_off4 + i] = t;
    }
}

}  // namespace

// The owner kernels should run UNDER the persistent GEMMs, whose CTAs configure their SM for the largest shared-memory
// carve-out: ask for the same carve-out, so that becoming resident on such an SM needs no reconfiguration.
template <typename Kern>
static void prefer_max_shared(Kern k) {
    static bool done = false;      // per kernel instantiation; the attribute is advisory, failures are ignored
    if (!done) { (void)cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); (void)cudaGetLastError(); done = true; }
}

cudaError_t p2p_gate_launch(const int* flags, int count, int value, cudaStream_t s) {
    gate_kernel<<<1, 32, 0, s>>>(flags, count, value);
    return cudaGetLastError();
}
cudaError_t p2p_gate_layers_launch(const int* flags_kind, uint64_t layer_mask, int world, int value, cudaStream_t s) {
    if (layer_mask == 0) return cudaSuccess;
    gate_layers_kernel<<<1, 128, 0, s>>>(flags_kind, static_cast<unsigned long long>(layer_mask), world, value);
    return cudaGetLastError();
}
cudaError_t p2p_small_scatter_launch(const P2pPeers& peers, const float* src, int64_t slot_off, int64_t count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const int64_t n4 = count / 4, want = (n4 + 255) / 256;
    small_scatter_kernel<<<unsigned(want < 148 ? want : 148), 256, 0, s>>>(peers, reinterpret_cast<const float4*>(src), slot_off / 4, n4);
    return cudaGetLastError();
}
cudaError_t p2p_small_reduce_update_launch(const float* slots, int64_t slot_stride, int world, float* theta, float* vel,
                                           int64_t count, float mu, float alpha, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    prefer_max_shared(small_reduce_update_kernel);
    const int64_t n4 = count / 4, want = (n4 + 255) / 256;
    small_reduce_update_kernel<<<unsigned(want < 148 ? want : 148), 256, 0, s>>>(reinterpret_cast<const float4*>(slots), slot_stride / 4,
                                                                               world, reinterpret_cast<float4*>(theta),
                                                                               reinterpret_cast<float4*>(vel), n4, mu, alpha);
    return cudaGetLastError();
}
cudaError_t p2p_signal_launch(const P2pPeers& peers, int index, int value, cudaStream_t s) {
    signal_kernel<<<1, 32, 0, s>>>(peers, index, value);
    return cudaGetLastError();
}
cudaError_t p2p_reduce_update_launch(const P2pPeers& peers, const float* partial, int64_t slice_floats, float* vel,
                                     int64_t w_off, int64_t count, float mu, float alpha, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    prefer_max_shared(reduce_update_allgather_kernel);
    const int64_t n4 = count / 4;
    const int64_t want = (n4 + 255) / 256;
    const unsigned grid = unsigned(want < 148 * 4 ? want : 148 * 4);
    reduce_update_allgather_kernel<<<grid, 256, 0, s>>>(peers, reinterpret_cast<const float4*>(partial), slice_floats / 4,
                                                       reinterpret_cast<float4*>(vel), w_off / 4, n4, mu, alpha);
    return cudaGetLastError();
}

}  // namespace b2
