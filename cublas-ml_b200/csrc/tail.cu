// tail.cu — the skinny last layer as ONE HBM-streaming kernel.
//
// When the output layer is narrow (K <= 32 units: 10 classes in configs C1/C3/C4/C5, 1 target in C2) its three
// contractions are "GEMMs" with N = K or K-depth = K: a tensor-core tile is >90 % padding and the work is pure memory
// streaming of the last hidden activation a (rows x (H+1)).  The reference spends seven launches there
// (cublasSgemm N,N :658-659; activationOutput :606; vecVecSubtract :666; cost :613-614; activationDerivative :673;
// cublasSgemm N,T :674-675; vecVecElementMultiply :676).  Three kernels share the work (tail_launch picks one):
// tail_tile_kernel (training, small hidden layer: the 32-sample tile of a lives in shared memory), tail_small_kernel
// (inference, one warp per 32 samples) and tail_kernel, which does, per block of 32 samples x 16 warps:
//   pass 1   z = [a,1] W^T           (fp32 FFMA, W chunked through shared memory, h split over the 16 warps)
//            h = softmax / sigmoid / identity (z);  delta = h - y;  cost, argmax, error count
//   pass 2   delta_prev = (delta W)[no bias] .* f'(a)          (second read of a comes from L2)
//            optionally G = delta^T [a,1] for small layers (warp-shuffle reduction over the 32 samples, block-local
//            accumulation in shared memory, one atomic flush per block)
// Lanes walk consecutive samples, so every global access is a 128-byte line.  Arithmetic is exact fp32 in every
// precision mode.
#include "engine.h"
#include <cstdlib>
#include "pdl.cuh"
#include "epilogue.cuh"
#include "../../include/b2nn.h"

namespace b2 {
namespace {

constexpr int TAIL_THREADS = 512;
constexpr int TAIL_WARPS = TAIL_THREADS / 32;
constexpr int HC = 512;                   // hidden units per shared-memory chunk of W (multiple of TAIL_WARPS)

template <int KMAX, bool TRAIN>
__global__ void __launch_bounds__(TAIL_THREADS) tail_kernel(const TailArgs p) {
    pdl_enter();
    extern __shared__ float smem[];
    float* W_s = smem;                                    // [KMAX][HC + 1]
    float* red = W_s + KMAX * (HC + 1);                   // [TAIL_WARPS][KMAX][32]
    float* G_s = red + TAIL_WARPS * KMAX * 32;            // [K][H + 1] when p.G_inline
    __shared__ double cost_blk, err_blk;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = p.K, H1 = p.H + 1;
    const bool inline_g = TRAIN && p.G_inline != nullptr;
    if (inline_g) for (int idx = threadIdx.x; idx < K * H1; idx += TAIL_THREADS) G_s[idx] = 0.f;
    if (threadIdx.x == 0) { cost_blk = 0.0; err_blk = 0.0; }
    __syncthreads();

    const int64_t n_groups = (p.rows + 31) / 32;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t i = grp * 32 + lane;
        const bool ok = i < p.rows;
        const int64_t ic = ok ? i : p.rows - 1;            // clamped: loads stay in bounds, results are discarded

        // ---------------- pass 1: z = [a, 1] W^T ----------------
        float zp[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) zp[k] = 0.f;
        for (int h0 = 0; h0 < H1; h0 += HC) {
            const int hc = min(HC, H1 - h0);
            for (int idx = threadIdx.x; idx < K * hc; idx += TAIL_THREADS) {
                const int k = idx / hc, hh = idx - k * hc;
                W_s[k * (HC + 1) + hh] = __ldg(p.W + int64_t(k) * p.ldw + h0 + hh);
            }
            __syncthreads();
            const float* ap = p.a + int64_t(h0) * p.lda + ic;
            // 8 independent 128-byte loads in flight per warp before the first FMA consumes one: this loop is pure HBM
            // streaming and lives on memory-level parallelism
            int hh = warp;
            for (; hh + 7 * TAIL_WARPS < hc; hh += 8 * TAIL_WARPS) {
                float av[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) av[u] = __ldg(ap + int64_t(hh + u * TAIL_WARPS) * p.lda);
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) if (k < K) zp[k] = fmaf(av[u], W_s[k * (HC + 1) + hh + u * TAIL_WARPS], zp[k]);
            }
            for (; hh < hc; hh += TAIL_WARPS) {
                const float av = __ldg(ap + int64_t(hh) * p.lda);
#pragma unroll
                for (int k = 0; k < KMAX; ++k) if (k < K) zp[k] = fmaf(av, W_s[k * (HC + 1) + hh], zp[k]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K) red[(warp * KMAX + k) * 32 + lane] = zp[k];
        __syncthreads();
        float dl[KMAX];                                    // delta of this lane's sample (every warp computes it)
        {
            float z[KMAX];
            float zmax = -INFINITY;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                float s = 0.f;
                if (k < K) {
#pragma unroll
                    for (int w = 0; w < TAIL_WARPS; ++w) s += red[(w * KMAX + k) * 32 + lane];
                    zmax = fmaxf(zmax, s);
                }
                z[k] = s;
            }
            float total = 0.f;
            if (p.out_kind == B2NN_OUT_SOFTMAX) {
#pragma unroll
                for (int k = 0; k < KMAX; ++k) if (k < K) total += expf(z[k] - zmax);
            }
            float best_h = 0.f, best_y = 0.f; int arg_h = 0, arg_y = 0;
            double cost = 0.0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                dl[k] = 0.f;
                if (k < K) {
                    float hv;
                    if (p.out_kind == B2NN_OUT_SOFTMAX) hv = expf(z[k] - zmax) / total;
                    else if (p.out_kind == B2NN_OUT_SIGMOID) hv = sigmoid_f(z[k]);
                    else hv = z[k];
                    if (hv > best_h) { best_h = hv; arg_h = k; }
                    float yv = 0.f;
                    if (p.Y) {
                        yv = __ldg(p.Y + int64_t(k) * p.ldy + ic);
                        if (yv > best_y) { best_y = yv; arg_y = k; }
                    }
                    const float dv = hv - yv;
                    dl[k] = ok ? dv : 0.f;
                    if (warp == 0 && ok) {
                        if (p.h_out) p.h_out[int64_t(k) * p.ldh + i] = hv;
                        if (p.delta_last) p.delta_last[int64_t(k) * p.ldd + i] = dv;
                        if (p.cost_sum && p.Y) {
                            float c = 0.f;
                            if (p.cost_kind == B2NN_COST_SQUARED) c = dv * dv;
                            else {
                                if (yv != 0.f) c += -yv * logf(hv);
                                if (p.cost_kind == B2NN_COST_NEGLNMAX && yv != 1.f) c += -(1.f - yv) * logf(1.f - hv);
                                c = fabsf(c);
                            }
                            cost += double(c);
                        }
                    }
                }
            }
            if (warp == 0) {
                double err = 0.0;
                if (ok) {
                    if (p.pred) p.pred[i] = float(arg_h);
                    if (p.err_sum) {
                        if (p.y_labels) err = (float(arg_h) != __ldg(p.y_labels + i)) ? 1.0 : 0.0;
                        else if (p.Y) err = (arg_h != arg_y) ? 1.0 : 0.0;
                    }
                }
                if (p.cost_sum || p.err_sum) {
                    for (int o = 16; o > 0; o >>= 1) {
                        cost += __shfl_xor_sync(0xffffffffu, cost, o);
                        err += __shfl_xor_sync(0xffffffffu, err, o);
                    }
                    if (lane == 0) { cost_blk += cost; err_blk += err; }
                }
            }
        }

        // ---------------- pass 2: delta_prev = (delta W) .* f'(a)   [+ G = delta^T [a,1]] ----------------
        if (TRAIN) {
            for (int h0 = 0; h0 < H1; h0 += HC) {
                const int hc = min(HC, H1 - h0);
                __syncthreads();                           // red[] / W_s reuse
                for (int idx = threadIdx.x; idx < K * hc; idx += TAIL_THREADS) {
                    const int k = idx / hc, hh = idx - k * hc;
                    W_s[k * (HC + 1) + hh] = __ldg(p.W + int64_t(k) * p.ldw + h0 + hh);
                }
                __syncthreads();
                const float* ap = p.a + int64_t(h0) * p.lda + ic;
#pragma unroll 4
                for (int hh = warp; hh < hc; hh += TAIL_WARPS) {
                    const int h = h0 + hh;
                    const float av = __ldg(ap + int64_t(hh) * p.lda);
                    if (h < p.H && p.delta_prev) {
                        float s = 0.f;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) if (k < K) s = fmaf(dl[k], W_s[k * (HC + 1) + hh], s);
                        const float d = (p.act_kind == B2NN_ACT_SIGMOID) ? av * (1.0f - av) : 1.0f - av * av;
                        if (ok) p.delta_prev[int64_t(h) * p.ldp + i] = s * d;
                    }
                    if (inline_g) {
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) {
                            if (k < K) {
                                float v = dl[k] * av;      // dl is 0 for out-of-range lanes
                                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                                if (lane == 0) G_s[k * H1 + h] += v;   // (k, h) is owned by exactly one warp: h % TAIL_WARPS == warp
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    if (inline_g) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < K * H1; idx += TAIL_THREADS) {
            const int k = idx / H1, h = idx - k * H1;
            const float v = G_s[idx];
            if (v != 0.f) atomicAdd(p.G_inline + int64_t(k) * p.ldw + h, v);
        }
    }
    if (threadIdx.x == 0) {
        if (p.cost_sum && cost_blk != 0.0) atomicAdd(p.cost_sum, cost_blk);
        if (p.err_sum && err_blk != 0.0) atomicAdd(p.err_sum, err_blk);
    }
}


// Small hidden layer (H+1 <= 512, K*(H+1) <= 8192 floats: configs C1, C2, C5): the whole W lives in shared memory and
// ONE WARP owns a group of 32 samples end to end — no block-level synchronisation per group (the kernel above spends
// its time in __syncthreads when each of its 16 warps only has three rows of a to read).  The second read of a hits L1.
constexpr int SMALL_THREADS = 128;
template <int KMAX, bool TRAIN>
__global__ void __launch_bounds__(SMALL_THREADS) tail_small_kernel(const TailArgs p) {
    pdl_enter();
    extern __shared__ float smem[];
    const int K = p.K, H1 = p.H + 1;
    float* W_s = smem;                    // [K][H1]
    float* G_s = W_s + K * H1;            // [K][H1] when the gradient is accumulated in-kernel
    const bool inline_g = TRAIN && p.G_inline != nullptr;
    for (int idx = threadIdx.x; idx < K * H1; idx += SMALL_THREADS) {
        const int k = idx / H1, h = idx - k * H1;
        W_s[idx] = __ldg(p.W + int64_t(k) * p.ldw + h);
        if (inline_g) G_s[idx] = 0.f;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t warp_global = int64_t(blockIdx.x) * (SMALL_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t n_warps = int64_t(gridDim.x) * (SMALL_THREADS / 32);
    const int64_t n_groups = (p.rows + 31) / 32;
    double cost = 0.0, err = 0.0;
    for (int64_t grp = warp_global; grp < n_groups; grp += n_warps) {
        const int64_t i = grp * 32 + lane;
        const bool ok = i < p.rows;
        const int64_t ic = ok ? i : p.rows - 1;
        const float* ap = p.a + ic;
        float z[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) z[k] = 0.f;
        int h = 0;
        for (; h + 8 <= H1; h += 8) {
            float av[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) av[u] = __ldg(ap + int64_t(h + u) * p.lda);
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int k = 0; k < KMAX; ++k) if (k < K) z[k] = fmaf(av[u], W_s[k * H1 + h + u], z[k]);
        }
        for (; h < H1; ++h) {
            const float av = __ldg(ap + int64_t(h) * p.lda);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) if (k < K) z[k] = fmaf(av, W_s[k * H1 + h], z[k]);
        }
        float zmax = -INFINITY, total = 0.f;
        if (p.out_kind == B2NN_OUT_SOFTMAX) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) if (k < K) zmax = fmaxf(zmax, z[k]);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) if (k < K) total += expf(z[k] - zmax);
        }
        float dl[KMAX];
        float best_h = 0.f, best_y = 0.f; int arg_h = 0, arg_y = 0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            dl[k] = 0.f;
            if (k < K) {
                float hv;
                if (p.out_kind == B2NN_OUT_SOFTMAX) hv = expf(z[k] - zmax) / total;
                else if (p.out_kind == B2NN_OUT_SIGMOID) hv = sigmoid_f(z[k]);
                else hv = z[k];
                if (hv > best_h) { best_h = hv; arg_h = k; }
                float yv = 0.f;
                if (p.Y) {
                    yv = __ldg(p.Y + int64_t(k) * p.ldy + ic);
                    if (yv > best_y) { best_y = yv; arg_y = k; }
                }
                const float dv = hv - yv;
                if (ok) {
                    dl[k] = dv;
                    if (p.h_out) p.h_out[int64_t(k) * p.ldh + i] = hv;
                    if (p.delta_last) p.delta_last[int64_t(k) * p.ldd + i] = dv;
                    if (p.cost_sum && p.Y) {
                        float c = 0.f;
                        if (p.cost_kind == B2NN_COST_SQUARED) c = dv * dv;
                        else {
                            if (yv != 0.f) c += -yv * logf(hv);
                            if (p.cost_kind == B2NN_COST_NEGLNMAX && yv != 1.f) c += -(1.f - yv) * logf(1.f - hv);
                            c = fabsf(c);
                        }
                        cost += double(c);
                    }
                }
            }
        }
        if (ok) {
            if (p.pred) p.pred[i] = float(arg_h);
            if (p.err_sum) {
                if (p.y_labels) err += (float(arg_h) != __ldg(p.y_labels + i)) ? 1.0 : 0.0;
                else if (p.Y) err += (arg_h != arg_y) ? 1.0 : 0.0;
            }
        }
        if (TRAIN) {
            for (int hh = 0; hh < H1; ++hh) {
                const float av = __ldg(ap + int64_t(hh) * p.lda);          // L1 hit: read in pass 1 by the same warp
                if (hh < p.H && p.delta_prev) {
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) if (k < K) s = fmaf(dl[k], W_s[k * H1 + hh], s);
                    const float d = (p.act_kind == B2NN_ACT_SIGMOID) ? av * (1.0f - av) : 1.0f - av * av;
                    if (ok) p.delta_prev[int64_t(hh) * p.ldp + i] = s * d;
                }
                if (inline_g) {
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) {
                        if (k < K) {
                            float v = dl[k] * av;
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            if (lane == 0) atomicAdd(&G_s[k * H1 + hh], v);
                        }
                    }
                }
            }
        }
    }
    if (p.cost_sum || p.err_sum) {
        for (int o = 16; o > 0; o >>= 1) {
            cost += __shfl_xor_sync(0xffffffffu, cost, o);
            err += __shfl_xor_sync(0xffffffffu, err, o);
        }
        if (lane == 0) {
            if (p.cost_sum && cost != 0.0) atomicAdd(p.cost_sum, cost);
            if (p.err_sum && err != 0.0) atomicAdd(p.err_sum, err);
        }
    }
    if (inline_g) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < K * H1; idx += SMALL_THREADS) {
            const int k = idx / H1, h = idx - k * H1;
            const float v = G_s[idx];
            if (v != 0.f) atomicAdd(p.G_inline + int64_t(k) * p.ldw + h, v);
        }
    }
}

// Training step of a small last layer (H+1 <= 512, K*(H+1) <= 2048: configs C1, C2).  The two kernels above are
// latency-bound there — each warp walks ~3000 dependent instructions for 32 samples (ncu: 33 us for C1's 4200 rows at
// 0.23 IPC).  Here a block stages its 32-sample tile of a in shared memory ONCE and every phase is spread over all
// eight warps with independent accumulators:
//   stage   a_s[h][s] (row pitch 33: conflict-free both along s and along h), Y prefetched by warp 0
//   z       warps split the K outputs                     -> z_s[k][s]
//   output  warp 0: activation, delta, cost, argmax       -> d_s[k][s]
//   dprev   warps split the H rows: (delta W) .* f'(a)    -> global
//   G       threads own (k,h) outputs: sum over the 32 samples from shared memory, kept in registers across the
//           block's groups, one atomic per output at the end (no warp shuffles)
constexpr int TILE_THREADS = 256;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr int TILE_PITCH = 33;
constexpr int TILE_GMAX = 8;               // (k,h) outputs per thread: K*(H+1) <= 2048
// Code size matters here: every warp walks the kernel once per group, so the first version (everything unrolled over
// KMAX = 16 and 8 x 32 gradient terms, 5.7 k SASS instructions) spent its 20 us fetching instructions.  Loops over k
// are rolled and work out of shared memory; GMAX is a template parameter (C1: 2 outputs per thread, C2: 1).
template <int GMAX>
__global__ void __launch_bounds__(TILE_THREADS) tail_tile_kernel(const TailArgs p) {
    pdl_enter();
    extern __shared__ float smem[];
    const int K = p.K, H1 = p.H + 1;
    float* W_s = smem;                                 // [K][H1]
    float* a_s = W_s + K * H1;                         // [H1][33]
    float* z_s = a_s + H1 * TILE_PITCH;                // [K][33]  z, then delta
    float* y_s = z_s + K * TILE_PITCH;                 // [K][33]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool inline_g = p.G_inline != nullptr;

    for (int idx = threadIdx.x; idx < K * H1; idx += TILE_THREADS) {
        const int k = idx / H1, h = idx - k * H1;
        W_s[idx] = __ldg(p.W + int64_t(k) * p.ldw + h);
    }
    // the (k,h) gradient outputs this thread owns, as shared-memory row offsets
    int off_d[GMAX], off_a[GMAX];
    float acc[GMAX];
#pragma unroll
    for (int j = 0; j < GMAX; ++j) {
        const int o = threadIdx.x + j * TILE_THREADS;
        const int k = o < K * H1 ? o / H1 : 0, h = o < K * H1 ? o - k * H1 : 0;
        off_d[j] = k * TILE_PITCH; off_a[j] = h * TILE_PITCH; acc[j] = 0.f;
    }
    double cost = 0.0, err = 0.0;

    const int64_t n_groups = (p.rows + 31) / 32;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t i = grp * 32 + lane;
        const bool ok = i < p.rows;
        const int64_t ic = ok ? i : p.rows - 1;
        // ---- stage the tile of a (4 rows in flight per warp) and of Y ----
        const float* ap = p.a + ic;
        int h = warp;
#pragma unroll 1
        for (; h + 3 * TILE_WARPS < H1; h += 4 * TILE_WARPS) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(ap + int64_t(h + u * TILE_WARPS) * p.lda);
#pragma unroll
            for (int u = 0; u < 4; ++u) a_s[(h + u * TILE_WARPS) * TILE_PITCH + lane] = v[u];
        }
#pragma unroll 1
        for (; h < H1; h += TILE_WARPS) a_s[h * TILE_PITCH + lane] = __ldg(ap + int64_t(h) * p.lda);
#pragma unroll 1
        for (int k = warp; k < K; k += TILE_WARPS) y_s[k * TILE_PITCH + lane] = p.Y ? __ldg(p.Y + int64_t(k) * p.ldy + ic) : 0.f;
        __syncthreads();
        // ---- z = [a,1] W^T ----
#pragma unroll 1
        for (int k = warp; k < K; k += TILE_WARPS) {
            float z0 = 0.f, z1 = 0.f;
            const float* wk = W_s + k * H1;
            int hh = 0;
#pragma unroll 4
            for (; hh + 1 < H1; hh += 2) {
                z0 = fmaf(a_s[hh * TILE_PITCH + lane], wk[hh], z0);
                z1 = fmaf(a_s[(hh + 1) * TILE_PITCH + lane], wk[hh + 1], z1);
            }
            if (hh < H1) z0 = fmaf(a_s[hh * TILE_PITCH + lane], wk[hh], z0);
            z_s[k * TILE_PITCH + lane] = z0 + z1;
        }
        __syncthreads();
        // ---- output activation, delta, cost, argmax: one lane per sample, rolled over k ----
        if (warp == 0) {
            float zmax = -INFINITY, total = 1.f;
            if (p.out_kind == B2NN_OUT_SOFTMAX) {
#pragma unroll 1
                for (int k = 0; k < K; ++k) zmax = fmaxf(zmax, z_s[k * TILE_PITCH + lane]);
                total = 0.f;
#pragma unroll 1
                for (int k = 0; k < K; ++k) { const float e = expf(z_s[k * TILE_PITCH + lane] - zmax); z_s[k * TILE_PITCH + lane] = e; total += e; }
            }
            float best_h = 0.f, best_y = 0.f; int arg_h = 0, arg_y = 0;
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                const float zv = z_s[k * TILE_PITCH + lane], yv = y_s[k * TILE_PITCH + lane];
                float hv;
                if (p.out_kind == B2NN_OUT_SOFTMAX) hv = zv / total;
                else if (p.out_kind == B2NN_OUT_SIGMOID) hv = sigmoid_f(zv);
                else hv = zv;
                if (hv > best_h) { best_h = hv; arg_h = k; }
                if (p.Y && yv > best_y) { best_y = yv; arg_y = k; }
                const float dv = hv - yv;
                z_s[k * TILE_PITCH + lane] = ok ? dv : 0.f;
                if (ok) {
                    if (p.h_out) p.h_out[int64_t(k) * p.ldh + i] = hv;
                    if (p.delta_last) p.delta_last[int64_t(k) * p.ldd + i] = dv;
                    if (p.cost_sum && p.Y) {
                        float c = 0.f;
                        if (p.cost_kind == B2NN_COST_SQUARED) c = dv * dv;
                        else {
                            if (yv != 0.f) c += -yv * logf(hv);
                            if (p.cost_kind == B2NN_COST_NEGLNMAX && yv != 1.f) c += -(1.f - yv) * logf(1.f - hv);
                            c = fabsf(c);
                        }
                        cost += double(c);
                    }
                }
            }
            if (ok) {
                if (p.pred) p.pred[i] = float(arg_h);
                if (p.err_sum) {
                    if (p.y_labels) err += (float(arg_h) != __ldg(p.y_labels + i)) ? 1.0 : 0.0;
                    else if (p.Y) err += (arg_h != arg_y) ? 1.0 : 0.0;
                }
            }
        }
        __syncthreads();
        // ---- delta_prev = (delta W)[no bias] .* f'(a) ----
        if (p.delta_prev) {
#pragma unroll 1
            for (int hh = warp; hh < p.H; hh += TILE_WARPS) {
                float s0 = 0.f, s1 = 0.f;
                int k = 0;
#pragma unroll 2
                for (; k + 1 < K; k += 2) {
                    s0 = fmaf(z_s[k * TILE_PITCH + lane], W_s[k * H1 + hh], s0);
                    s1 = fmaf(z_s[(k + 1) * TILE_PITCH + lane], W_s[(k + 1) * H1 + hh], s1);
                }
                if (k < K) s0 = fmaf(z_s[k * TILE_PITCH + lane], W_s[k * H1 + hh], s0);
                const float av = a_s[hh * TILE_PITCH + lane];
                const float d = (p.act_kind == B2NN_ACT_SIGMOID) ? av * (1.0f - av) : 1.0f - av * av;
                if (ok) p.delta_prev[int64_t(hh) * p.ldp + i] = (s0 + s1) * d;
            }
        }
        // ---- G(k,h) += sum_s delta(k,s) a(h,s) ----
        if (inline_g) {
#pragma unroll
            for (int j = 0; j < GMAX; ++j) {
                if (threadIdx.x + j * TILE_THREADS < K * H1) {
                    const float* dp = z_s + off_d[j];
                    const float* aq = a_s + off_a[j];
                    float g0 = 0.f, g1 = 0.f;
#pragma unroll 4
                    for (int t = 0; t < 32; t += 2) { g0 = fmaf(dp[t], aq[t], g0); g1 = fmaf(dp[t + 1], aq[t + 1], g1); }
                    acc[j] += g0 + g1;
                }
            }
        }
        __syncthreads();                                   // the next group overwrites a_s / z_s / y_s
    }

    if (inline_g) {
#pragma unroll
        for (int j = 0; j < GMAX; ++j) {
            const int o = threadIdx.x + j * TILE_THREADS;
            if (o < K * H1 && acc[j] != 0.f) {
                const int k = o / H1, h = o - k * H1;
                atomicAdd(p.G_inline + int64_t(k) * p.ldw + h, acc[j]);
            }
        }
    }
    if ((p.cost_sum || p.err_sum) && warp == 0) {
        for (int o = 16; o > 0; o >>= 1) {
            cost += __shfl_xor_sync(0xffffffffu, cost, o);
            err += __shfl_xor_sync(0xffffffffu, err, o);
        }
        if (lane == 0) {
            if (p.cost_sum && cost != 0.0) atomicAdd(p.cost_sum, cost);
            if (p.err_sum && err != 0.0) atomicAdd(p.err_sum, err);
        }
    }
}

template <int GMAX>
cudaError_t launch_tile(const TailArgs& a, unsigned grid, size_t smem, cudaStream_t s) {
    static std::atomic<unsigned long long> done{0};
        if (cudaError_t e = ensure_dyn_smem(tail_tile_kernel<GMAX>, 100 * 1024, done)) return e;
    return launch_pdl(tail_tile_kernel<GMAX>, dim3(grid), dim3(TILE_THREADS), smem, s, a);
}

template <int KMAX>
cudaError_t launch_small(const TailArgs& a, bool train, unsigned grid, size_t smem, cudaStream_t s) {
    if (train) {
        static std::atomic<unsigned long long> done{0};
        if (cudaError_t e = ensure_dyn_smem(tail_small_kernel<KMAX, true>, 100 * 1024, done)) return e;
        return launch_pdl(tail_small_kernel<KMAX, true>, dim3(grid), dim3(SMALL_THREADS), smem, s, a);
    } else {
        static std::atomic<unsigned long long> done{0};
        if (cudaError_t e = ensure_dyn_smem(tail_small_kernel<KMAX, false>, 100 * 1024, done)) return e;
        return launch_pdl(tail_small_kernel<KMAX, false>, dim3(grid), dim3(SMALL_THREADS), smem, s, a);
    }
}

template <int KMAX>
cudaError_t launch_k(const TailArgs& a, bool train, unsigned grid, size_t smem, cudaStream_t s) {
    if (train) {
        static std::atomic<unsigned long long> done{0};
        if (cudaError_t e = ensure_dyn_smem(tail_kernel<KMAX, true>, 200 * 1024, done)) return e;
        return launch_pdl(tail_kernel<KMAX, true>, dim3(grid), dim3(TAIL_THREADS), smem, s, a);
    } else {
        static std::atomic<unsigned long long> done{0};
        if (cudaError_t e = ensure_dyn_smem(tail_kernel<KMAX, false>, 200 * 1024, done)) return e;
        return launch_pdl(tail_kernel<KMAX, false>, dim3(grid), dim3(TAIL_THREADS), smem, s, a);
    }
}

}  // namespace

bool tail_supported(int K, int H, bool want_inline_g) {
    if (K < 1 || K > 32) return false;
    // Measured on config C3 (H = 4096, 8192 rows): 344 us against 154 us for the three separate kernels it replaces —
    // with a wide hidden layer the two sequential passes over a are latency bound.  Small nets (C1, C2, C5) gain.
    if (H > 1024) return false;
    if (want_inline_g && int64_t(K) * (H + 1) > 4096) return false;
    return true;
}

cudaError_t tail_launch(const TailArgs& a, bool train, cudaStream_t s) {
    if (a.rows <= 0) return cudaSuccess;
    const int kmax = a.K <= 4 ? 4 : a.K <= 16 ? 16 : 32;
    // Forward-only calls take the warp-per-group kernel (C5 at 2^20 rows: 1121 -> 335 us).  Training keeps the
    // block-cooperative kernel: its in-kernel gradient reduction is spread over 16 warps instead of serialised in one
    // (C1: 28 us against 142 us).
    if (train && a.H + 1 <= 512 && int64_t(a.K) * (a.H + 1) <= TILE_GMAX * TILE_THREADS && !std::getenv("B2NN_NO_TAIL_TILE")) {
        const int H1 = a.H + 1;
        const size_t sm = sizeof(float) * (size_t(a.K) * H1 + size_t(H1) * TILE_PITCH + 2 * size_t(a.K) * TILE_PITCH);
        const int64_t grp = (a.rows + 31) / 32;
        const unsigned g = unsigned(grp < 148 * 2 ? grp : 148 * 2);
        const int per_thread = int((int64_t(a.K) * H1 + TILE_THREADS - 1) / TILE_THREADS);
        if (per_thread <= 1) return launch_tile<1>(a, g, sm, s);
        if (per_thread <= 2) return launch_tile<2>(a, g, sm, s);
        if (per_thread <= 4) return launch_tile<4>(a, g, sm, s);
        return launch_tile<8>(a, g, sm, s);
    }
    if (!train && a.H + 1 <= 512 && int64_t(a.K) * (a.H + 1) <= 8192) {
        const bool ig = train && a.G_inline != nullptr;
        const size_t sm = sizeof(float) * size_t(a.K) * (a.H + 1) * (ig ? 2 : 1);
        const int64_t grp = (a.rows + 31) / 32;
        const int64_t blocks = (grp + SMALL_THREADS / 32 - 1) / (SMALL_THREADS / 32);
        const int64_t capb = ig ? 148 * 2 : 148 * 16;
        const unsigned g = unsigned(blocks < capb ? blocks : capb);
        switch (kmax) {
            case 4:  return launch_small<4>(a, train, g, sm, s);
            case 16: return launch_small<16>(a, train, g, sm, s);
            default: return launch_small<32>(a, train, g, sm, s);
        }
    }
    size_t smem = sizeof(float) * (size_t(kmax) * (HC + 1) + size_t(TAIL_WARPS) * kmax * 32);
    const bool inline_g = train && a.G_inline != nullptr;
    if (inline_g) smem += sizeof(float) * size_t(a.K) * (a.H + 1);
    const int64_t groups = (a.rows + 31) / 32;
    // with block-local gradient accumulation fewer, longer-lived blocks mean fewer atomic flushes
    const int64_t cap = inline_g ? 148 : 148 * 4;
    const unsigned grid = unsigned(groups < cap ? groups : cap);
    switch (kmax) {
        case 4:  return launch_k<4>(a, train, grid, smem, s);
        case 16: return launch_k<16>(a, train, grid, smem, s);
        default: return launch_k<32>(a, train, grid, smem, s);
    }
}

}  // namespace b2
