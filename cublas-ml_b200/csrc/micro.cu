// micro.cu — a whole training step (minus the update) of a TINY network as one kernel.
//
// Config C2 (bikeshare, 10-40-160-1: 7 161 weights, 4 337 rows) spends its step in seven dependent launches of a few
// microseconds each; the reference spends 23 (cublasNN.cpp:526-579, 642-685).  When every weight of the net fits in
// shared memory the chain disappears: a block keeps the weights resident, takes 32 samples at a time (lane = sample)
// through every layer forward (activations stay in shared memory), applies the output activation / cost / delta, walks
// back through the layers (delta backprop and the gradient of each layer while its delta is live), and accumulates the
// gradient of ALL layers in shared memory across the samples it owns.  One atomic flush per block at the end; the
// momentum update (which also re-zeroes the gradient arena) is the only other launch of the step.
//
//   forward   z_j = sum_k W[j][k] a[k][s]       warps split the output units in blocks of 4 (one read of a per 4 FMA)
//   backward  d_k = f'(a_k) sum_j W[j][k] d[j]  warps split the input units in blocks of 4 (W row chunks are one 16-byte
//                                                broadcast read)
//   gradient  G[j][k] += sum_s d[j][s] a[k][s]  threads own (j,k) pairs: conflict-free 16-byte reads (row pitch 36), no shuffles
// Exact fp32.  Same arithmetic per element as the separate kernels; only the summation order over samples differs.
#include "engine.h"
#include <cstdlib>
#include "pdl.cuh"
#include "epilogue.cuh"
#include "../../include/b2nn.h"

namespace b2 {
namespace {

constexpr int MC_THREADS = 512;             // measured on C2: 256 / 512 / 1024 threads -> 40.8 / 32.8 / 34.8 us per step
constexpr int MC_WARPS = MC_THREADS / 32;
constexpr int MC_PITCH = MICRO_PITCH;  // 36 floats per row: 16-byte aligned rows, conflict-free for lane = sample and for
                                       // 16-byte reads with lane = row (bank = 4 row + s mod 32)

__device__ __forceinline__ float act_fwd(int kind, float z) {
    return kind == B2NN_ACT_SIGMOID ? sigmoid_f(z) : tanhf(z);
}
__device__ __forceinline__ float act_grad_from_a(int kind, float a) {
    return kind == B2NN_ACT_SIGMOID ? a * (1.0f - a) : 1.0f - a * a;
}

__global__ void __launch_bounds__(MC_THREADS) micro_mlp_kernel(const MicroArgs p) {
    pdl_enter();
    extern __shared__ __align__(16) float mc_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = p.nw;
    float* W_s = mc_smem;                          // the internal weight arena, as is
    float* G_s = W_s + p.total;                    // gradient of every layer, same layout
    float* act_s = G_s + p.total;                  // inputs of layer l at act_s + p.act_off[l]: (in_l + 1) rows of 33
    float* d_s = act_s + p.act_total;              // two delta buffers of max_width rows
    const int dbuf = p.max_width * MC_PITCH;

    for (int i = threadIdx.x; i < p.total; i += MC_THREADS) { W_s[i] = __ldg(p.W + i); G_s[i] = 0.f; }
    // the ones rows of the hidden activations never change
    for (int l = 1; l < nw; ++l)
        if (threadIdx.x < 32) act_s[p.act_off[l] + p.in[l] * MC_PITCH + lane] = 1.0f;
    double cost = 0.0;
    __syncthreads();

    const int64_t n_groups = (p.rows + 31) / 32;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t i = grp * 32 + lane;
        const bool ok = i < p.rows;
        const int64_t ic = ok ? i : p.rows - 1;
        // ---- stage the input tile (ones row included: X carries it) ----
        {
            const float* xp = p.X + ic;
            float* a0 = act_s + p.act_off[0];
            for (int r = warp; r <= p.in[0]; r += MC_WARPS) a0[r * MC_PITCH + lane] = __ldg(xp + int64_t(r) * p.ldx);
            // targets ride in delta buffer 1, which the backward pass only writes after the output phase has read them
            float* y_s = d_s + dbuf;
            for (int k = warp; k < p.out[nw - 1]; k += MC_WARPS) y_s[k * MC_PITCH + lane] = __ldg(p.Y + int64_t(k) * p.ldy + ic);
        }
        __syncthreads();
        // ---- forward ----
        for (int l = 0; l < nw; ++l) {
            const float* Wl = W_s + p.w_off[l];
            const int ldw = p.ldw[l], in1 = p.in[l] + 1, out = p.out[l];
            const float* al = act_s + p.act_off[l];
            const bool last = l == nw - 1;
            float* zout = last ? d_s : act_s + p.act_off[l + 1];       // last layer: raw z into delta buffer 0
            for (int j0 = warp * 4; j0 < out; j0 += MC_WARPS * 4) {
                float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
                const float* w0 = Wl + (j0 + 0) * ldw;
                const float* w1 = Wl + min(j0 + 1, out - 1) * ldw;
                const float* w2 = Wl + min(j0 + 2, out - 1) * ldw;
                const float* w3 = Wl + min(j0 + 3, out - 1) * ldw;
#pragma unroll 4
                for (int k = 0; k < in1; ++k) {
                    const float a = al[k * MC_PITCH + lane];
                    z0 = fmaf(w0[k], a, z0); z1 = fmaf(w1[k], a, z1); z2 = fmaf(w2[k], a, z2); z3 = fmaf(w3[k], a, z3);
                }
                const float zz[4] = {z0, z1, z2, z3};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (j0 + q < out) zout[(j0 + q) * MC_PITCH + lane] = last ? zz[q] : act_fwd(p.act_kind, zz[q]);
            }
            __syncthreads();
        }
        // ---- output activation, delta of the last layer, cost: one lane per sample (as tail.cu) ----
        const int K = p.out[nw - 1];
        if (warp == 0) {
            float zmax = -INFINITY, total = 1.f;
            if (p.out_kind == B2NN_OUT_SOFTMAX) {
                for (int k = 0; k < K; ++k) zmax = fmaxf(zmax, d_s[k * MC_PITCH + lane]);
                total = 0.f;
                for (int k = 0; k < K; ++k) { const float e = expf(d_s[k * MC_PITCH + lane] - zmax); d_s[k * MC_PITCH + lane] = e; total += e; }
            }
            for (int k = 0; k < K; ++k) {
                const float zv = d_s[k * MC_PITCH + lane];
                const float yv = d_s[dbuf + k * MC_PITCH + lane];
                float hv;
                if (p.out_kind == B2NN_OUT_SOFTMAX) hv = zv / total;
                else if (p.out_kind == B2NN_OUT_SIGMOID) hv = sigmoid_f(zv);
                else hv = zv;
                const float dv = hv - yv;
                d_s[k * MC_PITCH + lane] = ok ? dv : 0.f;
                if (ok && p.cost_sum) {
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
        __syncthreads();
        // ---- backward: gradient of layer l and delta of layer l-1 while delta_l is live ----
        int cur = 0;
        for (int l = nw - 1; l >= 0; --l) {
            const float* Wl = W_s + p.w_off[l];
            float* Gl = G_s + p.w_off[l];
            const int ldw = p.ldw[l], in = p.in[l], in1 = in + 1, out = p.out[l];
            const float* al = act_s + p.act_off[l];
            const float* dl = d_s + cur * dbuf;
            for (int q = threadIdx.x; q < out * in1; q += MC_THREADS) {
                const int j = q / in1, k = q - j * in1;
                const float* dp = dl + j * MC_PITCH;
                const float* ap = al + k * MC_PITCH;
                float g0 = 0.f, g1 = 0.f;
#pragma unroll
                for (int s = 0; s < 32; s += 4) {
                    const float4 dv = *reinterpret_cast<const float4*>(dp + s);
                    const float4 av = *reinterpret_cast<const float4*>(ap + s);
                    g0 = fmaf(dv.x, av.x, g0); g1 = fmaf(dv.y, av.y, g1); g0 = fmaf(dv.z, av.z, g0); g1 = fmaf(dv.w, av.w, g1);
                }
                Gl[j * ldw + k] += g0 + g1;
            }
            if (l > 0) {
                float* dn = d_s + (cur ^ 1) * dbuf;
                for (int k0 = warp * 4; k0 < in; k0 += MC_WARPS * 4) {
                    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
                    for (int j = 0; j < out; ++j) {
                        const float4 w = *reinterpret_cast<const float4*>(Wl + j * ldw + k0);     // ldw % 4 == 0: in-row, aligned
                        const float dv = dl[j * MC_PITCH + lane];
                        t0 = fmaf(w.x, dv, t0); t1 = fmaf(w.y, dv, t1); t2 = fmaf(w.z, dv, t2); t3 = fmaf(w.w, dv, t3);
                    }
                    const float tt[4] = {t0, t1, t2, t3};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (k0 + q < in) dn[(k0 + q) * MC_PITCH + lane] = tt[q] * act_grad_from_a(p.act_kind, al[(k0 + q) * MC_PITCH + lane]);
                }
            }
            __syncthreads();
            cur ^= 1;
        }
    }
    // ---- flush: one atomic per weight per block ----
    for (int i = threadIdx.x; i < p.total; i += MC_THREADS) {
        const float v = G_s[i];
        if (v != 0.f) atomicAdd(p.G + i, v);
    }
    if (p.cost_sum && warp == 0) {
        for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
        if (lane == 0 && cost != 0.0) atomicAdd(p.cost_sum, cost);
    }
}

}  // namespace

size_t micro_smem_bytes(const MicroArgs& a) {
    return sizeof(float) * (2 * size_t(a.total) + size_t(a.act_total) + 2 * size_t(a.max_width) * MC_PITCH);
}

bool micro_supported(const MicroArgs& a) {
    if (std::getenv("B2NN_NO_MICRO")) return false;
    if (a.nw < 2 || a.nw > MICRO_MAX_LAYERS || a.total > 16384) return false;
    if (a.out[a.nw - 1] > 32) return false;
    return micro_smem_bytes(a) <= 200 * 1024;
}

// The caller has zeroed G (or knows it is zero).
cudaError_t micro_launch(const MicroArgs& a, cudaStream_t s) {
    if (a.rows <= 0) return cudaSuccess;
    const size_t smem = micro_smem_bytes(a);
    static std::atomic<unsigned long long> done{0};
    if (cudaError_t e = ensure_dyn_smem(micro_mlp_kernel, 200 * 1024, done)) return e;
    const int64_t groups = (a.rows + 31) / 32;
    const unsigned grid = unsigned(groups < 148 ? groups : 148);
    return launch_pdl(micro_mlp_kernel, dim3(grid), dim3(MC_THREADS), smem, s, a);
}

}  // namespace b2
