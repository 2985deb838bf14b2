// elementwise.cu — the bandwidth-bound kernels around the contractions: output layer (+ cost, argmax, error count),
// momentum update, layout conversion at the ABI edge, weight-init rescale.
#include "engine.h"
#include "pdl.cuh"
#include "epilogue.cuh"
#include "../../include/b2nn.h"

#include <cuda_bf16.h>

namespace b2 {
namespace {

constexpr int OUT_THREADS = 128;

__device__ __forceinline__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += sh[i];
    __syncthreads();
    return t;   // valid in thread 0
}

// One thread per sample, like kernelSoftmax (activations.cu:68-88): lanes walk consecutive rows so every access
// z[k*ld + i] is coalesced.  Fuses what the reference does in up to five launches:
//   output activation (activations.cu:9-12, 68-88), delta = h - y (kernels.cu:46-51),
//   per-element cost + Sasum/Sdot (costfunctions.cu:14-26, cublasNN.cpp:552-554, 613-614),
//   argmax (kernels.cu:74-86, with the out-of-bounds store fixed) and error count (kernels.cu:88-93).
// Softmax is max-shifted: identical to the reference whenever the reference does not overflow.
// cost_sum / err_sum are double accumulators.
// One sample (row i).  KMAX > 0: K <= KMAX, z and y are read into registers with every load in flight at once (with
// run-time loop bounds the loads went out one per iteration: 2 K dependent L2 round trips, ~12 us for 8192 rows).
template <int KMAX>
__device__ __forceinline__ void output_row(const OutputArgs& a, int64_t i, bool want_cost, bool want_err, double& cost, double& err) {
    const int K = a.K;
    float zv[KMAX], yv[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        zv[k] = k < K ? a.z[int64_t(k) * a.ldz + i] : -INFINITY;
        yv[k] = (k < K && a.y) ? a.y[int64_t(k) * a.ldy + i] : 0.f;
    }
    float zmax = -INFINITY, total = 0.f;
    if (a.out_kind == B2NN_OUT_SOFTMAX) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) zmax = fmaxf(zmax, zv[k]);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K) { zv[k] = expf(zv[k] - zmax); total += zv[k]; }
    }
    float best_h = 0.f, best_y = 0.f; int arg_h = 0, arg_y = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            float h;
            if (a.out_kind == B2NN_OUT_SOFTMAX) h = zv[k] / total;
            else if (a.out_kind == B2NN_OUT_SIGMOID) h = sigmoid_f(zv[k]);
            else h = zv[k];
            if (a.h) a.h[int64_t(k) * a.ldh + i] = h;
            if (h > best_h) { best_h = h; arg_h = k; }
            if (a.y) {
                if (yv[k] > best_y) { best_y = yv[k]; arg_y = k; }
                const float d = h - yv[k];
                if (a.delta) a.delta[int64_t(k) * a.ldd + i] = d;
                if (want_cost) {
                    float c = 0.f;
                    if (a.cost_kind == B2NN_COST_SQUARED) c = d * d;
                    else {
                        if (yv[k] != 0.f) c += -yv[k] * logf(h);
                        if (a.cost_kind == B2NN_COST_NEGLNMAX && yv[k] != 1.f) c += -(1.f - yv[k]) * logf(1.f - h);
                        c = fabsf(c);
                    }
                    cost += double(c);
                }
            }
        }
    }
    if (a.pred) a.pred[i] = float(arg_h);
    if (want_err) {
        if (a.y_labels) err = (float(arg_h) != a.y_labels[i]) ? 1.0 : 0.0;
        else if (a.y)   err = (arg_h != arg_y) ? 1.0 : 0.0;
    }
}

// Any K: the same row, walking memory (the generic path the register version was derived from).
__device__ __forceinline__ void output_row_any(const OutputArgs& a, int64_t i, bool want_cost, bool want_err, double& cost, double& err) {
    const int K = a.K;
    float zmax = -INFINITY, total = 0.f;
    if (a.out_kind == B2NN_OUT_SOFTMAX) {
        for (int k = 0; k < K; ++k) zmax = fmaxf(zmax, a.z[int64_t(k) * a.ldz + i]);
        for (int k = 0; k < K; ++k) total += expf(a.z[int64_t(k) * a.ldz + i] - zmax);
    }
    float best_h = 0.f, best_y = 0.f; int arg_h = 0, arg_y = 0;
    for (int k = 0; k < K; ++k) {
        const float zv = a.z[int64_t(k) * a.ldz + i];
        float h;
        if (a.out_kind == B2NN_OUT_SOFTMAX) h = expf(zv - zmax) / total;
        else if (a.out_kind == B2NN_OUT_SIGMOID) h = sigmoid_f(zv);
        else h = zv;
        if (a.h) a.h[int64_t(k) * a.ldh + i] = h;
        if (h > best_h) { best_h = h; arg_h = k; }
        if (a.y) {
            const float yv = a.y[int64_t(k) * a.ldy + i];
            if (yv > best_y) { best_y = yv; arg_y = k; }
            const float d = h - yv;
            if (a.delta) a.delta[int64_t(k) * a.ldd + i] = d;
            if (want_cost) {
                float c = 0.f;
                if (a.cost_kind == B2NN_COST_SQUARED) c = d * d;
                else {
                    if (yv != 0.f) c += -yv * logf(h);
                    if (a.cost_kind == B2NN_COST_NEGLNMAX && yv != 1.f) c += -(1.f - yv) * logf(1.f - h);
                    c = fabsf(c);
                }
                cost += double(c);
            }
        }
    }
    if (a.pred) a.pred[i] = float(arg_h);
    if (want_err) {
        if (a.y_labels) err = (float(arg_h) != a.y_labels[i]) ? 1.0 : 0.0;
        else if (a.y)   err = (arg_h != arg_y) ? 1.0 : 0.0;
    }
}

__global__ void __launch_bounds__(OUT_THREADS) output_layer_kernel(OutputArgs a, double* cost_sum, double* err_sum) {
    pdl_enter();
    __shared__ double sh[OUT_THREADS / 32];
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    double cost = 0.0, err = 0.0;
    if (i < a.rows) {
        if (a.K <= 16) output_row<16>(a, i, cost_sum != nullptr, err_sum != nullptr, cost, err);
        else output_row_any(a, i, cost_sum != nullptr, err_sum != nullptr, cost, err);
    }
    if (cost_sum) { const double t = block_sum(cost, sh); if (threadIdx.x == 0 && t != 0.0) atomicAdd(cost_sum, t); }
    if (err_sum)  { const double t = block_sum(err, sh);  if (threadIdx.x == 0 && t != 0.0) atomicAdd(err_sum, t); }
}

// v <- mu*v + alpha*g ; theta <- theta + v.  20 bytes per weight (read v, g, theta; write v, theta) against the
// reference's two cublasSgeam passes at 24 bytes per weight (cublasNN.cpp:558-566, 619-627).  The gradient is already
// in theta's layout, so the Sgeam transpose disappears.  128-bit accesses, grid-stride, one pass.
template <bool ZERO_G>
__global__ void __launch_bounds__(256) momentum_update_kernel(float4* __restrict__ theta, float4* __restrict__ vel,
                                                              float4* __restrict__ grad, int64_t n4, float mu, float alpha) {
    pdl_enter();
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = vel[i];
        const float4 g = __ldcs(grad + i);
        float4 t = theta[i];
        v.x = fmaf(mu, v.x, alpha * g.x); v.y = fmaf(mu, v.y, alpha * g.y);
        v.z = fmaf(mu, v.z, alpha * g.z); v.w = fmaf(mu, v.w, alpha * g.w);
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        vel[i] = v;
        theta[i] = t;
        if (ZERO_G) grad[i] = make_float4(0.f, 0.f, 0.f, 0.f);     // small nets: next step's split-K / in-kernel sums need no memset
    }
}
__global__ void momentum_update_tail_kernel(float* theta, float* vel, const float* grad, int64_t from, int64_t count,
                                            float mu, float alpha) {
    pdl_enter();
    const int64_t i = from + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) {
        const float v = fmaf(mu, vel[i], alpha * grad[i]);
        vel[i] = v;
        theta[i] += v;
    }
}

__global__ void theta_ref_to_internal_kernel(const float* __restrict__ ref, float* __restrict__ w, int in, int out,
                                             int64_t ldw) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t per = int64_t(in) + 1;
    if (idx >= per * out) return;
    const int64_t o = idx / per, r = idx % per;          // r: reference row (0 = bias)
    const int64_t c = (r == 0) ? in : r - 1;             // internal column (bias last)
    w[o * ldw + c] = ref[idx];
}
__global__ void theta_internal_to_ref_kernel(const float* __restrict__ w, float* __restrict__ ref, int in, int out,
                                             int64_t ldw) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t per = int64_t(in) + 1;
    if (idx >= per * out) return;
    const int64_t o = idx / per, r = idx % per;
    const int64_t c = (r == 0) ? in : r - 1;
    ref[idx] = w[o * ldw + c];
}

__global__ void stage_x_kernel(const float* __restrict__ x_ref, int64_t ld_ref, float* __restrict__ x_int,
                               int64_t ld_int, int64_t rows, int n) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;                             // internal feature row; f == n is the ones row
    if (i >= rows) return;
    x_int[int64_t(f) * ld_int + i] = (f == n) ? 1.0f : x_ref[int64_t(f + 1) * ld_ref + i];
}

// vecVecElementMultiplyGPU (kernels.cu:53-58); only the custom-activation path still needs it un-fused.
__global__ void multiply_kernel(const float* a, const float* b, float* out, int64_t n) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
}

__global__ void split_batches_kernel(const float* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst,
                                     int64_t rows, int64_t bs, int64_t bs_pad, int nb) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int64_t b = i / bs;
    if (b > nb - 1) b = nb - 1;
    const int64_t r = i - b * bs;
    const int64_t f = blockIdx.y;
    dst[f * ld_dst + b * bs_pad + r] = src[f * ld_src + i];
}

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(256) split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, int64_t count) {
    pdl_enter();
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) lo[i] = tf32_lo(x[i]);
}

__global__ void __launch_bounds__(256) momentum_update_lo_kernel(float* __restrict__ theta, float* __restrict__ vel,
                                                                 const float* __restrict__ grad, float* __restrict__ theta_lo,
                                                                 int64_t count, float mu, float alpha) {
    pdl_enter();
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float v = fmaf(mu, vel[i], alpha * grad[i]);
        const float t = theta[i] + v;
        vel[i] = v;
        theta[i] = t;
        theta_lo[i] = tf32_lo(t);
    }
}

// bf16 copies (round to nearest even) of an fp32 array: the operands of the bf16 tensor-core mode.
__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ h, int64_t count) {
    pdl_enter();
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) h[i] = __float2bfloat16_rn(x[i]);
}
// 22 B per weight: read v, g, theta (12), write v, theta (8) and the bf16 copy (2); four weights per thread and access.
__global__ void __launch_bounds__(256) momentum_update_bf16_kernel(float* __restrict__ theta, float* __restrict__ vel,
                                                                   const float* __restrict__ grad, __nv_bfloat16* __restrict__ theta_h,
                                                                   int64_t count, float mu, float alpha) {
    pdl_enter();
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    const int64_t n4 = count >> 2;
    float4* th4 = reinterpret_cast<float4*>(theta);
    float4* v4 = reinterpret_cast<float4*>(vel);
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    uint2* h4 = reinterpret_cast<uint2*>(theta_h);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = v4[i], t = th4[i];
        const float4 g = __ldcs(g4 + i);
        v.x = fmaf(mu, v.x, alpha * g.x); v.y = fmaf(mu, v.y, alpha * g.y);
        v.z = fmaf(mu, v.z, alpha * g.z); v.w = fmaf(mu, v.w, alpha * g.w);
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        v4[i] = v;
        th4[i] = t;
        const __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
        h4[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
    for (int64_t i = (n4 << 2) + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float v = fmaf(mu, vel[i], alpha * grad[i]);
        const float t = theta[i] + v;
        vel[i] = v;
        theta[i] = t;
        theta_h[i] = __float2bfloat16_rn(t);
    }
}

__global__ void __launch_bounds__(256) abs_sum_kernel(const float* __restrict__ J, int64_t ld, int64_t rows, int K, double* out) {
    __shared__ double sh[8];
    double s = 0.0;
    const int64_t total = rows * K, stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t k = e / rows, i = e - k * rows;
        s += double(fabsf(J[k * ld + i]));
    }
    const double t = block_sum(s, sh);
    if (threadIdx.x == 0 && t != 0.0) atomicAdd(out, t);
}

__global__ void fill_kernel(float* p, float v, int64_t count) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) p[i] = v;
}

// theta = u*2*eps - eps (randinitweights.cu:13-31).  nvcc emits FADD (u+u) + FFMA for the reference expression, and
// IEEE sqrt / division; the same operations are spelled out here so theta_0 is bit-identical.
__global__ void init_scale_kernel(float* theta, int in, int out, int64_t count, int kind) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float eps = (kind == B2NN_INIT_1) ? __fdiv_rn(1.0f, __fsqrt_rn(float(in)))
                                            : __fdiv_rn(__fsqrt_rn(6.0f), __fsqrt_rn(float(in + out)));
    theta[i] = __fmaf_rn(__fadd_rn(theta[i], theta[i]), eps, -eps);
}

// classToBin (cublasNN.cpp:135-146) on the device, integer indexing (Q10).
__global__ void labels_to_onehot_kernel(const float* __restrict__ labels, float* __restrict__ y, int64_t ldy,
                                        int64_t rows, int K) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int c = int(labels[i]);
    for (int k = 0; k < K; ++k) y[int64_t(k) * ldy + i] = (k == c) ? 1.0f : 0.0f;
}

// normaliseData (cublasNN.cpp:214-268) on the device: per-feature mean, then the (m-1)-normalised standard deviation of
// the deviations from that mean — the reference's two passes, accumulated in double instead of its sequential float sum.
// One block per feature row of the engine-layout matrix (samples contiguous).
__global__ void __launch_bounds__(256) feature_stats_kernel(const float* __restrict__ x, int64_t ld, int64_t rows,
                                                            float* __restrict__ mean, float* __restrict__ stddev) {
    __shared__ double sh[8];
    __shared__ float mu_sh;
    const float* row = x + int64_t(blockIdx.x) * ld;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) s += double(row[i]);
    const double tot = block_sum(s, sh);
    if (threadIdx.x == 0) mu_sh = float(tot / double(rows));
    __syncthreads();
    const float mu = mu_sh;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) { const float d = row[i] - mu; q += double(d) * double(d); }
    const double qt = block_sum(q, sh);
    if (threadIdx.x == 0) { mean[blockIdx.x] = mu; stddev[blockIdx.x] = float(sqrt(qt / double(rows - 1))); }
}

// normalise (cublasNN.cpp:199-212): (x - mean) / stddev with IEEE division, 0 where the feature is constant.
__global__ void normalise_rows_kernel(float* __restrict__ x, int64_t ld, int64_t rows, const float* __restrict__ mean,
                                      const float* __restrict__ stddev) {
    const int f = blockIdx.y;
    const float mu = mean[f], sd = stddev[f];
    float* row = x + int64_t(f) * ld;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride)
        row[i] = (sd != 0.f) ? __fdiv_rn(row[i] - mu, sd) : 0.f;
}

inline unsigned blocks_for(int64_t n, int threads) { return unsigned((n + threads - 1) / threads); }

}  // namespace

cudaError_t output_layer_launch(const OutputArgs& a, cudaStream_t stream) {
    if (a.rows <= 0) return cudaSuccess;
    return launch_pdl(output_layer_kernel, dim3(blocks_for(a.rows, OUT_THREADS)), dim3(OUT_THREADS), 0, stream, a, a.cost_sum,
                      a.err_sum);
}

cudaError_t momentum_update_launch(float* theta, float* vel, const float* grad, int64_t count, float mu, float alpha,
                                   cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    const bool aligned = ((reinterpret_cast<uintptr_t>(theta) | reinterpret_cast<uintptr_t>(vel) |
                           reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
    const int64_t n4 = aligned ? count / 4 : 0;
    if (n4 > 0) {
        // 148 SMs x 8 resident blocks of 256 threads; grid-stride covers the rest.
        const int64_t want = (n4 + 255) / 256;
        const unsigned grid = unsigned(want < 148 * 8 ? want : 148 * 8);
        cudaError_t e = launch_pdl(momentum_update_kernel<false>, dim3(grid), dim3(256), 0, stream, reinterpret_cast<float4*>(theta),
                                   reinterpret_cast<float4*>(vel), reinterpret_cast<float4*>(const_cast<float*>(grad)), n4, mu, alpha);
        if (e != cudaSuccess) return e;
    }
    const int64_t done = n4 * 4;
    if (done < count)
        return launch_pdl(momentum_update_tail_kernel, dim3(blocks_for(count - done, 256)), dim3(256), 0, stream, theta, vel, grad,
                          done, count, mu, alpha);
    return cudaSuccess;
}

// Same update, and the gradient arena is left all-zero (small networks: the producers of the next step accumulate
// atomically and would otherwise each need a memset node in front).  count % 4 == 0 and 16-byte aligned arenas.
cudaError_t momentum_update_zero_launch(float* theta, float* vel, float* grad, int64_t count, float mu, float alpha,
                                        cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    if ((count & 3) || ((reinterpret_cast<uintptr_t>(theta) | reinterpret_cast<uintptr_t>(vel) | reinterpret_cast<uintptr_t>(grad)) & 15))
        return cudaErrorInvalidValue;
    const int64_t n4 = count / 4, want = (n4 + 255) / 256;
    return launch_pdl(momentum_update_kernel<true>, dim3(unsigned(want < 148 * 8 ? want : 148 * 8)), dim3(256), 0, stream,
                      reinterpret_cast<float4*>(theta), reinterpret_cast<float4*>(vel), reinterpret_cast<float4*>(grad), n4, mu, alpha);
}

cudaError_t theta_ref_to_internal_launch(const float* ref, float* w, int in, int out, int64_t ldw, cudaStream_t s) {
    const int64_t n = (int64_t(in) + 1) * out;
    theta_ref_to_internal_kernel<<<blocks_for(n, 256), 256, 0, s>>>(ref, w, in, out, ldw);
    return cudaGetLastError();
}
cudaError_t theta_internal_to_ref_launch(const float* w, float* ref, int in, int out, int64_t ldw, cudaStream_t s) {
    const int64_t n = (int64_t(in) + 1) * out;
    theta_internal_to_ref_kernel<<<blocks_for(n, 256), 256, 0, s>>>(w, ref, in, out, ldw);
    return cudaGetLastError();
}
cudaError_t stage_x_launch(const float* x_ref, int64_t ld_ref, float* x_int, int64_t ld_int, int64_t rows, int n,
                           cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    dim3 grid(blocks_for(rows, 256), unsigned(n + 1));
    stage_x_kernel<<<grid, 256, 0, s>>>(x_ref, ld_ref, x_int, ld_int, rows, n);
    return cudaGetLastError();
}
cudaError_t split_batches_launch(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int64_t bs,
                                 int64_t bs_pad, int nb, int n_rows, cudaStream_t s) {
    if (rows <= 0 || n_rows <= 0) return cudaSuccess;
    dim3 grid(blocks_for(rows, 256), unsigned(n_rows));
    split_batches_kernel<<<grid, 256, 0, s>>>(src, ld_src, dst, ld_dst, rows, bs, bs_pad, nb);
    return cudaGetLastError();
}
cudaError_t split_lo_launch(const float* x, float* lo, int64_t count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const int64_t want = (count + 255) / 256;
    return launch_pdl(split_lo_kernel, dim3(unsigned(want < 148 * 8 ? want : 148 * 8)), dim3(256), 0, s, x, lo, count);
}
cudaError_t momentum_update_lo_launch(float* theta, float* vel, const float* grad, float* theta_lo, int64_t count, float mu,
                                      float alpha, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    const int64_t want = (count + 255) / 256;
    return launch_pdl(momentum_update_lo_kernel, dim3(unsigned(want < 148 * 16 ? want : 148 * 16)), dim3(256), 0, stream, theta, vel,
                      grad, theta_lo, count, mu, alpha);
}
cudaError_t to_bf16_launch(const float* x, void* h, int64_t count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const int64_t want = (count + 255) / 256;
    return launch_pdl(to_bf16_kernel, dim3(unsigned(want < 148 * 8 ? want : 148 * 8)), dim3(256), 0, s, x, static_cast<__nv_bfloat16*>(h), count);
}
cudaError_t momentum_update_bf16_launch(float* theta, float* vel, const float* grad, void* theta_h, int64_t count, float mu,
                                        float alpha, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    // the vector path needs 16-byte aligned fp32 arenas and an 8-byte aligned bf16 copy: layer offsets are multiples of 8 floats
    if ((reinterpret_cast<uintptr_t>(theta) | reinterpret_cast<uintptr_t>(vel) | reinterpret_cast<uintptr_t>(grad)) & 15 ||
        (reinterpret_cast<uintptr_t>(theta_h) & 7)) return cudaErrorInvalidValue;
    const int64_t want = (count / 4 + 255) / 256 + 1;
    return launch_pdl(momentum_update_bf16_kernel, dim3(unsigned(want < 148 * 8 ? want : 148 * 8)), dim3(256), 0, stream, theta, vel,
                      grad, static_cast<__nv_bfloat16*>(theta_h), count, mu, alpha);
}
cudaError_t abs_sum_launch(const float* J, int64_t ld, int64_t rows, int K, double* out, cudaStream_t s) {
    if (rows <= 0 || K <= 0) return cudaSuccess;
    const int64_t want = (rows * K + 255) / 256;
    abs_sum_kernel<<<unsigned(want < 296 ? want : 296), 256, 0, s>>>(J, ld, rows, K, out);
    return cudaGetLastError();
}
cudaError_t fill_launch(float* p, float v, int64_t count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    const int64_t want = (count + 255) / 256;
    fill_kernel<<<unsigned(want < 148 * 8 ? want : 148 * 8), 256, 0, s>>>(p, v, count);
    return cudaGetLastError();
}
cudaError_t multiply_launch(const float* a, const float* b, float* out, int64_t count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    multiply_kernel<<<blocks_for(count, 256), 256, 0, s>>>(a, b, out, count);
    return cudaGetLastError();
}
cudaError_t init_scale_launch(float* theta, int in, int out, int64_t count, int kind, cudaStream_t s) {
    init_scale_kernel<<<blocks_for(count, 256), 256, 0, s>>>(theta, in, out, count, kind);
    return cudaGetLastError();
}
cudaError_t feature_stats_launch(const float* x, int64_t ld, int64_t rows, int n, float* mean, float* stddev, cudaStream_t s) {
    if (n <= 0 || rows <= 0) return cudaSuccess;
    feature_stats_kernel<<<unsigned(n), 256, 0, s>>>(x, ld, rows, mean, stddev);
    return cudaGetLastError();
}
cudaError_t normalise_rows_launch(float* x, int64_t ld, int64_t rows, int n, const float* mean, const float* stddev, cudaStream_t s) {
    if (n <= 0 || rows <= 0) return cudaSuccess;
    const int64_t want = (rows + 255) / 256;
    dim3 grid(unsigned(want < 64 ? want : 64), unsigned(n));
    normalise_rows_kernel<<<grid, 256, 0, s>>>(x, ld, rows, mean, stddev);
    return cudaGetLastError();
}
cudaError_t labels_to_onehot_launch(const float* labels, float* y, int64_t ldy, int64_t rows, int K, cudaStream_t s) {
    labels_to_onehot_kernel<<<blocks_for(rows, 256), 256, 0, s>>>(labels, y, ldy, rows, K);
    return cudaGetLastError();
}

}  // namespace b2
