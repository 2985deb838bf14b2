// builtins.cu — the reference's pluggable device ops as stand-alone callables.
//
// main.cpp passes these by address (tanhGPU, sechSqGPU, softmaxGPU, crossEntropyCostGPU, randInitWeights2GPU ...:
// main.cpp:75, 96, 127, 157).  The engine recognises the addresses (b2nn_builtin_kind) and runs the fused epilogues
// instead; calling them directly still works on any device buffer, on the legacy default stream like the reference
// (activations.cu:4-32, costfunctions.cu:4-12, kernels.cu:4-37, randinitweights.cu:3-11), so user code that composes
// them into its own activation keeps working through the generic function-pointer path.
#include "../../include/b2nn.h"
#include "../../include/activations.h"
#include "../../include/costfunctions.h"
#include "../../include/kernels.h"
#include "../../include/randinitweights.h"

#include <cuda_runtime.h>
#include <cstdint>

namespace {

constexpr int TPB = 256;
inline unsigned nblocks(long long n) {
    long long b = (n + TPB - 1) / TPB;
    const long long cap = 148LL * 16;
    return unsigned(b < 1 ? 1 : (b > cap ? cap : b));
}

enum Op { OP_SIGMOID, OP_SIGMOID_GRAD, OP_TANH, OP_SECH_SQ, OP_ABS };

template <int OP>
__global__ void __launch_bounds__(TPB) map_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = in[i];
        float r;
        if (OP == OP_SIGMOID) r = 1.0f / (1.0f + expf(-v));
        else if (OP == OP_SIGMOID_GRAD) { const float a = 1.0f / (1.0f + expf(-v)); r = a * (1.0f - a); }
        else if (OP == OP_TANH) r = tanhf(v);
        else if (OP == OP_SECH_SQ) { const float c = coshf(v); r = 1.0f / (c * c); }
        else r = fabsf(v);
        out[i] = r;
    }
}

__global__ void __launch_bounds__(TPB) add_scalar_kernel(const float* __restrict__ in, float alpha, float* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i] + alpha;
}

enum Bin { BIN_SUB, BIN_MUL, BIN_NEQ, BIN_COST_NLM, BIN_COST_CE };

template <int OP>
__global__ void __launch_bounds__(TPB) zip_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = a[i], y = b[i];
        float r;
        if (OP == BIN_SUB) r = x - y;
        else if (OP == BIN_MUL) r = y * x;
        else if (OP == BIN_NEQ) r = (x != y) ? 1.0f : 0.0f;
        else if (OP == BIN_COST_NLM) r = -y * logf(x) - (1.0f - y) * logf(1.0f - x);   // x = h, y = target
        else r = -y * logf(x);
        out[i] = r;
    }
}

// Row softmax of a rows x cols column-major matrix, one thread per row, max-shifted for stability.
__global__ void __launch_bounds__(TPB) softmax_rows_kernel(const float* __restrict__ z, float* __restrict__ h, int rows, int cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    float zmax = -INFINITY, total = 0.f;
    for (int j = 0; j < cols; ++j) zmax = fmaxf(zmax, z[(long long)j * rows + i]);
    for (int j = 0; j < cols; ++j) total += expf(z[(long long)j * rows + i] - zmax);
    for (int j = 0; j < cols; ++j) h[(long long)j * rows + i] = expf(z[(long long)j * rows + i] - zmax) / total;
}

__global__ void __launch_bounds__(TPB) ones_kernel(float* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1.0f;
}

// First maximum of each row; the store stays inside the bounds guard (the reference's does not, kernels.cu:85).
__global__ void __launch_bounds__(TPB) argmax_rows_kernel(const float* __restrict__ p, float* __restrict__ out, int rows, int cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    float best = 0.f; int idx = 0;
    for (int j = 0; j < cols; ++j) {
        const float v = p[(long long)j * rows + i];
        if (v > best) { best = v; idx = j; }
    }
    out[i] = float(idx);
}

__global__ void __launch_bounds__(TPB) init_rescale_kernel(float* theta, int in, int out, int n, int kind) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float eps = (kind == 1) ? __fdiv_rn(1.0f, __fsqrt_rn(float(in)))
                                  : __fdiv_rn(__fsqrt_rn(6.0f), __fsqrt_rn(float(in + out)));
    theta[i] = __fmaf_rn(__fadd_rn(theta[i], theta[i]), eps, -eps);
}

}  // namespace

// ---- activations.h ---------------------------------------------------------------------------------------------------
void sigmoidGPU(const float* A, float* B, int M) { if (M > 0) map_kernel<OP_SIGMOID><<<nblocks(M), TPB>>>(A, B, M); }
void sigmoidOutputGPU(const float* A, float* B, int M, int N) {
    const long long n = (long long)M * N;
    if (n > 0) map_kernel<OP_SIGMOID><<<nblocks(n), TPB>>>(A, B, n);
}
void sigmoidGradGPU(const float* A, float* B, int M) { if (M > 0) map_kernel<OP_SIGMOID_GRAD><<<nblocks(M), TPB>>>(A, B, M); }
void tanhGPU(const float* A, float* B, int M) { if (M > 0) map_kernel<OP_TANH><<<nblocks(M), TPB>>>(A, B, M); }
void sechSqGPU(const float* A, float* B, int M) { if (M > 0) map_kernel<OP_SECH_SQ><<<nblocks(M), TPB>>>(A, B, M); }
void softmaxGPU(const float* A, float* B, int M, int N) {
    if (M > 0 && N > 0) softmax_rows_kernel<<<(M + TPB - 1) / TPB, TPB>>>(A, B, M, N);
}

// ---- costfunctions.h -------------------------------------------------------------------------------------------------
void negLnMaxCostGPU(float* h, float* y, float* J, int M) { if (M > 0) zip_kernel<BIN_COST_NLM><<<nblocks(M), TPB>>>(h, y, J, M); }
void crossEntropyCostGPU(float* h, float* y, float* J, int M) { if (M > 0) zip_kernel<BIN_COST_CE><<<nblocks(M), TPB>>>(h, y, J, M); }

// ---- kernels.h -------------------------------------------------------------------------------------------------------
void scaVecAddGPU(const float* A, const float alpha, float* B, int M) { if (M > 0) add_scalar_kernel<<<nblocks(M), TPB>>>(A, alpha, B, M); }
void vecVecSubtractGPU(const float* A, float* B, float* C, int M) { if (M > 0) zip_kernel<BIN_SUB><<<nblocks(M), TPB>>>(A, B, C, M); }
void vecVecElementMultiplyGPU(const float* A, float* B, float* C, int M) { if (M > 0) zip_kernel<BIN_MUL><<<nblocks(M), TPB>>>(A, B, C, M); }
void absGPU(const float* A, float* B, int M) { if (M > 0) map_kernel<OP_ABS><<<nblocks(M), TPB>>>(A, B, M); }
void absVecGPU(const float* A, float* B, int M) { absGPU(A, B, M); }     // the reference declares one name and defines the other (Q11)
void addBiasMatGPU(float* A, int M) { if (M > 0) ones_kernel<<<(M + TPB - 1) / TPB, TPB>>>(A, M); }
void probToNumGPU(float* hProb, float* hNum, int M, int N) { if (M > 0) argmax_rows_kernel<<<(M + TPB - 1) / TPB, TPB>>>(hProb, hNum, M, N); }
void countErrorGPU(float* h, float* y, float* errors, int M) { if (M > 0) zip_kernel<BIN_NEQ><<<nblocks(M), TPB>>>(h, y, errors, M); }

// ---- randinitweights.h -----------------------------------------------------------------------------------------------
void randInitWeights1GPU(float* theta, int in, int out, int M) { if (M > 0) init_rescale_kernel<<<(M + TPB - 1) / TPB, TPB>>>(theta, in, out, M, 1); }
void randInitWeights2GPU(float* theta, int in, int out, int M) { if (M > 0) init_rescale_kernel<<<(M + TPB - 1) / TPB, TPB>>>(theta, in, out, M, 2); }

// ---- recognition -----------------------------------------------------------------------------------------------------
extern "C" int b2nn_builtin_kind(const void* fn, int family) {
    if (!fn) return -1;
    switch (family) {
        case 0:
            if (fn == (const void*)&sigmoidGPU) return B2NN_ACT_SIGMOID;
            if (fn == (const void*)&tanhGPU) return B2NN_ACT_TANH;
            return -1;
        case 1:
            if (fn == (const void*)&sigmoidGradGPU) return B2NN_ACT_SIGMOID;
            if (fn == (const void*)&sechSqGPU) return B2NN_ACT_TANH;
            return -1;
        case 2:
            if (fn == (const void*)&sigmoidOutputGPU) return B2NN_OUT_SIGMOID;
            if (fn == (const void*)&softmaxGPU) return B2NN_OUT_SOFTMAX;
            return -1;
        case 3:
            if (fn == (const void*)&negLnMaxCostGPU) return B2NN_COST_NEGLNMAX;
            if (fn == (const void*)&crossEntropyCostGPU) return B2NN_COST_CROSSENTROPY;
            return -1;
        case 4:
            if (fn == (const void*)&randInitWeights1GPU) return B2NN_INIT_1;
            if (fn == (const void*)&randInitWeights2GPU) return B2NN_INIT_2;
            return -1;
        default:
            return -1;
    }
}
