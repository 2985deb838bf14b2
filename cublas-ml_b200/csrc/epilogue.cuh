// epilogue.cuh — element-wise math fused into the contraction kernels.
// Same libm calls as the reference kernels (expf / tanhf, IEEE division, no fast-math: activations.cu:34-66); the
// derivatives are evaluated from the stored activation instead of recomputing it from z:
//   sigmoid'(z) = a(1-a)            (reference: a = sigmoid(z); a*(1-a), activations.cu:41-49 — identical)
//   sech^2(z)   = 1 - tanh(z)^2     (reference: 1/cosh(z)^2, activations.cu:58-66 — equal up to fp32 rounding)
#pragma once
#include "engine.h"

namespace b2 {

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

__device__ __forceinline__ float apply_epilogue(int epi, float acc, float a) {
    switch (epi) {
        case EPI_SIGMOID:  return sigmoid_f(acc);
        case EPI_TANH:     return tanhf(acc);
        case EPI_DSIGMOID: return acc * (a * (1.0f - a));
        case EPI_DTANH:    return acc * (1.0f - a * a);
        default:           return acc;
    }
}

template <int EPI>
__device__ __forceinline__ float apply_epilogue_t(float acc, float a) {
    if (EPI == EPI_SIGMOID) return sigmoid_f(acc);
    if (EPI == EPI_TANH) return tanhf(acc);
    if (EPI == EPI_DSIGMOID) return acc * (a * (1.0f - a));
    if (EPI == EPI_DTANH) return acc * (1.0f - a * a);
    return acc;
}

}  // namespace b2
