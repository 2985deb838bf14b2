/* costfunctions.h — per-element reported costs of the reference API (costfunctions.h:8-9), B200 build.
 * J[i] = cost(h[i], y[i]) over `count` elements, device pointers, legacy default stream.  As in the reference the
 * cost only feeds the printed J; gradients always use h - y (cublasNN.cpp:666-667). */
#ifndef B2NN_COSTFUNCTIONS_H
#define B2NN_COSTFUNCTIONS_H

#include "cudadefs.h"

void negLnMaxCostGPU(float* h, float* y, float* J, int count);      /* -y ln h - (1-y) ln(1-h) */
void crossEntropyCostGPU(float* h, float* y, float* J, int count);  /* -y ln h                 */

#endif
