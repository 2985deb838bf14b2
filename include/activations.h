/* activations.h — hidden / output activations of the reference API (activations.h:8-13), B200 build.
 *
 * Each symbol is (a) a token: pass it to cublasNN::train*Momentum / validate* / predict* and the engine runs the
 * matching fused tcgen05 epilogue; (b) a callable: host function, DEVICE pointers, legacy default stream, same
 * argument meaning as the reference.  Hidden activations take an element count; output activations take
 * (rows, cols) of a column-major matrix. */
#ifndef B2NN_ACTIVATIONS_H
#define B2NN_ACTIVATIONS_H

#include "cudadefs.h"

void sigmoidGPU(const float* in, float* out, int count);                 /* 1 / (1 + exp(-x))                     */
void sigmoidGradGPU(const float* in, float* out, int count);             /* s(x) (1 - s(x)), evaluated from x     */
void tanhGPU(const float* in, float* out, int count);                    /* tanh(x)                               */
void sechSqGPU(const float* in, float* out, int count);                  /* 1 / cosh(x)^2, evaluated from x       */
void sigmoidOutputGPU(const float* in, float* out, int rows, int cols);  /* element-wise sigmoid over rows*cols   */
void softmaxGPU(const float* in, float* out, int rows, int cols);        /* row softmax (max-shifted)             */

#endif
