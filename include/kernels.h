/* kernels.h — small vector helpers of the reference API (kernels.h:8-14), B200 build.
 * The train loop no longer launches any of them (they are fused into the contraction epilogues and the output
 * kernel); they remain for user code.  Device pointers, legacy default stream. */
#ifndef B2NN_KERNELS_H
#define B2NN_KERNELS_H

#include "cudadefs.h"

void scaVecAddGPU(const float* a, const float alpha, float* out, int count);   /* out = a + alpha        */
void vecVecSubtractGPU(const float* a, float* b, float* out, int count);       /* out = a - b            */
void vecVecElementMultiplyGPU(const float* a, float* b, float* out, int count);/* out = a .* b           */
void absGPU(const float* a, float* out, int count);                            /* out = |a|              */
void absVecGPU(const float* a, float* out, int count);                         /* same (reference Q11)   */
void addBiasMatGPU(float* a, int rows);                                        /* first column := 1      */
void probToNumGPU(float* probs, float* out, int rows, int cols);               /* row argmax as float    */
void countErrorGPU(float* h, float* y, float* errors, int count);              /* errors = (h != y)      */

#endif
