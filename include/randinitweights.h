/* randinitweights.h — weight initialisers of the reference API (randinitweights.h:8-9), B200 build.
 * theta holds `count` uniform (0,1] draws on entry and is rescaled in place to [-eps, +eps];
 * `in` counts the bias unit (cublasNN.cpp:317). */
#ifndef B2NN_RANDINITWEIGHTS_H
#define B2NN_RANDINITWEIGHTS_H

#include "cudadefs.h"

void randInitWeights1GPU(float* theta, int in, int out, int count);  /* eps = 1 / sqrt(in)            */
void randInitWeights2GPU(float* theta, int in, int out, int count);  /* eps = sqrt(6) / sqrt(in+out)  */

#endif
