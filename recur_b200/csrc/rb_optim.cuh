/* recur-b200: one element of a13, the seven optimisers of rnn_apply_learning
 * (reference recur-nn.c:454-593), shared by the plain elementwise kernel and
 * the tensor engine's fused reduce + update + operand-plane kernel.          */
#ifndef RB_OPTIM_CUH
#define RB_OPTIM_CUH

#include "rb_internal.h"

/* returns the new weight; momentums[i] / aux[i] are updated in place.
   `method` is the kernel-level method: the three momentum styles arrive as
   RNN_MOMENTUM_WEIGHTED with their momentum_weight (recur-nn.c:601-678). */
__device__ __forceinline__ float
rb_optimiser_step(int method, float w, float d, float *__restrict__ momentums,
    float *__restrict__ aux, size_t i, float rate, float momentum, float momentum_weight)
{
  if (method == RNN_MOMENTUM_NESTEROV) {
    float t = d * rate;
    float m = (momentums[i] + t) * momentum;
    w += t;
    w += m;
    momentums[i] = m;
  }
  else if (method == RNN_ADAGRAD) {
    float a = momentums[i] + d * d;
    w += d * rate / sqrtf(a);
    momentums[i] = a;
  }
  else if (method == RNN_ADADELTA) {
    const float decay = momentum, renewal = 1.0f - decay;
    float gacc = momentums[i] * decay;
    float sacc = aux[i] * decay;
    gacc += fabsf(d) * renewal + rate;
    float step = sacc / gacc * d;
    sacc += fabsf(step) * renewal + rate;
    momentums[i] = gacc;
    aux[i] = sacc;
    w += step;
  }
  else if (method == RNN_RPROP) {
    const float max_step = 1.0f * rate;
    const float min_step = (float)(1e-6 * (double)rate);
    float p = momentums[i];
    float step = aux[i];
    if (d * p > 0.0f) {
      step = fminf(step * 1.2f, max_step);
    }
    else if (d * p < 0.0f) {
      step = fmaxf(step * 0.5f, min_step);
      d = 0.0f;
    }
    if (d > 0.0f)
      w += step;
    else
      w -= step;
    aux[i] = step;
    momentums[i] = d;
  }
  else { /* weighted / simplified Nesterov / classical: recur-nn.c:482-487 */
    float t = d * rate;
    float m = momentums[i];
    w += t + m * momentum_weight;
    momentums[i] = (m + t) * momentum;
  }
  return w;
}

#endif
