// the optimizer step of the training loop (reference: Trainer.optimize -> torch.optim.Adam.step,
// core/trainer.py:451-483, core/raycasters.py:116) as ONE launch over all parameter tensors instead of ~10 foreach
// kernels per step (or ~4 per tensor in the single-tensor path).
//
// torch.optim.Adam semantics (amsgrad = False, maximize = False):
//   g' = g + weight_decay * p;  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#pragma once
#include <cuda_runtime.h>

namespace anerf {
namespace optim {

constexpr int kMaxTensors = 64;

struct AdamArgs {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  long long size[kMaxTensors];
  int n;
  float lr, beta1, beta2, eps, weight_decay;
  float omb1, omb2;                // 1 - beta1, 1 - beta2 rounded from double (what torch's lerp_ / addcmul_ receive)
  float bc1, bc2_sqrt;             // 1 - beta1^t, sqrt(1 - beta2^t)
  float grad_scale;                // gradients are multiplied by this first (1/world after a sum all-reduce; 1 otherwise)
};

#ifdef __CUDACC__
// grid = (blocks per tensor, tensors): blockIdx.y picks the tensor, the x dimension strides over its elements
__global__ void adam_step_kernel(const __grid_constant__ AdamArgs a) {
  const int t = blockIdx.y;
  const long long size = a.size[t];
  float* __restrict__ p = a.p[t];
  const float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  const float step_size = a.lr / a.bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < size; i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(a.weight_decay, pi, g[i] * a.grad_scale);
    const float mi = fmaf(a.omb1, gi - m[i], m[i]);                // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(a.omb2 * gi, gi, a.beta2 * v[i]);        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / a.bc2_sqrt + a.eps));
  }
}
#endif

}  // namespace optim
}  // namespace anerf
