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

// Loss of the reference's trainer and its gradient seed in one pass (Trainer._compute_nerf_loss, core/trainer.py:352-381;
// img2mse / img2l1, :9-41):  pred = rgb + (1 - acc) * bg  (use_background),  loss = mean over N x 3 of |pred - target|
// (L1) or (pred - target)^2 (MSE), times `weight` (coarse_weight for the coarse pair).
struct LossArgs {
  const float* rgb;       // [N,3]
  const float* acc;       // [N]
  const float* target;    // [N,3]
  const float* bg;        // [N,3] or NULL (bg_const for every pixel); ignored when !use_bg
  float bg_const;
  int use_bg, mse, N;
  float weight;
  float* g_rgb;           // [N,3] out: d loss / d rgb
  float* g_acc;           // [N]   out: d loss / d acc
  float* sums;            // [2]   accumulated: sum of the per-element loss terms, sum of squared errors (for the PSNR)
};

#ifdef __CUDACC__
__global__ void loss_seed_kernel(const LossArgs a) {
  float s_loss = 0.f, s_sq = 0.f;
  const float inv = a.weight / (3.0f * (float)a.N);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += gridDim.x * blockDim.x) {
    const float one_m_acc = a.use_bg ? 1.0f - a.acc[i] : 0.f;
    float ga = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float bg = a.use_bg ? (a.bg ? a.bg[i * 3 + c] : a.bg_const) : 0.f;
      const float d = a.rgb[i * 3 + c] + one_m_acc * bg - a.target[i * 3 + c];
      const float g = a.mse ? 2.0f * d * inv : (d > 0.f ? inv : (d < 0.f ? -inv : 0.f));
      a.g_rgb[i * 3 + c] = g;
      ga -= g * bg;
      s_loss += a.mse ? d * d : fabsf(d);
      s_sq += d * d;
    }
    a.g_acc[i] = ga;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s_loss += __shfl_xor_sync(0xffffffffu, s_loss, o); s_sq += __shfl_xor_sync(0xffffffffu, s_sq, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(a.sums, s_loss); atomicAdd(a.sums + 1, s_sq); }
}

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
