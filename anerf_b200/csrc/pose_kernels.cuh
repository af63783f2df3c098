// the pose chain of pose-refinement training, fused with the reduction of the renderer's
// d/d skts.  Reference: PoseOptLayer.calculate_kinematic + unrolled_kinematic_chain + torch.inverse
// (core/pose_opt.py:372-445, 482-521).
//
//   l2w_root = [R_root | rest_root],  l2w_j = l2w_parent(j) [R_j | rest_j - rest_parent(j)]      (homogeneous 4x4)
//   l2w_j[:3, 3] += pelvis                                                                      (after the chain)
//   skt_j = l2w_j^-1,   kp_j = l2w_j[:3, 3]
//
// Forward: one thread per pose walks the tree in joint order (parents precede children, as in the SMPL tree the
// reference unrolls by hand); the inverse of a rigid transform is written in closed form ([R^T | -R^T t]) instead of
// the reference's LU `torch.inverse` (same value up to fp32 rounding for the orthonormal R that rot6d / axis-angle
// parametrisations produce).  Backward: the per-pose cotangent of skts (already segment-summed over the rays of each
// pose by the renderer's backward, which adds into [P,J,4,4] with atomics when it is given a ray -> pose index) is
// pulled through the inverse with the exact derivative of matrix inversion (-S^T G S^T, so gradients w.r.t. the
// entries of `rots` equal autograd's through torch.inverse, not only their projection on SO(3)), then through the chain
// in reverse joint order.
#pragma once
#include <cuda_runtime.h>

namespace anerf {
namespace pose {

constexpr int kMaxPoseJoints = 32;

struct ChainArgs {
  int P, J;
  int parent[kMaxPoseJoints];     // parent[root] == root; parent[j] < j otherwise
  int root;
  const float* rots;              // [P,J,3,3]
  const float* rest;              // [Pr,J,3]
  long long rest_stride;          // J*3, or 0 when one rest pose serves all
  const float* pelvis;            // [P,3]
  float* l2ws;                    // [P,J,4,4]  (shifted by the pelvis, as the reference returns them)
  float* skts;                    // [P,J,4,4]
  float* kps;                     // [P,J,3]
  // backward
  const float* g_skts;            // [P,J,4,4] or NULL
  const float* g_l2ws;            // [P,J,4,4] or NULL
  const float* g_kps;             // [P,J,3] or NULL
  float* g_rots;                  // [P,J,3,3]
  float* g_pelvis;                // [P,3]
  float* scratch;                 // [P,J,12] backward: accumulated cotangents of the unshifted l2w (rows 0..2)
};

#ifdef __CUDACC__

// order in which joints are visited: index order with the root first (parents precede children)
__global__ void pose_chain_fwd_kernel(const ChainArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.P) return;
  const int J = a.J;
  const float* R = a.rots + (size_t)p * J * 9;
  const float* rest = a.rest + (size_t)p * a.rest_stride;
  const float px = a.pelvis[p * 3], py = a.pelvis[p * 3 + 1], pz = a.pelvis[p * 3 + 2];
  float* L = a.l2ws + (size_t)p * J * 16;
  // pass 1: unshifted chain, written to l2ws (read back by the same thread for the children)
  for (int step = 0; step < J; ++step) {
    const int j = step == 0 ? a.root : (step <= a.root ? step - 1 : step);
    const float* r = R + j * 9;
    float* out = L + j * 16;
    if (j == a.root) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        out[4 * i] = r[3 * i]; out[4 * i + 1] = r[3 * i + 1]; out[4 * i + 2] = r[3 * i + 2];
        out[4 * i + 3] = rest[j * 3 + i];
      }
    } else {
      const int q = a.parent[j];
      const float* Pm = L + q * 16;
      const float tx = rest[j * 3] - rest[q * 3], ty = rest[j * 3 + 1] - rest[q * 3 + 1], tz = rest[j * 3 + 2] - rest[q * 3 + 2];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float a0 = Pm[4 * i], a1 = Pm[4 * i + 1], a2 = Pm[4 * i + 2], a3 = Pm[4 * i + 3];
        out[4 * i] = a0 * r[0] + a1 * r[3] + a2 * r[6];
        out[4 * i + 1] = a0 * r[1] + a1 * r[4] + a2 * r[7];
        out[4 * i + 2] = a0 * r[2] + a1 * r[5] + a2 * r[8];
        out[4 * i + 3] = a0 * tx + a1 * ty + a2 * tz + a3;
      }
    }
    out[12] = 0.f; out[13] = 0.f; out[14] = 0.f; out[15] = 1.f;
  }
  // pass 2: pelvis shift, keypoints, closed-form inverse
  for (int j = 0; j < J; ++j) {
    float* m = L + j * 16;
    m[3] += px; m[7] += py; m[11] += pz;
    const float t0 = m[3], t1 = m[7], t2 = m[11];
    if (a.kps) { float* k = a.kps + ((size_t)p * J + j) * 3; k[0] = t0; k[1] = t1; k[2] = t2; }
    float* s = a.skts + ((size_t)p * J + j) * 16;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float c0 = m[i], c1 = m[4 + i], c2 = m[8 + i];         // column i of R = row i of R^T
      s[4 * i] = c0; s[4 * i + 1] = c1; s[4 * i + 2] = c2;
      s[4 * i + 3] = -(c0 * t0 + c1 * t1 + c2 * t2);
    }
    s[12] = 0.f; s[13] = 0.f; s[14] = 0.f; s[15] = 1.f;
  }
}

__global__ void pose_chain_bwd_kernel(const ChainArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.P) return;
  const int J = a.J;
  const float* R = a.rots + (size_t)p * J * 9;
  const float* rest = a.rest + (size_t)p * a.rest_stride;
  const float pel[3] = {a.pelvis[p * 3], a.pelvis[p * 3 + 1], a.pelvis[p * 3 + 2]};
  const float* L = a.l2ws + (size_t)p * J * 16;
  float* G = a.scratch + (size_t)p * J * 12;        // rows 0..2 of the cotangent of l2w_j (the bottom row is constant)
  float gp[3] = {0.f, 0.f, 0.f};
  // cotangent of the (shifted) l2w_j from its three consumers: l2ws itself, kps, and skts = l2w^-1
  for (int j = 0; j < J; ++j) {
    float g[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) g[i] = a.g_l2ws ? a.g_l2ws[((size_t)p * J + j) * 16 + i] : 0.f;
    if (a.g_kps) {
      const float* k = a.g_kps + ((size_t)p * J + j) * 3;
      g[3] += k[0]; g[7] += k[1]; g[11] += k[2];
    }
    if (a.g_skts) {
      // d(M^-1): G_M = -S^T G_S S^T with S = M^-1 (rows 0..2 needed; S's bottom row is [0 0 0 1])
      const float* S = a.skts + ((size_t)p * J + j) * 16;
      const float* GS = a.g_skts + ((size_t)p * J + j) * 16;
      float T[4][4];                                   // T = G_S S^T
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) acc = fmaf(GS[4 * r + k], S[4 * c + k], acc);
          T[r][c] = acc;
        }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) acc = fmaf(S[4 * k + r], T[k][c], acc);     // (S^T T)[r][c]
          g[4 * r + c] -= acc;
        }
    }
    gp[0] += g[3]; gp[1] += g[7]; gp[2] += g[11];      // the shift adds the pelvis to every joint's translation
#pragma unroll
    for (int i = 0; i < 12; ++i) G[j * 12 + i] = g[i];
  }
  a.g_pelvis[p * 3] = gp[0]; a.g_pelvis[p * 3 + 1] = gp[1]; a.g_pelvis[p * 3 + 2] = gp[2];
  // chain in reverse order: l2w_j = l2w_q rel_j  =>  G_rel = l2w_q^T G_j (its 3x3 block is d/d R_j),  G_q += G_j rel_j^T
  for (int step = J - 1; step >= 0; --step) {
    const int j = step == 0 ? a.root : (step <= a.root ? step - 1 : step);
    const float* g = G + j * 12;
    float* gr = a.g_rots + ((size_t)p * J + j) * 9;
    if (j == a.root) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { gr[3 * i] = g[4 * i]; gr[3 * i + 1] = g[4 * i + 1]; gr[3 * i + 2] = g[4 * i + 2]; }
      continue;
    }
    const int q = a.parent[j];
    const float* Pm = L + q * 16;                      // shifted; only its rotation part and (unshifted) translation matter:
    // G_rel[:3,:3] = R_q^T G_j[:3,:3] uses the rotation block only, so the pelvis shift of the stored matrix is harmless
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        gr[3 * r + c] = Pm[r] * g[c] + Pm[4 + r] * g[4 + c] + Pm[8 + r] * g[8 + c];
    // G_q[:3,:] += G_j[:3,:] rel_j^T  with rel_j = [R_j | t_j; 0 0 0 1]
    const float* r9 = R + j * 9;
    const float t[3] = {rest[j * 3] - rest[q * 3], rest[j * 3 + 1] - rest[q * 3 + 1], rest[j * 3 + 2] - rest[q * 3 + 2]};
    float* gq = G + q * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float g0 = g[4 * r], g1 = g[4 * r + 1], g2 = g[4 * r + 2], g3 = g[4 * r + 3];
#pragma unroll
      for (int c = 0; c < 3; ++c) gq[4 * r + c] += g0 * r9[3 * c] + g1 * r9[3 * c + 1] + g2 * r9[3 * c + 2] + g3 * t[c];
      gq[4 * r + 3] += g3;
    }
    (void)pel;
  }
}

#endif  // __CUDACC__
}  // namespace pose
}  // namespace anerf
