// Host build of anerf_b200/csrc/path_math.cuh for tests/test_host_layout.py (TEST INFRASTRUCTURE).
// Mirrors the order in which the kernel's producers emit A-operand values so that the K permutation
// of the packed weights (layer_ref_col) can be checked against the oracle on the CPU.  Not a compute
// path of the product.
#include <cstring>
#include "../../anerf_b200/csrc/path_math.cuh"

using namespace anerf;

extern "C" {

int h_layer_chunks(int J, int D, int W, int skip, int fc, int l) {
  NetDims d{J, D, W, skip, fc, fc ? 4 : 0, kFv};
  return layer_chunks(d, l);
}
int h_layer_ref_col(int J, int D, int W, int skip, int fc, int l, int k) {
  NetDims d{J, D, W, skip, fc, fc ? 4 : 0, kFv};
  return layer_ref_col(d, l, k);
}
// emission order of produce_pts_chunks: group g encodes joints g, g+4, ... two at a time (36 values + 4 zeros)
// and its stream fills the chunks g, g+4, g+8, ... of the part
void h_emit_pts(const float* skt /*[J][12]*/, const float* p, float tau, const float* cut, int J, float* out) {
  NetDims d{J, 8, 256, 4, 0, 0, kFv};
  int n = pts_chunks(d) * kKC;
  memset(out, 0, n * sizeof(float));
  for (int g = 0; g < kGroups; ++g) {
    int pos = 0;
    for (int pair = 0; pair < pts_pairs(d); ++pair) {
      float v[kPtsPairK];
      memset(v, 0, sizeof(v));
      for (int jj = 0; jj < 2; ++jj) {
        int j = g + kGroups * (2 * pair + jj);
        if (j < J) encode_joint_pts(skt + j * 12, p, tau, cut[j], v + jj * kPtsPerJoint);
      }
      for (int q = 0; q < kPtsPairK; ++q, ++pos) out[((pos / kKC) * kGroups + g) * kKC + pos % kKC] = v[q];
    }
  }
}
int h_view_weight_col(int J, int D, int W, int skip, int fc, int j, int q) {
  NetDims d{J, D, W, skip, fc, fc ? 4 : 0, kFv};
  return view_weight_col(d, j, q);
}
// per-ray direction features of one joint (the table the kernel contracts the view weights with)
void h_view_table(const float* skt, const float* dir, int J, float* out /*[J][27]*/) {
  for (int j = 0; j < J; ++j) encode_joint_viewdir(skt + j * 12, dir, out + j * kViewPerJoint);
}
float h_cutoff_w(const float* skt12, const float* p, float tau, float cut) { return cutoff_w(joint_dist(skt12, p), tau, cut); }
float h_linspace01(int i, int n) { return linspace01(i, n); }
// a3: coarse depths of one ray (optionally jittered / linear in disparity), as the training path recomputes them
void h_coarse_depths(float near, float far, int Sc, int lindisp, const float* t_rand_or_null, float* out) {
  for (int s = 0; s < Sc; ++s) out[s] = coarse_depth(near, far, s, Sc, lindisp, t_rand_or_null);
}
// in-kernel ray generation (frame mode): rays of n consecutive pixels starting at pixel0
void h_pixel_rays(const float* c2w12, float fx, float fy, float cx, float cy, int W, int pixel0, int n, float* out /*[n][8]*/) {
  RayGen g{};
  for (int i = 0; i < 12; ++i) g.c2w[i] = c2w12[i];
  g.fx = fx; g.fy = fy; g.cx = cx; g.cy = cy; g.near = 0.f; g.far = 1.f; g.W = W; g.pixel0 = pixel0; g.pixels = nullptr;
  for (int r = 0; r < n; ++r) pixel_ray(g, r, out + 8 * r);
}
void h_near_far(const float* o, const float* d, const float* cyl, float near, float far, float* out) {
  bool miss;
  near_far_cylinder(o, d, cyl, near, far, out[0], out[1], miss);
  out[2] = miss ? 1.f : 0.f;
}
}
