// Scalar stages of the A-NeRF hot path as __host__ __device__ functions, shared by the CUDA kernels
// and by the host-side layout tests (tests/test_host_layout.py compiles this header with g++ to check
// the K-permutation of the packed weights against the encoders' emission order -- the kernels
// themselves never run on the CPU).
//
// Reference lines each function follows are cited at the function.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ANERF_HD __host__ __device__ __forceinline__
#else
#define ANERF_HD inline
#endif

namespace anerf {

// ------------------------------------------------------------------------------------------------
// static shape of the encodings (all shipped configs: multires=7, multires_views=4;
// reference configs/*/*.txt, run_nerf.py:273-283)
// ------------------------------------------------------------------------------------------------
constexpr int kF = 7;                        // distance frequencies
constexpr int kFv = 4;                       // view-direction frequencies
constexpr int kPtsPerJoint = 1 + 2 * kF + 3; // 18: [v, (sin,cos) x 7] * w  ++  r(3)
constexpr int kViewPerJoint = 3 * (1 + 2 * kFv);  // 27: [d, (sin,cos) x 4] x 3 components, * w
constexpr int kGroups = 4;                   // worker groups; group g owns the chunks c with c % 4 == g of every operand part
constexpr int kPtsPairK = 40;                // 2 joints x 18 values + 4 zeros = 5 x 8
constexpr int kMaxJoints = 24;
constexpr int kKC = 32;                      // K elements per operand chunk (two K=16 MMA slabs)
constexpr int kTileM = 128;                  // rows (samples) per tile = UMMA M per CTA

ANERF_HD int ceil_div(int a, int b) { return (a + b - 1) / b; }
ANERF_HD int round_up(int a, int b) { return ceil_div(a, b) * b; }

struct NetDims {
  int J;       // joints
  int D;       // trunk depth (pts_linears)
  int W;       // trunk width
  int skip;    // index i such that layer i+1 takes cat[enc, h] (reference skips=[4]); -1 = none
  int fc_ch;   // per-frame appearance code channels appended to the view input (0 or 16)
  int n_fc;    // rows of the framecode table
  int fv;      // view-direction frequencies of the NETWORK (multires_views): kFv = 4, or 0 (configs/surreal/surreal_single.txt:
               // the view input is the 3 raw bone-local direction components per joint, times the cutoff weight)
};
// view inputs per joint that the network actually has: 3 * (1 + 2 fv) of the kViewPerJoint = 27 the kernels tabulate
ANERF_HD int view_per_joint(const NetDims& d) { return 3 * (1 + 2 * d.fv); }

// ------------------------------------------------------------------------------------------------
// K layout of the A operand.  Four worker groups produce the chunks of an operand part concurrently:
// group g owns the chunks whose index inside the part is congruent to g mod 4 (whole 32-wide chunks, so
// that one publish per chunk amortises the barrier / proxy-fence latency).
//   pts part : group g encodes joints g, g+4, g+8, ... two at a time (2 x 18 values + 4 zeros = 40
//              = 5 x 8) into its own value stream, zero padded to `pts_group_chunks` chunks; the
//              part has 4 x pts_group_chunks chunks;
//   view part: contracted per ray.  The 27 J view inputs of a sample are table[ray][j][q] * w_j(sample), so
//              Wv[:, view] x = sum_j w_j(sample) G[ray][:, j] with G[ray][n][j] = sum_q Wv[n][(j, q)] table[ray][j][q]
//              (a per-ray 128 x (J+1) matrix; the pseudo joint J carries the framecode with weight 1).  The
//              operand is one chunk per "ray slot" of the CTA pair (2R slots): a row puts [w_0..w_{J-1}, 1, 0..]
//              into the chunk of its own ray and zeros elsewhere; the B side of those chunks is G, built
//              in shared memory per item and network (not streamed from the packed image);
//   hidden   : chunk cb = accumulator columns [32cb, 32cb+32).
// ------------------------------------------------------------------------------------------------
ANERF_HD int pts_group_joints(const NetDims& d) { return ceil_div(d.J, kGroups); }
ANERF_HD int pts_pairs(const NetDims& d) { return ceil_div(pts_group_joints(d), 2); }
ANERF_HD int pts_group_chunks(const NetDims& d) { return ceil_div(pts_pairs(d) * kPtsPairK, kKC); }
ANERF_HD int pts_chunks(const NetDims& d) { return kGroups * pts_group_chunks(d); }
// every part is padded to a multiple of 4 chunks: group g then always fills ring stage g (4 stages), i.e. each
// group follows the phases of ONE "stage empty" barrier in order (a parity wait is only valid one phase ahead)
ANERF_HD int slot_chunks(int rays_per_item) { return round_up(2 * rays_per_item, kGroups); }   // ray-slot chunks of the views layer
ANERF_HD int hid_chunks(const NetDims& d) { return round_up(d.W / kKC, kGroups); }
ANERF_HD int in_pts_ref(const NetDims& d) { return d.J * (1 + 2 * kF) + d.J * 3; }
ANERF_HD int in_views_ref(const NetDims& d) { return d.J * view_per_joint(d); }

// Layer program: l in [0,D) trunk, l == D the views layer.  feature_linear has no activation behind it
// (nerf.py:121-125), so it is folded into views_linears[0] when the weights are packed:
//   views(cat[feature(h), x_view]) = (Wv[:, :W] Wf) h + Wv[:, W:] x_view + (bv + Wv[:, :W] bf)
// -- the same function up to fp32 re-association, one 256x256 GEMM per sample less.
ANERF_HD int layer_n(const NetDims& d, int l) { return l < d.D ? d.W : d.W / 2; }
// chunks of layer l that are streamed from the packed image (the views layer's ray-slot chunks are not)
ANERF_HD int layer_chunks(const NetDims& d, int l) {
  if (l == 0) return pts_chunks(d);
  if (l < d.D) return hid_chunks(d) + ((l - 1) == d.skip ? pts_chunks(d) : 0);
  return hid_chunks(d);
}

// Map packed K index of layer l -> column of the reference weight matrix (or -1 for zero padding).
// Reference column orders: pts input = [k*J + j (k = 0 raw, 1+2f sin, 2+2f cos) | 15J + 3j + c]
// (cutoff_embedder.py:147-172, raycasters.py:560-569); skip layer input = cat[pts input, h]
// (nerf.py:100-101); (folded) views layer input = cat[h, k*3J + 3j + c, framecode] (nerf.py:121-125).
ANERF_HD int pts_part_ref_col(const NetDims& d, int k) {
  int c = k / kKC, g = c % kGroups;
  int p = (c / kGroups) * kKC + (k % kKC);       // position in group g's value stream
  if (p >= pts_pairs(d) * kPtsPairK) return -1;
  int pair = p / kPtsPairK, within = p % kPtsPairK;
  if (within >= 2 * kPtsPerJoint) return -1;
  int j = g + kGroups * (2 * pair + within / kPtsPerJoint);
  int q = within % kPtsPerJoint;
  if (j >= d.J) return -1;
  return q < 1 + 2 * kF ? q * d.J + j : (1 + 2 * kF) * d.J + 3 * j + (q - (1 + 2 * kF));
}
// column of the (folded) views weight matrix that multiplies feature q (= 3*kk + c, kk = 0 raw, 1+2f sin,
// 2+2f cos) of joint j; j == J is the framecode pseudo joint (q < fc_ch).  -1 = no such input.
ANERF_HD int view_weight_col(const NetDims& d, int j, int q) {
  if (j < d.J) return q < view_per_joint(d) ? d.W + (q / 3) * 3 * d.J + 3 * j + (q % 3) : -1;
  return (j == d.J && q < d.fc_ch) ? d.W + in_views_ref(d) + q : -1;
}
ANERF_HD int layer_ref_col(const NetDims& d, int l, int k) {
  int P = pts_chunks(d) * kKC;
  if (l == 0) return pts_part_ref_col(d, k);
  if (l < d.D) {
    if ((l - 1) == d.skip) return k < P ? pts_part_ref_col(d, k) : ((k - P) < d.W ? in_pts_ref(d) + (k - P) : -1);
    return k < d.W ? k : -1;
  }
  return k < d.W ? k : -1;   // views layer, h part (columns of the folded Wv[:, :W] Wf)
}

// ------------------------------------------------------------------------------------------------
// torch.linspace(0, 1, n)[i] in fp32: symmetric evaluation, as ATen's range factory computes it
// (used by ray_utils.py:176,218 for t_vals and the deterministic importance draws u).
// ------------------------------------------------------------------------------------------------
ANERF_HD float linspace01(int i, int n) {
  if (n == 1) return 0.f;
  float step = 1.0f / (float)(n - 1);
  return i < n / 2 ? step * (float)i : 1.0f - step * (float)(n - 1 - i);
}

// get_rays (ray_utils.py:6-28) for one pixel: dirs = [(i - cx) / fx, -(j - cy) / fy, -1], d = c2w[:3,:3] dirs,
// o = c2w[:3,3].  c2w = rows 0..2 of the camera-to-world matrix, row-major [3][4]; pix = j * W + i.
struct RayGen {
  float c2w[12];
  float fx, fy, cx, cy;
  float near, far;
  int W;
  int pixel0;            // pixel of ray 0 when `pixels` is NULL
  const int* pixels;     // optional list of flat pixel indices, one per ray
};
ANERF_HD void pixel_ray(const RayGen& g, int ray, float r[8]) {
  const int pix = g.pixels ? g.pixels[ray] : g.pixel0 + ray;
  const float i = (float)(pix % g.W), j = (float)(pix / g.W);
  const float dx = (i - g.cx) / g.fx, dy = -(j - g.cy) / g.fy, dz = -1.0f;
#if defined(__CUDA_ARCH__)       // products and sums rounded separately, as the reference's elementwise torch code does
#pragma unroll
  for (int a = 0; a < 3; ++a)
    r[3 + a] = __fadd_rn(__fadd_rn(__fmul_rn(dx, g.c2w[4 * a]), __fmul_rn(dy, g.c2w[4 * a + 1])), __fmul_rn(dz, g.c2w[4 * a + 2]));
#else
  for (int a = 0; a < 3; ++a) {
    volatile float p0 = dx * g.c2w[4 * a], p1 = dy * g.c2w[4 * a + 1], p2 = dz * g.c2w[4 * a + 2];
    volatile float s01 = p0 + p1;
    r[3 + a] = s01 + p2;
  }
#endif
  r[0] = g.c2w[3]; r[1] = g.c2w[7]; r[2] = g.c2w[11];
  r[6] = g.near; r[7] = g.far;
}

// a3 (ray_utils.py:204-251): coarse depth `sidx` of a ray; `t_rand_row` (the ray's Sc draws) switches the
// stratified jitter on.  Same arithmetic as the fused kernel's stage (2).
ANERF_HD float coarse_depth(float near, float far, int sidx, int Sc, int lindisp, const float* t_rand_row) {
  auto zat = [&](int k) {
    float t = linspace01(k, Sc);
    return lindisp ? 1.f / (1.f / near * (1.f - t) + 1.f / far * t) : near * (1.f - t) + far * t;
  };
  float z = zat(sidx);
  if (t_rand_row) {
    float lower = sidx == 0 ? z : 0.5f * (zat(sidx - 1) + z);
    float upper = sidx == Sc - 1 ? z : 0.5f * (z + zat(sidx + 1));
    z = lower + (upper - lower) * t_rand_row[sidx];
  }
  return z;
}

// ------------------------------------------------------------------------------------------------
// a2: ray / bounding-cylinder intersection in the x-z plane (ray_utils.py:292-326).
// Returns near', far' (NaN when the ray's projection misses the circle) .
// ------------------------------------------------------------------------------------------------
ANERF_HD void near_far_cylinder(const float o[3], const float d[3], const float cyl[3], float near, float far,
                                float& nn, float& ff, bool& miss) {
  float pnx = o[0] + d[0] * near, pnz = o[2] + d[2] * near;
  float pfx = o[0] + d[0] * far, pfz = o[2] + d[2] * far;
  float ncx = cyl[0] - pnx, ncz = cyl[1] - pnz;
  float sx = pfx - pnx, sz = pfz - pnz;
  float seg = sqrtf(sx * sx + sz * sz);
  float scale = sqrtf(d[0] * d[0] + d[2] * d[2]);
  float cross = ncx * sz - ncz * sx;
  float dl = fabsf(cross) / seg;
  float q2 = cyl[2] * cyl[2] - dl * dl;
  float Q = sqrtf(q2);                       // NaN for q2 < 0, like tensor.pow(0.5)
  float K = (ncx * sx + ncz * sz) / seg;
  float outside = (Q < K) ? 1.f : 0.f;
  nn = near + outside * (K - Q) / scale;
  ff = near + (K + Q) / scale;
  miss = Q != Q;
}

// ------------------------------------------------------------------------------------------------
// a4-a8: one joint of the distance/bone encoding of a world point (encoders.py:8-23,120,189;
// cutoff_embedder.py:111-174 with dist_inputs=False, cutoff_inputs=True).
// skt = rows 0..2 of the 4x4 world->bone transform, row-major (12 floats).
// out[18] = [v w, sin(2^f v) w, cos(2^f v) w (f=0..6), r0, r1, r2];  also returns v.
// ------------------------------------------------------------------------------------------------
// 1 - sigmoid(tau (v - c))  (cutoff_embedder.py:138-146).  Device: ex2/rcp based (relative error ~2^-21 on a
// factor in [0,1]); host (layout tests): libm.
ANERF_HD float cutoff_w(float v, float tau, float cut) {
  float a = tau * (v - cut);
#if defined(__CUDA_ARCH__)
  return 1.0f - __fdividef(1.0f, 1.0f + __expf(-a));
#else
  return 1.0f - 1.0f / (1.0f + expf(-a));
#endif
}

ANERF_HD void bone_local(const float* skt, const float p[3], float x[3]) {
  x[0] = skt[0] * p[0] + skt[1] * p[1] + skt[2] * p[2] + skt[3];
  x[1] = skt[4] * p[0] + skt[5] * p[1] + skt[6] * p[2] + skt[7];
  x[2] = skt[8] * p[0] + skt[9] * p[1] + skt[10] * p[2] + skt[11];
}

// sin and cos of x = 2^f v (exact in fp32), |x| up to a few hundred.  Device: two-term Cody-Waite reduction by
// 2 pi to [-pi, pi], then the SFU (sin.approx / cos.approx, absolute error <= 2^-20.9 there): ~5e-7 absolute,
// 5 instructions.  (sincosf costs ~35 instructions per call; two calls + double-angle steps were 1.4e-6.)
ANERF_HD void sincos_2pi(float x, float& s, float& c) {
#if defined(__CUDA_ARCH__)
  float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, x);          // 2 pi, high part (fp32)
  r = fmaf(-k, -1.7484555314695172e-07f, r);            // 2 pi - high part
  s = __sinf(r);
  c = __cosf(r);
#else
  s = sinf(x); c = cosf(x);
#endif
}

ANERF_HD float encode_joint_pts(const float* skt, const float p[3], float tau, float cut, float* out) {
  float x[3];
  bone_local(skt, p, x);
  float v = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
#if defined(__CUDA_ARCH__)
  float inv = __frcp_rn(fmaxf(v, 1e-12f));
#else
  float inv = 1.0f / fmaxf(v, 1e-12f);
#endif
  float w = cutoff_w(v, tau, cut);
  out[0] = v * w;
#pragma unroll
  for (int f = 0; f < kF; ++f) {
    float s, c;
    sincos_2pi(v * (float)(1 << f), s, c);
    out[1 + 2 * f] = s * w;
    out[2 + 2 * f] = c * w;
  }
  out[15] = x[0] * inv;
  out[16] = x[1] * inv;
  out[17] = x[2] * inv;
  return v;
}

// distance of a world point to one joint only (for the view-direction cutoff weight)
ANERF_HD float joint_dist(const float* skt, const float p[3]) {
  float x[3];
  bone_local(skt, p, x);
  return sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
}

// ------------------------------------------------------------------------------------------------
// a5, a7, a10 (per-ray part): unit bone-local ray direction and its sin/cos features
// (encoders.py:25-37,189; cutoff_embedder.py:116-124).  out[27] = [d_c, sin(2^f d_c), cos(2^f d_c)]
// at index 3*kk + c, kk = 0 raw, 1+2f sin, 2+2f cos.  The per-sample factor w_j is applied later.
// ------------------------------------------------------------------------------------------------
ANERF_HD void encode_joint_viewdir(const float* skt, const float d[3], float* out) {
  float x[3];
  x[0] = skt[0] * d[0] + skt[1] * d[1] + skt[2] * d[2];
  x[1] = skt[4] * d[0] + skt[5] * d[1] + skt[6] * d[2];
  x[2] = skt[8] * d[0] + skt[9] * d[1] + skt[10] * d[2];
  float n = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  float inv = 1.0f / fmaxf(n, 1e-12f);
  for (int c = 0; c < 3; ++c) {
    float u = x[c] * inv;
    out[c] = u;
    for (int f = 0; f < kFv; ++f) {
      float a = u * (float)(1 << f);
      out[3 * (1 + 2 * f) + c] = sinf(a);
      out[3 * (2 + 2 * f) + c] = cosf(a);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// a12: per-sample part of raw2outputs (nerf.py:150-190)
// ------------------------------------------------------------------------------------------------
ANERF_HD float density_act(float raw_sigma, float B, float noise, int softplus, float shift) {
  float x = raw_sigma / B + noise;   // raw / B + noise (nerf.py:153)
  if (!softplus) return fmaxf(x, 0.f);
  x -= shift;
  return x > 20.f ? x : log1pf(expf(x));
}
ANERF_HD float sigmoid_rgb(float x) { return (1.0f / (1.0f + expf(-x))) * 1.002f - 0.001f; }

}  // namespace anerf
