// Kernels of the training (backward) path: layer-wise recomputation of one network's activations in
// fp32 and the gradients of everything A-NeRF trains through the ray caster (network weights,
// framecodes, and the per-ray bone transforms `skts` for pose refinement).
//
// Reference autograd graph this restates by hand (paths relative to the reference root):
//   NeRF.raw2outputs      core/networks/nerf.py:150-205     -> composite_bwd_kernel
//   NeRF.forward          core/networks/nerf.py:94-148      -> sgemm_kernel (forward / dgrad / wgrad forms),
//                                                              head_fwd_kernel, head_bwd_kernel, colsum_kernel
//   Optcodes.forward      core/networks/embedding.py:17-34  -> framecode_bwd_kernel
//   encoders + embedders  core/encoders.py:8-37,101-122,172-193, core/cutoff_embedder.py:111-174
//                                                           -> encode_rows_kernel, encode_bwd_kernel
//
// Written in a portable SIMT subset (thread/block indices, static __shared__, __syncthreads, float
// atomicAdd, __ldg; no warp intrinsics, no PTX) so that tests/host/simt_emu.h can execute the very same
// kernels and launch sequence on the CPU of the build container against the oracle's autograd.  The
// product only ever runs them as CUDA kernels (anerf_api.cu); nothing here has a CPU path in the library.
#pragma once
#include "path_math.cuh"

namespace anerf {
namespace train {

// ------------------------------------------------------------------------------------------------
// C[M,N] (op)= A[M,K] * B[K,N] with explicit element strides, fp32 FMA.  Three uses:
//   forward  H  = relu(X W^T + b)        A = X  [rows, K]  (k contiguous), B(k,n) = W[n*ldb + k]   (BT)
//   dgrad    dX = (G W) . (H_prev > 0)   A = G  [rows, N'] (k contiguous), B(k,n) = W[k*ldb + n]
//   wgrad    dW += G^T X                 A(m,k) = G[k*lda + m] (AT), B(k,n) = X[k*ldb + n], k = rows, split
//                                         over blockIdx.z and added atomically
// ------------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A; long long lda;
  const float* B; long long ldb;
  float* C; long long ldc;
  int M, N, K;
  int k_chunk;              // K range of one blockIdx.z slice (multiple of 16); gridDim.z = ceil(K / k_chunk)
  const float* bias;        // [N] added to the product (NULL = none)
  const float* mask;        // NULL, or [M, ldmask]: the result is zeroed where mask <= 0 (ReLU derivative)
  long long ldmask;
  int relu;                 // clamp the result at 0
  int mode;                 // 0: C = r   1: C = (C + r) then mask/relu   2: atomicAdd(C, r)
};

constexpr int kBM = 128, kBN = 128, kBK = 16, kGemmThreads = 256, kPad = 4;

// 8 consecutive elements along the contiguous direction, guarded; vector loads when aligned and in range
__device__ __forceinline__ void load8(const float* p, int valid, float (&v)[8]) {
  if (valid >= 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < valid ? __ldg(p + i) : 0.f;
  }
}

template <bool AT, bool BT>
__global__ void __launch_bounds__(kGemmThreads) sgemm_kernel(GemmArgs g) {
  __shared__ float As[kBK][kBM + kPad];
  __shared__ float Bs[kBK][kBN + kPad];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int kbeg = blockIdx.z * g.k_chunk;
  const int kend = (kbeg + g.k_chunk < g.K) ? kbeg + g.k_chunk : g.K;
  const int ty = t / 16, tx = t % 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += kBK) {
    float v[8];
    // ---- A tile -> As[k][m]
    if (!AT) {                       // A[m*lda + k]: thread = (row, 8-wide k group)
      const int r = t >> 1, kq = (t & 1) * 8;
      const int m = m0 + r, k = k0 + kq;
      int valid = (m < g.M) ? (kend - k) : 0;
      load8(g.A + (long long)(m < g.M ? m : 0) * g.lda + k, valid < 0 ? 0 : valid, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) As[kq + i][r] = v[i];
    } else {                         // A[k*lda + m]: thread = (k, 8-wide m group)
      const int kk = t >> 4, mq = (t & 15) * 8;
      const int k = k0 + kk, m = m0 + mq;
      int valid = (k < kend) ? (g.M - m) : 0;
      load8(g.A + (long long)(k < kend ? k : 0) * g.lda + m, valid < 0 ? 0 : valid, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) As[kk][mq + i] = v[i];
    }
    // ---- B tile -> Bs[k][n]
    if (BT) {                        // B[n*ldb + k]
      const int r = t >> 1, kq = (t & 1) * 8;
      const int n = n0 + r, k = k0 + kq;
      int valid = (n < g.N) ? (kend - k) : 0;
      load8(g.B + (long long)(n < g.N ? n : 0) * g.ldb + k, valid < 0 ? 0 : valid, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) Bs[kq + i][r] = v[i];
    } else {                         // B[k*ldb + n]
      const int kk = t >> 4, nq = (t & 15) * 8;
      const int k = k0 + kk, n = n0 + nq;
      int valid = (k < kend) ? (g.N - n) : 0;
      load8(g.B + (long long)(k < kend ? k : 0) * g.ldb + n, valid < 0 ? 0 : valid, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) Bs[kk][nq + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[kk][ty * 4 + i];
        a[4 + i] = As[kk][64 + ty * 4 + i];
        b[i] = Bs[kk][tx * 4 + i];
        b[4 + i] = Bs[kk][64 + tx * 4 + i];
      }
#if defined(ANERF_EMU_BF16X3) || defined(ANERF_EMU_FP16X3)
      // host experiment only (tests/host): the arithmetic of the tensor-core engine (operands split into 16-bit hi + lo,
      // three products, fp32 accumulation) to separate its rounding from bugs when gradients disagree
      {
#if defined(ANERF_EMU_FP16X3)
        auto bf = [](float x) { return (float)(_Float16)x; };
#else
        auto bf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u; float r; memcpy(&r, &u, 4); return r; };
#endif
        for (int i = 0; i < 8; ++i) {
          const float ah = bf(a[i]), al = bf(a[i] - ah);
          for (int j = 0; j < 8; ++j) {
            const float bh = bf(b[j]), bl = bf(b[j] - bh);
            acc[i][j] += al * bh;
            acc[i][j] += ah * bl;
            acc[i][j] += ah * bh;
          }
        }
        continue;
      }
#endif
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float r = acc[i][j];
      float* c = g.C + (long long)m * g.ldc + n;
      if (g.mode == 2) { atomicAdd(c, r); continue; }
      if (g.bias) r += g.bias[n];
      if (g.mode == 1) r += *c;
      if (g.relu) r = fmaxf(r, 0.f);
      if (g.mask && !(g.mask[(long long)m * g.ldmask + n] > 0.f)) r = 0.f;
      *c = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// narrow heads (alpha_linear: 1 output, rgb_linear: 3 outputs)
// ------------------------------------------------------------------------------------------------
// out[row*ldo + i] = H[row, :] . W[i, :] + b[i], i < NO; one thread per row
template <int NO>
__global__ void head_fwd_kernel(const float* __restrict__ H, long long ldh, int K, const float* __restrict__ W,
                                const float* __restrict__ b, long long rows, float* __restrict__ out, long long ldo) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float acc[NO];
#pragma unroll
  for (int i = 0; i < NO; ++i) acc[i] = 0.f;
  const float* h = H + row * ldh;
  int k = 0;
  // 16-byte loads when the row is aligned (it is for every buffer of the training workspace): a thread walks its own
  // row, so wide loads cut the load instructions per row by four; same summation order as the scalar loop
  if ((ldh & 3) == 0 && (K & 3) == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0) {
    const float4* h4 = reinterpret_cast<const float4*>(h);
    for (; k < K; k += 4) {
      const float4 x = h4[k >> 2];
#pragma unroll
      for (int i = 0; i < NO; ++i) {
        acc[i] = fmaf(x.x, __ldg(W + i * K + k), acc[i]);
        acc[i] = fmaf(x.y, __ldg(W + i * K + k + 1), acc[i]);
        acc[i] = fmaf(x.z, __ldg(W + i * K + k + 2), acc[i]);
        acc[i] = fmaf(x.w, __ldg(W + i * K + k + 3), acc[i]);
      }
    }
  }
  for (; k < K; ++k) {
    const float x = h[k];
#pragma unroll
    for (int i = 0; i < NO; ++i) acc[i] = fmaf(x, __ldg(W + i * K + k), acc[i]);
  }
#pragma unroll
  for (int i = 0; i < NO; ++i) out[row * ldo + i] = acc[i] + b[i];
}

// Backward of a narrow head over a slice of rows per block (blockDim.x == K):
//   dH[row, t] = sum_i g[row, i] W[i, t]   (zeroed where H <= 0 when mask_self: the head reads a ReLU output)
//   dW[i, t]  += sum_rows g[row, i] H[row, t],  db[i] += sum_rows g[row, i]      (atomic across blocks)
template <int NO>
__global__ void head_bwd_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ H, long long ldh,
                                int K, const float* __restrict__ W, long long rows, int rows_per_block, int mask_self,
                                float* __restrict__ dH, long long lddh, float* __restrict__ dW, float* __restrict__ db) {
  const int t = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
  float w[NO], accw[NO];
#pragma unroll
  for (int i = 0; i < NO; ++i) { w[i] = t < K ? W[i * K + t] : 0.f; accw[i] = 0.f; }
  if (t < K) {
    for (long long row = r0; row < r1; ++row) {
      const float h = H[row * ldh + t];
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < NO; ++i) {
        const float gi = __ldg(G + row * ldg + i);
        accw[i] = fmaf(gi, h, accw[i]);
        d = fmaf(gi, w[i], d);
      }
      if (mask_self && !(h > 0.f)) d = 0.f;
      dH[row * lddh + t] = d;
    }
    if (dW) {
#pragma unroll
      for (int i = 0; i < NO; ++i) atomicAdd(dW + i * K + t, accw[i]);
    }
  }
  if (db && t < NO) {
    float s = 0.f;
    for (long long row = r0; row < r1; ++row) s += G[row * ldg + t];
    atomicAdd(db + t, s);
  }
}

// db[n] += sum_rows G[row*ld + n]; blockDim.x >= N, a slice of rows per block
__global__ void colsum_kernel(const float* __restrict__ G, long long ld, int N, long long rows, int rows_per_block,
                              float* __restrict__ db) {
  const int n = threadIdx.x;
  if (n >= N) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
  float s = 0.f;
  for (long long row = r0; row < r1; ++row) s += G[row * ld + n];
  atomicAdd(db + n, s);
}

// Pose gradient: the two encoding-part dgrads (skip layer's [encoding | h] input and layer 0) as ONE product with the
// contraction dimensions concatenated: wcat [2W, P] = [W_skip[:, :P] ; W_0]  (row k = output unit k of the layer).
__global__ void stack_enc_weights_kernel(const float* __restrict__ w_skip, int ld_skip, const float* __restrict__ w0, int W, int P,
                                         float* __restrict__ wcat) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * W * P) return;
  const int k = idx / P, n = idx % P;
  wcat[idx] = k < W ? w_skip[(long long)k * ld_skip + n] : w0[(long long)(k - W) * P + n];
}

// The feature layer has no activation, so it folds into the views layer (as in the fused render kernel):
//   hv = relu([h | enc] [Wv_f Wf | Wv_e]^T + (Wv_f bf + bv)),   Wv = [Wv_f | Wv_e]  ([H, W + E]),  Wf [W, W]
// -> one GEMM with K = W + E instead of two, and no feature-layer dgrad / wgrad in the backward pass.
// wvf [H, LV] = [Wv_f Wf | Wv_e], bvf [H] = Wv_f bf + bv.  One thread per element of wvf, then of bvf.
__global__ void fold_views_train_kernel(const float* __restrict__ wv, const float* __restrict__ bv, const float* __restrict__ wf,
                                        const float* __restrict__ bf, int H, int W, int LV, float* __restrict__ wvf,
                                        float* __restrict__ bvf) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < H * LV) {
    const int h = idx / LV, c = idx % LV;
    if (c >= W) { wvf[idx] = wv[idx]; return; }
    float acc = 0.f;
    for (int m = 0; m < W; ++m) acc = fmaf(wv[(long long)h * LV + m], wf[(long long)m * W + c], acc);
    wvf[idx] = acc;
  } else if (idx < H * LV + H) {
    const int h = idx - H * LV;
    float acc = bv[h];
    for (int m = 0; m < W; ++m) acc = fmaf(wv[(long long)h * LV + m], bf[m], acc);
    bvf[h] = acc;
  }
}
// Backward of the fold, the parts that are not matrix products (those run as two small GEMMs, train_path.cuh):
//   g_wv[h, c >= W] += dwvf[h, c];   g_wv[h, m < W] += dbvf[h] bf[m];   g_bf[m] += sum_h wv[h, m] dbvf[h];   g_bv[h] += dbvf[h]
// One thread per output element: [H * LV | W | H]; any output pointer may be NULL.
__global__ void unfold_views_grads_kernel(const float* __restrict__ dwvf, const float* __restrict__ dbvf, const float* __restrict__ wv,
                                          const float* __restrict__ bf, int H, int W, int LV, float* g_wv, float* g_bv, float* g_bf) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < H * LV) {
    if (!g_wv) return;
    const int h = idx / LV, c = idx % LV;
    g_wv[idx] += c >= W ? dwvf[idx] : dbvf[h] * bf[c];
    return;
  }
  idx -= H * LV;
  if (idx < W) {
    if (!g_bf) return;
    float acc = 0.f;
    for (int h = 0; h < H; ++h) acc = fmaf(wv[(long long)h * LV + idx], dbvf[h], acc);
    g_bf[idx] += acc;
    return;
  }
  idx -= W;
  if (idx < H && g_bv) g_bv[idx] += dbvf[idx];
}

// ------------------------------------------------------------------------------------------------
// sampling positions
// ------------------------------------------------------------------------------------------------
// z_coarse[ray, s] from the forward pass's repaired near/far (a3)
__global__ void coarse_depths_kernel(const float* __restrict__ nearfar, const float* __restrict__ t_rand, int n_rays,
                                     int Sc, int lindisp, float* __restrict__ z) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_rays * Sc) return;
  const int ray = (int)(i / Sc), s = (int)(i % Sc);
  z[i] = coarse_depth(nearfar[2 * ray], nearfar[2 * ray + 1], s, Sc, lindisp, t_rand ? t_rand + (long long)ray * Sc : nullptr);
}

// ------------------------------------------------------------------------------------------------
// encodings of a block of rows (a4-a10), in the REFERENCE's column order so that the fp32 weight
// matrices are used as they are:
//   XS [row, 0 .. 18J)           = [k*J + j (k = 0 raw, 1+2f sin, 2+2f cos) | 15J + 3j + c]
//   VIN[row, W .. W+27J (+fc))   = [kk*3J + 3j + c | framecode]
// one thread per (row, joint)
// ------------------------------------------------------------------------------------------------
struct EncodeArgs {
  const float* rays;      // [N,8]
  const float* skts;      // [N,J,16], or [P,J,16] read through pose_idx
  const int* pose_idx;    // optional [N]
  int n_poses;            // rows of skts when pose_idx is set (indices are clamped); 0 = unchecked
  const float* z;         // [N,S] depths of this network's pass
  const float* cams;      // [N] or NULL
  const float* codes;     // [n_fc, fc_ch] or NULL
  int ray0, n_rays_blk, S, J, W, fc_ch, n_fc, vq;    // vq = view inputs per joint, 3 (1 + 2 multires_views)
  float tau_p, tau_v;
  float cut_p[kMaxJoints], cut_v[kMaxJoints];
  float* XS; long long ldxs;
  float* VIN; long long ldv;
};

__global__ void encode_rows_kernel(EncodeArgs e) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows = (long long)e.n_rays_blk * e.S;
  if (idx >= rows * e.J) return;
  const long long row = idx / e.J;
  const int j = (int)(idx % e.J), J = e.J;
  const int ray = e.ray0 + (int)(row / e.S), s = (int)(row % e.S);
  const float* rp = e.rays + (long long)ray * 8;
  const float zz = e.z[(long long)ray * e.S + s];
  const float p[3] = {rp[0] + rp[3] * zz, rp[1] + rp[4] * zz, rp[2] + rp[5] * zz};
  long long prow = e.pose_idx ? (long long)e.pose_idx[ray] : (long long)ray;
  if (e.pose_idx && e.n_poses > 0) prow = prow < 0 ? 0 : (prow >= e.n_poses ? e.n_poses - 1 : prow);
  const float* skt = e.skts + (prow * J + j) * 16;
  float f[kPtsPerJoint];
  const float v = encode_joint_pts(skt, p, e.tau_p, e.cut_p[j], f);
  float* xs = e.XS + row * e.ldxs;
#pragma unroll
  for (int k = 0; k < 1 + 2 * kF; ++k) xs[k * J + j] = f[k];
#pragma unroll
  for (int c = 0; c < 3; ++c) xs[(1 + 2 * kF) * J + 3 * j + c] = f[1 + 2 * kF + c];
  float T[kViewPerJoint];
  encode_joint_viewdir(skt, rp + 3, T);
  const float wv = cutoff_w(v, e.tau_v, e.cut_v[j]);
  float* vin = e.VIN + row * e.ldv + e.W;
#pragma unroll
  for (int q = 0; q < kViewPerJoint; ++q)
    if (q < e.vq) vin[(q / 3) * 3 * J + 3 * j + (q % 3)] = T[q] * wv;
  if (j == 0 && e.fc_ch > 0) {
    int cam = (int)e.cams[ray];
    cam = cam < 0 ? 0 : (cam >= e.n_fc ? e.n_fc - 1 : cam);
    for (int q = 0; q < e.fc_ch; ++q) vin[e.vq * J + q] = e.codes[(long long)cam * e.fc_ch + q];
  }
}

// ------------------------------------------------------------------------------------------------
// a12 backward: one thread per ray.  Recomputes alpha / transmittance / weights from raw, then walks the
// ray back to front.  Formulas = autograd of nerf.py:150-205 (cumprod backward in its division form,
// which is what torch uses when no factor is zero; the factors here are >= 1e-10).
// ------------------------------------------------------------------------------------------------
struct CompositeBwdArgs {
  const float* raw;       // [rays_blk * S, 4] (r, g, b, sigma) of this block's rows
  const float* z;         // [N,S]
  const float* rays;      // [N,8]
  const float* noise;     // [N,S] or NULL
  const float *g_rgb, *g_disp, *g_acc, *g_alpha;   // dL/d(outputs) of this pass: [N,3], [N], [N], [N,S]; NULL = 0
  int ray0, n_rays_blk, S, softplus;
  float B, shift;
  float* scratch;         // [rays_blk * S, 2] (alpha, transmittance)
  float* g_raw;           // [rays_blk * S, 4]
};

__global__ void composite_bwd_kernel(CompositeBwdArgs a) {
  const int rl = blockIdx.x * blockDim.x + threadIdx.x;
  if (rl >= a.n_rays_blk) return;
  const int ray = a.ray0 + rl, S = a.S;
  const float* rp = a.rays + (long long)ray * 8;
  const float dnorm = sqrtf(rp[3] * rp[3] + rp[4] * rp[4] + rp[5] * rp[5]);
  const float4* raw = reinterpret_cast<const float4*>(a.raw) + (long long)rl * S;
  const float* z = a.z + (long long)ray * S;
  const float* nz = a.noise ? a.noise + (long long)ray * S : nullptr;
  float* sc = a.scratch + (long long)rl * S * 2;
  float T = 1.f, depth = 0.f, wsum = 0.f;
  for (int i = 0; i < S; ++i) {
    const float dist = (i + 1 < S ? z[i + 1] - z[i] : 1e10f) * dnorm;
    const float sg = density_act(raw[i].w, a.B, nz ? nz[i] : 0.f, a.softplus, a.shift);
    const float al = 1.f - expf(-sg * dist);
    sc[2 * i] = al;
    sc[2 * i + 1] = T;
    const float w = al * T;
    depth = fmaf(w, z[i], depth);
    wsum += w;
    T *= (1.f - al + 1e-10f);
  }
  const float g_r = a.g_rgb ? a.g_rgb[(long long)ray * 3] : 0.f, g_g = a.g_rgb ? a.g_rgb[(long long)ray * 3 + 1] : 0.f,
              g_b = a.g_rgb ? a.g_rgb[(long long)ray * 3 + 2] : 0.f;
  // disp = 1 / max(1e-10, depth / (wsum + 1e-10)) (zero, with zero gradient, where wsum ~ 0); acc = min(wsum, 1)
  float g_depth = 0.f, g_wsum = 0.f;
  if (a.g_disp && !(fabsf(wsum) <= 1e-8f)) {
    const float den = wsum + 1e-10f, q = depth / den;
    if (q > 1e-10f) {
      const float gq = -a.g_disp[ray] / (q * q);
      g_depth = gq / den;
      g_wsum = -gq * depth / (den * den);
    }
  }
  if (a.g_acc && wsum < 1.f) g_wsum += a.g_acc[ray];
  const float* ga = a.g_alpha ? a.g_alpha + (long long)ray * S : nullptr;
  float4* gout = reinterpret_cast<float4*>(a.g_raw) + (long long)rl * S;
  float suffix = 0.f;                       // sum_{k > i} g_w[k] * w[k]
  for (int i = S - 1; i >= 0; --i) {
    const float al = sc[2 * i], Ti = sc[2 * i + 1], w = al * Ti;
    const float4 r = raw[i];
    const float sr = 1.f / (1.f + expf(-r.x)), sgn = 1.f / (1.f + expf(-r.y)), sb = 1.f / (1.f + expf(-r.z));
    const float cr = sr * 1.002f - 0.001f, cg = sgn * 1.002f - 0.001f, cb = sb * 1.002f - 0.001f;
    const float g_w = g_r * cr + g_g * cg + g_b * cb + g_depth * z[i] + g_wsum;
    const float g_al = (ga ? ga[i] : 0.f) + g_w * Ti - suffix / (1.f - al + 1e-10f);
    suffix = fmaf(g_w, w, suffix);
    const float dist = (i + 1 < S ? z[i + 1] - z[i] : 1e10f) * dnorm;
    const float pre = r.w / a.B + (nz ? nz[i] : 0.f);
    float dact;
    if (!a.softplus) dact = pre > 0.f ? 1.f : 0.f;
    else { const float x = pre - a.shift; dact = x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }
    // d alpha / d sigma = dist * exp(-sigma dist) = dist * (1 - alpha)
    const float g_sig = dact == 0.f ? 0.f : g_al * (dist * (1.f - al)) * dact / a.B;
    gout[i] = make_float4(w * g_r * 1.002f * sr * (1.f - sr), w * g_g * 1.002f * sgn * (1.f - sgn),
                          w * g_b * 1.002f * sb * (1.f - sb), g_sig);
  }
}

// ------------------------------------------------------------------------------------------------
// backward of the encodings into the bone transforms: one thread per (ray, joint), loop over the ray's
// samples.  gXS / gVIN are dL/d(XS) and dL/d(VIN) in the layouts of encode_rows_kernel.
//   d/dv   of [v, sin(2^f v), cos(2^f v)] w(v)  and of the view cutoff weight  w_v(v),  w = 1 - sigmoid(tau (v - c))
//   d/dx   of v = |x| and r = x / max(|x|, 1e-12)           (x = skt [p; 1])
//   d/dskt of x (rows 0..2, columns 0..3) and of the bone-local ray direction R d (rotation part)
// g_skts [N,J,16] is ADDED to (row 3 stays untouched).
// ------------------------------------------------------------------------------------------------
struct EncodeBwdArgs {
  const float* rays; const float* skts; const float* z;
  const int* pose_idx;    // optional [N]: skts / g_skts are then per pose
  int n_poses;
  int ray0, n_rays_blk, S, J, W, vq;
  float tau_p, tau_v;
  float cut_p[kMaxJoints], cut_v[kMaxJoints];
  const float* gXS; long long ldxs;
  const float* gVIN; long long ldv;
  float* g_skts;
};

__global__ void encode_bwd_kernel(EncodeBwdArgs e) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= e.n_rays_blk * e.J) return;
  const int rl = idx / e.J, j = idx % e.J, J = e.J, S = e.S;
  const int ray = e.ray0 + rl;
  const float* rp = e.rays + (long long)ray * 8;
  long long prow = e.pose_idx ? (long long)e.pose_idx[ray] : (long long)ray;
  if (e.pose_idx && e.n_poses > 0) prow = prow < 0 ? 0 : (prow >= e.n_poses ? e.n_poses - 1 : prow);
  const float* skt = e.skts + (prow * J + j) * 16;
  float T[kViewPerJoint], gT[kViewPerJoint];
  encode_joint_viewdir(skt, rp + 3, T);
#pragma unroll
  for (int q = 0; q < kViewPerJoint; ++q) gT[q] = 0.f;
  float gs[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) gs[i] = 0.f;
  for (int s = 0; s < S; ++s) {
    const long long row = (long long)rl * S + s;
    const float zz = e.z[(long long)ray * S + s];
    const float p[3] = {rp[0] + rp[3] * zz, rp[1] + rp[4] * zz, rp[2] + rp[5] * zz};
    float x[3];
    bone_local(skt, p, x);
    const float v = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const float inv = 1.0f / fmaxf(v, 1e-12f);
    const float w = cutoff_w(v, e.tau_p, e.cut_p[j]);
    const float dw = -e.tau_p * (1.f - w) * w;
    const float* gx = e.gXS + row * e.ldxs;
    // distance features
    float g0 = gx[j];
    float gv_w = g0, gv_dw = g0 * v;          // coefficients of w and of w' in dL/dv
#pragma unroll
    for (int f = 0; f < kF; ++f) {
      const float a = (float)(1 << f);
      float sn, cs;
      sincos_2pi(v * a, sn, cs);
      const float gsn = gx[(1 + 2 * f) * J + j], gcs = gx[(2 + 2 * f) * J + j];
      gv_w += a * (gsn * cs - gcs * sn);
      gv_dw += gsn * sn + gcs * cs;
    }
    float gv = gv_w * w + gv_dw * dw;
    // view features: VIN = T[q] * w_v(v)
    const float wv = cutoff_w(v, e.tau_v, e.cut_v[j]);
    const float dwv = -e.tau_v * (1.f - wv) * wv;
    const float* gvn = e.gVIN + row * e.ldv + e.W;
    float gwv = 0.f;
#pragma unroll
    for (int q = 0; q < kViewPerJoint; ++q) {
      const float gq = q < e.vq ? gvn[(q / 3) * 3 * J + 3 * j + (q % 3)] : 0.f;
      gwv = fmaf(gq, T[q], gwv);
      gT[q] = fmaf(gq, wv, gT[q]);
    }
    gv += gwv * dwv;
    // r = x * inv
    const float gr0 = gx[(1 + 2 * kF) * J + 3 * j], gr1 = gx[(1 + 2 * kF) * J + 3 * j + 1], gr2 = gx[(1 + 2 * kF) * J + 3 * j + 2];
    const float r0 = x[0] * inv, r1 = x[1] * inv, r2 = x[2] * inv;
    const float rg = r0 * gr0 + r1 * gr1 + r2 * gr2;
    const float live = v > 1e-12f ? 1.f : 0.f;       // below the clamp r = x * 1e12 and v's subgradient is 0
    float gxv[3];
    gxv[0] = live * ((gr0 - r0 * rg) * inv + gv * r0) + (1.f - live) * gr0 * inv;
    gxv[1] = live * ((gr1 - r1 * rg) * inv + gv * r1) + (1.f - live) * gr1 * inv;
    gxv[2] = live * ((gr2 - r2 * rg) * inv + gv * r2) + (1.f - live) * gr2 * inv;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gs[4 * i + 0] = fmaf(gxv[i], p[0], gs[4 * i + 0]);
      gs[4 * i + 1] = fmaf(gxv[i], p[1], gs[4 * i + 1]);
      gs[4 * i + 2] = fmaf(gxv[i], p[2], gs[4 * i + 2]);
      gs[4 * i + 3] += gxv[i];
    }
  }
  // per-ray direction features T[3*kk + c] of u = normalize(R d)
  {
    const float* d = rp + 3;
    float y[3];
    y[0] = skt[0] * d[0] + skt[1] * d[1] + skt[2] * d[2];
    y[1] = skt[4] * d[0] + skt[5] * d[1] + skt[6] * d[2];
    y[2] = skt[8] * d[0] + skt[9] * d[1] + skt[10] * d[2];
    const float n = sqrtf(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    const float inv = 1.0f / fmaxf(n, 1e-12f);
    float gu[3], u[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      u[c] = y[c] * inv;
      float g = gT[c];
#pragma unroll
      for (int f = 0; f < kFv; ++f) {
        const float a = (float)(1 << f);
        // T[3(1+2f)+c] = sin(a u), T[3(2+2f)+c] = cos(a u)
        g += a * (gT[3 * (1 + 2 * f) + c] * T[3 * (2 + 2 * f) + c] - gT[3 * (2 + 2 * f) + c] * T[3 * (1 + 2 * f) + c]);
      }
      gu[c] = g;
    }
    const float ug = u[0] * gu[0] + u[1] * gu[1] + u[2] * gu[2];
    const float live = n > 1e-12f ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float gy = (gu[i] - live * u[i] * ug) * inv;
      gs[4 * i + 0] = fmaf(gy, d[0], gs[4 * i + 0]);
      gs[4 * i + 1] = fmaf(gy, d[1], gs[4 * i + 1]);
      gs[4 * i + 2] = fmaf(gy, d[2], gs[4 * i + 2]);
    }
  }
  float* out = e.g_skts + (prow * J + j) * 16;
  if (e.pose_idx) {         // several rays share the pose: the segment sum of the reference's expand-backward, in place
#pragma unroll
    for (int i = 0; i < 12; ++i) atomicAdd(out + i, gs[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) out[i] += gs[i];
  }
}

// d codes[cam, q] += sum over the ray's samples of gVIN[row, W + 27J + q]; one thread per (ray, q)
__global__ void framecode_bwd_kernel(const float* __restrict__ gVIN, long long ldv, int col0, const float* __restrict__ cams,
                                     int ray0, int n_rays_blk, int S, int fc_ch, int n_fc, float* __restrict__ g_codes) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays_blk * fc_ch) return;
  const int rl = idx / fc_ch, q = idx % fc_ch;
  float s = 0.f;
  for (int i = 0; i < S; ++i) s += gVIN[((long long)rl * S + i) * ldv + col0 + q];
  int cam = (int)cams[ray0 + rl];
  cam = cam < 0 ? 0 : (cam >= n_fc ? n_fc - 1 : cam);
  atomicAdd(g_codes + (long long)cam * fc_ch + q, s);
}

}  // namespace train
}  // namespace anerf
