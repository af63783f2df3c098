// SURVEY.md 8(f) row 4 (second half): the training-ray sampler on the device (reference: BaseH5Dataset.__getitem__ ->
// sample_pixels / get_rays / get_img_data on the host, one image at a time through h5py and a DataLoader,
// core/dataset.py:57-105, 277-322, 346-362).  Images, masks, backgrounds and cameras stay resident in HBM (uint8, as the
// .h5 files store them); one CTA per selected image draws `k` DISTINCT pixels uniformly from the pixels whose sampling
// mask is set, returns them in increasing order (like np.sort(sampled_idxs)), and writes the rays and the colours.
//
// Sampling without replacement: every valid pixel gets the key hash(seed, image, pixel); the k smallest keys are taken
// (a uniformly random k-subset).  The CTA finds the k-th smallest key by bisection on the key value (32 counting sweeps
// over the mask), then compacts the selected pixels in index order with a block scan.  The random stream is ours
// (counter-based hash), not numpy's Mersenne twister: identical statistics, different draws -- unpinned by design.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "path_math.cuh"

namespace anerf {
namespace sampler {

struct SampleArgs {
  const uint8_t* masks;       // [F, HW] sampling masks (> 0 = may be sampled)
  const uint8_t* imgs;        // [F, HW, 3]
  const uint8_t* fgs;         // [F, HW] foreground masks (0/1 or 0/255 -> reported as value > 0 ? 1 : 0 ... see fg_scale) or NULL
  const uint8_t* bgs;         // [B, HW, 3] or NULL
  const int* bg_idx;          // [F] background of each image, or NULL (image index)
  const float* c2ws;          // [F, 3, 4] (rows 0..2 of the camera-to-world matrix)
  const float* focals;        // [F, 2] (fx, fy)
  const float* centers;       // [F, 2] or NULL (image centre)
  const int* frames;          // [n_img] selected image indices
  int n_img, k, H, W, n_frames;
  unsigned long long seed;
  float fg_scale;             // 1/255 when masks are stored as 0/255, 1 when 0/1
  int mask_img;               // compose img * fg + (1 - fg) * bg like the reference's mask_img option
  // outputs, n_img * k rows
  float* rays;                // [N, 8]: o, d, near = 0, far = 1
  float* target;              // [N, 3]
  float* fg_out;              // [N] or NULL
  float* bg_out;              // [N, 3] or NULL
  int* pixel_idx;             // [N]
  int* frame_of_ray;          // [N]
  int* status;                // [n_img]: number of valid pixels of the image (k > valid is an error the host reports)
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t pix_key(unsigned long long seed, int frame, int pix) {
  // splitmix64 finaliser over (seed, frame, pixel): a different, well-mixed 32-bit key per (image, pixel, call)
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(frame + 1) + ((unsigned long long)pix << 32 | (unsigned)pix);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}

__device__ __forceinline__ int block_sum(int v, int* s_warp) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_warp[i];
  return t;
}

// one CTA (1024 threads) per selected image; thread t owns the contiguous pixel range [t * per, (t + 1) * per)
__global__ void __launch_bounds__(1024, 1) sample_rays_kernel(const SampleArgs a) {
  __shared__ int s_warp[32];
  __shared__ int s_scan[1024];
  const int img = blockIdx.x, tid = threadIdx.x;
  const int frame = a.frames[img];
  if (a.n_frames > 0 && (frame < 0 || frame >= a.n_frames)) {      // a bad image index: report, touch nothing
    if (tid == 0) a.status[img] = -1;
    return;
  }
  const int HW = a.H * a.W;
  const int per = (HW + blockDim.x - 1) / blockDim.x;
  const int p0 = tid * per, p1 = min(p0 + per, HW);
  const uint8_t* mask = a.masks + (size_t)frame * HW;
  int valid = 0;
  for (int p = p0; p < p1; ++p) valid += mask[p] > 0;
  const int n_valid = block_sum(valid, s_warp);
  if (tid == 0) a.status[img] = n_valid;
  if (n_valid < a.k) return;                       // the host raises (the reference's np.random.choice would, too)
  // smallest threshold T with count(key <= T) >= k
  uint32_t lo = 0u, hi = 0xFFFFFFFFu;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    int c = 0;
    for (int p = p0; p < p1; ++p) c += (mask[p] > 0 && pix_key(a.seed, frame, p) <= mid);
    c = block_sum(c, s_warp);
    if (c >= a.k) hi = mid; else lo = mid + 1;
  }
  const uint32_t T = lo;
  // keys below T are all taken; of the keys equal to T (hash collisions) the first few in index order
  int below = 0, equal = 0;
  for (int p = p0; p < p1; ++p)
    if (mask[p] > 0) { const uint32_t key = pix_key(a.seed, frame, p); below += key < T; equal += key == T; }
  const int n_below = block_sum(below, s_warp);
  const int take_equal = a.k - n_below;            // >= 1
  // exclusive scans over the threads (index order) of `below` and `equal`
  auto excl_scan = [&](int v) {
    __syncthreads();
    s_scan[tid] = v;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
      int t = tid >= o ? s_scan[tid - o] : 0;
      __syncthreads();
      s_scan[tid] += t;
      __syncthreads();
    }
    return s_scan[tid] - v;
  };
  const int eq_before = excl_scan(equal);
  int mine = below + max(0, min(equal, take_equal - eq_before));
  int out = excl_scan(mine);
  int eq_seen = eq_before;
  RayGen g;
  const float* c2w = a.c2ws + (size_t)frame * 12;
#pragma unroll
  for (int i = 0; i < 12; ++i) g.c2w[i] = c2w[i];
  g.fx = a.focals[frame * 2]; g.fy = a.focals[frame * 2 + 1];
  g.cx = a.centers ? a.centers[frame * 2] : a.W * 0.5f;
  g.cy = a.centers ? a.centers[frame * 2 + 1] : a.H * 0.5f;
  g.near = 0.f; g.far = 1.f; g.W = a.W; g.pixel0 = 0; g.pixels = nullptr;
  const uint8_t* im = a.imgs + (size_t)frame * HW * 3;
  const uint8_t* fgp = a.fgs ? a.fgs + (size_t)frame * HW : nullptr;
  const uint8_t* bgp = a.bgs ? a.bgs + (size_t)(a.bg_idx ? a.bg_idx[frame] : frame) * HW * 3 : nullptr;
  for (int p = p0; p < p1; ++p) {
    if (!(mask[p] > 0)) continue;
    const uint32_t key = pix_key(a.seed, frame, p);
    bool take = key < T;
    if (key == T) { take = eq_seen < take_equal; ++eq_seen; }
    if (!take) continue;
    const size_t row = (size_t)img * a.k + out++;
    float r[8];
    pixel_ray(g, p, r);
    float* ro = a.rays + row * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) ro[i] = r[i];
    float rgb[3] = {im[(size_t)p * 3] / 255.f, im[(size_t)p * 3 + 1] / 255.f, im[(size_t)p * 3 + 2] / 255.f};
    const float fg = fgp ? (float)fgp[p] * a.fg_scale : 1.f;
    float bg[3] = {0.f, 0.f, 0.f};
    if (bgp) { bg[0] = bgp[(size_t)p * 3] / 255.f; bg[1] = bgp[(size_t)p * 3 + 1] / 255.f; bg[2] = bgp[(size_t)p * 3 + 2] / 255.f; }
    if (a.mask_img && fgp && bgp) {
#pragma unroll
      for (int i = 0; i < 3; ++i) rgb[i] = rgb[i] * fg + (1.f - fg) * bg[i];
    }
    a.target[row * 3] = rgb[0]; a.target[row * 3 + 1] = rgb[1]; a.target[row * 3 + 2] = rgb[2];
    if (a.fg_out) a.fg_out[row] = fg;
    if (a.bg_out) { a.bg_out[row * 3] = bg[0]; a.bg_out[row * 3 + 1] = bg[1]; a.bg_out[row * 3 + 2] = bg[2]; }
    a.pixel_idx[row] = p;
    a.frame_of_ray[row] = frame;
  }
}
#endif

}  // namespace sampler
}  // namespace anerf
