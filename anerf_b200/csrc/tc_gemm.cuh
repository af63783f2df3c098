// Split-precision tensor-core GEMM for the training path: C (op)= A * B^T with fp32 operands in global
// memory, each product issued as three bf16 tcgen05 MMAs (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM),
// built from the same pipeline pieces as the fused render kernel (render_kernels.cuh: A-operand ring filled by
// the 16 worker warps, weight/B ring streamed by one bulk-copy thread per CTA, cta_group::2 MMAs over a CTA
// pair with M = 256, two TMEM accumulator regions).
//
//   A(m, k)  fp32 at A[m*a_ms + k*a_ks]  (either stride may be 1; rows >= M and columns >= K read as zero)
//   B(n, k)  pre-packed by tc_pack_b_kernel into the UMMA K-major core-matrix layout, hi and lo parts,
//            in tiles of NT <= 256 rows (n) and chunks of 32 (k); rows >= N and columns >= K are zero
//   C(m, n)  fp32 at C[m*c_ms + n*c_ns]; epilogue: + bias[n], ReLU, ReLU-derivative mask, store / add / atomic add
//
// Work items = (k slice, m tile of 256 rows, n tile); CTA pair p takes items p, p + pairs, ...  The three
// GEMM forms of the backward pass map onto it as
//   forward  H  = relu(X W^T + b)      A = X (a_ks = 1),            B = W   [N = out, K = in]
//   dgrad    dX = (G W) . mask         A = G (a_ks = 1),            B = W^T [N = in,  K = out]
//   wgrad    dW^T = X^T G  (split-K)   A = X^T (a_ms = 1, rows = input features), B = G^T [N = out, K = rows],
//                                      C written transposed (c_ms = 1, c_ns = ld of dW) with atomic adds
#pragma once
#include "render_kernels.cuh"

namespace anerf {

struct TcGemmArgs {
  const float* A; long long a_ms, a_ks; int M, K;
  const uint8_t* Bp; int N, NT, n_tiles;
  int chunks_total;         // round_up(K, 128) / 32
  int k_slices, slice_chunks;   // split-K: slice s covers chunks [s*slice_chunks, min((s+1)*slice_chunks, chunks_total)); multiple of 4
  float* C; long long c_ms, c_ns;
  const float* bias;
  const float* mask; long long mask_ms;   // mask(m, n) at mask[m*mask_ms + n]
  int relu, mode;           // mode 0: store, 1: C += r (then relu / mask), 2: atomicAdd
  DeviceStatus* status;
};

inline __host__ __device__ int tc_n_tiles(int N) { return (N + 255) / 256; }
inline __host__ __device__ int tc_tile_width(int N) { return round_up(ceil_div(N, tc_n_tiles(N)), 32); }
inline __host__ __device__ int tc_chunks(int K) { return round_up(K, kGroups * kKC) / kKC; }
inline __host__ __device__ size_t tc_packed_bytes(int N, int K) {
  return (size_t)tc_n_tiles(N) * tc_chunks(K) * tc_tile_width(N) * 128;
}

#ifdef __CUDACC__

// B(n, k) = src[n*s_n + k*s_k] (zero outside [0,N) x [0,K)) -> packed tiles.  One thread per (tile, chunk, 8-wide k
// group, row of the tile); rows vary fastest so that both the strided reads (s_n == 1) and the 16-byte writes coalesce.
template <int FMT>
__global__ void tc_pack_b_kernel(const float* __restrict__ src, long long s_n, long long s_k, int N, int K, int NT,
                                 int n_tiles, int chunks_total, uint8_t* __restrict__ out) {
  const long long total = (long long)n_tiles * chunks_total * 4 * NT;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t % NT);
    const int g = (int)((t / NT) & 3);
    const long long cc = t / (4LL * NT);
    const int c = (int)(cc % chunks_total), tile = (int)(cc / chunks_total);
    const int n = tile * NT + r;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = c * kKC + g * 8 + i;
      x[i] = (n < N && k < K) ? __ldg(src + (long long)n * s_n + (long long)k * s_k) : 0.f;
    }
    uint4 hi, lo;
    Split<FMT>::pair(x[0], x[1], hi.x, lo.x);
    Split<FMT>::pair(x[2], x[3], hi.y, lo.y);
    Split<FMT>::pair(x[4], x[5], hi.z, lo.z);
    Split<FMT>::pair(x[6], x[7], hi.w, lo.w);
    // chunk = [half 0: hi, lo][half 1: hi, lo]; a half holds NT/2 rows (what one CTA of the pair feeds)
    const int nh = NT >> 1;
    uint8_t* chunk = out + ((size_t)tile * chunks_total + c) * NT * 128;
    uint8_t* half = chunk + (size_t)(r / nh) * NT * 64;
    const int rr = r % nh;
    const size_t off = (size_t)g * nh * 16 + (size_t)(rr >> 3) * 128 + (size_t)(rr & 7) * 16;
    *reinterpret_cast<uint4*>(half + off) = hi;
    *reinterpret_cast<uint4*>(half + (size_t)nh * 64 + off) = lo;
  }
}

template <int FMT>
__global__ void __launch_bounds__(kThreads, 1) tc_gemm_kernel(const __grid_constant__ TcGemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAStages * kAStageBytes + kBStages * kBStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float* zero_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);
  float* stage_all = zero_bias + 256;            // per worker warp: 32 rows x 36 floats (epilogue transpose)
  Pipe pp;
  pipe_init(pp, smem, smem + kAStages * kAStageBytes, bars, g.status);
  if (tid == 0) pipe_init_barriers(pp);
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  for (int i = tid; i < 256; i += kThreads) zero_bias[i] = 0.f;
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  pp.tmem_base = *tmem_slot;

  const int m_tiles = ceil_div(g.M, 2 * kTileM);
  const int items = g.k_slices * m_tiles * g.n_tiles;
  const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
  const int NT = g.NT;
  auto slice_len = [&](int ks) {
    const int c0 = ks * g.slice_chunks;
    return (c0 + g.slice_chunks <= g.chunks_total) ? g.slice_chunks : g.chunks_total - c0;
  };

  if (warp == kMmaWarp) {
    uint32_t a_seq = 0, b_seq = 0;
    int it = 0;
    if (pp.rank == 0) {
      for (int item = pair; item < items; item += n_pairs, ++it)
        mma_layer<FMT>(pp, a_seq, b_seq, NT, slice_len(item / (m_tiles * g.n_tiles)), it & 1);
    } else if (lane == 0) {
      for (int item = pair; item < items; item += n_pairs) relay_layer(pp, b_seq, slice_len(item / (m_tiles * g.n_tiles)));
    }
    __syncwarp();
  } else if (warp == kLoadWarp) {
    if (lane == 0) {
      uint32_t b_seq = 0;
      for (int item = pair; item < items; item += n_pairs) {
        const int ks = item / (m_tiles * g.n_tiles), nt = item % g.n_tiles;
        const uint8_t* src = g.Bp + ((size_t)nt * g.chunks_total + (size_t)ks * g.slice_chunks) * NT * 128;
        load_layer(pp, b_seq, src, NT, slice_len(ks));
      }
    }
    __syncwarp();
  } else {
    const int grp = warp >> 2, quarter = warp & 3, row = quarter * 32 + lane;
    AProducer<FMT> ap(pp, row);
    uint32_t d_cnt[2] = {0u, 0u};
    const bool vec_ok = g.a_ks == 1 && (g.a_ms & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    float* stg = stage_all + warp * (32 * 36);
    // vector epilogue: row-major output whose rows, bias and mask are 16-byte aligned, plain store / add modes
    const bool vec_out = g.c_ns == 1 && g.mode != 2 && (g.c_ms & 3) == 0 && (g.N & 3) == 0 &&
                         ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                         (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) &&
                         (!g.mask || ((g.mask_ms & 3) == 0 && (reinterpret_cast<uintptr_t>(g.mask) & 15) == 0));
    // epilogue of one finished item.  A thread holds one row of the accumulator (TMEM lane); when the output is
    // row-major (c_ns == 1) each 32 x 32 block goes through shared memory so that a warp writes (and, for the add
    // mode and the mask, reads) 128 contiguous bytes of one row per instruction; when the output is column-major
    // (wgrad, c_ms == 1) the lanes' rows are already adjacent in memory.
    auto drain_item = [&](int item, int region) {
      const int rem = item % (m_tiles * g.n_tiles);
      const int mt = rem / g.n_tiles, nt = rem % g.n_tiles;
      const int m_warp = mt * 2 * kTileM + (int)pp.rank * kTileM + quarter * 32;
      const int m = m_warp + lane;
      const int n0 = nt * NT;
      const bool transposed = g.c_ns == 1;
      drain_region<FMT, false, false>(ap, pp, d_cnt, region, NT, zero_bias, 1.0f, quarter, grp,
                                      [&](int col0, const float (&x)[8]) {
        if (!transposed) {
          if (m >= g.M) return;
          float* c = g.C + (long long)m * g.c_ms + (long long)(n0 + col0) * g.c_ns;
#pragma unroll
          for (int i = 0; i < 8; ++i, c += g.c_ns) {
            const int n = n0 + col0 + i;
            if (n >= g.N) continue;
            float r = x[i];
            if (g.mode == 2) { atomicAdd(c, r); continue; }
            if (g.bias) r += __ldg(g.bias + n);
            if (g.mode == 1) r += *c;
            if (g.relu) r = fmaxf(r, 0.f);
            if (g.mask && !(__ldg(g.mask + (long long)m * g.mask_ms + n) > 0.f)) r = 0.f;
            *c = r;
          }
          return;
        }
        // stage the thread's 8 columns (row stride 36 floats: 16-byte aligned and conflict-free for both phases)
        const int cin = col0 & 31;
        *reinterpret_cast<float4*>(stg + lane * 36 + cin) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(stg + lane * 36 + cin + 4) = make_float4(x[4], x[5], x[6], x[7]);
        if (cin != 24) return;                 // the 32-column block is complete after its fourth call
        __syncwarp();
        const int nb = n0 + (col0 - 24);       // first column of the block
        const int rmax = g.M - m_warp < 32 ? g.M - m_warp : 32;
        if (vec_out) {
          // a warp instruction covers 4 rows x 128 contiguous bytes: lane = (row % 4, 4-column group)
          const int r4 = lane >> 3, n = nb + (lane & 7) * 4;
          const bool n_ok = n < g.N;            // N is a multiple of 4 here: a group is inside or outside as a whole
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g.bias && n_ok) b = __ldg(reinterpret_cast<const float4*>(g.bias + n));
          float* crow = g.C + (long long)(m_warp + r4) * g.c_ms + n;
          const float* mrow = g.mask ? g.mask + (long long)(m_warp + r4) * g.mask_ms + n : nullptr;
          const long long cstep = 4 * g.c_ms, mstep = g.mask ? 4 * g.mask_ms : 0;
#pragma unroll 2
          for (int r = r4; r < rmax; r += 4, crow += cstep, mrow += mstep) {
            float4 v = *reinterpret_cast<const float4*>(stg + r * 36 + (lane & 7) * 4);
            if (!n_ok) continue;
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            if (g.mode == 1) { const float4 o = *reinterpret_cast<const float4*>(crow); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            if (g.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if (g.mask) {
              const float4 k = __ldg(reinterpret_cast<const float4*>(mrow));
              v.x = k.x > 0.f ? v.x : 0.f; v.y = k.y > 0.f ? v.y : 0.f; v.z = k.z > 0.f ? v.z : 0.f; v.w = k.w > 0.f ? v.w : 0.f;
            }
            *reinterpret_cast<float4*>(crow) = v;
          }
        } else {
          const int n = nb + lane;
          const bool n_ok = n < g.N;
          const float b = (g.bias && n_ok) ? __ldg(g.bias + n) : 0.f;
#pragma unroll 4
          for (int r = 0; r < rmax; ++r) {
            float v = stg[r * 36 + lane];
            if (n_ok) {
              float* c = g.C + (long long)(m_warp + r) * g.c_ms + n;
              if (g.mode == 2) { atomicAdd(c, v); continue; }
              v += b;
              if (g.mode == 1) v += *c;
              if (g.relu) v = fmaxf(v, 0.f);
              if (g.mask && !(__ldg(g.mask + (long long)(m_warp + r) * g.mask_ms + n) > 0.f)) v = 0.f;
              *c = v;
            }
          }
        }
        __syncwarp();
      });
    };
    int it = 0, prev = -1;
    for (int item = pair; item < items; item += n_pairs, ++it) {
      const int ks = item / (m_tiles * g.n_tiles);
      const int mt = (item % (m_tiles * g.n_tiles)) / g.n_tiles;
      const int m = mt * 2 * kTileM + (int)pp.rank * kTileM + row;
      const int chunks = slice_len(ks);
      const int k_base = ks * g.slice_chunks * kKC;
      const bool m_ok = m < g.M;
      const float* arow = g.A + (long long)(m_ok ? m : 0) * g.a_ms;
      // ---- this group's chunks of the A operand
#pragma unroll 1
      for (int c = grp; c < chunks; c += kGroups) {
        ap.begin(c);
        const int k0 = k_base + c * kKC;
        if (vec_ok) {
          // row-major A: a warp instruction reads 8 rows x 128 contiguous bytes (lane = (row % 8, 8-wide k group)) and
          // writes one 128-byte core-matrix row group per k group, instead of 32 scattered 16-byte pieces
          const int t = lane & 3;
          const int k = k0 + t * 8;
          uint8_t* st0 = pp.a_ring + ap.cur * kAStageBytes + (t >> 1) * 4096 + (t & 1) * 2048;
          const int m_q = mt * 2 * kTileM + (int)pp.rank * kTileM + quarter * 32;
#pragma unroll
          for (int it4 = 0; it4 < 4; ++it4) {
            const int rl = it4 * 8 + (lane >> 2);            // row inside the warp's 32
            const int mm = m_q + rl;
            float x[8];
            if (mm < g.M && k + 8 <= g.K) {
              const float4* src = reinterpret_cast<const float4*>(g.A + (long long)mm * g.a_ms + k);
              const float4 u = __ldg(src), v = __ldg(src + 1);
              x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = (mm < g.M && k + i < g.K) ? __ldg(g.A + (long long)mm * g.a_ms + k + i) : 0.f;
            }
            uint4 hi, lo;
            Split<FMT>::pair(x[0], x[1], hi.x, lo.x);
            Split<FMT>::pair(x[2], x[3], hi.y, lo.y);
            Split<FMT>::pair(x[4], x[5], hi.z, lo.z);
            Split<FMT>::pair(x[6], x[7], hi.w, lo.w);
            const int rr = quarter * 32 + rl;
            uint8_t* p = st0 + (rr >> 3) * 128 + (rr & 7) * 16;
            *reinterpret_cast<uint4*>(p) = hi;
            *reinterpret_cast<uint4*>(p + kAHalfBytes) = lo;
          }
          ap.end();
          continue;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float x[8];
          const int k = k0 + t * 8;
          if (!m_ok || k >= g.K) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 0.f;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = (k + i < g.K) ? __ldg(arow + (long long)(k + i) * g.a_ks) : 0.f;
          }
          ap.store8(t, x);
        }
        ap.end();
      }
      ap.base += chunks;
      // ---- epilogue of the previous item while this item's MMAs run
      if (prev >= 0) drain_item(prev, (it - 1) & 1);
      prev = item;
      // the region of item it+1 is the one just drained: every group must be done with it before any group
      // publishes a chunk of item it+1 (whose first MMA overwrites that region)
      worker_sync();
    }
    if (prev >= 0) drain_item(prev, (it - 1) & 1);
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc(pp.tmem_base, kTmemCols);
}

#endif  // __CUDACC__
}  // namespace anerf
