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
                                 int n_tiles, int chunks_total, uint8_t* __restrict__ out, float* __restrict__ rowsum) {
  // rowsum (optional, single-tile operands only): rowsum[n] += sum_k B(n, k) -- the bias gradient when B is a
  // gradient matrix G^T; the launch keeps gridDim*blockDim a multiple of NT, so a thread's row n never changes
  float acc = 0.f;
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
    acc += ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
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
  if (rowsum) {
    const int n = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % NT);
    if (n < N && acc != 0.f) atomicAdd(rowsum + n, acc);
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
  if (tid == 0) {
    // as pipe_init_barriers(), except that an A chunk is produced by all 16 worker warps of both CTAs here
    // (in the fused kernel: by one group of 4 warps per CTA)
    for (int i = 0; i < kAStages; ++i) { mbar_init(&pp.a_full[i], 2 * kWorkerWarps); mbar_init(&pp.a_empty[i], 1); }
    for (int i = 0; i < kBStages; ++i) { mbar_init(&pp.b_full[i], 1); mbar_init(&pp.b_empty[i], 1); mbar_init(&pp.peer_b[i], 1); }
    mbar_init(&pp.d_full[0], 1);
    mbar_init(&pp.d_full[1], 1);
    fence_mbar_init();
  }
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
          const float lo = g.relu ? 0.f : -3.0e38f;          // branch-free ReLU
          const float* sp = stg + r4 * 36 + (lane & 7) * 4;
          if (n_ok) {
            if (!g.mask && g.mode == 0) {                     // forward form: bias, ReLU, store
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (r4 + 4 * i < rmax) {
                  float4 v = *reinterpret_cast<const float4*>(sp + i * 4 * 36);
                  v.x = fmaxf(v.x + b.x, lo); v.y = fmaxf(v.y + b.y, lo); v.z = fmaxf(v.z + b.z, lo); v.w = fmaxf(v.w + b.w, lo);
                  *reinterpret_cast<float4*>(crow + i * cstep) = v;
                }
              }
            } else {                                          // dgrad form: optional add to C, ReLU-derivative mask
#pragma unroll 4
              for (int i = 0; i < 8; ++i) {
                if (r4 + 4 * i < rmax) {
                  float4 v = *reinterpret_cast<const float4*>(sp + i * 4 * 36);
                  v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                  if (g.mode == 1) { const float4 o = *reinterpret_cast<const float4*>(crow + i * cstep); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                  v.x = fmaxf(v.x, lo); v.y = fmaxf(v.y, lo); v.z = fmaxf(v.z, lo); v.w = fmaxf(v.w, lo);
                  if (g.mask) {
                    const float4 k = __ldg(reinterpret_cast<const float4*>(mrow + i * mstep));
                    v.x = k.x > 0.f ? v.x : 0.f; v.y = k.y > 0.f ? v.y : 0.f; v.z = k.z > 0.f ? v.z : 0.f; v.w = k.w > 0.f ? v.w : 0.f;
                  }
                  *reinterpret_cast<float4*>(crow + i * cstep) = v;
                }
              }
            }
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
    // ---- A-operand production.  Every worker warp contributes 1/16 of EVERY chunk (8 values per lane), so a lane
    // holds only 8 registers per chunk and keeps the loads of the next three chunks in flight; loads go to
    // registers and are issued before the wait for the ring stage, so the global-memory round trip overlaps the
    // tensor core consuming earlier chunks.  (Each chunk's "full" barrier therefore counts all 16 warps of both CTAs.)
    //   row-major A (a_ks == 1): lane = (row % 8, 8-wide k group): a warp reads 8 rows x 128 contiguous bytes
    //   otherwise              : lane = row within a 32-row quarter, warp / 4 = k group (coalesced when a_ms == 1)
    const int a_t = vec_ok ? (lane & 3) : (warp >> 2);
    const int a_row = vec_ok ? (warp * 8 + (lane >> 2)) : ((warp & 3) * 32 + lane);
    const uint32_t a_off = (uint32_t)((a_t >> 1) * 4096 + (a_t & 1) * 2048 + (a_row >> 3) * 128 + (a_row & 7) * 16);
    uint32_t a_seq = 0;                              // chunks published so far (all items)
    // per item: pointer to this lane's 8 values of chunk 0 (NULL for a row past M) and the k index they start at
    const long long a_cstep = (long long)kKC * g.a_ks;      // one chunk further along k
    auto item_ptr = [&](int item, int& k0) -> const float* {
      const int ks = item / (m_tiles * g.n_tiles);
      const int mt = (item % (m_tiles * g.n_tiles)) / g.n_tiles;
      const int mrow = mt * 2 * kTileM + (int)pp.rank * kTileM + a_row;
      k0 = ks * g.slice_chunks * kKC + a_t * 8;
      return mrow < g.M ? g.A + (long long)mrow * g.a_ms + (long long)k0 * g.a_ks : nullptr;
    };
    auto load8 = [&](const float* src, int k, float (&x)[8]) {
      if (src == nullptr || k >= g.K) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
      } else if (vec_ok && k + 8 <= g.K) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(src)), v = __ldg(reinterpret_cast<const float4*>(src) + 1);
        x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = (k + i < g.K) ? __ldg(src + (long long)i * g.a_ks) : 0.f;
      }
    };
    // pull a line that will be loaded a few chunks from now into L2 (no register cost)
    auto prefetch = [&](const float* src, int k) {
      if (src != nullptr && k < g.K) asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
    };
    auto emit = [&](const float (&x)[8]) {
      const uint32_t stage = a_seq % kAStages;
      mbar_wait_warp(&pp.a_empty[stage], ((a_seq / kAStages) & 1) ^ 1, pp.st, 100 + stage);
      uint4 hi, lo;
      Split<FMT>::pair(x[0], x[1], hi.x, lo.x);
      Split<FMT>::pair(x[2], x[3], hi.y, lo.y);
      Split<FMT>::pair(x[4], x[5], hi.z, lo.z);
      Split<FMT>::pair(x[6], x[7], hi.w, lo.w);
      uint8_t* p = pp.a_ring + stage * kAStageBytes + a_off;
      *reinterpret_cast<uint4*>(p) = hi;
      *reinterpret_cast<uint4*>(p + kAHalfBytes) = lo;
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) a_chunk_ready(pp, stage);
      __syncwarp();
      ++a_seq;
    };
    int it = 0, prev = -1;
    float x0[8], x1[8], x2[8], x3[8];
    int k0 = 0, k0n = 0;
    const float* ap0 = nullptr;                      // this item's lane pointer
    const float* apn = nullptr;                      // next item's
    if (pair < items) {                              // first three chunks of the first item
      ap0 = item_ptr(pair, k0);
      load8(ap0, k0, x0); load8(ap0 ? ap0 + a_cstep : nullptr, k0 + kKC, x1); load8(ap0 ? ap0 + 2 * a_cstep : nullptr, k0 + 2 * kKC, x2);
    }
    for (int item = pair; item < items; item += n_pairs, ++it) {
      const int ks = item / (m_tiles * g.n_tiles);
      const int chunks = slice_len(ks);              // multiple of 4
      const int nxt = item + n_pairs;
      apn = nxt < items ? item_ptr(nxt, k0n) : nullptr;
      auto at = [&](int c) -> const float* { return ap0 ? ap0 + (long long)c * a_cstep : nullptr; };
#pragma unroll 1
      for (int c = 0; c < chunks; c += 4) {
        // L2 prefetch two rounds ahead: this item's chunks c+8..c+11, or the head of the next item
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cc = c + 8 + u;
          if (cc < chunks) prefetch(at(cc), k0 + cc * kKC);
          else if (apn && cc - chunks < 8) prefetch(apn + (long long)(cc - chunks) * a_cstep, k0n + (cc - chunks) * kKC);
        }
        load8(at(c + 3), k0 + (c + 3) * kKC, x3);
        emit(x0);
        if (c + 4 < chunks) load8(at(c + 4), k0 + (c + 4) * kKC, x0);
        emit(x1);
        if (c + 5 < chunks) load8(at(c + 5), k0 + (c + 5) * kKC, x1);
        emit(x2);
        if (c + 6 < chunks) load8(at(c + 6), k0 + (c + 6) * kKC, x2);
        emit(x3);
      }
      // the next item's first chunks are requested before the epilogue below, which hides their latency
      if (nxt < items) {
        load8(apn, k0n, x0); load8(apn ? apn + a_cstep : nullptr, k0n + kKC, x1); load8(apn ? apn + 2 * a_cstep : nullptr, k0n + 2 * kKC, x2);
      }
      ap0 = apn; k0 = k0n;
      // ---- epilogue of the previous item while this item's MMAs run
      if (prev >= 0) drain_item(prev, (it - 1) & 1);
      prev = item;
      // the region of item it+1 is the one just drained: every warp must be done with it before any warp
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
