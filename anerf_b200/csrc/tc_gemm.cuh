// Split-precision tensor-core GEMM for the training path: C (op)= A * B^T with fp32 operands in global
// memory, each product issued as three 16-bit tcgen05 MMAs (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM).
// Operand format (template FMT): 0 = fp16 hi/lo (22 mantissa bits per value) with a power-of-two scale PER OPERAND
// MATRIX chosen on the device from the matrix's max |x| (every producer kernel leaves that maximum in an "amax slot",
// see AmaxRef), so that gradients of any magnitude sit in fp16's range with normal lo parts; 1 = bf16 hi/lo (16 bits,
// no scaling needed; round 1's engine, kept as ANERF_TRAIN_GEMM=bf16).  The epilogue undoes the scales exactly and
// multiplies by the expected loss of the tensor core's truncating accumulation (render_kernels.cuh: trunc_comp).  Kernel
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

// max |x| of an operand matrix, kept on the device (written with atomicMax on the bit pattern by the kernel that
// produced the matrix); an operand made of two buffers (cat[encoding, h]) takes the larger of two slots
struct AmaxRef {
  const float* p0;
  const float* p1;
};
// power of two that brings `amax` into [2^12, 2^13) (exact to apply and to undo); 1 for an empty / non-finite matrix
inline __host__ __device__ float tc_operand_scale(float amax) {
  if (!(amax > 0.f) || !(amax < 3.0e38f)) return 1.0f;
  int e;
  frexpf(amax, &e);                        // amax = f * 2^e, f in [0.5, 1)
  int k = 13 - e;
  k = k < -100 ? -100 : (k > 100 ? 100 : k);
  return ldexpf(1.0f, k);
}

struct TcGemmArgs {
  const float* A; long long a_ms, a_ks; int M, K;
  AmaxRef a_amax, b_amax;   // fp16 format: device maxima the operand scales are derived from (NULL pointers: scale 1)
  float* c_amax;            // optional: max |C| of what this launch stores (modes 0 / 1) is atomically folded in here
  float comp;               // epilogue factor: 1 + expected relative loss of the truncating accumulation
  const uint8_t* Bp; int N, NT, n_tiles;
  int chunks_total;         // round_up(K, 128) / 32
  int k_slices, slice_chunks;   // split-K: slice s covers chunks [s*slice_chunks, min((s+1)*slice_chunks, chunks_total)); multiple of 4
  float* C; long long c_ms, c_ns;
  const float* bias;
  const float* mask; long long mask_ms;   // mask(m, n) at mask[m*mask_ms + n]
  int relu, mode;           // mode 0: store, 1: C += r (then relu / mask), 2: atomicAdd
  DeviceStatus* status;
  int debug;                // timing experiments only (ANERF_TC_DEBUG; results are WRONG): bit 0 skips the epilogue's global stores, bit 1 the producers' global loads, bit 2 copies 16 bytes per B chunk
  long long* trace;         // optional debug timeline of CTA 0 (anerf_debug_set_trace): stream 0 MMA warp, 1 worker warp 0, 2 worker warp 5
};

inline __host__ __device__ int tc_n_tiles(int N) { return (N + 255) / 256; }
inline __host__ __device__ int tc_tile_width(int N) { return round_up(ceil_div(N, tc_n_tiles(N)), 32); }
inline __host__ __device__ int tc_chunks(int K) { return round_up(K, kGroups * kKC) / kKC; }
inline __host__ __device__ size_t tc_packed_bytes(int N, int K) {
  return (size_t)tc_n_tiles(N) * tc_chunks(K) * tc_tile_width(N) * 128;
}

#ifdef __CUDACC__

inline __host__ __device__ int tc_gemm_smem_bytes() {
  return kAStages * kAStageBytes + kBStages * kBStageBytes + 8 * (kNumBars + 2) + 16 + 64 + 8 * 32 * 36 * 4;
}

constexpr int kTcProdWarps = 8;    // worker warps 0..7 produce the A operand
constexpr int kTcDrainWarps = 8;   // worker warps 8..15 drain the accumulators
static_assert(kTcProdWarps + kTcDrainWarps == kWorkerWarps, "worker warp roles");

// B(n, k) = src[n*s_n + k*s_k] (zero outside [0,N) x [0,K)) -> packed tiles.  One thread per (tile, chunk, 8-wide k
// group, row of the tile); rows vary fastest so that both the strided reads (s_n == 1) and the 16-byte writes coalesce.
// explicit shared-memory accesses for the epilogue's transposition buffer: its pointer is derived through an integer
// alignment cast, which made the compiler emit GENERIC loads / stores (LD.E / ST.E, ~250 cycles each in a dependent chain:
// the store loop of one 32 x 32 block took ~2000 cycles, tools/trace_gemm.py)
__device__ __forceinline__ void sts_f4(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float amax_value(const AmaxRef& r) {
  float a = r.p0 ? __ldg(r.p0) : 0.f;
  if (r.p1) a = fmaxf(a, __ldg(r.p1));
  return a;
}
__device__ __forceinline__ void amax_fold(float* slot, float v) {      // v >= 0: the bit patterns order like the values
  atomicMax(reinterpret_cast<int*>(slot), __float_as_int(v));
}

// max |B(n, k)| of a strided matrix -> slot (weights: run once per pack; activations and gradients get theirs from
// the kernels that produce them)
__global__ void tc_absmax_kernel(const float* __restrict__ src, long long s_n, long long s_k, int N, int K, float* __restrict__ slot) {
  float m = 0.f;
  const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
  if (s_k == 1 && (K & 3) == 0 && (s_n & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // rows of K contiguous floats: 16-byte loads, one division per four elements
    const int k4 = K >> 2;
    const long long total = (long long)N * k4;
    for (long long t = gid; t < total; t += gsz) {
      const long long n = t / k4;
      const int q = (int)(t - n * k4);
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + n * s_n) + q);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  } else {
    const long long total = (long long)N * K;
    for (long long t = gid; t < total; t += gsz) {
      const long long n = s_k == 1 ? t / K : t % N, k = s_k == 1 ? t % K : t / N;
      m = fmaxf(m, fabsf(__ldg(src + n * s_n + k * s_k)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) amax_fold(slot, m);
}

template <int FMT>
__global__ void tc_pack_b_kernel(const float* __restrict__ src, long long s_n, long long s_k, int N, int K, int NT,
                                 int n_tiles, int chunks_total, uint8_t* __restrict__ out, float* __restrict__ rowsum,
                                 AmaxRef amax) {
  const float sb = FMT == 0 ? tc_operand_scale(amax_value(amax)) : 1.0f;
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
    if (FMT == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] *= sb;
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
  uint64_t* r_free = bars + kNumBars;            // [2] (leader's copy is used): the drain warps of the pair are done with accumulator region r
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_free + 2);
  float* stage_all = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);
  Pipe pp;
  pipe_init(pp, smem, smem + kAStages * kAStageBytes, bars, g.status);
  if (tid == 0) {
    pipe_init_barriers(pp);                      // an A chunk is produced by one group of 4 warps per CTA, as in the fused kernel
    mbar_init(&r_free[0], 2 * kTcDrainWarps);      // the drain warps of BOTH CTAs arrive at the leader's barrier
    mbar_init(&r_free[1], 2 * kTcDrainWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  pp.tmem_base = *tmem_slot;

  // operand scales (fp16 format) and the epilogue factor that undoes them
  const float sa = FMT == 0 ? tc_operand_scale(amax_value(g.a_amax)) : 1.0f;
  const float out_scale = (FMT == 0 ? (1.0f / sa) * (1.0f / tc_operand_scale(amax_value(g.b_amax))) : 1.0f) * g.comp;

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
    Trace trc; trc.init(lane == 0 ? g.trace : nullptr, 0);
    if (pp.rank == 0) {
      for (int item = pair; item < items; item += n_pairs, ++it) {
        // the accumulator region this item's MMAs overwrite must have been drained (two items ago) in both CTAs.  The
        // wait sits HERE, not in front of the producers: they run ahead into the A ring while the drain finishes
        // (with the wait on their side they idled ~5 k of every ~31 k cycles, tools/trace_gemm.py)
        if (it >= 2) {
          mbar_wait_warp(&r_free[it & 1], ((it >> 1) - 1) & 1, pp.st, 700 + (it & 1));
          tc_fence_after_sync();
        }
        mma_layer<FMT>(pp, a_seq, b_seq, NT, slice_len(item / (m_tiles * g.n_tiles)), it & 1, trc.p ? &trc : nullptr);
      }
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
        load_layer(pp, b_seq, src, NT, slice_len(ks), (g.debug & 4) ? 16u : 0u);
      }
    }
    __syncwarp();
  } else if (warp < kTcProdWarps) {
    // ------------------------------------------------------------------------------------------------------
    // producers: two groups of 4 warps; group pg fills the chunks c with c % 2 == pg of every item, so two chunks
    // are in production at any time.  The loads of a chunk go to registers BEFORE the wait for its ring stage and
    // are consumed before the arrive (whose release semantics would otherwise wait for loads still in flight);
    // everything further ahead is only pulled into L2 (prefetch.global.L2, no register or ordering cost).
    //   row-major A (a_ks == 1): lane = (8-wide k group, row % 8): a warp reads 8 rows x 128 contiguous bytes
    //   otherwise              : lane = row, 4 k groups per lane (coalesced along the rows when a_ms == 1)
    // ------------------------------------------------------------------------------------------------------
    const int pg = warp >> 2;
    const int lane_g = (warp & 3) * 32 + lane;                       // 0..127 within the group
    const bool vec_ok = g.a_ks == 1 && (g.a_ms & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    Trace trc; trc.init((lane == 0 && warp == 0) ? g.trace : nullptr, 1);
    Trace* tr = trc.p ? &trc : nullptr;
    // slot j (0..3) of this lane: row a_row(j), k group a_t(j)
    // (row-major: the 8 lanes of a quarter-warp take 8 DIFFERENT rows of one k group -- the packed stores of a phase then
    // fall into 8 different 16-byte bank groups; with lane = (row, k group) in the other order every packed store was a
    // 4-way bank conflict, 5.8 M conflict cycles per 196,608-row GEMM in ncu)
    auto a_row = [&](int j) { return vec_ok ? (32 * j + 8 * (lane_g >> 5) + (lane_g & 7)) : lane_g; };
    auto a_t = [&](int j) { return vec_ok ? ((lane_g >> 3) & 3) : j; };
    const long long a_cstep = (long long)kKC * g.a_ks;
    auto load8 = [&](const float* src, int k, float (&x)[8]) {
      if (src == nullptr || k >= g.K || (g.debug & 2)) {
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
    uint32_t seq_base = 0;
    int it = 0;
    // where slot j lands inside a packed A stage: slot 0's place + a per-slot constant
    const int q0 = (a_t(0) >> 1) * 4096 + (a_t(0) & 1) * 2048 + (a_row(0) >> 3) * 128 + (a_row(0) & 7) * 16;
    auto qoff = [&](int j) { return q0 + (vec_ok ? j * 512 : (j >> 1) * 4096 + (j & 1) * 2048); };
    for (int item = pair; item < items; item += n_pairs, ++it) {
      const int ks = item / (m_tiles * g.n_tiles);
      const int mt = (item % (m_tiles * g.n_tiles)) / g.n_tiles;
      const int chunks = slice_len(ks);
      const int m0 = mt * 2 * kTileM + (int)pp.rank * kTileM;
      const int kbase = ks * g.slice_chunks * kKC;
      // per slot: pointer to the 8 values of chunk 0 (NULL past the last row) and their k index
      const float* sp[4];
      int sk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int mrow = m0 + a_row(j);
        sk[j] = kbase + a_t(j) * 8;
        sp[j] = mrow < g.M ? g.A + (long long)mrow * g.a_ms + (long long)sk[j] * g.a_ks : nullptr;
      }
      // next item's head, for the L2 prefetch
      const int nxt = item + n_pairs;
      const float* np[4] = {nullptr, nullptr, nullptr, nullptr};
      int nk[4] = {0, 0, 0, 0};
      if (nxt < items) {
        const int ks2 = nxt / (m_tiles * g.n_tiles), mt2 = (nxt % (m_tiles * g.n_tiles)) / g.n_tiles;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int mrow2 = mt2 * 2 * kTileM + (int)pp.rank * kTileM + a_row(j);
          nk[j] = ks2 * g.slice_chunks * kKC + a_t(j) * 8;
          if (mrow2 < g.M) np[j] = g.A + (long long)mrow2 * g.a_ms + (long long)nk[j] * g.a_ks;
        }
      }
      if (tr) tr->mark(50);
      // interior item (every row of the tile and every k of the slice exists): no per-element predicates, one running
      // pointer -- the loop head was ~1000 of a chunk's ~2800 cycles with the general address arithmetic
      const bool interior = m0 + kTileM <= g.M && kbase + chunks * kKC <= g.K && !(g.debug & 2) && (vec_ok || g.a_ms == 1);
      const float* rp = interior ? sp[0] + (long long)pg * a_cstep : nullptr;     // slot 0, this group's first chunk
      const long long slot_step = vec_ok ? 32 * g.a_ms : 8 * g.a_ks;              // slot j is slot 0 + j * slot_step
#pragma unroll 1
      for (int c = pg; c < chunks; c += 2) {
        float x[4][8];
        if (interior) {
          if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4* q4 = reinterpret_cast<const float4*>(rp + j * slot_step);
              const float4 u = __ldg(q4), v = __ldg(q4 + 1);
              x[j][0] = u.x; x[j][1] = u.y; x[j][2] = u.z; x[j][3] = u.w; x[j][4] = v.x; x[j][5] = v.y; x[j][6] = v.z; x[j][7] = v.w;
            }
            if (c + 4 < chunks) {
#pragma unroll
              for (int j = 0; j < 4; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + j * slot_step + 4 * kKC));
            }
          } else {                               // transposed operand: 32 consecutive k of this thread's row, a_ks apart
            const float* t = rp;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { x[j][i] = __ldg(t); t += g.a_ks; }
            }
            // two rounds ahead: lane i asks for the 128-byte line that holds the warp's 32 rows at k + i
            if (c + 4 < chunks) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 4 * a_cstep + (long long)lane * (g.a_ks - 1)));
          }
          rp += 2 * a_cstep;
        } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) load8(sp[j] ? sp[j] + (long long)c * a_cstep : nullptr, sk[j] + c * kKC, x[j]);
        // L2 prefetch, two of this group's rounds ahead (or the head of the next item)
        {
          const int cc = c + 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float* q = nullptr;
            if (cc < chunks) { if (sp[j] && sk[j] + cc * kKC < g.K) q = sp[j] + (long long)cc * a_cstep; }
            else if (np[j] && cc - chunks < 4 && nk[j] + (cc - chunks) * kKC < g.K) q = np[j] + (long long)(cc - chunks) * a_cstep;
            if (q) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
          }
        }
        }
        const uint32_t seq = seq_base + (uint32_t)c;
        const uint32_t stage = seq % kAStages;
        if (tr) tr->mark(52);
        mbar_wait_warp(&pp.a_empty[stage], ((seq / kAStages) & 1) ^ 1, pp.st, 100 + stage);
        if (tr) tr->mark(53);
        uint8_t* st0 = pp.a_ring + stage * kAStageBytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (FMT == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[j][i] *= sa;
          }
          uint4 hi, lo;
          Split<FMT>::pair(x[j][0], x[j][1], hi.x, lo.x);
          Split<FMT>::pair(x[j][2], x[j][3], hi.y, lo.y);
          Split<FMT>::pair(x[j][4], x[j][5], hi.z, lo.z);
          Split<FMT>::pair(x[j][6], x[j][7], hi.w, lo.w);
          uint8_t* q = st0 + qoff(j);
          *reinterpret_cast<uint4*>(q) = hi;
          *reinterpret_cast<uint4*>(q + kAHalfBytes) = lo;
        }
        if (tr) tr->mark(54);
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) a_chunk_ready(pp, stage);
        __syncwarp();
        if (tr) tr->mark(55);
      }
      seq_base += (uint32_t)chunks;
      if (tr) tr->mark(51);
    }
  } else {
    // ------------------------------------------------------------------------------------------------------
    // drain warps (8): warp = (TMEM lane quarter, column half); they take the finished accumulators of item it-1
    // while the producers and the tensor core work on item it.  A thread holds one row (TMEM lane), 32 columns at
    // a time; for a row-major output each 32 x 32 block goes through shared memory so that a warp writes (and, for
    // the add mode and the mask, reads) 4 rows x 128 contiguous bytes per instruction; for the column-major output
    // of the wgrad form the lanes' rows are already adjacent in memory (coalesced red.global).
    // ------------------------------------------------------------------------------------------------------
    const int dw = warp - kTcProdWarps;
    const int quarter = warp & 3, half = dw >> 2;
    const uint32_t stg = smem_u32(stage_all + dw * (32 * 36));        // byte address in shared memory
    Trace trc; trc.init((lane == 0 && dw == 0) ? g.trace : nullptr, 2);
    Trace* tr = trc.p ? &trc : nullptr;
    const bool transposed = g.c_ns == 1;
    const bool vec_out = transposed && g.mode != 2 && (g.c_ms & 3) == 0 && (g.N & 3) == 0 &&
                         ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                         (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) &&
                         (!g.mask || ((g.mask_ms & 3) == 0 && (reinterpret_cast<uintptr_t>(g.mask) & 15) == 0));
    const int nblk = NT / 32;
    uint32_t d_cnt[2] = {0u, 0u};
    int it = 0;
    float cmax = 0.f;                                  // max |C| of what this thread stored (c_amax)
    for (int item = pair; item < items; item += n_pairs, ++it) {
      const int region = it & 1;
      const int rem = item % (m_tiles * g.n_tiles);
      const int mt = rem / g.n_tiles, nt = rem % g.n_tiles;
      const int m_warp = mt * 2 * kTileM + (int)pp.rank * kTileM + quarter * 32;
      const int m = m_warp + lane;
      const int n0 = nt * NT;
      // while the item's MMAs still run: pull what the epilogue will read (ReLU mask, C in the add mode) into L2
      if (transposed && (g.mask || g.mode == 1) && m < g.M) {
        for (int cb = half; cb < nblk; cb += 2) {
          const int nb = n0 + cb * 32;
          if (nb < g.N) {
            if (g.mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.mask + (long long)m * g.mask_ms + nb));
            if (g.mode == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.C + (long long)m * g.c_ms + nb));
          }
        }
      }
      if (tr) tr->mark(10);
      mbar_wait_warp(&pp.d_full[region], d_cnt[region] & 1, pp.st, 500 + region);
      ++d_cnt[region];
      tc_fence_after_sync();
      if (tr) tr->mark(11);
      const uint32_t taddr = pp.tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)region * 256u;
#pragma unroll 1
      for (int cb = half; cb < nblk; cb += 2) {
        // ReLU-derivative mask of this block (dgrad form): the 8 x 16 B per lane are requested BEFORE the accumulators are
        // fetched and staged, so that their L2 latency hides behind the TMEM load and the transpose (issued inside the
        // write loop they were the exposed part of the epilogue: the dgrad form ran at 1.8x the forward form's time)
        float4 mk[8];
        const bool early_mask = vec_out && g.mask != nullptr;
        if (early_mask) {
          const int r4e = lane >> 3, ne = n0 + cb * 32 + (lane & 7) * 4;
          const int rmaxe = g.M - m_warp < 32 ? g.M - m_warp : 32;
          const float* mrowe = g.mask + (long long)(m_warp + r4e) * g.mask_ms + ne;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            mk[i] = (ne < g.N && r4e + 4 * i < rmaxe) ? __ldg(reinterpret_cast<const float4*>(mrowe + (long long)i * 4 * g.mask_ms)) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
        uint32_t v[32];
        tmem_ld32(taddr + cb * 32, v);
        tmem_ld_wait();
        if (tr) tr->mark(13);
        const int nb = n0 + cb * 32;             // first column of the block
        // ---- fast path: a full 32 x 32 block of a row-major output that is stored (forward and dgrad forms, the bulk of
        // the training step): no per-row / per-column predicates, the operand scale folded into one FMA with the bias,
        // one running output pointer.  (The general path below issued ~230 instructions per block at ~14 cycles each.)
        if (vec_out && g.mode == 0 && m_warp + 32 <= g.M && nb + 32 <= g.N) {
#pragma unroll
          for (int q = 0; q < 8; ++q)              // row stride 36 floats: 16-byte aligned, conflict-free in both phases
            sts_f4(stg + (uint32_t)(lane * 36 + 4 * q) * 4u, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                   __uint_as_float(v[4 * q + 3]));
          __syncwarp();
          if (tr) tr->mark(14);
          const int r4 = lane >> 3, n = nb + (lane & 7) * 4;
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g.bias) b = __ldg(reinterpret_cast<const float4*>(g.bias + n));
          float* cp = g.C + (long long)(m_warp + r4) * g.c_ms + n;
          const long long cstep = 4 * g.c_ms;
          const float lo = g.relu ? 0.f : -3.0e38f;          // branch-free ReLU
          const uint32_t sp = stg + (uint32_t)(r4 * 36 + (lane & 7) * 4) * 4u;
          const bool has_mask = g.mask != nullptr;
#pragma unroll
          for (int i = 0; i < 8; ++i, cp += cstep) {
            float4 w = lds_f4(sp + (uint32_t)(i * 4 * 36) * 4u);
            w.x = fmaxf(fmaf(w.x, out_scale, b.x), lo); w.y = fmaxf(fmaf(w.y, out_scale, b.y), lo);
            w.z = fmaxf(fmaf(w.z, out_scale, b.z), lo); w.w = fmaxf(fmaf(w.w, out_scale, b.w), lo);
            if (has_mask) {
              const float4 k = mk[i];
              w.x = k.x > 0.f ? w.x : 0.f; w.y = k.y > 0.f ? w.y : 0.f; w.z = k.z > 0.f ? w.z : 0.f; w.w = k.w > 0.f ? w.w : 0.f;
            }
            if (!(g.debug & 1)) *reinterpret_cast<float4*>(cp) = w;
            cmax = fmaxf(fmaxf(cmax, fmaxf(fabsf(w.x), fabsf(w.y))), fmaxf(fabsf(w.z), fabsf(w.w)));
          }
          __syncwarp();
          if (tr) tr->mark(12);
          continue;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * out_scale);
        if (!transposed) {
          if (m < g.M) {
            float* c = g.C + (long long)m * g.c_ms + (long long)nb * g.c_ns;
#pragma unroll
            for (int i = 0; i < 32; ++i, c += g.c_ns) {
              if (nb + i >= g.N) break;
              float r = __uint_as_float(v[i]);
              if (g.mode == 2) { atomicAdd(c, r); continue; }
              if (g.bias) r += __ldg(g.bias + nb + i);
              if (g.mode == 1) r += *c;
              if (g.relu) r = fmaxf(r, 0.f);
              if (g.mask && !(__ldg(g.mask + (long long)m * g.mask_ms + nb + i) > 0.f)) r = 0.f;
              *c = r;
              cmax = fmaxf(cmax, fabsf(r));
            }
          }
          continue;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)              // row stride 36 floats: 16-byte aligned, conflict-free in both phases
          sts_f4(stg + (uint32_t)(lane * 36 + 4 * q) * 4u, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                 __uint_as_float(v[4 * q + 3]));
        __syncwarp();
        if (tr) tr->mark(14);
        const int rmax = g.M - m_warp < 32 ? g.M - m_warp : 32;
        if (vec_out) {
          // a warp instruction covers 4 rows x 128 contiguous bytes: lane = (row % 4, 4-column group)
          const int r4 = lane >> 3, n = nb + (lane & 7) * 4;
          const bool n_ok = n < g.N;            // N is a multiple of 4 here: a group is inside or outside as a whole
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g.bias && n_ok) b = __ldg(reinterpret_cast<const float4*>(g.bias + n));
          float* crow = g.C + (long long)(m_warp + r4) * g.c_ms + n;
          const long long cstep = 4 * g.c_ms;
          const float lo = g.relu ? 0.f : -3.0e38f;          // branch-free ReLU
          const uint32_t sp = stg + (uint32_t)(r4 * 36 + (lane & 7) * 4) * 4u;
          if (n_ok) {
            // general path (edge tiles, C += forms); fully unrolled: mk[] must stay in registers
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (r4 + 4 * i < rmax) {
                float4 w = lds_f4(sp + (uint32_t)(i * 4 * 36) * 4u);
                w.x += b.x; w.y += b.y; w.z += b.z; w.w += b.w;
                if (g.mode == 1) { const float4 o = *reinterpret_cast<const float4*>(crow + i * cstep); w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w; }
                w.x = fmaxf(w.x, lo); w.y = fmaxf(w.y, lo); w.z = fmaxf(w.z, lo); w.w = fmaxf(w.w, lo);
                if (g.mask) {
                  const float4 k = mk[i];
                  w.x = k.x > 0.f ? w.x : 0.f; w.y = k.y > 0.f ? w.y : 0.f; w.z = k.z > 0.f ? w.z : 0.f; w.w = k.w > 0.f ? w.w : 0.f;
                }
                *reinterpret_cast<float4*>(crow + i * cstep) = w;
                cmax = fmaxf(fmaxf(cmax, fmaxf(fabsf(w.x), fabsf(w.y))), fmaxf(fabsf(w.z), fabsf(w.w)));
              }
            }
          }
        } else {
          const int n = nb + lane;
          const bool n_ok = n < g.N;
          const float b = (g.bias && n_ok) ? __ldg(g.bias + n) : 0.f;
#pragma unroll 4
          for (int r = 0; r < rmax; ++r) {
            float w = lds_f(stg + (uint32_t)(r * 36 + lane) * 4u);
            if (n_ok) {
              float* c = g.C + (long long)(m_warp + r) * g.c_ms + n;
              if (g.mode == 2) { atomicAdd(c, w); continue; }
              w += b;
              if (g.mode == 1) w += *c;
              if (g.relu) w = fmaxf(w, 0.f);
              if (g.mask && !(__ldg(g.mask + (long long)(m_warp + r) * g.mask_ms + n) > 0.f)) w = 0.f;
              *c = w;
              cmax = fmaxf(cmax, fabsf(w));
            }
          }
        }
        __syncwarp();
        if (tr) tr->mark(12);
      }
      // hand the region back to the producers (its next MMAs overwrite it)
      tc_fence_before_sync();
      __syncwarp();
      if (elect_one()) { if (pp.rank == 0) mbar_arrive(&r_free[region]); else mbar_arrive_remote_cnt(&r_free[region], 0, 1); }
      __syncwarp();
      if (tr) tr->mark(61);
    }
    if (g.c_amax) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
      if (lane == 0 && cmax > 0.f && cmax < 3.0e38f) amax_fold(g.c_amax, cmax);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc(pp.tmem_base, kTmemCols);
}

#endif  // __CUDACC__
}  // namespace anerf
