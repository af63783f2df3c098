// The fused A-NeRF ray-marching kernel for sm_100a and its helpers.
//
// One persistent CTA per SM, CTAs paired (cluster of 2, cta_group::2), processes "items" of R rays.  Per item:
//   sample coarse depths -> [coarse net over R*Sc rows] -> composite -> inverse-CDF importance
//   sampling + merge -> [fine net over R*(Sc+Si) rows] -> composite -> outputs.
// A "net pass" over one tile of 128 rows (samples) runs the whole 8x256 MLP with activations never
// leaving the SM:
//   * 16 worker warps in four groups of 4 (thread == row == TMEM lane; the groups share the rows and own the operand
//     chunks c with c % 4 == group) produce the A operand in 32-wide K chunks, either by computing the encodings of
//     their sample (bone-local transform, cutoff positional encoding) or by draining the previous layer's accumulators
//     from TMEM (scale incl. the truncation compensation, bias, ReLU), split every value into hi + lo 16-bit parts and
//     store them in the UMMA K-major core-matrix layout (A ring, 4 stages);
//   * 1 loader thread per CTA streams this CTA's half (N/2 rows) of the pre-packed weight chunks
//     (same layout, hi + lo) from L2 with 1-D bulk TMA copies into the B ring;
//   * the MMA warp of the leader CTA issues, per chunk, 2 K-slabs x 3 tcgen05.mma (lo*hi + hi*lo + hi*hi) of shape
//     M=256 (both CTAs' 128 rows) x N into 128 x N fp32 accumulators in each CTA's TMEM; every SM reads its own A rows
//     and only half of B from its shared memory.  Accumulators ping-pong between two 256-column regions so the drain
//     of layer l overlaps the MMAs of layer l+1 chunk by chunk.  The peer CTA's spare warp relays "my weight half has
//     landed" to the leader.
// The view branch is contracted per ray (per-ray matrices G as a B operand built in shared memory, path_math.cuh).
// Layer program and K layout: path_math.cuh.  Protocol: mbarrier full/empty rings, bounded waits (a protocol bug
// records a site id in pinned host memory and traps; anerf_check_status reports it).
#pragma once
#include "tc_sm100.cuh"
#include "path_math.cuh"

namespace anerf {

constexpr int kAStages = 4;
#ifndef ANERF_B_STAGES
#define ANERF_B_STAGES 4
#endif
constexpr int kBStages = ANERF_B_STAGES;   // weight ring depth (compile-time knob for tools/ab_variants.py)
constexpr int kAHalfBytes = kTileM * kKC * 2;        // 8 KB: hi (or lo) part of one A chunk
constexpr int kAStageBytes = 2 * kAHalfBytes;        // 16 KB
constexpr int kBStageBytes = 128 * kKC * 2 * 2;      // 16 KB: this CTA's half (N/2 <= 128 rows) of a weight chunk, hi + lo
constexpr int kNumBars = 2 * kAStages + 3 * kBStages + 2;
constexpr int kMaxLayers = 10;
constexpr int kWorkerWarps = 4 * kGroups;            // group g = warp / 4 owns K elements [8g, 8g+8) of every chunk
constexpr int kWorkerThreads = kWorkerWarps * 32;
constexpr int kMmaWarp = kWorkerWarps;
constexpr int kLoadWarp = kWorkerWarps + 1;
constexpr int kThreads = (kWorkerWarps + 2) * 32;
constexpr int kTmemCols = 512;
constexpr int kSmallsHeader = 16;                    // floats: per-layer output scales
constexpr int kMaxRaysPerItem = 8;
#ifndef ANERF_SPLIT_FIRST_CHUNK
#define ANERF_SPLIT_FIRST_CHUNK 1
#endif
constexpr bool kSplitFirstChunk = ANERF_SPLIT_FIRST_CHUNK != 0;   // compile-time knob for tools/ab_variants.py
#ifndef ANERF_SORT_ON_RAY_WARP
#define ANERF_SORT_ON_RAY_WARP 0
#endif
constexpr bool kSortOnRayWarp = ANERF_SORT_ON_RAY_WARP != 0;      // importance sampling + merge by the ray's compositing warp
                                                                  // (measured 3 % slower than all 512 threads after a barrier: off)
constexpr int kAFullCount = 32;                      // arrivals (weighted) that complete an A chunk, see pipe_init_barriers

struct LayerProg {
  int n;            // output features (UMMA N)
  int chunks;       // K chunks of 32
  unsigned w_off;   // byte offset of the layer's first chunk in the packed image
};

// offsets (in floats) inside the fp32 "smalls" block of a packed image
struct SmallsLayout {
  int bias[kMaxLayers];
  int alpha_w, alpha_b, rgb_w, rgb_b;
  int fixed_floats;     // everything above (copied to shared memory)
  int framecodes;       // [n_fc + 1][fc_ch] (last row = mean code); stays in global memory
  int gw;               // [W/2][27][J+1] view weights of the folded views layer, regrouped per joint; global memory
  int total_floats;
};

struct NetProgram {
  NetDims dims;
  int n_layers;         // D + 1 (trunk layers + the views layer with feature_linear folded in)
  LayerProg layer[kMaxLayers];
  unsigned smalls_off;  // byte offset of the smalls block in the packed image
  SmallsLayout sm;
  unsigned packed_bytes;
};

inline __host__ __device__ int align_up(int x, int a) { return (x + a - 1) / a * a; }

// tcgen05.mma accumulates in fp32 with truncation toward zero (sign-symmetric; probed on the B200, tools/probes/
// trunc_probe.py): every MMA instruction that adds into an accumulator loses on average half an ulp of the running sum's
// magnitude.  A layer with K real input columns issues 3 * K/16 such MMAs (hi*hi + the two cross terms per K=16 slab), so
// its outputs come out smaller in magnitude by about kappa * 1.5 * (K/16) ulp -- a systematic shrink of ~2e-6 per 256-wide
// layer that compounds over the 9 layers of a network (measured: -1.9e-5 mean on sigma at the benchmark shape, four times
// the fp32 reference's own rounding noise).  The drains multiply by the layer's output scale anyway, so the expected loss
// is folded into that factor: E[ulp(x)/|x|] ~ 8.6e-8 for log-uniform mantissas, and kappa < 1 accounts for the partial
// sums being smaller than the final one while they are accumulated.  kappa is calibrated on 4096 rays of the benchmark
// frame against the fp64 oracle (profiles/r2_parity.md): fp16 operands 0.4 (mean error of sigma -1.9e-5 -> -4e-6, mean
// |error| 4.6e-5 -> 2.1e-5, rgb0 2.3e-5 -> 8e-6; the fp32 reference itself: 1.3e-5 / 5e-6); bf16 operands carry a further
// systematic shrink of the same form (their 16-bit hi+lo representation) and calibrate to 1.4.
constexpr float kTruncKappaFp16 = 0.4f;
constexpr float kTruncKappaBf16 = 1.4f;
inline __host__ float trunc_comp(int k_real, float kappa) {
  return 1.0f + kappa * 1.5f * (float)((k_real + 15) / 16) * 8.6e-8f;
}

inline __host__ NetProgram make_program(const NetDims& d) {
  NetProgram p{};
  p.dims = d;
  p.n_layers = d.D + 1;
  unsigned off = 0;
  for (int l = 0; l < p.n_layers; ++l) {
    p.layer[l].n = layer_n(d, l);
    p.layer[l].chunks = layer_chunks(d, l);
    p.layer[l].w_off = off;
    off += (unsigned)p.layer[l].chunks * (unsigned)p.layer[l].n * 128u;   // N * 32 * 2 B * (hi + lo)
  }
  p.smalls_off = off;
  int f = kSmallsHeader;
  for (int l = 0; l < p.n_layers; ++l) { p.sm.bias[l] = f; f += align_up(p.layer[l].n, 4); }
  p.sm.alpha_w = f; f += d.W;
  p.sm.alpha_b = f; f += 4;
  p.sm.rgb_w = f; f += 3 * (d.W / 2);
  p.sm.rgb_b = f; f += 4;
  p.sm.fixed_floats = f;
  p.sm.framecodes = f; f += (d.n_fc + 1) * d.fc_ch;
  p.sm.gw = f; f += (d.W / 2) * (d.J + 1) * kViewPerJoint;
  p.sm.total_floats = align_up(f, 4);
  p.packed_bytes = off + (unsigned)p.sm.total_floats * 4u;
  return p;
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up of the fused kernel
// ------------------------------------------------------------------------------------------------
struct SmemLayout {
  int a_ring, b_ring, g_buf, smalls0, smalls1, ray, skt, view_tab, fcode, task_ctr, z_coarse, z_all, raw, part, wj, wts, cdf, bars, tmem_ptr;
  int total;
};
// view table T[joint][ray slot][kViewPad]: the 27 direction features of (ray, joint) padded to 28 floats so a
// thread reads them as seven float4; the joint stride is padded by 4 floats, which makes the float4 reads of
// eight neighbouring joints hit eight different bank groups for every even slot count.
constexpr int kViewPad = 28;
inline __host__ __device__ int view_tab_jstride(int R) { return 2 * R * kViewPad + 4; }
// ray_s: 12 floats per ray: o(3) d(3) near far |d| pad(3)
inline __host__ __device__ SmemLayout make_smem_layout(const NetDims& d, int smalls_fixed_floats, int R, int Sc, int Sf) {
  SmemLayout L{};
  int off = 0;
  L.a_ring = off; off += kAStages * kAStageBytes;
  L.b_ring = off; off += kBStages * kBStageBytes;
  L.g_buf = off; off += slot_chunks(R) * (d.W / 2) * 64;   // per-ray view matrices G as B-operand chunks (this CTA's half)
  L.smalls0 = off; off += align_up(smalls_fixed_floats * 4, 16);
  L.smalls1 = off; off += align_up(smalls_fixed_floats * 4, 16);
  L.ray = off; off += R * 12 * 4;
  L.skt = off; off += align_up(R * d.J * 12 * 4, 16);
  L.view_tab = off; off += align_up(d.J * view_tab_jstride(R) * 4, 16);     // [joint][ray slot of the pair][kViewPad]
  L.fcode = off; off += align_up(2 * R * 4, 16);   // [ray slot] framecode row index
  L.task_ctr = off; off += 16;                     // work counter of build_view_matrices
  int rows = R * (Sf > Sc ? Sf : Sc);
  L.z_coarse = off; off += align_up(R * Sc * 4, 16);
  L.z_all = off; off += align_up(rows * 4, 16);
  L.raw = off; off += rows * 16;             // network outputs (r,g,b,sigma) of the item's samples
  L.part = off; off += kGroups * kTileM * 16; // per-group partial outputs of the current tile
  L.wj = off; off += d.J * kTileM * 4;        // view cutoff weights of the current tile [joint][row]
  L.wts = off; off += align_up(rows * 4, 16);
  L.cdf = off; off += align_up(R * Sc * 4, 16);
  L.bars = off; off += 8 * kNumBars;
  L.tmem_ptr = off; off += 16;
  L.total = off;
  return L;
}

struct RenderKParams {
  NetProgram prog;
  const uint8_t* packed[2];
  // problem
  int n_rays, Sc, Si, Sf, R, tilesC, tilesF, n_items, slotc;
  int lindisp, softplus, eval_mean_fc, blur_is;
  float B, shift, tau_p, tau_v;
  float cut_p[kMaxJoints], cut_v[kMaxJoints];
  const float *rays, *skts, *cams, *t_rand, *u_rand, *noise0, *noise1;
  // frame mode (anerf_render_frame): rays == NULL, generated per pixel from `gen`; ONE pose / camera index for all rays
  RayGen gen;
  long long skt_stride;     // floats between the poses of consecutive rays: J*16, or 0 in frame mode
  const int* pose_idx;      // optional [N]: ray -> row of `skts` (one transform set per pose instead of per ray)
  int n_poses;              // rows of `skts` when pose_idx is set (indices are clamped); 0 = unchecked
  float cam_const;          // frame mode: the frame's camera index (framecodes)
  const float* nearfar;     // [N,2] from the near/far pre-kernel
  float *rgb_map, *disp_map, *acc_map, *alpha, *rgb0, *disp0, *acc0, *alpha0, *z_all_out, *raw_out;
  // density-only mode (mesh grid): one pose; explicit points, or (pts == NULL) the points [grid_first, grid_first +
  // n_points) of the reference's flattened (res+1)^3 grid generated in the kernel (core/raycasters.py:579-595)
  const float* pts;
  float* sigma;
  long long n_points;
  long long grid_first;
  int grid_n1;              // res + 1
  double grid_start, grid_step;   // np.linspace(-radius, radius, res + 1): start + i * step in fp64, last = radius
  double grid_stop;
  const float* grid_origin; // device pointer to kps[0, 0] (3 floats)
  DeviceStatus* status;
  long long* trace;         // optional debug timeline (anerf_debug_set_trace): [3 streams][1024] clock64 stamps of CTA 0
  SmemLayout sl;
};

#ifdef __CUDACC__

// row of `skts` that ray `gr` reads: the ray itself, or its pose through the (clamped) index
__device__ __forceinline__ size_t pose_row(const RenderKParams& P, int gr) {
  if (!P.pose_idx) return (size_t)gr;
  int p = __ldg(P.pose_idx + gr);
  if (P.n_poses > 0) p = min(max(p, 0), P.n_poses - 1);
  return (size_t)p;
}

// ------------------------------------------------------------------------------------------------
// pipeline context shared by the three roles
// ------------------------------------------------------------------------------------------------
struct Pipe {
  uint8_t* a_ring;
  uint8_t* b_ring;
  uint64_t* a_full;
  uint64_t* a_empty;
  uint64_t* b_full;
  uint64_t* b_empty;
  uint64_t* peer_b;    // leader only: the peer CTA's weight half of stage s has landed
  uint64_t* d_full;    // [2]
  uint32_t tmem_base;
  uint32_t rank;       // CTA rank in the pair (0 = leader)
  DeviceStatus* st;
};

__device__ __forceinline__ void pipe_init(Pipe& pp, uint8_t* a_ring, uint8_t* b_ring, uint64_t* bars, DeviceStatus* st) {
  pp.a_ring = a_ring;
  pp.b_ring = b_ring;
  pp.a_full = bars;
  pp.a_empty = bars + kAStages;
  pp.b_full = bars + 2 * kAStages;
  pp.b_empty = bars + 2 * kAStages + kBStages;
  pp.peer_b = bars + 2 * kAStages + 2 * kBStages;
  pp.d_full = bars + 2 * kAStages + 3 * kBStages;
  pp.rank = cluster_ctarank();
  pp.st = st;
}
// thread 0 of each CTA; followed by a cluster-wide sync
__device__ __forceinline__ void pipe_init_barriers(const Pipe& pp) {
  // an A chunk is complete after kAFullCount arrivals: normally the 4 warps of the owning group in both CTAs, each
  // arriving with weight 4; a chunk shared by all four groups (first chunk of a drained operand) gets weight 1 from
  // each of the 16 worker warps of both CTAs
  for (int i = 0; i < kAStages; ++i) { mbar_init(&pp.a_full[i], kAFullCount); mbar_init(&pp.a_empty[i], 1); }
  for (int i = 0; i < kBStages; ++i) { mbar_init(&pp.b_full[i], 1); mbar_init(&pp.b_empty[i], 1); mbar_init(&pp.peer_b[i], 1); }
  mbar_init(&pp.d_full[0], 1);
  mbar_init(&pp.d_full[1], 1);
  fence_mbar_init();
}
// a warp's share of a chunk of this CTA's A rows is complete: tell the leader's MMA thread
__device__ __forceinline__ void a_chunk_ready(const Pipe& pp, uint32_t stage, uint32_t weight = 4) {
  if (pp.rank == 0) mbar_arrive_cnt(&pp.a_full[stage], weight); else mbar_arrive_remote_cnt(&pp.a_full[stage], 0, weight);
}

__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// whole-warp wait: every lane probes (warp-uniform control flow; a single-lane region would make the
// compiler wrap the barrier instructions in ELECT loops)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, DeviceStatus* st, unsigned site) {
  mbar_wait(bar, parity, st, site);
  __syncwarp();
}

// A-operand producer state of one worker thread (= one row of the tile).  Group g fills the chunks of an
// operand part whose index inside the part is congruent to g mod 4; `base` is the sequence number of the
// part's first chunk (every worker advances it identically, part by part).
template <int FMT>
struct AProducer {
  const Pipe& pp;
  uint32_t base;       // sequence number (monotonic over the kernel's lifetime) of the current part's chunk 0
  uint32_t row_off;    // (row/8)*128 + (row%8)*16
  uint32_t cur;        // stage of the open chunk
  uint8_t* stage;
  __device__ AProducer(const Pipe& p, int row) : pp(p), base(0), row_off((row >> 3) * 128 + (row & 7) * 16), cur(0), stage(nullptr) {}

  // open chunk `c` of the current part: wait until the tensor core has released its ring stage
  __device__ __forceinline__ void begin(uint32_t c) {
    const uint32_t seq = base + c;
    cur = seq % kAStages;
    mbar_wait_warp(&pp.a_empty[cur], ((seq / kAStages) & 1) ^ 1, pp.st, 100 + cur);
    stage = pp.a_ring + cur * kAStageBytes + row_off;
  }
  // K elements [8t, 8t+8) of the open chunk
  __device__ __forceinline__ void store8(int t, const float (&x)[8]) {
    uint4 hi, lo;
    Split<FMT>::pair(x[0], x[1], hi.x, lo.x);
    Split<FMT>::pair(x[2], x[3], hi.y, lo.y);
    Split<FMT>::pair(x[4], x[5], hi.z, lo.z);
    Split<FMT>::pair(x[6], x[7], hi.w, lo.w);
    uint8_t* p = stage + (t >> 1) * 4096 + (t & 1) * 2048;
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + kAHalfBytes) = lo;
  }
  __device__ __forceinline__ void zero8(int t) {
    uint8_t* p = stage + (t >> 1) * 4096 + (t & 1) * 2048;
    *reinterpret_cast<uint4*>(p) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(p + kAHalfBytes) = make_uint4(0u, 0u, 0u, 0u);
  }
  // publish the open chunk: one proxy fence per thread, one barrier arrival per warp
  __device__ __forceinline__ void end() {
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) a_chunk_ready(pp, cur);
    __syncwarp();
  }
  // same for a chunk that all four groups fill together (each warp contributes a quarter of the owner's weight)
  __device__ __forceinline__ void end_shared() {
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) a_chunk_ready(pp, cur, 1);
    __syncwarp();
  }
};

// debug timeline: stream 0 = MMA thread, 1 = worker warp 0 (group 0), 2 = worker warp 4 (group 1); CTA 0 only
struct Trace {
  long long* p;
  int n;
  __device__ __forceinline__ void init(long long* base, int stream) { p = (base && blockIdx.x == 0) ? base + stream * 1024 : nullptr; n = 0; }
  __device__ __forceinline__ void mark(int tag) {
    if (p && n < 1022) { p[n++] = ((long long)tag << 48) | (clock64() & 0xFFFFFFFFFFFFLL); p[1023] = n; }
  }
};

// ------------------------------------------------------------------------------------------------
// MMA issuer: one thread of the leader CTA.  All chunks of one layer, for both CTAs of the pair.
// ------------------------------------------------------------------------------------------------
// Executed by the WHOLE MMA warp of the leader CTA (warp-uniform control flow keeps addresses and loop
// state in uniform registers; a divergent single-lane region makes the compiler wrap every UTCHMMA in an
// ELECT/BRA.U.ANY loop, ~100 cycles per instruction); one elected lane issues.
template <int FMT>
__device__ __forceinline__ void mma_layer(const Pipe& pp, uint32_t& a_seq, uint32_t& b_seq, int N, int chunks,
                                          int region, Trace* tr = nullptr, int g_chunks = 0, uint32_t g_base = 0) {
  // `chunks` operand chunks in total; the first `g_chunks` take their B operand from shared memory at g_base
  // (per-ray view matrices, N * 64 bytes each) instead of the weight ring.
  const uint32_t id = make_idesc_f16(fmt_hi(FMT), fmt_hi(FMT), 2 * kTileM, N);
  const uint32_t dcol = pp.tmem_base + (uint32_t)region * 256u;
  const uint32_t a_base = smem_u32(pp.a_ring), b_base = smem_u32(pp.b_ring);
  const uint32_t NH = (uint32_t)N / 2;           // B rows held by each CTA
  // Narrow layers (N <= 128: 64-cycle MMAs) take two chunks per iteration so that the fixed cost of an
  // iteration (barrier probes, commits) stays below the tensor time of what it issues.  Chunk counts are
  // multiples of 4, so pairs never straddle a layer.
  const int step = N <= 128 ? 2 : 1;
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int c = 0; c < chunks; c += step) {
    {
      // the "operand ready" barriers of this iteration's chunks are probed concurrently by different lanes
      const int which = c < g_chunks ? 0 : lane % 3, sub = (lane / 3) % step;
      const uint32_t qa = a_seq + sub, qb = b_seq + sub;
      uint64_t* bar = which == 0 ? &pp.a_full[qa % kAStages] : (which == 1 ? &pp.b_full[qb % kBStages] : &pp.peer_b[qb % kBStages]);
      const uint32_t par = which == 0 ? ((qa / kAStages) & 1) : ((qb / kBStages) & 1);
      mbar_wait(bar, par, pp.st, 200 + 100 * which);
      __syncwarp();
    }
    tc_fence_after_sync();
    if (tr) tr->mark(c == 0 ? 1 : 2);        // 1: first chunk of a layer ready, 2: later chunk ready
    if (elect_one()) {
      const bool from_g = c < g_chunks;
      for (int u = 0; u < step; ++u) {
        const uint32_t sa = (a_seq + u) % kAStages, sb = (b_seq + u) % kBStages;
        const uint32_t a_hi = a_base + sa * kAStageBytes, a_lo = a_hi + kAHalfBytes;
        const uint32_t b_hi = from_g ? g_base + (uint32_t)(c + u) * NH * 128u : b_base + sb * kBStageBytes, b_lo = b_hi + NH * 64u;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          uint64_t da_hi = smem_desc(a_hi + s * 4096, 2048, 128);
          uint64_t da_lo = smem_desc(a_lo + s * 4096, 2048, 128);
          uint64_t db_hi = smem_desc(b_hi + s * NH * 32u, NH * 16u, 128);
          uint64_t db_lo = smem_desc(b_lo + s * NH * 32u, NH * 16u, 128);
          // the two small cross terms first, then the dominant one
          umma_f16(dcol, da_lo, db_hi, id, (c + u > 0 || s > 0) ? 1u : 0u);
          umma_f16(dcol, da_hi, db_lo, id, 1u);
          umma_f16(dcol, da_hi, db_hi, id, 1u);
        }
        umma_commit(&pp.a_empty[sa]);
        if (!from_g) umma_commit(&pp.b_empty[sb]);
      }
      if (c + step >= chunks) umma_commit(&pp.d_full[region]);
    }
    __syncwarp();
    a_seq += step;
    if (c >= g_chunks) b_seq += step;
  }
  if (tr) tr->mark(3);                       // 3: layer fully issued
}

// weight loader: one thread per CTA.  This CTA's half of every chunk of one layer.
// packed chunk = [half 0: hi, lo][half 1: hi, lo], each half N*64 bytes.
__device__ __forceinline__ void load_layer(const Pipe& pp, uint32_t& b_seq, const uint8_t* src, int N, int chunks,
                                           uint32_t debug_copy_bytes = 0) {
  const uint32_t stride = (uint32_t)N * 64u;
  const uint32_t bytes = debug_copy_bytes ? debug_copy_bytes : stride;      // (timing experiments copy less than a chunk)
  src += (size_t)pp.rank * stride;
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    uint32_t sb = b_seq % kBStages;
    mbar_wait(&pp.b_empty[sb], ((b_seq / kBStages) & 1) ^ 1, pp.st, 400 + sb);
    mbar_arrive_expect_tx(&pp.b_full[sb], bytes);
    bulk_g2s(pp.b_ring + sb * kBStageBytes, src + (size_t)c * 2 * stride, bytes, &pp.b_full[sb]);
    ++b_seq;
  }
}

// peer CTA only, one thread: forward "my half of stage s has landed" to the leader's MMA thread
__device__ __forceinline__ void relay_layer(const Pipe& pp, uint32_t& b_seq, int chunks) {
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    uint32_t sb = b_seq % kBStages;
    mbar_wait(&pp.b_full[sb], (b_seq / kBStages) & 1, pp.st, 600 + sb);
    mbar_arrive_remote(&pp.peer_b[sb], 0);
    ++b_seq;
  }
}

// ------------------------------------------------------------------------------------------------
// worker-side building blocks
// ------------------------------------------------------------------------------------------------
struct RowCtx {
  float p[3];          // world position of this row's sample
  const float* skt;    // this row's ray: [J][12] in shared memory
  int slot;            // ray slot of this row's ray inside the CTA pair (rank * R + ray index in the item)
};

// layer-0 / skip-layer part: distance + bone encodings of the row's sample for this group's joints
// (g, g+4, g+8, ...), two joints at a time (36 values + 4 zeros = 5 slots of 8), streamed into the
// group's chunks g, g+4, ... of the part
template <int FMT>
__device__ __forceinline__ void produce_pts_chunks(AProducer<FMT>& ap, const RowCtx& rc, const RenderKParams& P, int grp) {
  const int J = P.prog.dims.J;
  const int pairs = pts_pairs(P.prog.dims);
  int slot = 0;
#pragma unroll 1
  for (int pr = 0; pr < pairs; ++pr) {
    float vals[kPtsPairK];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      int j = grp + kGroups * (2 * pr + jj);
      if (j < J) {
        encode_joint_pts(rc.skt + j * 12, rc.p, P.tau_p, P.cut_p[j], &vals[jj * kPtsPerJoint]);
      } else {
#pragma unroll
        for (int q = 0; q < kPtsPerJoint; ++q) vals[jj * kPtsPerJoint + q] = 0.f;
      }
    }
#pragma unroll
    for (int q = 2 * kPtsPerJoint; q < kPtsPairK; ++q) vals[q] = 0.f;
#pragma unroll
    for (int t = 0; t < kPtsPairK / 8; ++t) {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = vals[8 * t + i];
      if ((slot & 3) == 0) ap.begin((slot >> 2) * kGroups + grp);
      ap.store8(slot & 3, x);
      if ((slot & 3) == 3) ap.end();
      ++slot;
    }
  }
  const int total = pts_group_chunks(P.prog.dims) * 4;       // zero padding up to whole chunks
  for (; slot < total; ++slot) {
    if ((slot & 3) == 0) ap.begin((slot >> 2) * kGroups + grp);
    ap.zero8(slot & 3);
    if ((slot & 3) == 3) ap.end();
  }
  ap.base += pts_chunks(P.prog.dims);
}

// views-layer part (view branch contracted per ray, see path_math.cuh): one chunk per ray slot of the CTA pair.
// A row writes [w_0 .. w_{J-1}, 1 (framecode pseudo joint), 0 ...] into the chunk of its own ray and zeros into the
// others; this group takes the chunks g, g+4, ...
template <int FMT>
__device__ __forceinline__ void produce_slot_chunks(AProducer<FMT>& ap, const RowCtx& rc, const RenderKParams& P, int grp,
                                                    int row, const float* wj) {
  const int J = P.prog.dims.J;
#pragma unroll 1
  for (int c = grp; c < P.slotc; c += kGroups) {
    ap.begin(c);
    if (c == rc.slot) {
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = 8 * t + i;
          x[i] = j < J ? wj[j * kTileM + row] : ((j == J && P.prog.dims.fc_ch > 0) ? 1.0f : 0.0f);
        }
        ap.store8(t, x);
      }
    } else {
      for (int t = 0; t < 4; ++t) ap.zero8(t);
    }
    ap.end();
  }
  ap.base += P.slotc;
}

// cutoff weights of the view branch for this tile: group g computes joints g, g+4, ... of its row into shared
// memory (wj[joint][row]); the group that owns the row's ray-slot chunk reads all of them after a barrier
__device__ __forceinline__ void compute_view_weights(const RowCtx& rc, const RenderKParams& P, int grp, int row, float* wj) {
  const int J = P.prog.dims.J;
#pragma unroll 1
  for (int j = grp; j < J; j += kGroups)
    wj[j * kTileM + row] = cutoff_w(joint_dist(rc.skt + j * 12, rc.p), P.tau_v, P.cut_v[j]);
}

// Per-ray view matrices of one network for all 2R ray slots of the CTA pair, written as B-operand chunks
// (this CTA's half of the N = W/2 output features): G[slot][n][j] = scale * sum_q gw[n][q][j] * T[j][slot][q],
// T = per-ray direction features (j < J) or the ray's framecode (j == J).  Work is handed out to warps in
// groups of 32 (n, j) tasks through a shared counter, so warps that arrive late (the ones compositing a ray)
// simply take fewer groups -- any subset of the worker warps may call this.  `ctr` must be zero on entry (reset
// behind a barrier); callers synchronise afterwards (the proxy fence is inside).
template <int FMT>
__device__ __forceinline__ void build_view_matrices(const RenderKParams& P, int net, uint8_t* g_buf, const float* vtab_s,
                                                    const int* fc_row_s, int* ctr, int rank, float scale) {
  const NetDims& d = P.prog.dims;
  const int NH = d.W / 4;                       // rows of the views layer held by this CTA
  const int J = d.J, J1 = d.J + 1;
  const int S = 2 * P.R;                        // ray slots (<= 2 * kMaxRaysPerItem), even
  const int jstride = view_tab_jstride(P.R);
  const int lane = threadIdx.x & 31;
  const float* smalls = reinterpret_cast<const float*>(P.packed[net] + P.prog.smalls_off);
  const float* gw = smalls + P.prog.sm.gw + (size_t)rank * NH * kViewPerJoint * J1;   // [n][q][j], j fastest: coalesced
  const int n_tasks = NH * J, n_groups = (n_tasks + 31) / 32;
  const int fc_tasks = d.fc_ch > 0 ? NH * S : 0, fc_groups = (fc_tasks + 31) / 32;   // framecode pseudo joint j == J
  auto put = [&](int slot, uint32_t off, float v) {
    uint32_t hi, lo;
    Split<FMT>::pair(v * scale, 0.f, hi, lo);
    uint8_t* chunk = g_buf + (size_t)slot * NH * 128;
    *reinterpret_cast<uint16_t*>(chunk + off) = (uint16_t)(hi & 0xFFFFu);
    *reinterpret_cast<uint16_t*>(chunk + NH * 64 + off) = (uint16_t)(lo & 0xFFFFu);
  };
  auto chunk_off = [&](int n, int j) {
    return (uint32_t)(j >> 3) * NH * 16 + (uint32_t)(n >> 3) * 128 + (uint32_t)(n & 7) * 16 + (uint32_t)(j & 7) * 2;
  };
#pragma unroll 1
  for (;;) {
    int g = 0;
    if (lane == 0) g = atomicAdd(ctr, 1);
    g = __shfl_sync(0xFFFFFFFFu, g, 0);
    if (g >= n_groups + fc_groups) break;
    if (g >= n_groups) {                                   // one (n, slot) of the framecode pseudo joint per lane
      const int i = (g - n_groups) * 32 + lane;
      if (i < fc_tasks) {
        const int n = i / S, sl = i % S;
        const float* w = gw + (size_t)n * kViewPerJoint * J1 + J;
        const float* c = smalls + P.prog.sm.framecodes + (size_t)fc_row_s[sl] * d.fc_ch;
        float acc = 0.f;
        for (int q = 0; q < d.fc_ch; ++q) acc = fmaf(__ldg(w + q * J1), __ldg(c + q), acc);
        put(sl, chunk_off(n, J), acc);
      }
      continue;
    }
    const int i = g * 32 + lane;
    if (i < n_tasks) {
      const int j = i % J, n = i / J;
      const float* w = gw + (size_t)n * kViewPerJoint * J1 + j;
      float wq[kViewPad];                                  // the joint's 27 weights, all loads in flight at once
#pragma unroll
      for (int q = 0; q < kViewPerJoint; ++q) wq[q] = __ldg(w + q * J1);
      wq[kViewPerJoint] = 0.f;
      const uint32_t off = chunk_off(n, j);
      const float4* t = reinterpret_cast<const float4*>(vtab_s + j * jstride);
#pragma unroll 1
      for (int sl = 0; sl < S; sl += 2) {                  // two ray slots at a time
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
        for (int v = 0; v < kViewPad / 4; ++v) {
          const float4 x = t[sl * (kViewPad / 4) + v], y = t[(sl + 1) * (kViewPad / 4) + v];
          a0 = fmaf(wq[4 * v], x.x, a0); a1 = fmaf(wq[4 * v + 1], x.y, a1);
          a0 = fmaf(wq[4 * v + 2], x.z, a0); a1 = fmaf(wq[4 * v + 3], x.w, a1);
          b0 = fmaf(wq[4 * v], y.x, b0); b1 = fmaf(wq[4 * v + 1], y.y, b1);
          b0 = fmaf(wq[4 * v + 2], y.z, b0); b1 = fmaf(wq[4 * v + 3], y.w, b1);
        }
        put(sl, off, a0 + a1);
        put(sl + 1, off, b0 + b1);
      }
    }
  }
  fence_proxy_async_smem();
}

// Wait for the accumulators of the layer that used `region`, then walk this group's column blocks
// (cb = g, g+4, ...; 32 columns each, 8 at a time).  acc(col0, x[8]) sees scale*acc + bias (ReLU applied
// when RELU); with EMIT the block also becomes chunk cb of the next layer's operand.
// `split0` (EMIT only): columns [0, 32) -- chunk 0 of the next operand, the one the tensor core is waiting for right after
// "layer done" -- are drained by all four groups together, 8 columns each, instead of by group 0 alone: the first MMA
// of the next layer can start after one TMEM load per warp instead of four in sequence.
template <int FMT, bool RELU, bool EMIT, typename F>
__device__ __forceinline__ void drain_region(AProducer<FMT>& ap, const Pipe& pp, uint32_t (&d_cnt)[2], int region, int N,
                                             const float* bias, float scale, int quarter, int grp, F&& acc,
                                             Trace* tr = nullptr, bool split0 = false) {
  if (tr) tr->mark(10);                      // 10: start waiting for the layer's accumulators
  mbar_wait_warp(&pp.d_full[region], d_cnt[region] & 1, pp.st, 500 + region);
  ++d_cnt[region];
  tc_fence_after_sync();
  if (tr) tr->mark(11);                      // 11: accumulators ready
  const uint32_t taddr = pp.tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)region * 256u;
  const int nblk = N / 32;
  // All 16 warps start loading at once (TMEM reads run at ~64 B/cycle per SM; letting the groups start in turn,
  // so that group 0's chunk comes out sooner, measured 1 % slower overall: tools/ab_variants.py, r1 notes).
  uint32_t v[8];
  int cb0 = grp;
  if (EMIT && split0) {
    tmem_ld8(taddr + grp * 8, v);
    ap.begin(0);
    tmem_ld_wait();
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __uint_as_float(v[i]);
    cb0 = grp == 0 ? kGroups : grp;            // group 0's block 0 is the shared one
    if (cb0 < nblk) tmem_ld8(taddr + cb0 * 32, v);
    const int col0 = grp * 8;
    const float4 b0 = *reinterpret_cast<const float4*>(bias + col0), b1 = *reinterpret_cast<const float4*>(bias + col0 + 4);
    float x[8];
    x[0] = fmaf(a[0], scale, b0.x); x[1] = fmaf(a[1], scale, b0.y); x[2] = fmaf(a[2], scale, b0.z); x[3] = fmaf(a[3], scale, b0.w);
    x[4] = fmaf(a[4], scale, b1.x); x[5] = fmaf(a[5], scale, b1.y); x[6] = fmaf(a[6], scale, b1.z); x[7] = fmaf(a[7], scale, b1.w);
    if (RELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
    }
    acc(col0, x);
    ap.store8(grp, x);
    ap.end_shared();
    if (tr) tr->mark(12);
  } else if (grp < nblk) {
    tmem_ld8(taddr + grp * 32, v);
  }
#pragma unroll 1
  for (int cb = cb0; cb < nblk; cb += kGroups) {
    if (EMIT) ap.begin(cb);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      tmem_ld_wait();
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __uint_as_float(v[i]);
      // next 8 columns in flight while these are processed
      if (t < 3) tmem_ld8(taddr + cb * 32 + (t + 1) * 8, v);
      else if (cb + kGroups < nblk) tmem_ld8(taddr + (cb + kGroups) * 32, v);
      const int col0 = cb * 32 + t * 8;
      const float4 b0 = *reinterpret_cast<const float4*>(bias + col0), b1 = *reinterpret_cast<const float4*>(bias + col0 + 4);
      float x[8];
      x[0] = fmaf(a[0], scale, b0.x); x[1] = fmaf(a[1], scale, b0.y); x[2] = fmaf(a[2], scale, b0.z); x[3] = fmaf(a[3], scale, b0.w);
      x[4] = fmaf(a[4], scale, b1.x); x[5] = fmaf(a[5], scale, b1.y); x[6] = fmaf(a[6], scale, b1.z); x[7] = fmaf(a[7], scale, b1.w);
      if (RELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
      }
      acc(col0, x);
      if (EMIT) ap.store8(t, x);
    }
    if (EMIT) ap.end();
    if (tr) tr->mark(12);                    // 12: one column block drained (and its chunk published)
  }
  if (EMIT) {
    const int padded = round_up(nblk, kGroups);            // hidden parts are padded to a multiple of 4 chunks
    for (int cb = nblk + ((grp - nblk) % kGroups + kGroups) % kGroups; cb < padded; cb += kGroups) {
      ap.begin(cb);
      for (int t = 0; t < 4; ++t) ap.zero8(t);
      ap.end();
    }
    ap.base += padded;
  }
  tc_fence_before_sync();
}

// One tile (128 rows) through one network, this group's share.  Returns this group's partial
// (rgb logits, raw sigma) of the thread's row; the four groups' partials add up to the result
// (biases are added by group 0).  DENSITY: trunk + alpha only.
template <int FMT, bool DENSITY>
__device__ __forceinline__ float4 worker_net_pass(AProducer<FMT>& ap, const Pipe& pp, uint32_t (&d_cnt)[2],
                                                  const RowCtx& rc, const RenderKParams& P, const float* sm,
                                                  int quarter, int grp, int row, float* wj, Trace* tr = nullptr) {
  const NetProgram& pg = P.prog;
  const int D = pg.dims.D, W = pg.dims.W;
  float sigma = 0.f;
  if (!DENSITY) compute_view_weights(rc, P, grp, row, wj);     // consumed after the trunk, behind a barrier
  const float* wa = sm + pg.sm.alpha_w;
  // operands of trunk layers 0..D-1: [encoding part] + drain of layer l-1
#pragma unroll 1
  for (int l = 0; l < D; ++l) {
    if (l == 0 || (l - 1) == pg.dims.skip) {
      if (tr) tr->mark(20);                  // 20/21: encoding part begin/end
      produce_pts_chunks<FMT>(ap, rc, P, grp);
      if (tr) tr->mark(21);
    }
    if (l > 0)
      drain_region<FMT, true, true>(ap, pp, d_cnt, (l - 1) & 1, W, sm + pg.sm.bias[l - 1], sm[l - 1], quarter, grp,
                                    [](int, const float (&)[8]) {}, tr, /*split0=*/kSplitFirstChunk && (l - 1) != pg.dims.skip);
  }
  // h of the last trunk layer: alpha_linear in fp32 on the way; operand of feature_linear unless DENSITY
  auto alpha_acc = [&](int col0, const float (&x)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sigma = fmaf(x[i], wa[col0 + i], sigma);
  };
  if (DENSITY) {
    drain_region<FMT, true, false>(ap, pp, d_cnt, (D - 1) & 1, W, sm + pg.sm.bias[D - 1], sm[D - 1], quarter, grp, alpha_acc, tr);
    if (grp == 0) sigma += sm[pg.sm.alpha_b];
    return make_float4(0.f, 0.f, 0.f, sigma);
  }
  // operand of the views layer (feature_linear folded in): the view encoding first -- it does not depend on the
  // trunk, so it is produced while the last trunk layer's MMAs run -- then h of the last trunk layer
  if (tr) tr->mark(22);
  asm volatile("bar.sync 5, 512;" ::: "memory");              // all groups' view weights are in shared memory
  produce_slot_chunks<FMT>(ap, rc, P, grp, row, wj);
  if (tr) tr->mark(23);
  drain_region<FMT, true, true>(ap, pp, d_cnt, (D - 1) & 1, W, sm + pg.sm.bias[D - 1], sm[D - 1], quarter, grp, alpha_acc, tr);
  if (grp == 0) sigma += sm[pg.sm.alpha_b];
  // views layer output -> rgb_linear in fp32
  float r0 = 0.f, r1 = 0.f, r2 = 0.f;
  const float* wr = sm + pg.sm.rgb_w;
  const int H = W / 2;
  drain_region<FMT, true, false>(ap, pp, d_cnt, D & 1, H, sm + pg.sm.bias[D], sm[D], quarter, grp,
                                 [&](int col0, const float (&x)[8]) {
#pragma unroll
                                   for (int i = 0; i < 8; ++i) {
                                     r0 = fmaf(x[i], wr[col0 + i], r0);
                                     r1 = fmaf(x[i], wr[H + col0 + i], r1);
                                     r2 = fmaf(x[i], wr[2 * H + col0 + i], r2);
                                   }
                                 }, tr);
  if (grp == 0) {
    const float* br = sm + pg.sm.rgb_b;
    r0 += br[0]; r1 += br[1]; r2 += br[2];
  }
  return make_float4(r0, r1, r2, sigma);
}

// ------------------------------------------------------------------------------------------------
// per-ray stages
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// a12 (nerf.py:150-205): alpha compositing of one ray by one warp.  z, raw: shared [S]; wts: shared [S] (out).
// P: anything with the density options B (density_scale), softplus, shift (the fused kernel passes its parameter block)
template <typename PT>
__device__ __forceinline__ void composite_ray(int lane, int S, const float* z, const float4* raw,
                                              float dnorm, const float* noise, const PT& P, float* wts,
                                              float* alpha_out, float* rgb_out, float* disp_out, float* acc_out) {
  const int per = (S + 31) / 32;
  const int i0 = lane * per;
  float prod = 1.f;
  for (int k = 0; k < per; ++k) {
    int i = i0 + k;
    if (i < S) {
      float dist = (i + 1 < S ? z[i + 1] - z[i] : 1e10f) * dnorm;
      float sg = density_act(raw[i].w, P.B, noise ? noise[i] : 0.f, P.softplus, P.shift);
      float a = 1.f - expf(-sg * dist);
      wts[i] = a;
      prod *= (1.f - a + 1e-10f);
    }
  }
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= t;
  }
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.f;
  float cr = 0.f, cg = 0.f, cb = 0.f, depth = 0.f, wsum = 0.f;
  for (int k = 0; k < per; ++k) {
    int i = i0 + k;
    if (i < S) {
      float a = wts[i];
      float w = a * T;
      T *= (1.f - a + 1e-10f);
      if (alpha_out) alpha_out[i] = a;
      wts[i] = w;
      float4 r = raw[i];
      cr = fmaf(w, sigmoid_rgb(r.x), cr);
      cg = fmaf(w, sigmoid_rgb(r.y), cg);
      cb = fmaf(w, sigmoid_rgb(r.z), cb);
      depth = fmaf(w, z[i], depth);
      wsum += w;
    }
  }
  cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); depth = warp_sum(depth); wsum = warp_sum(wsum);
  if (lane == 0) {
    if (rgb_out) { rgb_out[0] = cr; rgb_out[1] = cg; rgb_out[2] = cb; }
    float disp = 1.f / fmaxf(1e-10f, depth / (wsum + 1e-10f));
    if (fabsf(wsum) <= 1e-8f) disp = 0.f;             // torch.isclose(sum, 0): atol 1e-8
    if (disp_out) *disp_out = disp;
    if (acc_out) *acc_out = fminf(wsum, 1.f);
  }
  __syncwarp();
}

// a13, first part (ray_utils.py:157-166): cdf of the coarse weights of one ray, by one warp.
// w: shared [Sc]; cdf: shared [Sc] (Sc-1 entries used).  fp64 running sum like torch's CPU cumsum.
// `blur` (single_net, ray_utils.py:271-277): the pdf is 0.5 (max(w_l, w_k) + max(w_k, w_u)) + 0.01 instead of w_k.
__device__ __forceinline__ void importance_cdf(int lane, int Sc, const float* w, float* cdf, int blur) {
  const int nw = Sc - 2;
  auto wt = [&](int i) {
    const float wk = blur ? 0.5f * (fmaxf(w[i], w[1 + i]) + fmaxf(w[1 + i], w[2 + i])) + 0.01f : w[1 + i];
    return wk + 1e-5f;
  };
  double part = 0.0;
  for (int i = lane; i < nw; i += 32) part += (double)wt(i);
  const float total = (float)warp_sum(part);
  const int per = (nw + 31) / 32;
  const int i0 = lane * per;
  double loc = 0.0;
  for (int k = 0; k < per; ++k) {
    int i = i0 + k;
    if (i < nw) loc += (double)(wt(i) / total);
  }
  double incl = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  double run = incl - loc;
  for (int k = 0; k < per; ++k) {
    int i = i0 + k;
    if (i < nw) {
      run += (double)(wt(i) / total);
      cdf[i + 1] = (float)run;
    }
  }
  if (lane == 0) cdf[0] = 0.f;
}
// a13, second part (ray_utils.py:168-201): one inverse-CDF sample.  searchsorted(cdf, u, right=True).
__device__ __forceinline__ float importance_sample(float u, int Sc, const float* zc, const float* cdf) {
  const int nb = Sc - 1;
  int lo = 0, hi = nb;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  int below = max(lo - 1, 0), above = min(lo, nb - 1);
  float cb = cdf[below], ca = cdf[above];
  float bb = 0.5f * (zc[below + 1] + zc[below]), ba = 0.5f * (zc[above + 1] + zc[above]);
  float denom = ca - cb;
  if (denom < 1e-5f) denom = 1.f;
  return bb + (u - cb) / denom * (ba - bb);
}

// ------------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------------
template <int FMT, bool DENSITY>
// 18 warps: registers are allocated for 20 (warp count rounded to a multiple of 4), hence the cap of 96 per thread
__global__ void __launch_bounds__(kThreads, 1) anerf_fused_kernel(const __grid_constant__ RenderKParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemLayout& L = P.sl;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const NetProgram& pg = P.prog;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  Pipe pp;
  pipe_init(pp, smem + L.a_ring, smem + L.b_ring, bars, P.status);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem_ptr);

  if (tid == 0) pipe_init_barriers(pp);
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  // both networks' small fp32 parameters -> shared memory
  {
    const int nf = pg.sm.fixed_floats;
    for (int net = 0; net < 2; ++net) {
      const float* src = reinterpret_cast<const float*>(P.packed[net] + pg.smalls_off);
      float* dst = reinterpret_cast<float*>(smem + (net ? L.smalls1 : L.smalls0));
      for (int i = tid; i < nf; i += kThreads) dst[i] = src[i];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();            // barrier inits of both CTAs are visible before any remote arrive
  tc_fence_after_sync();
  pp.tmem_base = *tmem_slot;

  // both CTAs of a pair run the same number of items in lockstep (the MMAs span the pair); an item index
  // past the end is processed as a dummy (clamped inputs, no outputs)
  const int n_iter = (P.n_items + (int)gridDim.x - 1) / (int)gridDim.x;
  const int passes = DENSITY ? 1 : (P.tilesC + P.tilesF);
  const int nl = DENSITY ? pg.dims.D : pg.n_layers;

  if (warp == kMmaWarp) {
    uint32_t a_seq = 0, b_seq = 0;
    Trace trc; trc.init(lane == 0 ? P.trace : nullptr, 0);
    if (pp.rank == 0) {            // leader: the whole warp walks the program, one elected lane issues
      for (int it = 0; it < n_iter; ++it)
        for (int ps = 0; ps < passes; ++ps)
          for (int l = 0; l < nl; ++l)
          {
            const bool views = !DENSITY && l == pg.dims.D;      // leading ray-slot chunks read G from shared memory
            mma_layer<FMT>(pp, a_seq, b_seq, pg.layer[l].n, pg.layer[l].chunks + (views ? P.slotc : 0), l & 1,
                           trc.p ? &trc : nullptr, views ? P.slotc : 0, smem_u32(smem + L.g_buf));
          }
    } else if (lane == 0) {        // peer: relay "my weight half has landed"
      for (int it = 0; it < n_iter; ++it)
        for (int ps = 0; ps < passes; ++ps)
          for (int l = 0; l < nl; ++l) relay_layer(pp, b_seq, pg.layer[l].chunks);
    }
    __syncwarp();
  } else if (warp == kLoadWarp) {
    if (lane == 0) {
      uint32_t b_seq = 0;
      for (int it = 0; it < n_iter; ++it)
        for (int ps = 0; ps < passes; ++ps) {
          const uint8_t* img = P.packed[(!DENSITY && ps >= P.tilesC) ? 1 : 0];
          for (int l = 0; l < nl; ++l) load_layer(pp, b_seq, img + pg.layer[l].w_off, pg.layer[l].n, pg.layer[l].chunks);
        }
    }
    __syncwarp();
  } else {
    // ------------------------------- worker warps (rows) -------------------------------------
    const int grp = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    AProducer<FMT> ap(pp, row);
    uint32_t d_cnt[2] = {0u, 0u};
    Trace trc; trc.init((quarter == 0 && lane == 0 && grp < 2) ? P.trace : nullptr, 1 + grp);
    Trace* tr = trc.p ? &trc : nullptr;
    float* ray_s = reinterpret_cast<float*>(smem + L.ray);
    float* skt_s = reinterpret_cast<float*>(smem + L.skt);
    float* vtab_s = reinterpret_cast<float*>(smem + L.view_tab);
    int* fcrow_s = reinterpret_cast<int*>(smem + L.fcode);
    int* gctr_s = reinterpret_cast<int*>(smem + L.task_ctr);
    uint8_t* g_buf = smem + L.g_buf;
    float* wj_s = reinterpret_cast<float*>(smem + L.wj);
    float* zc_s = reinterpret_cast<float*>(smem + L.z_coarse);
    float* za_s = reinterpret_cast<float*>(smem + L.z_all);
    float4* raw_s = reinterpret_cast<float4*>(smem + L.raw);
    float4* part_s = reinterpret_cast<float4*>(smem + L.part);
    float* w_s = reinterpret_cast<float*>(smem + L.wts);
    float* cdf_s = reinterpret_cast<float*>(smem + L.cdf);
    const float* sm0 = reinterpret_cast<const float*>(smem + L.smalls0);
    const float* sm1 = reinterpret_cast<const float*>(smem + L.smalls1);
    const int J = pg.dims.J;

    if (DENSITY) {
      // one pose for the whole launch
      for (int i = tid; i < J * 12; i += kWorkerThreads) skt_s[i] = P.skts[(i / 12) * 16 + (i % 12)];
      worker_sync();
      for (int it = 0; it < n_iter; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        long long idx = (long long)item * kTileM + row;
        bool valid = idx < P.n_points;
        long long ci = valid ? idx : (P.n_points - 1);
        RowCtx rc;
        if (P.pts) {
          rc.p[0] = P.pts[ci * 3 + 0]; rc.p[1] = P.pts[ci * 3 + 1]; rc.p[2] = P.pts[ci * 3 + 2];
        } else {
          // flat index -> (a, b, c) of np.meshgrid(t, t, t) ('xy' indexing: x = t[b], y = t[a], z = t[c]); the
          // coordinate is np.linspace's fp64 value rounded to fp32, then the fp32 add of the root joint
          const long long f = P.grid_first + ci;
          const int n1 = P.grid_n1;
          const int a = (int)(f / ((long long)n1 * n1)), rem = (int)(f % ((long long)n1 * n1));
          const int b = rem / n1, c = rem % n1;
          auto tv = [&](int i) { return i == n1 - 1 ? (float)P.grid_stop : (float)__dadd_rn(__dmul_rn((double)i, P.grid_step), P.grid_start); };   // two roundings, like numpy
          rc.p[0] = tv(b) + __ldg(P.grid_origin); rc.p[1] = tv(a) + __ldg(P.grid_origin + 1); rc.p[2] = tv(c) + __ldg(P.grid_origin + 2);
        }
        rc.skt = skt_s; rc.slot = 0;
        float4 r = worker_net_pass<FMT, true>(ap, pp, d_cnt, rc, P, sm0, quarter, grp, row, nullptr);
        part_s[grp * kTileM + row] = r;
        worker_sync();
        if (grp == 0 && valid)
          P.sigma[idx] = part_s[row].w + part_s[kTileM + row].w + part_s[2 * kTileM + row].w + part_s[3 * kTileM + row].w;
        worker_sync();
      }
    } else {
      const int R = P.R, Sc = P.Sc, Sf = P.Sf, Si = P.Si;
      const bool fine = Si > 0;
      const int rank = (int)pp.rank;
      {   // K rows of the ray-slot chunks that no joint uses (and whole padding chunks) stay zero for the whole launch
        uint4* z = reinterpret_cast<uint4*>(g_buf);
        for (int i = tid; i < P.slotc * (pg.dims.W / 2) * 4; i += kWorkerThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
      }
      // ---- (1)(2) per-ray inputs of item `it_`: rays, bone transforms, framecode rows, view-direction table of
      // the pair's 2R ray slots, coarse depths.  Run by threads t0, t0 + nt, ... (all workers before the first
      // item; afterwards the warps that are not compositing the previous item's rays).  No barrier inside.
      auto stage_inputs = [&](int it_, int t0, int nt) {
        const int ray0 = (blockIdx.x + it_ * (int)gridDim.x) * R;   // >= n_rays for a dummy item: every load clamps
        // ray slots of the CTA pair: slot = owner rank * R + ray index; the peer's item is the neighbouring one
        auto slot_ray = [&](int slot) {
          const int o = slot / R, r = slot % R;
          const int it_o = (int)blockIdx.x - rank + o + it_ * (int)gridDim.x;
          return min(it_o * R + r, P.n_rays - 1);
        };
        if (t0 == 0) gctr_s[1] = 0;        // counter of the fine network's build (mid-item, many barriers away)
        for (int i = t0; i < R; i += nt) {
          int gr = min(ray0 + i, P.n_rays - 1);
          float rbuf[8];
          const float* rp = rbuf;
          if (P.rays) rp = P.rays + (size_t)gr * 8; else pixel_ray(P.gen, gr, rbuf);
          float* d = ray_s + i * 12;
          d[0] = rp[0]; d[1] = rp[1]; d[2] = rp[2]; d[3] = rp[3]; d[4] = rp[4]; d[5] = rp[5];
          d[6] = P.nearfar[gr * 2]; d[7] = P.nearfar[gr * 2 + 1];
          d[8] = sqrtf(rp[3] * rp[3] + rp[4] * rp[4] + rp[5] * rp[5]);
        }
        for (int i = t0; i < R * J * 12; i += nt) {
          int r = i / (J * 12), e = i % (J * 12);
          int gr = min(ray0 + r, P.n_rays - 1);
          const size_t prow = pose_row(P, gr);
          skt_s[i] = P.skts[prow * P.skt_stride + (e / 12) * 16 + (e % 12)];
        }
        if (pg.dims.fc_ch > 0)
          for (int i = t0; i < 2 * R; i += nt) {
            int cam = P.eval_mean_fc ? pg.dims.n_fc : (int)(P.cams ? P.cams[slot_ray(i)] : P.cam_const);
            fcrow_s[i] = min(max(cam, 0), pg.dims.n_fc);
          }
        const int vstride = view_tab_jstride(R);
        for (int u = t0; u < 2 * R * J; u += nt) {
          const int slot = u / J, j = u % J;
          const int gr = slot_ray(slot);
          float f[kViewPerJoint];
          float rbuf[8];
          const float* rd = rbuf + 3;
          if (P.rays) rd = P.rays + (size_t)gr * 8 + 3; else pixel_ray(P.gen, gr, rbuf);
          const size_t prow = pose_row(P, gr);
          encode_joint_viewdir(P.skts + prow * P.skt_stride + (size_t)j * 16, rd, f);
#pragma unroll
          for (int q = 0; q < kViewPerJoint; ++q) vtab_s[j * vstride + slot * kViewPad + q] = f[q];
          vtab_s[j * vstride + slot * kViewPad + kViewPerJoint] = 0.f;
        }
        for (int i = t0; i < R * Sc; i += nt) {
          int r = i / Sc, sidx = i % Sc;
          const int gr = min(ray0 + r, P.n_rays - 1);
          float near = P.nearfar[gr * 2], far = P.nearfar[gr * 2 + 1];
          auto zat = [&](int k) {
            float t = linspace01(k, Sc);
            return P.lindisp ? 1.f / (1.f / near * (1.f - t) + 1.f / far * t) : near * (1.f - t) + far * t;
          };
          float z = zat(sidx);
          if (P.t_rand) {
            float lower = sidx == 0 ? z : 0.5f * (zat(sidx - 1) + z);
            float upper = sidx == Sc - 1 ? z : 0.5f * (z + zat(sidx + 1));
            z = lower + (upper - lower) * P.t_rand[(size_t)gr * Sc + sidx];
          }
          zc_s[i] = z;
        }
      };
      if (tid == 0) gctr_s[0] = 0;
      stage_inputs(0, tid, kWorkerThreads);
      worker_sync();
      build_view_matrices<FMT>(P, 0, g_buf, vtab_s, fcrow_s, gctr_s, rank, 1.0f / sm0[pg.dims.D]);
      worker_sync();
      for (int it = 0; it < n_iter; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int ray0 = item * R;       // >= n_rays for a dummy item: every load clamps, every store is guarded
        // ---- (3)/(6) network passes: tilesC coarse tiles, then tilesF fine tiles; (4)(5)(7) between ------
#pragma unroll 1
        for (int ps = 0; ps < passes; ++ps) {
          const bool is_fine = ps >= P.tilesC;
          const int S = is_fine ? Sf : Sc;
          const float* zsrc = is_fine ? za_s : zc_s;
          int g = (is_fine ? ps - P.tilesC : ps) * kTileM + row;
          int r = g / S, s = g % S;
          bool valid = r < R;
          if (!valid) { r = R - 1; s = S - 1; }
          float z = zsrc[r * S + s];
          RowCtx rc;
          const float* rr = ray_s + r * 12;
          rc.p[0] = rr[0] + rr[3] * z; rc.p[1] = rr[1] + rr[4] * z; rc.p[2] = rr[2] + rr[5] * z;
          rc.skt = skt_s + r * J * 12; rc.slot = rank * R + r;
          if (tr) tr->mark(30 + ps);
          float4 o = worker_net_pass<FMT, false>(ap, pp, d_cnt, rc, P, is_fine ? sm1 : sm0, quarter, grp, row, wj_s, tr);
          if (tr) tr->mark(40 + ps);
          part_s[grp * kTileM + row] = o;
          worker_sync();
          if (grp == 0 && valid)
            raw_s[g] = add4(add4(part_s[row], part_s[kTileM + row]), add4(part_s[2 * kTileM + row], part_s[3 * kTileM + row]));
          if (ps == P.tilesC - 1) {
            worker_sync();
            if (tid == 0) gctr_s[0] = 0;   // counter of the next item's coarse build (item end, many barriers away)
            // ---- (4) composite coarse: one warp per ray ----------------------------------------------
            for (int q = warp; q < R; q += kWorkerWarps) {
              int gr = ray0 + q;
              bool live = gr < P.n_rays;
              int grc = min(gr, P.n_rays - 1);
              float* a_out = fine ? P.alpha0 : P.alpha;
              float* rgb_o = fine ? P.rgb0 : P.rgb_map;
              float* disp_o = fine ? P.disp0 : P.disp_map;
              float* acc_o = fine ? P.acc0 : P.acc_map;
              composite_ray(lane, Sc, zc_s + q * Sc, raw_s + q * Sc, ray_s[q * 12 + 8],
                            P.noise0 ? P.noise0 + (size_t)grc * Sc : nullptr, P, w_s + q * Sc,
                            (live && a_out) ? a_out + (size_t)gr * Sc : nullptr,
                            (live && rgb_o) ? rgb_o + (size_t)gr * 3 : nullptr,
                            (live && disp_o) ? disp_o + gr : nullptr, (live && acc_o) ? acc_o + gr : nullptr);
              if (!fine && live && P.raw_out)
                for (int i = lane; i < Sc; i += 32) reinterpret_cast<float4*>(P.raw_out)[(size_t)gr * Sc + i] = raw_s[q * Sc + i];
              if (fine) {
                importance_cdf(lane, Sc, w_s + q * Sc, cdf_s + q * Sc, P.blur_is);
                if (kSortOnRayWarp) {
                  // ---- (5) importance sampling + sorted merge of this ray by its own warp, while the other warps
                  // build the fine network's view matrices (part_s is free between the passes: scratch)
                  __syncwarp();
                  float* t = reinterpret_cast<float*>(part_s) + q * Sf;
                  for (int e = lane; e < Sf; e += 32) {
                    float v;
                    if (e < Sc) {
                      v = zc_s[q * Sc + e];
                    } else {
                      const int m = e - Sc;
                      const float u = P.u_rand ? P.u_rand[(size_t)grc * Si + m] : linspace01(m, Si);
                      v = importance_sample(u, Sc, zc_s + q * Sc, cdf_s + q * Sc);
                    }
                    t[e] = v;
                  }
                  __syncwarp();
                  for (int e = lane; e < Sf; e += 32) {        // rank sort (stable) == torch.sort on values
                    const float x = t[e];
                    int rk = 0;
                    for (int k = 0; k < Sf; ++k) {
                      const float y = t[k];
                      rk += (y < x || (y == x && k < e)) ? 1 : 0;
                    }
                    za_s[q * Sf + rk] = x;
                  }
                  __syncwarp();
                  if (P.z_all_out && live)
                    for (int e = lane; e < Sf; e += 32) P.z_all_out[(size_t)gr * Sf + e] = za_s[q * Sf + e];
                }
              }
            }
            if (tr) tr->mark(52);
            if (fine) build_view_matrices<FMT>(P, 1, g_buf, vtab_s, fcrow_s, gctr_s + 1, rank, 1.0f / sm1[pg.dims.D]);
            if (tr) tr->mark(53);
            worker_sync();   // raw_s is free from here (scratch for the merge below)
            if (fine && !kSortOnRayWarp) {
              // ---- (5) importance sampling: every worker thread takes samples, then ranks -------------
              float* tmp = reinterpret_cast<float*>(raw_s);
              for (int i = tid; i < R * Sf; i += kWorkerThreads) {
                int q = i / Sf, e = i % Sf;
                float v;
                if (e < Sc) {
                  v = zc_s[q * Sc + e];
                } else {
                  int m = e - Sc;
                  int grc = min(ray0 + q, P.n_rays - 1);
                  float u = P.u_rand ? P.u_rand[(size_t)grc * Si + m] : linspace01(m, Si);
                  v = importance_sample(u, Sc, zc_s + q * Sc, cdf_s + q * Sc);
                }
                tmp[i] = v;
              }
              worker_sync();
              for (int i = tid; i < R * Sf; i += kWorkerThreads) {   // rank sort (stable) == torch.sort on values
                int q = i / Sf, e = i % Sf;
                const float* t = tmp + q * Sf;
                float x = t[e];
                int rk = 0;
                for (int k = 0; k < Sf; ++k) {
                  float y = t[k];
                  rk += (y < x || (y == x && k < e)) ? 1 : 0;
                }
                za_s[q * Sf + rk] = x;
              }
              worker_sync();
              if (P.z_all_out)
                for (int i = tid; i < R * Sf; i += kWorkerThreads)
                  if (ray0 + i / Sf < P.n_rays) P.z_all_out[(size_t)ray0 * Sf + i] = za_s[i];
            }
          }
        }
        // ---- item end: (7) composite fine (one warp per ray) while the other warps stage the next item's inputs
        // and build its coarse-network view matrices (the last pass's MMAs are done: g_buf is free) ----------------
        const bool has_next = it + 1 < n_iter;
        static_assert(kMaxRaysPerItem < kWorkerWarps, "one compositing warp per ray, and warps left for staging");
        if (fine) {
          const float dnorm = warp < R ? ray_s[warp * 12 + 8] : 0.f;    // ray_s is about to be overwritten
          worker_sync();
          if (warp < R) {
            if (tr) tr->mark(54);
            const int q = warp;
            int gr = ray0 + q;
            bool live = gr < P.n_rays;
            int grc = min(gr, P.n_rays - 1);
            composite_ray(lane, Sf, za_s + q * Sf, raw_s + q * Sf, dnorm,
                          P.noise1 ? P.noise1 + (size_t)grc * Sf : nullptr, P, w_s + q * Sf,
                          (live && P.alpha) ? P.alpha + (size_t)gr * Sf : nullptr,
                          (live && P.rgb_map) ? P.rgb_map + (size_t)gr * 3 : nullptr,
                          (live && P.disp_map) ? P.disp_map + gr : nullptr,
                          (live && P.acc_map) ? P.acc_map + gr : nullptr);
            if (live && P.raw_out)
              for (int i = lane; i < Sf; i += 32) reinterpret_cast<float4*>(P.raw_out)[(size_t)gr * Sf + i] = raw_s[q * Sf + i];
            if (tr) tr->mark(55);
          } else if (has_next) {
            const int nt = kWorkerThreads - 32 * R;
            if (tr) tr->mark(50);
            stage_inputs(it + 1, tid - 32 * R, nt);
            asm volatile("bar.sync 6, %0;" ::"r"(nt) : "memory");        // the staging warps only
            build_view_matrices<FMT>(P, 0, g_buf, vtab_s, fcrow_s, gctr_s, rank, 1.0f / sm0[pg.dims.D]);
            if (tr) tr->mark(51);
          }
          worker_sync();
        } else if (has_next) {
          stage_inputs(it + 1, tid, kWorkerThreads);
          worker_sync();
          build_view_matrices<FMT>(P, 0, g_buf, vtab_s, fcrow_s, gctr_s, rank, 1.0f / sm0[pg.dims.D]);
          worker_sync();
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();            // the peer may still be reading this CTA's barriers / TMEM pair state
  if (warp == kMmaWarp) tmem_dealloc(pp.tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// a2 pre-kernel: near/far of every ray of the chunk + the chunk-wide nanmean repair
// (ray_utils.py:292-344).  One CTA; the reduction is what makes the result chunk-dependent.
// ------------------------------------------------------------------------------------------------
// rays == NULL: frame mode, rays generated per pixel from `gen`; cyl_stride = 5 (per-ray cylinders) or 0 (one per frame)
__global__ void __launch_bounds__(1024, 1) anerf_nearfar_kernel(const float* __restrict__ rays, const RayGen gen,
                                                                const float* __restrict__ cyls, int cyl_stride, int n,
                                                                float* __restrict__ nearfar) {
  __shared__ double s_sum[2][32];
  __shared__ int s_cnt[2][32];
  __shared__ float s_mean[2];
  __shared__ int s_any;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double sn = 0.0, sf = 0.0;
  int cn = 0, cf = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    float rbuf[8];
    const float* r = rbuf;
    if (rays) r = rays + (size_t)i * 8; else pixel_ray(gen, i, rbuf);
    float nn, ff;
    bool miss;
    near_far_cylinder(r, r + 3, cyls + (size_t)i * cyl_stride, r[6], r[7], nn, ff, miss);
    nearfar[2 * i] = nn;
    nearfar[2 * i + 1] = ff;
    if (nn == nn) { sn += nn; ++cn; }
    if (ff == ff) { sf += ff; ++cf; }
  }
  sn = warp_sum(sn); sf = warp_sum(sf);
  cn = __reduce_add_sync(0xffffffffu, cn); cf = __reduce_add_sync(0xffffffffu, cf);
  if (lane == 0) { s_sum[0][warp] = sn; s_sum[1][warp] = sf; s_cnt[0][warp] = cn; s_cnt[1][warp] = cf; }
  __syncthreads();
  if (warp == 0) {
    int nwarps = blockDim.x >> 5;
    double a = lane < nwarps ? s_sum[0][lane] : 0.0, b = lane < nwarps ? s_sum[1][lane] : 0.0;
    int ca = lane < nwarps ? s_cnt[0][lane] : 0, cb = lane < nwarps ? s_cnt[1][lane] : 0;
    a = warp_sum(a); b = warp_sum(b);
    ca = __reduce_add_sync(0xffffffffu, ca); cb = __reduce_add_sync(0xffffffffu, cb);
    if (lane == 0) {
      s_mean[0] = ca > 0 ? (float)(a / ca) : nanf("");
      s_mean[1] = cb > 0 ? (float)(b / cb) : nanf("");
      s_any = (ca < n) ? 1 : 0;       // torch.isnan(new_near).any()
    }
  }
  __syncthreads();
  if (!s_any) return;
  const float mn = s_mean[0], mf = s_mean[1];
  for (int i = tid; i < n; i += blockDim.x) {
    float rbuf[8];
    const float* r = rbuf;
    if (rays) r = rays + (size_t)i * 8; else pixel_ray(gen, i, rbuf);
    float nn, ff;
    bool miss;
    near_far_cylinder(r, r + 3, cyls + (size_t)i * cyl_stride, r[6], r[7], nn, ff, miss);
    if (miss) {     // rows where Q is NaN get the chunk mean (or the original bound if no ray hit)
      nearfar[2 * i] = (mn == mn) ? mn : r[6];
      nearfar[2 * i + 1] = (mf == mf) ? mf : r[7];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 [N_out, K_in] -> chunks of 2 halves x [hi: 4 x (N/2)/8 x (8 x 8)] [lo: same], K permuted by kmap
// ------------------------------------------------------------------------------------------------
// Per-layer operand scale.  bf16 operands: 1.  fp16 operands: the power of two that brings max|W| into
// [2^12, 2^13), so that the lo parts (|lo| <= 2^-11 |hi|) stay normal fp16 numbers; the fused kernel
// multiplies the accumulators by the exact inverse (`inv_scale`, smalls header) before adding the bias.
// The drain scale written to the smalls header also carries `comp` = 1 + expected relative loss of the tensor core's
// fp32 accumulation (trunc_comp() below); `pure_scale` (what the pack kernel divides the weights by) stays a power of two.
template <int FMT>
__global__ void anerf_pack_layer_kernel(const float* __restrict__ w, int k_in, const int* __restrict__ kmap,
                                        int n, int chunks, const float* __restrict__ inv_scale,
                                        uint8_t* __restrict__ out) {
  const float scale = 1.0f / *inv_scale;   // exact: power of two
  // one thread per (chunk, n, 8-wide k group): writes 16 B hi + 16 B lo
  long long total = (long long)chunks * n * 4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int g = (int)(t & 3);
    int nn = (int)((t >> 2) % n);
    int c = (int)((t >> 2) / n);
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int src = kmap[c * kKC + g * 8 + i];
      x[i] = src >= 0 ? w[(size_t)nn * k_in + src] * scale : 0.f;
    }
    uint4 hi, lo;
    Split<FMT>::pair(x[0], x[1], hi.x, lo.x);
    Split<FMT>::pair(x[2], x[3], hi.y, lo.y);
    Split<FMT>::pair(x[4], x[5], hi.z, lo.z);
    Split<FMT>::pair(x[6], x[7], hi.w, lo.w);
    // chunk = [half 0: hi, lo][half 1: hi, lo]; a half holds n/2 rows (the B rows one CTA of the pair feeds)
    const int nh = n >> 1;
    uint8_t* half = out + (size_t)c * n * 128 + (size_t)(nn / nh) * n * 64;
    const int r = nn % nh;
    size_t off = (size_t)g * nh * 16 + (size_t)(r >> 3) * 128 + (size_t)(r & 7) * 16;
    *reinterpret_cast<uint4*>(half + off) = hi;
    *reinterpret_cast<uint4*>(half + (size_t)nh * 64 + off) = lo;
  }
}

// The whole image in a handful of launches (the optimizer re-packs after every step: SURVEY.md 8(f) row 3): one
// launch computes every layer's scale, one packs every layer, one copies the small fp32 parameters.
struct PackAllArgs {
  const float* w[kMaxLayers];      // fp32 [n, k_in] (the views layer: the folded matrix)
  const float* b[kMaxLayers];      // fp32 [n]
  const int* kmap[kMaxLayers];
  int k_in[kMaxLayers], n[kMaxLayers], chunks[kMaxLayers];
  unsigned w_off[kMaxLayers];
  int bias_off[kMaxLayers];        // offsets in floats inside the smalls block
  float comp[kMaxLayers];
  int n_layers, fmt;
  float* pure_scale;               // [kMaxLayers] scratch
  uint8_t* img;
  float* smalls;
  // heads
  const float *alpha_w, *alpha_b, *rgb_w, *rgb_b;
  int alpha_w_off, alpha_b_off, rgb_w_off, rgb_b_off, W;
};

// grid = n_layers blocks of 1024 threads: per-layer operand scale (fp16 operands: the power of two that brings max|W| into [2^12, 2^13); bf16: 1)
__global__ void anerf_all_scales_kernel(const __grid_constant__ PackAllArgs a) {
  __shared__ float s_max[32];
  const int l = blockIdx.x;
  const float* w = a.w[l];
  const long long count = (long long)a.n[l] * a.k_in[l];
  float m = 0.f;
  if (a.fmt == 0)
    for (long long i = threadIdx.x; i < count; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, s_max[i]);
    int k = 0;
    if (a.fmt == 0 && m > 0.f && m < 3.0e38f) {
      int e;
      frexpf(m, &e);
      k = min(max(13 - e, -24), 24);
    }
    a.pure_scale[l] = ldexpf(1.0f, -k);
    a.smalls[l] = ldexpf(1.0f, -k) * a.comp[l];
  }
}

// grid = (blocks, n_layers): blockIdx.y picks the layer; also copies the layer's bias
template <int FMT>
__global__ void anerf_pack_all_kernel(const __grid_constant__ PackAllArgs a) {
  const int l = blockIdx.y;
  const float* __restrict__ w = a.w[l];
  const int* __restrict__ kmap = a.kmap[l];
  const int n = a.n[l], k_in = a.k_in[l], chunks = a.chunks[l];
  uint8_t* out = a.img + a.w_off[l];
  const float scale = 1.0f / a.pure_scale[l];   // exact: power of two
  const long long total = (long long)chunks * n * 4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t & 3);
    const int nn = (int)((t >> 2) % n);
    const int c = (int)((t >> 2) / n);
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int src = kmap[c * kKC + g * 8 + i];
      x[i] = src >= 0 ? w[(size_t)nn * k_in + src] * scale : 0.f;
    }
    uint4 hi, lo;
    Split<FMT>::pair(x[0], x[1], hi.x, lo.x);
    Split<FMT>::pair(x[2], x[3], hi.y, lo.y);
    Split<FMT>::pair(x[4], x[5], hi.z, lo.z);
    Split<FMT>::pair(x[6], x[7], hi.w, lo.w);
    const int nh = n >> 1;
    uint8_t* half = out + (size_t)c * n * 128 + (size_t)(nn / nh) * n * 64;
    const int r = nn % nh;
    const size_t off = (size_t)g * nh * 16 + (size_t)(r >> 3) * 128 + (size_t)(r & 7) * 16;
    *reinterpret_cast<uint4*>(half + off) = hi;
    *reinterpret_cast<uint4*>(half + (size_t)nh * 64 + off) = lo;
  }
  // biases and (layer 0's blocks) the heads
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a.smalls[a.bias_off[l] + i] = a.b[l][i];
  if (l == 0) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.W; i += gridDim.x * blockDim.x) a.smalls[a.alpha_w_off + i] = a.alpha_w[i];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * (a.W / 2); i += gridDim.x * blockDim.x) a.smalls[a.rgb_w_off + i] = a.rgb_w[i];
    if (blockIdx.x == 0 && threadIdx.x < 3) a.smalls[a.rgb_b_off + threadIdx.x] = a.rgb_b[threadIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 3) a.smalls[a.alpha_b_off] = a.alpha_b[0];
  }
}

// feature_linear folded into views_linears[0] (path_math.cuh, layer program):
//   out_w [H, W + V] = [ Wv[:, :W] * Wf  |  Wv[:, W:] ],   out_b [H] = bv + Wv[:, :W] * bf      (fp64 accumulation)
__global__ void anerf_fold_views_kernel(const float* __restrict__ wv, const float* __restrict__ bv,
                                        const float* __restrict__ wf, const float* __restrict__ bf, int H, int W, int V,
                                        float* __restrict__ out_w, float* __restrict__ out_b) {
  const int cols = W + V;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * (cols + 1); i += gridDim.x * blockDim.x) {
    const int n = i / (cols + 1), c = i % (cols + 1);
    const float* wrow = wv + (size_t)n * cols;
    if (c == cols) {
      double acc = (double)bv[n];
      for (int k = 0; k < W; ++k) acc += (double)wrow[k] * (double)bf[k];
      out_b[n] = (float)acc;
    } else if (c < W) {
      double acc = 0.0;
      for (int k = 0; k < W; ++k) acc += (double)wrow[k] * (double)wf[(size_t)k * W + c];
      out_w[(size_t)n * cols + c] = (float)acc;
    } else {
      out_w[(size_t)n * cols + c] = wrow[c];
    }
  }
}

// view weights of the folded views layer regrouped per joint for the per-ray contraction:
// gw[n][q][j] = fold_w[n][view_weight_col(j, q)] (0 where the joint has no such input); j fastest
__global__ void anerf_pack_view_weights_kernel(const float* __restrict__ fold_w, int cols, NetDims d, float* __restrict__ gw) {
  const int H = d.W / 2, JJ = d.J + 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * JJ * kViewPerJoint; i += gridDim.x * blockDim.x) {
    const int j = i % JJ, q = (i / JJ) % kViewPerJoint, n = i / (kViewPerJoint * JJ);     // [n][q][j]
    const int c = view_weight_col(d, j, q);
    gw[i] = c >= 0 ? fold_w[(size_t)n * cols + c] : 0.f;
  }
}

// framecode mean row (embedding.py:22): codes [n, ch] -> dst[(n)*ch + q] = mean over rows; also copies rows
__global__ void anerf_pack_framecodes_kernel(const float* __restrict__ codes, int n, int ch, float* __restrict__ dst) {
  int q = threadIdx.x;
  if (q >= ch) return;
  float s = 0.f;
  for (int i = 0; i < n; ++i) { float v = codes[i * ch + q]; dst[i * ch + q] = v; s += v; }
  dst[n * ch + q] = s / (float)n;
}

// ------------------------------------------------------------------------------------------------
// self test: D[256,N] = A[256,K] * B[N,K]^T on one CTA pair through exactly the producer / loader / relay /
// MMA / drain code
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(kThreads, 1) anerf_selftest_gemm_kernel(const float* __restrict__ A,
                                                                          const uint8_t* __restrict__ Bpacked,
                                                                          float* __restrict__ Dout, int N, int K,
                                                                          DeviceStatus* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAStages * kAStageBytes + kBStages * kBStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float* zero_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);
  Pipe pp;
  pipe_init(pp, smem, smem + kAStages * kAStageBytes, bars, status);
  if (tid == 0) pipe_init_barriers(pp);
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  for (int i = tid; i < 256; i += kThreads) zero_bias[i] = 0.f;
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  pp.tmem_base = *tmem_slot;
  const int chunks = K / kKC;
  // two back-to-back "layers" into regions 0 and 1 exercise ring wrap-around and both TMEM regions
  if (warp == kMmaWarp) {
    uint32_t a_seq = 0, b_seq = 0;
    if (pp.rank == 0) {
      for (int rep = 0; rep < 2; ++rep) mma_layer<FMT>(pp, a_seq, b_seq, N, chunks, rep);
    } else if (lane == 0) {
      for (int rep = 0; rep < 2; ++rep) relay_layer(pp, b_seq, chunks);
    }
    __syncwarp();
  } else if (warp == kLoadWarp) {
    if (lane == 0) {
      uint32_t b_seq = 0;
      for (int rep = 0; rep < 2; ++rep) load_layer(pp, b_seq, Bpacked, N, chunks);
    }
    __syncwarp();
  } else {
    const int grp = warp >> 2, quarter = warp & 3, row = quarter * 32 + lane;
    const size_t grow = (size_t)pp.rank * kTileM + row;          // row of the 256-row problem
    AProducer<FMT> ap(pp, row);
    uint32_t d_cnt[2] = {0u, 0u};
    for (int rep = 0; rep < 2; ++rep) {
      for (int c = grp; c < chunks; c += kGroups) {
        ap.begin(c);
        for (int t = 0; t < 4; ++t) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = A[grow * K + c * kKC + t * 8 + i];
          ap.store8(t, x);
        }
        ap.end();
      }
      ap.base += chunks;
    }
    for (int rep = 0; rep < 2; ++rep) {
      drain_region<FMT, false, false>(ap, pp, d_cnt, rep, N, zero_bias, 1.0f, quarter, grp, [&](int col0, const float (&x)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) Dout[(size_t)rep * 2 * kTileM * N + grow * N + col0 + i] = x[i];
      });
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc(pp.tmem_base, kTmemCols);
}

#endif  // __CUDACC__
}  // namespace anerf
