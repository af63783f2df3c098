// Launch sequences of the training path (host side).  Per network pass:
//   encodings -> layer-wise forward (activations kept in the workspace) -> compositing backward ->
//   layer-wise backward (weight / bias gradients accumulated atomically, input gradients masked by the
//   ReLU derivative) -> encoding backward into the bone transforms / framecodes.
// Two routes use it: anerf_render_bwd recomputes the forward half per block of rays right before the backward half
// (train_backward); anerf_render_fwd_train runs the forward half of both passes as the step's forward, keeps the
// activations in a state buffer, and anerf_render_bwd_saved runs only the backward half (train_forward /
// train_backward_saved at the end of this file).
// Reference autograd graph: core/raycasters.py:361-474 (render_rays), core/networks/nerf.py:94-205.
//
// The same code drives the CUDA kernels in the library (anerf_api.cu: anerf_render_bwd) and, in the host
// tests only, their CPU emulation (tests/host/simt_emu.h) -- the ANERF_TLAUNCH / ANERF_TZERO macros are the
// only difference.
#pragma once
#include "../../include/anerf_b200.h"
#include "train_kernels.cuh"

#if !defined(ANERF_SIMT_EMU)
#include <map>
#include <tuple>
#include "tc_gemm.cuh"
#endif

#if defined(ANERF_SIMT_EMU)
#define ANERF_TLAUNCH(kernel, grid, block, stream, ...) simt_emu::launch(grid, block, kernel, __VA_ARGS__)
#define ANERF_TZERO(ptr, bytes, stream) memset(ptr, 0, bytes)
typedef void* anerf_tstream;
#else
#define ANERF_TLAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#define ANERF_TZERO(ptr, bytes, stream) cudaMemsetAsync(ptr, 0, bytes, stream)
typedef cudaStream_t anerf_tstream;
#endif

namespace anerf {
namespace train {

constexpr long long kRowsTarget = 262144;   // rows (samples) per block of rays: ~6 GB of fp32 activations + gradients at W=256 (of 180 GB)
constexpr long long kStateRowsTarget = 524288;   // saved-activation training state: one block per pass up to this many rows (2 x 12 GB)

#if !defined(ANERF_SIMT_EMU)
// Tensor-core GEMM engine of the backward pass (tc_gemm.cuh): launch helper + a per-pass cache of packed
// weight operands.  Not part of the emulated build: the host tests run the SIMT kernels, the GPU tests compare
// both engines with the oracle.
__global__ void tc_set_slot_kernel(float* slot, float v) { *slot = v; }

struct TcEngine {
  int n_sm;
  DeviceStatus* status;
  long long* trace;
  int fmt;                  // operand format of the GEMMs: 0 = fp16 hi/lo with per-matrix power-of-two scales, 1 = bf16 hi/lo
  int wgrad_slice_chunks;   // rows per wgrad work item / 32: each item accumulates this many chunks in TMEM, then adds to dW in fp32
  uint8_t* wpack; size_t wpack_bytes, wpack_used;      // packed weight operands of the current network pass
  uint8_t* gpack; size_t gpack_bytes;                   // packed gradient operand of the current wgrad
  std::map<std::tuple<const float*, long long, long long, int, int>, std::pair<const uint8_t*, float*>> cache;
  // amax slots (fp16 format): one device float per operand matrix, folded into by the kernel that produces the matrix.
  // Slots [0, kWeightSlots) belong to the weights of the current pass, the rest to the activation / gradient buffers
  // of the current block of rays (reset per block).
  static constexpr int kWeightSlots = 64, kSlots = 256;
  float* slots;
  int w_used, b_used;
  std::map<const float*, float*> amax_of;              // base pointer of a buffer -> slot of its latest contents
  std::map<std::tuple<const float*, long long>, float*> wmax;   // (weight base pointer, elements) -> slot of max |w| (this pass)
  int error;
  bool dry;                 // replay of the host-side bookkeeping only (slots, packed-operand offsets), no launches: how the
                            // backward of anerf_render_fwd_train finds what its forward left in the state buffer

  float* weight_slot() { if (w_used >= kWeightSlots) { error = 4; return slots; } return slots + w_used++; }
  // a fresh (zeroed) slot for the buffer starting at `base`, replacing whatever was known about it
  float* produce(const float* base) {
    if (fmt != 0) return nullptr;
    if (b_used >= kSlots) { error = 4; return slots + kWeightSlots; }
    float* sl = slots + b_used++;
    amax_of[base] = sl;
    return sl;
  }
  // a buffer whose bound is known analytically (encodings): no reduction needed
  void produce_const(cudaStream_t st, const float* base, float bound) {
    float* sl = produce(base);
    if (sl && !dry) tc_set_slot_kernel<<<1, 1, 0, st>>>(sl, bound);
  }
  // a buffer written by a kernel that does not fold its maximum: one extra read pass
  void produce_scan(cudaStream_t st, const float* base, long long ld, int cols, long long rows) {
    float* sl = produce(base);
    if (sl && !dry) tc_absmax_kernel<<<148 * 4, 256, 0, st>>>(base, ld, 1, (int)rows, cols, sl);
  }
  // slots of every buffer that starts inside [p, p + width) of a row: an operand may be cat[encoding, h]
  AmaxRef ref(const float* p, int width) {
    AmaxRef r{nullptr, nullptr};
    if (fmt != 0) return r;
    for (auto& kv : amax_of)
      if (kv.first >= p && kv.first < p + width) {
        if (!r.p0) r.p0 = kv.second; else if (!r.p1) r.p1 = kv.second; else error = 5;
      }
    if (!r.p0) error = 6;        // an operand nobody registered: the scale would silently be 1
    return r;
  }
  void new_block(cudaStream_t st) {
    if (fmt != 0) return;
    amax_of.clear();
    b_used = kWeightSlots;
    if (!dry) cudaMemsetAsync(slots + kWeightSlots, 0, (kSlots - kWeightSlots) * sizeof(float), st);
  }
  // after a dry replay of the forward: the slots the backward is about to hand out start from zero again
  void reset_unused_slots(cudaStream_t st) {
    if (fmt != 0) return;
    if (w_used < kWeightSlots) cudaMemsetAsync(slots + w_used, 0, (kWeightSlots - w_used) * sizeof(float), st);
    if (b_used < kSlots) cudaMemsetAsync(slots + b_used, 0, (kSlots - b_used) * sizeof(float), st);
  }

  const uint8_t* pack(cudaStream_t st, const float* src, long long s_n, long long s_k, int N, int K, uint8_t* dst,
                      float* rowsum, AmaxRef amax) {
    const int nt = tc_n_tiles(N), NT = tc_tile_width(N), ch = tc_chunks(K);
    const long long total = (long long)nt * ch * 4 * NT;
    long long blocks = (total + 255) / 256;
    const long long cap = rowsum ? 148 * 8 : 148 * 32;   // fewer, longer threads when they also reduce (one atomic each)
    if (blocks > cap) blocks = cap;
    if (rowsum && (nt != 1 || (blocks * 256) % NT != 0)) { error = 3; return nullptr; }
    if (dry) return dst;
    if (fmt == 0) tc_pack_b_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(src, s_n, s_k, N, K, NT, nt, ch, dst, rowsum, amax);
    else tc_pack_b_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(src, s_n, s_k, N, K, NT, nt, ch, dst, rowsum, amax);
    return dst;
  }
  // packed form of a weight matrix (cached per pass) + the slot holding its max |w|
  std::pair<const uint8_t*, float*> packed_weight(cudaStream_t st, const float* src, long long s_n, long long s_k, int N, int K) {
    auto key = std::make_tuple(src, s_n, s_k, N, K);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const size_t need = (tc_packed_bytes(N, K) + 255) & ~(size_t)255;
    if (wpack_used + need > wpack_bytes) { error = 1; return {nullptr, nullptr}; }
    uint8_t* dst = wpack + wpack_used;
    wpack_used += need;
    float* sl = nullptr;
    if (fmt == 0) {
      // max |w| does not depend on the form (W or W^T share it): one reduction per weight buffer and pass.  Sub-blocks
      // of a weight (the skip layer's [encoding | h] column ranges) have other base pointers and get their own.
      auto key2 = std::make_tuple(src, N * (long long)K);
      auto f = wmax.find(key2);
      if (f != wmax.end()) sl = f->second;
      else {
        sl = weight_slot();
        if (!dry) tc_absmax_kernel<<<64, 256, 0, st>>>(src, s_n, s_k, N, K, sl);
        wmax[key2] = sl;
      }
    }
    pack(st, src, s_n, s_k, N, K, dst, nullptr, AmaxRef{sl, nullptr});
    cache[key] = {dst, sl};
    return {dst, sl};
  }
  void new_pass(cudaStream_t st) {
    cache.clear(); wmax.clear(); wpack_used = 0; w_used = 0;
    if (fmt == 0 && !dry) cudaMemsetAsync(slots, 0, kWeightSlots * sizeof(float), st);
  }

  // k_real: real length of the contraction (for the truncation compensation)
  void run(cudaStream_t st, const float* A, long long a_ms, long long a_ks, int M, int K, const uint8_t* Bp, int N,
           float* C, long long c_ms, long long c_ns, const float* bias, int relu, const float* mask, long long mask_ms,
           int mode, int slice_chunks, AmaxRef a_amax = AmaxRef{nullptr, nullptr}, AmaxRef b_amax = AmaxRef{nullptr, nullptr},
           float* c_amax = nullptr) {
    if (!Bp) { error = 1; return; }
    if (dry) return;
    TcGemmArgs g{};
    g.A = A; g.a_ms = a_ms; g.a_ks = a_ks; g.M = M; g.K = K;
    g.a_amax = a_amax; g.b_amax = b_amax; g.c_amax = c_amax;
    g.Bp = Bp; g.N = N; g.NT = tc_tile_width(N); g.n_tiles = tc_n_tiles(N);
    g.chunks_total = tc_chunks(K);
    g.slice_chunks = slice_chunks > 0 ? slice_chunks : g.chunks_total;
    g.k_slices = ceil_div(g.chunks_total, g.slice_chunks);
    // each accumulator sees 3 truncating MMAs per K = 16 slab of its slice (render_kernels.cuh: trunc_comp)
    {
      const int k_acc = g.k_slices > 1 ? g.slice_chunks * kKC : K;
      g.comp = trunc_comp(k_acc < K ? k_acc : K, fmt == 0 ? kTruncKappaFp16 : kTruncKappaBf16);
    }
    g.C = C; g.c_ms = c_ms; g.c_ns = c_ns; g.bias = bias; g.mask = mask; g.mask_ms = mask_ms; g.relu = relu; g.mode = mode;
    g.status = status;
    g.trace = trace;
    { static const int dbg = getenv("ANERF_TC_DEBUG") ? atoi(getenv("ANERF_TC_DEBUG")) : 0; g.debug = dbg; }
    const int items = g.k_slices * ceil_div(M, 2 * kTileM) * g.n_tiles;
    int pairs = n_sm / 2;
    if (pairs > items) pairs = items;
    if (pairs < 1) return;
    const int smem = tc_gemm_smem_bytes();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    {   // function attributes are per device
      static bool attr_set[64] = {};
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaFuncSetAttribute(tc_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(tc_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
      }
    }
    if (fmt == 0) { if (cudaLaunchKernelEx(&cfg, tc_gemm_kernel<0>, g) != cudaSuccess) error = 2; }
    else { if (cudaLaunchKernelEx(&cfg, tc_gemm_kernel<1>, g) != cudaSuccess) error = 2; }
  }
};
#else
struct TcEngine;
#endif

struct TrainCall {
  NetDims dims;
  int n_rays, Sc, Si;
  const anerf_render_opts* opts;
  const anerf_render_inputs* in;
  const float* nearfar;          // [N,2] written by the forward pass
  const float* z_all;            // [N,Sc+Si] sorted depths of the fine pass (forward tap); unused when Si == 0
  const anerf_render_grads* gout;
  const anerf_net_params* net[2];
  const anerf_net_grads* grad[2];
  float* g_skts;                 // [N,J,16] accumulated, or NULL
  float* workspace;
  size_t workspace_floats;
  TcEngine* tc;                  // tensor-core GEMM engine, or NULL: SIMT fp32 GEMMs (debug knob, and the emulated host tests)
  int pass_mask;                 // bit 0: coarse pass, bit 1: fine pass (0 = both): lets the caller exchange the coarse
                                 // network's gradients while the fine pass still runs
};

struct Workspace {
  long long rb;                  // rows per block
  long long z_coarse, xs, h[8], vin, hv, raw, graw, cs, ga, gb, gxs, gvin, ghv, tc_w, tc_g, tc_slots, total;   // offsets in floats
  long long gcat, wcat;             // [rows, 2W] gradients of the skip consumer | of layer 0, and the stacked [2W, P] encoding weights (pose gradient)
  long long wvf, bvf, dwvf, dbvf;   // folded views weights [H, LV] / bias [H] of the pass and their gradients (train_kernels.cuh: fold_views_train_kernel)
  long long tc_w_floats, tc_g_floats;
  int P, LX, LV;                 // encoding width, leading dimension of XS (P + W), of VIN (W + 27J + fc)
};

inline int rays_per_block(int n_rays, int S, long long rows_target = kRowsTarget) {
  long long r = rows_target / S;
  if (r < 1) r = 1;
  if (r > n_rays) r = n_rays;
  return (int)r;
}

inline Workspace make_workspace(const NetDims& d, int n_rays, int Sc, int Si, long long rows_target = kRowsTarget) {
  Workspace w{};
  const int Sf = Sc + Si;
  long long rb = (long long)rays_per_block(n_rays, Sc, rows_target) * Sc;
  if (Si > 0) { long long r1 = (long long)rays_per_block(n_rays, Sf, rows_target) * Sf; if (r1 > rb) rb = r1; }
  w.rb = rb;
  w.P = in_pts_ref(d);
  w.LX = w.P + d.W;
  w.LV = d.W + in_views_ref(d) + d.fc_ch;
  long long off = 0;
  auto take = [&](long long n) { long long o = off; off += (n + 3) / 4 * 4; return o; };
  w.z_coarse = take((long long)n_rays * Sc);
  w.xs = take(rb * w.LX);
  for (int l = 0; l < 8; ++l) w.h[l] = l < d.D ? take(rb * d.W) : 0;
  w.vin = take(rb * w.LV);
  w.hv = take(rb * (d.W / 2));
  w.raw = take(rb * 4);
  w.graw = take(rb * 4);
  w.cs = take(rb * 2);
  w.ga = take(rb * d.W);
  w.gb = take(rb * d.W);
  w.gxs = take(rb * w.LX);
  w.gvin = take(rb * w.LV);
  w.ghv = take(rb * (d.W / 2));
  w.gcat = take(rb * 2 * d.W);
  w.wcat = take((long long)2 * d.W * w.P);
  w.wvf = take((long long)(d.W / 2) * w.LV);
  w.bvf = take(d.W / 2);
  w.dwvf = take((long long)(d.W / 2) * w.LV);
  w.dbvf = take(d.W / 2);
  // tensor-core engine scratch: packed weight operands of one network pass (forward + transposed forms) and the
  // packed gradient operand of one wgrad (bf16 hi + lo = 4 bytes per element, rows padded to 128)
#if !defined(ANERF_SIMT_EMU)
  {
    auto pb = [](int N, int K) { return (long long)((tc_packed_bytes(N, K) + 255) & ~(size_t)255); };
    const int W = d.W, H = d.W / 2;
    long long s = 0;
    for (int l = 0; l < d.D; ++l) s += pb(W, l == 0 ? w.P : ((l - 1) == d.skip ? w.LX : W));      // forward forms
    s += pb(W, W) + pb(H, w.LV);
    s += pb(w.LV, H) + pb(W, H) + pb(W, W);                                                        // transposed (dgrad) forms
    for (int l = 1; l < d.D; ++l) s += pb(W, W) + ((l - 1) == d.skip ? pb(w.P, W) : 0);
    s += pb(w.P, W);
    w.tc_w_floats = (s + 4096) / 4;
    w.tc_g_floats = (long long)round_up((int)rb, 128) * d.W + 1024;
  }
#endif
  w.tc_w = take(w.tc_w_floats);
  w.tc_g = take(w.tc_g_floats);
  w.tc_slots = take(256);
  w.total = off;
  return w;
}

inline size_t train_workspace_bytes(const NetDims& d, int n_rays, int Sc, int Si) {
  return (size_t)make_workspace(d, n_rays, Sc, Si).total * sizeof(float);
}

inline dim3 gemm_grid(int M, int N, int K, int k_chunk) {
  return dim3((unsigned)ceil_div(N, kBN), (unsigned)ceil_div(M, kBM), (unsigned)ceil_div(K, k_chunk));
}

// forward / dgrad form: C[rows, N] = A[rows, K] * B (+ bias, relu, mask)
template <bool BT>
inline void gemm_rows(TcEngine* tc, anerf_tstream st, const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                      long long rows, int N, int K, const float* bias, int relu, const float* mask, long long ldmask, int mode) {
#if !defined(ANERF_SIMT_EMU)
  if (tc) {      // B(n, k): forward W[n*ldb + k], dgrad W[k*ldb + n]
    auto bp = BT ? tc->packed_weight(st, B, ldb, 1, N, K) : tc->packed_weight(st, B, 1, ldb, N, K);
    const AmaxRef a_ref = tc->ref(A, K);                 // looked up BEFORE the output registers its slot (C may alias a row of A's buffer)
    float* c_slot = tc->produce(C);
    tc->run(st, A, lda, 1, (int)rows, K, bp.first, N, C, ldc, 1, bias, relu, mask, ldmask, mode, 0, a_ref, AmaxRef{bp.second, nullptr}, c_slot);
    return;
  }
#endif
  GemmArgs g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = (int)rows; g.N = N; g.K = K; g.k_chunk = round_up(K, kBK);
  g.bias = bias; g.relu = relu; g.mask = mask; g.ldmask = ldmask; g.mode = mode;
  auto k = sgemm_kernel<false, BT>;
  ANERF_TLAUNCH(k, gemm_grid(g.M, N, K, g.k_chunk), dim3(kGemmThreads), st, g);
}

inline void colsum(anerf_tstream st, const float* G, long long ld, int N, long long rows, float* db);

// wgrad form: dW[Nout, Kin] += G[rows, Nout]^T * X[rows, Kin] and db[Nout] += column sums of G (either may be NULL)
inline void gemm_wgrad(TcEngine* tc, anerf_tstream st, const float* G, long long ldg, const float* X, long long ldx, float* dW, long long lddw,
                       long long rows, int Nout, int Kin, float* db) {
#if !defined(ANERF_SIMT_EMU)
  if (tc && dW) {   // dW^T[k', n] += sum_rows X[row, k'] G[row, n]: A = X^T (rows of the MMA = input features), B = G^T;
                    // the pack of G^T reads every gradient once and leaves the bias gradient behind
    if (tc_packed_bytes(Nout, (int)rows) > tc->gpack_bytes) { tc->error = 1; return; }
    const AmaxRef g_ref = tc->ref(G, Nout);
    const uint8_t* bp = tc->pack(st, G, 1, ldg, Nout, (int)rows, tc->gpack, db, g_ref);
    tc->run(st, X, 1, ldx, Kin, (int)rows, bp, Nout, dW, 1, lddw, nullptr, 0, nullptr, 0, 2, tc->wgrad_slice_chunks, tc->ref(X, Kin), g_ref, nullptr);
    return;
  }
#endif
  colsum(st, G, ldg, Nout, rows, db);
  if (!dW) return;
  GemmArgs g{};
  g.A = G; g.lda = ldg; g.B = X; g.ldb = ldx; g.C = dW; g.ldc = lddw;
  g.M = Nout; g.N = Kin; g.K = (int)rows;
  int kc = round_up(ceil_div((int)rows, 64), kBK);
  if (kc < 256) kc = 256;
  g.k_chunk = kc;
  g.mode = 2;
  auto k = sgemm_kernel<true, false>;
  ANERF_TLAUNCH(k, gemm_grid(Nout, Kin, (int)rows, kc), dim3(kGemmThreads), st, g);
}

inline void colsum(anerf_tstream st, const float* G, long long ld, int N, long long rows, float* db) {
  if (!db) return;
  const int rpb = 128;
  auto k = colsum_kernel;
  ANERF_TLAUNCH(k, dim3((unsigned)((rows + rpb - 1) / rpb)), dim3((unsigned)round_up(N, 32)), st, G, ld, N, rows, rpb, db);
}

// Buffers of one network pass inside a workspace (the skip layer's output lives inside XS, next to the encoding)
struct PassView {
  const NetDims& d;
  const Workspace& w;
  float* ws;
  int D, W, H, P, LX, LV, J;
  float *XS, *VIN, *HV, *RAW;
  PassView(const NetDims& dims, const Workspace& wk, float* base)
      : d(dims), w(wk), ws(base), D(dims.D), W(dims.W), H(dims.W / 2), P(wk.P), LX(wk.LX), LV(wk.LV), J(dims.J),
        XS(base + wk.xs), VIN(base + wk.vin), HV(base + wk.hv), RAW(base + wk.raw) {}
  // the last trunk layer writes next to the view encoding (the views layer reads [h | enc] as one matrix: the feature
  // layer is folded into it), the skip layer next to the point encoding
  float* out_ptr(int l) const { return l == d.skip ? XS + P : (l == D - 1 ? VIN : ws + w.h[l]); }
  long long out_ld(int l) const { return l == d.skip ? LX : (l == D - 1 ? LV : W); }
  const float* in_ptr(int l) const { return (l == 0 || (l - 1) == d.skip) ? XS : out_ptr(l - 1); }
  long long in_ld(int l) const { return (l == 0 || (l - 1) == d.skip) ? LX : out_ld(l - 1); }
  int in_k(int l) const { return l == 0 ? P : ((l - 1) == d.skip ? P + W : W); }
};

// Forward half of one network pass over the rays [ray0, ray0 + nb): encodings, layer-wise forward with the
// activations kept in the workspace, heads -> RAW [rows,4].  `dry` (tensor-core engine only): nothing is launched, the
// engine's host-side bookkeeping (amax slots, packed-weight offsets) is replayed for a workspace that already holds
// the results of this very sequence (anerf_render_bwd_saved).
inline void pass_forward(const TrainCall& c, const Workspace& w, int net, int S, const float* z, int ray0, int nb,
                         anerf_tstream st, bool dry) {
  const NetDims& d = c.dims;
  const anerf_net_params& p = *c.net[net];
  const anerf_render_opts& o = *c.opts;
  const PassView v(d, w, c.workspace);
  const int D = v.D, W = v.W, H = v.H, LX = v.LX, LV = v.LV, J = v.J;
  float* XS = v.XS; float* VIN = v.VIN; float* HV = v.HV; float* RAW = v.RAW;
  const long long rows = (long long)nb * S;
  if (dry && !c.tc) return;
  // ---- encodings
  {
    EncodeArgs e{};
    e.rays = c.in->rays; e.skts = c.in->skts; e.z = z; e.cams = c.in->cams; e.codes = p.framecodes; e.pose_idx = c.in->pose_idx; e.n_poses = c.in->n_poses;
    e.ray0 = ray0; e.n_rays_blk = nb; e.S = S; e.J = J; e.W = W; e.fc_ch = d.fc_ch; e.n_fc = d.n_fc; e.vq = view_per_joint(d);
    e.tau_p = o.tau_pts; e.tau_v = o.tau_views;
    for (int j = 0; j < kMaxJoints; ++j) { e.cut_p[j] = o.cutoff_pts[j]; e.cut_v[j] = o.cutoff_views[j]; }
    e.XS = XS; e.ldxs = LX; e.VIN = VIN; e.ldv = LV;
    auto k = encode_rows_kernel;
    if (!dry) ANERF_TLAUNCH(k, dim3((unsigned)((rows * J + 127) / 128)), dim3(128), st, e);
#if !defined(ANERF_SIMT_EMU)
    if (c.tc) {
      // bounds of the encodings: |v w(v)|, |sin|, |cos|, |r|, |d| <= max(1, sup v w(v)); v w(v) peaks near the cutoff:
      // < cutoff + 2 / tau.  A bound within a factor of two of the true maximum costs at most one of the 22 bits.
      float cmax = 0.f;
      for (int j = 0; j < J; ++j) cmax = fmaxf(cmax, fmaxf(o.cutoff_pts[j] + 2.f / fmaxf(o.tau_pts, 1e-3f), 1.f));
      c.tc->produce_const(st, XS, cmax);
      if (d.fc_ch > 0) c.tc->produce_scan(st, VIN + W, LV, LV - W, rows);     // view encodings + framecodes (any magnitude)
      else c.tc->produce_const(st, VIN + W, 1.0f);
    }
#endif
  }
  // ---- forward, activations kept
  for (int l = 0; l < D; ++l)
    gemm_rows<true>(c.tc, st, v.in_ptr(l), v.in_ld(l), p.pts_w[l], v.in_k(l), v.out_ptr(l), v.out_ld(l), rows, W, v.in_k(l), p.pts_b[l], 1, nullptr, 0, 0);
  const float* HL = v.out_ptr(D - 1);
  const long long HLld = v.out_ld(D - 1);
  auto k1 = head_fwd_kernel<1>;
  if (!dry) ANERF_TLAUNCH(k1, dim3((unsigned)((rows + 127) / 128)), dim3(128), st, HL, HLld, W, p.alpha_w, p.alpha_b, rows, RAW + 3, (long long)4);
  // feature layer folded into the views layer: hv = relu([h | enc] WVF^T + BVF)
  float* WVF = c.workspace + w.wvf; float* BVF = c.workspace + w.bvf;
  auto kf = fold_views_train_kernel;
  if (!dry) ANERF_TLAUNCH(kf, dim3((unsigned)((H * LV + H + 127) / 128)), dim3(128), st, p.views_w, p.views_b, p.feature_w, p.feature_b, H, W, LV, WVF, BVF);
  gemm_rows<true>(c.tc, st, VIN, LV, WVF, LV, HV, H, rows, H, LV, BVF, 1, nullptr, 0, 0);
  auto k3 = head_fwd_kernel<3>;
  if (!dry) ANERF_TLAUNCH(k3, dim3((unsigned)((rows + 127) / 128)), dim3(128), st, (const float*)HV, (long long)H, H, p.rgb_w, p.rgb_b, rows, RAW, (long long)4);
}

// Backward half: compositing backward -> heads -> trunk (last layer first) -> encodings, on the activations
// pass_forward left in the workspace.
inline void pass_backward(const TrainCall& c, const Workspace& w, int net, int S, const float* z, const float* noise,
                          const float* g_rgb, const float* g_disp, const float* g_acc, const float* g_alpha,
                          int ray0, int nb, anerf_tstream st) {
  const NetDims& d = c.dims;
  const anerf_net_params& p = *c.net[net];
  static const anerf_net_grads no_grads{};
  const anerf_net_grads& gr = c.grad[net] ? *c.grad[net] : no_grads;
  const anerf_render_opts& o = *c.opts;
  float* ws = c.workspace;
  const PassView v(d, w, ws);
  const int D = v.D, W = v.W, H = v.H, P = v.P, LX = v.LX, LV = v.LV, J = v.J;
  const bool need_pose = c.g_skts != nullptr;
  const bool need_fc = d.fc_ch > 0 && gr.framecodes != nullptr;
  float* XS = v.XS; float* VIN = v.VIN; float* HV = v.HV; float* RAW = v.RAW;
  float* GRAW = ws + w.graw; float* GA = ws + w.ga; float* GB = ws + w.gb; float* GXS = ws + w.gxs;
  float* GVIN = ws + w.gvin; float* GHV = ws + w.ghv;
  const long long rows = (long long)nb * S;
  const float* HL = v.out_ptr(D - 1);
  const long long HLld = v.out_ld(D - 1);
  // ---- compositing backward -> dL/d raw
  {
    CompositeBwdArgs a{};
    a.raw = RAW; a.z = z; a.rays = c.in->rays; a.noise = noise;
    a.g_rgb = g_rgb; a.g_disp = g_disp; a.g_acc = g_acc; a.g_alpha = g_alpha;
    a.ray0 = ray0; a.n_rays_blk = nb; a.S = S; a.softplus = o.softplus; a.B = o.density_scale; a.shift = o.softplus_shift;
    a.scratch = ws + w.cs; a.g_raw = GRAW;
    auto k = composite_bwd_kernel;
    ANERF_TLAUNCH(k, dim3((unsigned)((nb + 63) / 64)), dim3(64), st, a);
  }
  // ---- heads and the views layer
  {
    auto k3 = head_bwd_kernel<3>;
    ANERF_TLAUNCH(k3, dim3((unsigned)((rows + 63) / 64)), dim3((unsigned)(H < 32 ? 32 : H)), st, (const float*)GRAW, (long long)4,
                  (const float*)HV, (long long)H, H, p.rgb_w, rows, 64, 1, GHV, (long long)H, gr.rgb_w, gr.rgb_b);
#if !defined(ANERF_SIMT_EMU)
    if (c.tc) c.tc->produce_scan(st, GHV, H, H, rows);
#endif
  }
  // folded views layer: gradients of [Wv_f Wf | Wv_e] and of the folded bias into scratch, then back to the four
  // parameter tensors in weight space (no row-sized work for the feature layer at all)
  float* WVF = ws + w.wvf; float* DWVF = ws + w.dwvf; float* DBVF = ws + w.dbvf;
  if (gr.views_w || gr.views_b || gr.feature_w || gr.feature_b) {
    ANERF_TZERO(DWVF, (size_t)((DBVF + H) - DWVF) * sizeof(float), st);              // dwvf and dbvf are adjacent in the workspace
    gemm_wgrad(c.tc, st, GHV, H, VIN, LV, DWVF, LV, rows, H, LV, DBVF);
    // d views_w[:, :W] += dWVF[:, :W] Wf   and   d feature_w += Wv_f^T dWVF[:, :W]: two [128 x 256 x 256] products on the
    // fp32 SIMT GEMM (weight-sized), the rest element-wise
    if (gr.views_w) gemm_rows<true>(nullptr, st, DWVF, LV, p.feature_w, W, gr.views_w, LV, H, W, W, nullptr, 0, nullptr, 0, 1);
    if (gr.feature_w) gemm_wgrad(nullptr, st, p.views_w, LV, DWVF, LV, gr.feature_w, W, H, W, W, nullptr);
    auto ku = unfold_views_grads_kernel;
    ANERF_TLAUNCH(ku, dim3((unsigned)((H * LV + W + H + 127) / 128)), dim3(128), st, (const float*)DWVF, (const float*)DBVF, p.views_w,
                  p.feature_b, H, W, LV, gr.views_w, gr.views_b, gr.feature_b);
  }
  {
    auto k1 = head_bwd_kernel<1>;
    ANERF_TLAUNCH(k1, dim3((unsigned)((rows + 63) / 64)), dim3((unsigned)W), st, (const float*)(GRAW + 3), (long long)4, HL, HLld, W,
                  p.alpha_w, rows, 64, 0, GA, (long long)W, gr.alpha_w, gr.alpha_b);
  }
  // (GA now holds g_sigma (x) w_alpha; the GEMM below adds to it and registers the maximum of the sum)
  // dL/dZ of the last trunk layer = (G_hv [Wv_f Wf] + g_sigma (x) w_alpha) . (h > 0)
  gemm_rows<false>(c.tc, st, GHV, H, WVF, LV, GA, W, rows, W, H, nullptr, 0, HL, HLld, 1);
  // the view-encoding (+ framecode) columns only when something consumes them
  if (need_pose || need_fc) gemm_rows<false>(c.tc, st, GHV, H, WVF + W, LV, GVIN + W, LV, rows, LV - W, H, nullptr, 0, nullptr, 0, 0);
  // ---- trunk, last layer first
  // Pose gradient: dL/d(encoding) = dZ_{skip+1} W_{skip+1}[:, :P] + dZ_0 W_0.  When both gradients come out of trunk dgrads
  // they are written side by side (GCAT [rows, 2W]) and the two products run as one GEMM with K = 2W at the end, instead
  // of a store and a read-add-store of the [rows, P] result.
  const bool cat = need_pose && d.skip >= 1 && d.skip + 2 <= D - 1;
  float* GCAT = ws + w.gcat;
  const float* cur = GA;
  long long curld = W;
  for (int l = D - 1; l >= 0; --l) {
    gemm_wgrad(c.tc, st, cur, curld, v.in_ptr(l), v.in_ld(l), gr.pts_w[l], v.in_k(l), rows, W, v.in_k(l), gr.pts_b[l]);
    if (l > 0) {
      if ((l - 1) == d.skip) {        // input = cat[encoding, h]: h part masked, encoding part kept for the pose gradient
        gemm_rows<false>(c.tc, st, cur, curld, p.pts_w[l] + P, v.in_k(l), GXS + P, LX, rows, W, W, nullptr, 0, XS + P, LX, 0);
        if (need_pose && !cat) gemm_rows<false>(c.tc, st, cur, curld, p.pts_w[l], v.in_k(l), GXS, LX, rows, P, W, nullptr, 0, nullptr, 0, 0);
        cur = GXS + P; curld = LX;
      } else {
        float* nxt = (cur == GA) ? GB : GA;
        long long nld = W;
        if (cat && l == d.skip + 2) { nxt = GCAT; nld = 2 * W; }              // dZ of the skip consumer
        else if (cat && l == 1) { nxt = GCAT + W; nld = 2 * W; }              // dZ of layer 0
        gemm_rows<false>(c.tc, st, cur, curld, p.pts_w[l], v.in_k(l), nxt, nld, rows, W, W, nullptr, 0, v.out_ptr(l - 1), v.out_ld(l - 1), 0);
        cur = nxt; curld = nld;
      }
    } else if (need_pose) {
      if (cat) {
        float* WCAT = ws + w.wcat;
        auto ks = stack_enc_weights_kernel;
        ANERF_TLAUNCH(ks, dim3((unsigned)((2 * W * P + 255) / 256)), dim3(256), st, p.pts_w[d.skip + 1], P + W, p.pts_w[0], W, P, WCAT);
        gemm_rows<false>(c.tc, st, GCAT, 2 * W, WCAT, P, GXS, LX, rows, P, 2 * W, nullptr, 0, nullptr, 0, 0);
      } else {
        gemm_rows<false>(c.tc, st, cur, curld, p.pts_w[0], P, GXS, LX, rows, P, W, nullptr, 0, nullptr, 0, d.skip >= 0 ? 1 : 0);
      }
    }
  }
  // ---- encodings backward
  if (need_pose) {
    EncodeBwdArgs e{};
    e.rays = c.in->rays; e.skts = c.in->skts; e.z = z; e.pose_idx = c.in->pose_idx; e.n_poses = c.in->n_poses;
    e.ray0 = ray0; e.n_rays_blk = nb; e.S = S; e.J = J; e.W = W; e.vq = view_per_joint(d);
    e.tau_p = o.tau_pts; e.tau_v = o.tau_views;
    for (int j = 0; j < kMaxJoints; ++j) { e.cut_p[j] = o.cutoff_pts[j]; e.cut_v[j] = o.cutoff_views[j]; }
    e.gXS = GXS; e.ldxs = LX; e.gVIN = GVIN; e.ldv = LV; e.g_skts = c.g_skts;
    auto k = encode_bwd_kernel;
    ANERF_TLAUNCH(k, dim3((unsigned)((nb * J + 63) / 64)), dim3(64), st, e);
  }
  if (need_fc) {
    auto k = framecode_bwd_kernel;
    ANERF_TLAUNCH(k, dim3((unsigned)((nb * d.fc_ch + 127) / 128)), dim3(128), st, (const float*)GVIN, (long long)LV, W + in_views_ref(d),
                  c.in->cams, ray0, nb, S, d.fc_ch, d.n_fc, gr.framecodes);
  }
}

// One network pass (net 0 on the coarse depths, or net 1 on the sorted fine depths) over all rays, block by block:
// forward recomputed, then backward.
inline void backward_pass(const TrainCall& c, const Workspace& w, int net, int S, const float* z, const float* noise,
                          const float* g_rgb, const float* g_disp, const float* g_acc, const float* g_alpha,
                          anerf_tstream st) {
#if !defined(ANERF_SIMT_EMU)
  if (c.tc) c.tc->new_pass(st);
#endif
  const int rpb = rays_per_block(c.n_rays, S);
  for (int ray0 = 0; ray0 < c.n_rays; ray0 += rpb) {
    const int nb = (ray0 + rpb <= c.n_rays) ? rpb : c.n_rays - ray0;
#if !defined(ANERF_SIMT_EMU)
    if (c.tc) c.tc->new_block(st);
#endif
    pass_forward(c, w, net, S, z, ray0, nb, st, false);
    pass_backward(c, w, net, S, z, noise, g_rgb, g_disp, g_acc, g_alpha, ray0, nb, st);
  }
}

// Whole backward of one chunk of rays.  Gradient buffers are accumulated into (the caller zero-fills them).
inline int train_backward(const TrainCall& c, anerf_tstream st) {
  const Workspace w = make_workspace(c.dims, c.n_rays, c.Sc, c.Si);
  if ((size_t)w.total > c.workspace_floats) return -1;
  float* zc = c.workspace + w.z_coarse;
  {
    auto k = coarse_depths_kernel;
    const long long n = (long long)c.n_rays * c.Sc;
    ANERF_TLAUNCH(k, dim3((unsigned)((n + 255) / 256)), dim3(256), st, c.nearfar, c.in->t_rand, c.n_rays, c.Sc, c.opts->lindisp, zc);
  }
  const anerf_render_grads& g = *c.gout;
  const int pm = c.pass_mask ? c.pass_mask : 3;
  if (c.Si > 0) {
    if (pm & 1) backward_pass(c, w, 0, c.Sc, zc, c.in->noise0, g.rgb0, g.disp0, g.acc0, g.alpha0, st);
    if (pm & 2) backward_pass(c, w, 1, c.Sc + c.Si, c.z_all, c.in->noise1, g.rgb_map, g.disp_map, g.acc_map, g.alpha, st);
  } else {
    backward_pass(c, w, 0, c.Sc, zc, c.in->noise0, g.rgb_map, g.disp_map, g.acc_map, g.alpha, st);
  }
  return 0;
}

#if !defined(ANERF_SIMT_EMU)
// ---------------------------------------------------------------------------------------------------------------
// Training forward that KEEPS its activations (anerf_render_fwd_train / anerf_render_bwd_saved): the step's forward is
// the layer-wise chain above instead of the fused render kernel, each pass in a workspace of its own inside one
// "state" buffer, and the backward starts from those activations instead of recomputing them -- three GEMM passes per
// step (forward, dgrad, wgrad) instead of four.  The per-ray stages between the network passes (a12 compositing, a13
// importance sampling + sorted merge) are the fused kernel's own device functions, one warp per ray.
// ---------------------------------------------------------------------------------------------------------------
struct RayStageArgs {
  const float* raw;      // [n_rays * S, 4] (r, g, b, sigma)
  const float* z;        // [n_rays, S]
  const float* rays;     // [n_rays, 8]
  const float* noise;    // [n_rays, S] or NULL
  int n_rays, S, softplus;
  float B, shift;
  float *alpha, *rgb, *disp, *acc;     // [n_rays,S] (or NULL), [n_rays,3], [n_rays], [n_rays]
  int Si, blur;                        // Si > 0: importance sampling from this pass's weights + merge
  const float* u_rand;                 // [n_rays, Si] or NULL (evenly spaced)
  float* z_all;                        // [n_rays, S + Si]
};

inline __host__ __device__ int ray_stage_smem_floats(int S, int Si) { return 4 * S + 3 * S + (Si > 0 ? 2 * (S + Si) : 0); }
constexpr int kRayStageWarps = 4;

__global__ void __launch_bounds__(kRayStageWarps * 32) ray_stage_kernel(const RayStageArgs a) {
  extern __shared__ __align__(16) float ray_stage_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kRayStageWarps + warp;
  if (ray >= a.n_rays) return;
  const int S = a.S, Sf = S + a.Si;
  float* base = ray_stage_smem + (size_t)warp * ray_stage_smem_floats(S, a.Si);
  float4* raw_s = reinterpret_cast<float4*>(base);
  float* z_s = base + 4 * S;
  float* w_s = z_s + S;
  float* cdf_s = w_s + S;
  float* t_s = cdf_s + S;
  float* za_s = t_s + Sf;
  for (int i = lane; i < S; i += 32) {
    raw_s[i] = __ldg(reinterpret_cast<const float4*>(a.raw) + (size_t)ray * S + i);
    z_s[i] = __ldg(a.z + (size_t)ray * S + i);
  }
  const float* rp = a.rays + (size_t)ray * 8;
  const float dnorm = sqrtf(rp[3] * rp[3] + rp[4] * rp[4] + rp[5] * rp[5]);
  __syncwarp();
  struct { float B; int softplus; float shift; } P{a.B, a.softplus, a.shift};
  composite_ray(lane, S, z_s, raw_s, dnorm, a.noise ? a.noise + (size_t)ray * S : nullptr, P, w_s,
                a.alpha ? a.alpha + (size_t)ray * S : nullptr, a.rgb + (size_t)ray * 3, a.disp + ray, a.acc + ray);
  if (a.Si > 0) {
    importance_cdf(lane, S, w_s, cdf_s, a.blur);
    __syncwarp();
    for (int e = lane; e < Sf; e += 32) {
      float v;
      if (e < S) v = z_s[e];
      else {
        const int m = e - S;
        const float u = a.u_rand ? a.u_rand[(size_t)ray * a.Si + m] : linspace01(m, a.Si);
        v = importance_sample(u, S, z_s, cdf_s);
      }
      t_s[e] = v;
    }
    __syncwarp();
    for (int e = lane; e < Sf; e += 32) {          // stable rank sort == torch.sort on the values
      const float x = t_s[e];
      int rk = 0;
      for (int k = 0; k < Sf; ++k) {
        const float y = t_s[k];
        rk += (y < x || (y == x && k < e)) ? 1 : 0;
      }
      za_s[rk] = x;
    }
    __syncwarp();
    for (int e = lane; e < Sf; e += 32) a.z_all[(size_t)ray * Sf + e] = za_s[e];
  }
}

inline int launch_ray_stage(const RayStageArgs& a, cudaStream_t st) {
  const int smem = kRayStageWarps * ray_stage_smem_floats(a.S, a.Si) * (int)sizeof(float);
  if (smem > 200 * 1024) return -1;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(ray_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
  ray_stage_kernel<<<(unsigned)ceil_div(a.n_rays, kRayStageWarps), kRayStageWarps * 32, smem, st>>>(a);
  return 0;
}

// Layout of the state buffer (floats): [workspace of the coarse pass][workspace of the fine pass][nearfar N*2][z_all N*Sf]
struct TrainState {
  Workspace w;
  long long ws1, nearfar, z_all, total;
  bool fits;             // every pass is a single block of rows (the activations of the whole batch stay resident)
};
inline TrainState make_train_state(const NetDims& d, int n_rays, int Sc, int Si) {
  TrainState t{};
  t.w = make_workspace(d, n_rays, Sc, Si, kStateRowsTarget);
  t.fits = rays_per_block(n_rays, Sc, kStateRowsTarget) >= n_rays && (Si == 0 || rays_per_block(n_rays, Sc + Si, kStateRowsTarget) >= n_rays);
  t.ws1 = t.w.total;
  t.nearfar = t.ws1 + (Si > 0 ? t.w.total : 0);
  t.z_all = t.nearfar + ((long long)n_rays * 2 + 3) / 4 * 4;
  t.total = t.z_all + ((long long)n_rays * (Sc + Si) + 3) / 4 * 4;
  return t;
}

struct TrainFwdOut {
  float *rgb_map, *disp_map, *acc_map, *alpha, *rgb0, *disp0, *acc0, *alpha0;
};

// c.workspace = base of the state buffer, c.nearfar = state + nearfar (already filled), tc0 / tc1: engines bound to the
// two workspaces (NULL: SIMT).  Returns 0, or a negative code.
inline int train_forward(TrainCall c, const TrainState& t, TcEngine* tc0, TcEngine* tc1, const TrainFwdOut& out, cudaStream_t st) {
  float* state = c.workspace;
  const anerf_render_opts& o = *c.opts;
  float* zc = state + t.w.z_coarse;
  float* z_all = state + t.z_all;
  {
    const long long n = (long long)c.n_rays * c.Sc;
    coarse_depths_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c.nearfar, c.in->t_rand, c.n_rays, c.Sc, o.lindisp, zc);
  }
  const bool fine = c.Si > 0;
  for (int pass = 0; pass < (fine ? 2 : 1); ++pass) {
    const int S = pass == 0 ? c.Sc : c.Sc + c.Si;
    const float* z = pass == 0 ? zc : z_all;
    c.workspace = state + (pass == 0 ? 0 : t.ws1);
    c.tc = pass == 0 ? tc0 : tc1;
    if (c.tc) { c.tc->new_pass(st); c.tc->new_block(st); }
    pass_forward(c, t.w, pass, S, z, 0, c.n_rays, st, false);
    RayStageArgs a{};
    a.raw = c.workspace + t.w.raw; a.z = z; a.rays = c.in->rays; a.noise = pass == 0 ? c.in->noise0 : c.in->noise1;
    a.n_rays = c.n_rays; a.S = S; a.softplus = o.softplus; a.B = o.density_scale; a.shift = o.softplus_shift;
    const bool last = pass == (fine ? 1 : 0);
    a.alpha = last ? out.alpha : out.alpha0; a.rgb = last ? out.rgb_map : out.rgb0;
    a.disp = last ? out.disp_map : out.disp0; a.acc = last ? out.acc_map : out.acc0;
    a.Si = (fine && pass == 0) ? c.Si : 0; a.blur = o.single_net; a.u_rand = c.in->u_rand; a.z_all = z_all;
    if (launch_ray_stage(a, st) != 0) return -2;
  }
  return 0;
}

// Backward from the state train_forward left behind.  Same engine selection as the forward (the dry replay follows the
// forward's bookkeeping step by step).
inline int train_backward_saved(TrainCall c, const TrainState& t, TcEngine* tc0, TcEngine* tc1, cudaStream_t st) {
  float* state = c.workspace;
  const anerf_render_grads& g = *c.gout;
  const int pm = c.pass_mask ? c.pass_mask : 3;
  const bool fine = c.Si > 0;
  for (int pass = 0; pass < (fine ? 2 : 1); ++pass) {
    if (!(pm & (1 << pass))) continue;
    const int S = pass == 0 ? c.Sc : c.Sc + c.Si;
    const float* z = pass == 0 ? state + t.w.z_coarse : state + t.z_all;
    c.workspace = state + (pass == 0 ? 0 : t.ws1);
    c.tc = pass == 0 ? tc0 : tc1;
    if (c.tc) {
      c.tc->dry = true;
      c.tc->new_pass(st); c.tc->new_block(st);
      pass_forward(c, t.w, pass, S, z, 0, c.n_rays, st, true);
      c.tc->dry = false;
      c.tc->reset_unused_slots(st);
    }
    const bool last = pass == (fine ? 1 : 0);
    pass_backward(c, t.w, pass, S, z, pass == 0 ? c.in->noise0 : c.in->noise1,
                  last ? g.rgb_map : g.rgb0, last ? g.disp_map : g.disp0, last ? g.acc_map : g.acc0, last ? g.alpha : g.alpha0,
                  0, c.n_rays, st);
  }
  return 0;
}
#endif  // !ANERF_SIMT_EMU

}  // namespace train
}  // namespace anerf
