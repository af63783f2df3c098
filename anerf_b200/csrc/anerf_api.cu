// C ABI of libanerf_b200.so (see include/anerf_b200.h).  Host-side logic only: argument checks,
// the layer program / K maps of a network configuration, launches.  No torch, no CPU compute path.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/anerf_b200.h"
#include "render_kernels.cuh"
#include "train_path.cuh"
#include "pose_kernels.cuh"
#include "optim_kernels.cuh"
#include "mesh_kernels.cuh"
#include "sampler_kernels.cuh"

using namespace anerf;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail(ANERF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// pinned status word shared by all launches of the process
DeviceStatus* g_status_host = nullptr;   // mapped pinned memory
DeviceStatus* g_status_dev = nullptr;
std::mutex g_status_mu;

int ensure_status() {
  std::lock_guard<std::mutex> lk(g_status_mu);
  if (g_status_host) return 0;
  CUDA_TRY(cudaHostAlloc((void**)&g_status_host, sizeof(DeviceStatus), cudaHostAllocMapped));
  memset(g_status_host, 0, sizeof(DeviceStatus));
  CUDA_TRY(cudaHostGetDevicePointer((void**)&g_status_dev, g_status_host, 0));
  return 0;
}

long long* g_trace = nullptr;

int check_device_status() {
  if (g_status_host && g_status_host->code != 0) {
    return fail(ANERF_ERR_DEVICE, "device protocol error code=%u site=%u block=%u thread=%u", g_status_host->code,
                g_status_host->where, g_status_host->block, g_status_host->thread);
  }
  return 0;
}

}  // namespace

struct anerf_plan {
  anerf_net_config cfg;
  NetDims dims;
  NetProgram prog;
  int* d_kmap[kMaxLayers];   // device: packed K index -> reference column (-1 = zero)
  int k_in[kMaxLayers];      // reference fan-in of each layer
  float* d_fold_w;           // [W/2, W + 27J + fc] views_linears[0] with feature_linear folded in
  float* d_fold_b;           // [W/2]
  float* d_scale;            // [kMaxLayers] power-of-two operand scale of each layer (scratch of anerf_pack_net)
  int n_sm;
  int max_smem;
};

// every entry point starts from a clean error string (a stale message must never be reported for a later failure)
#define ANERF_ENTRY() g_err.clear()

extern "C" {

const char* anerf_last_error(void) { return g_err.c_str(); }
/* Device status word of the calling process (pinned host memory the kernels write a protocol error into before they
 * trap): 0 when clean, ANERF_ERR_DEVICE + message otherwise.  The asynchronous entry points cannot consult it; callers
 * check it after synchronising the stream (tests, bench). */
int anerf_check_status(void) { ANERF_ENTRY(); return check_device_status(); }
int anerf_version(void) { return 100; }
/* debug: device buffer of 3 x 1024 int64 that the next launches fill with a clock64 timeline of CTA 0 (NULL = off) */
void anerf_debug_set_trace(long long* device_buffer) { g_trace = device_buffer; }

int anerf_plan_create(const anerf_net_config* cfg, anerf_plan** out) {
  ANERF_ENTRY();
  if (!cfg || !out) return fail(ANERF_ERR_INVALID, "null argument");
  if (cfg->n_joints < 1 || cfg->n_joints > kMaxJoints) return fail(ANERF_ERR_INVALID, "n_joints must be 1..24");
  if (cfg->width != 64 && cfg->width != 128 && cfg->width != 256) return fail(ANERF_ERR_INVALID, "width must be 64, 128 or 256");
  if (cfg->depth < 2 || cfg->depth > 8) return fail(ANERF_ERR_INVALID, "depth must be 2..8");
  if (cfg->framecode_ch != 0 && cfg->framecode_ch != 16) return fail(ANERF_ERR_INVALID, "framecode_ch must be 0 or 16");
  if (cfg->framecode_ch > 0 && cfg->n_framecodes < 1) return fail(ANERF_ERR_INVALID, "n_framecodes must be >= 1");
  if (cfg->view_freqs != 0 && cfg->view_freqs != kFv) return fail(ANERF_ERR_INVALID, "view_freqs (multires_views) must be 0 or 4");
  if (cfg->operand_format != 0 && cfg->operand_format != 1) return fail(ANERF_ERR_INVALID, "operand_format must be 0 (fp16) or 1 (bf16)");
  if (cfg->skip >= cfg->depth - 1 && cfg->skip != -1)
    return fail(ANERF_ERR_INVALID, "skip=%d: a skip connection after the last trunk layer is not supported (use -1 when skips >= depth-1)", cfg->skip);
  anerf_plan* p = new anerf_plan();
  p->cfg = *cfg;
  p->dims.J = cfg->n_joints;
  p->dims.D = cfg->depth;
  p->dims.W = cfg->width;
  p->dims.skip = cfg->skip < 0 ? -1 : cfg->skip;
  p->dims.fc_ch = cfg->framecode_ch;
  p->dims.n_fc = cfg->framecode_ch > 0 ? cfg->n_framecodes : 0;
  p->dims.fv = cfg->view_freqs;
  p->prog = make_program(p->dims);
  for (int l = 0; l < kMaxLayers; ++l) p->d_kmap[l] = nullptr;
  p->d_fold_w = p->d_fold_b = p->d_scale = nullptr;
  const NetDims& d = p->dims;
  for (int l = 0; l < p->prog.n_layers; ++l) {
    int kp = p->prog.layer[l].chunks * kKC;
    std::vector<int> km(kp);
    for (int k = 0; k < kp; ++k) km[k] = layer_ref_col(d, l, k);
    if (l == 0) p->k_in[l] = in_pts_ref(d);
    else if (l < d.D) p->k_in[l] = d.W + ((l - 1) == d.skip ? in_pts_ref(d) : 0);
    else p->k_in[l] = d.W + in_views_ref(d) + d.fc_ch;
    for (int k = 0; k < kp; ++k)
      if (km[k] >= p->k_in[l]) { delete p; return fail(ANERF_ERR_INVALID, "internal: kmap out of range"); }
    cudaError_t e = cudaMalloc((void**)&p->d_kmap[l], kp * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_kmap[l], km.data(), kp * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { anerf_plan_destroy(p); return fail(ANERF_ERR_CUDA, "plan upload failed: %s", cudaGetErrorString(e)); }
  }
  {
    const size_t cols = (size_t)d.W + in_views_ref(d) + d.fc_ch;
    cudaError_t e = cudaMalloc((void**)&p->d_fold_w, (size_t)(d.W / 2) * cols * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_fold_b, (size_t)(d.W / 2) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_scale, kMaxLayers * sizeof(float));
    if (e != cudaSuccess) { anerf_plan_destroy(p); return fail(ANERF_ERR_CUDA, "plan alloc failed: %s", cudaGetErrorString(e)); }
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&p->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  *out = p;
  return ANERF_OK;
}

void anerf_plan_destroy(anerf_plan* p) {
  if (!p) return;
  for (int l = 0; l < kMaxLayers; ++l)
    if (p->d_kmap[l]) cudaFree(p->d_kmap[l]);
  if (p->d_fold_w) cudaFree(p->d_fold_w);
  if (p->d_fold_b) cudaFree(p->d_fold_b);
  if (p->d_scale) cudaFree(p->d_scale);
  delete p;
}

size_t anerf_packed_bytes(const anerf_plan* plan) { return plan ? plan->prog.packed_bytes : 0; }

int anerf_pack_net(const anerf_plan* plan, const anerf_net_params* prm, void* packed, void* stream_) {
  ANERF_ENTRY();
  if (!plan || !prm || !packed) return fail(ANERF_ERR_INVALID, "null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const NetProgram& pg = plan->prog;
  const NetDims& d = plan->dims;
  uint8_t* img = (uint8_t*)packed;
  float* smalls = (float*)(img + pg.smalls_off);
  const int fmt = plan->cfg.operand_format;
  if (!prm->feature_w || !prm->feature_b || !prm->views_w || !prm->views_b) return fail(ANERF_ERR_INVALID, "missing feature/views parameters");
  // note: the plan's fold / scale buffers make concurrent anerf_pack_net calls on one plan unsafe across streams
  float kappa = fmt == 0 ? kTruncKappaFp16 : kTruncKappaBf16;
  if (const char* e = getenv("ANERF_TRUNC_KAPPA")) kappa = (float)atof(e);     // calibration knob (tools/parity_dump.py)
  anerf_fold_views_kernel<<<64, 256, 0, stream>>>(prm->views_w, prm->views_b, prm->feature_w, prm->feature_b, d.W / 2, d.W,
                                                  in_views_ref(d) + d.fc_ch, plan->d_fold_w, plan->d_fold_b);
  CUDA_TRY(cudaGetLastError());
  if (!prm->alpha_w || !prm->alpha_b || !prm->rgb_w || !prm->rgb_b) return fail(ANERF_ERR_INVALID, "missing alpha/rgb parameters");
  PackAllArgs a{};
  a.n_layers = pg.n_layers; a.fmt = fmt; a.pure_scale = plan->d_scale; a.img = img; a.smalls = smalls;
  long long most = 0;
  for (int l = 0; l < pg.n_layers; ++l) {
    a.w[l] = l < d.D ? prm->pts_w[l] : plan->d_fold_w;
    a.b[l] = l < d.D ? prm->pts_b[l] : plan->d_fold_b;
    if (!a.w[l] || !a.b[l]) return fail(ANERF_ERR_INVALID, "missing parameter pointer for layer %d", l);
    a.kmap[l] = plan->d_kmap[l]; a.k_in[l] = plan->k_in[l]; a.n[l] = pg.layer[l].n; a.chunks[l] = pg.layer[l].chunks;
    a.w_off[l] = pg.layer[l].w_off; a.bias_off[l] = pg.sm.bias[l];
    // real input columns of the layer as the kernel feeds it (the views layer: h + the row's own ray-slot chunk)
    a.comp[l] = trunc_comp(l < d.D ? plan->k_in[l] : d.W + kKC, kappa);
    const long long total = (long long)a.chunks[l] * a.n[l] * 4;
    if (total > most) most = total;
  }
  a.alpha_w = prm->alpha_w; a.alpha_b = prm->alpha_b; a.rgb_w = prm->rgb_w; a.rgb_b = prm->rgb_b;
  a.alpha_w_off = pg.sm.alpha_w; a.alpha_b_off = pg.sm.alpha_b; a.rgb_w_off = pg.sm.rgb_w; a.rgb_b_off = pg.sm.rgb_b; a.W = d.W;
  anerf_all_scales_kernel<<<pg.n_layers, 1024, 0, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  {
    int bx = (int)((most + 255) / 256);
    if (bx > 148) bx = 148;
    if (bx < 1) bx = 1;
    if (fmt == 1) anerf_pack_all_kernel<1><<<dim3(bx, pg.n_layers), 256, 0, stream>>>(a);
    else anerf_pack_all_kernel<0><<<dim3(bx, pg.n_layers), 256, 0, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
  }
  anerf_pack_view_weights_kernel<<<128, 256, 0, stream>>>(plan->d_fold_w, d.W + in_views_ref(d) + d.fc_ch, d, smalls + pg.sm.gw);
  CUDA_TRY(cudaGetLastError());
  if (d.fc_ch > 0) {
    if (!prm->framecodes) return fail(ANERF_ERR_INVALID, "framecodes pointer missing");
    anerf_pack_framecodes_kernel<<<1, 32, 0, stream>>>(prm->framecodes, d.n_fc, d.fc_ch, smalls + pg.sm.framecodes);
    CUDA_TRY(cudaGetLastError());
  }
  return ANERF_OK;
}

size_t anerf_render_workspace_bytes(int32_t n_rays) { return (size_t)(n_rays > 0 ? n_rays : 1) * 2 * sizeof(float); }

// rays per item: fill the 128-row tiles of both passes as well as possible
static int choose_rays_per_item(int Sc, int Sf, int n_rays) {
  int best = 1;
  double best_u = -1.0;
  for (int R = 1; R <= kMaxRaysPerItem; ++R) {
    int rows = R * (Sf > Sc ? Sf : Sc);
    if (rows > 768 && R > 1) break;
    int tiles = ceil_div(R * Sc, kTileM) + (Sf > Sc ? ceil_div(R * Sf, kTileM) : 0);
    double u = (double)(R * Sc + (Sf > Sc ? R * Sf : 0)) / (double)(tiles * kTileM);
    if (u > best_u + 1e-9) { best_u = u; best = R; }
  }
  (void)n_rays;
  return best;
}

static int launch_fused(const anerf_plan* plan, RenderKParams& P, bool density, cudaStream_t stream) {
  int rc = ensure_status();
  if (rc) return rc;
  P.status = g_status_dev;
  P.trace = g_trace;
  if (P.sl.total > plan->max_smem) return fail(ANERF_ERR_INVALID, "configuration needs %d B of shared memory (> %d)", P.sl.total, plan->max_smem);
  // CTAs run as pairs (cluster of 2, cta_group::2 MMAs): an even grid, at most one CTA per SM
  int grid = P.n_items < plan->n_sm ? P.n_items : plan->n_sm;
  if (grid < 1) return ANERF_OK;
  grid = (grid + 1) & ~1;
  if (grid > (plan->n_sm & ~1)) grid = plan->n_sm & ~1;
  const int fmt = plan->cfg.operand_format;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = P.sl.total;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define ANERF_LAUNCH(FMT, DENS)                                                                              \
  do {                                                                                                       \
    CUDA_TRY(cudaFuncSetAttribute(anerf_fused_kernel<FMT, DENS>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.sl.total)); \
    CUDA_TRY(cudaLaunchKernelEx(&cfg, anerf_fused_kernel<FMT, DENS>, P));                                    \
  } while (0)
  if (fmt == 1) { if (density) ANERF_LAUNCH(1, true); else ANERF_LAUNCH(1, false); }
  else          { if (density) ANERF_LAUNCH(0, true); else ANERF_LAUNCH(0, false); }
#undef ANERF_LAUNCH
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

static void fill_common(const anerf_plan* plan, const anerf_render_opts* o, RenderKParams& P) {
  P.prog = plan->prog;
  P.lindisp = o->lindisp;
  P.softplus = o->softplus;
  P.eval_mean_fc = o->eval_mean_framecode;
  P.blur_is = o->single_net;
  P.B = o->density_scale;
  P.shift = o->softplus_shift;
  P.tau_p = o->tau_pts;
  P.tau_v = o->tau_views;
  for (int j = 0; j < kMaxJoints; ++j) { P.cut_p[j] = o->cutoff_pts[j]; P.cut_v[j] = o->cutoff_views[j]; }
}

// shared by anerf_render_fwd (explicit per-ray inputs) and anerf_render_frame (rays generated per pixel, one pose)
static int render_common(const anerf_plan* plan, const void* packed_coarse, const void* packed_fine, const anerf_render_opts* o,
                         const float* rays, const RayGen& gen, const float* skts, long long skt_stride, const float* cyls,
                         int cyl_stride, const float* cams, float cam_const, const anerf_render_inputs* draws,
                         const anerf_render_outputs* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int N = o->n_rays, Sc = o->n_samples, Si = o->n_importance;
  if (N == 0) return ANERF_OK;
  if (N < 0 || Sc < 4 || Si < 0) return fail(ANERF_ERR_INVALID, "bad sizes n_rays=%d n_samples=%d n_importance=%d", N, Sc, Si);
  if (Sc + Si > 512) return fail(ANERF_ERR_INVALID, "n_samples + n_importance must be <= 512");
  if (Si > 0 && !packed_fine) return fail(ANERF_ERR_INVALID, "packed_fine missing");
  if (!out->rgb_map || !out->disp_map || !out->acc_map) return fail(ANERF_ERR_INVALID, "rgb/disp/acc outputs missing");
  if (!workspace || workspace_bytes < anerf_render_workspace_bytes(N)) return fail(ANERF_ERR_INVALID, "workspace too small");
  if (!(o->density_scale != 0.f)) return fail(ANERF_ERR_INVALID, "density_scale must be non-zero");

  anerf_nearfar_kernel<<<1, 1024, 0, stream>>>(rays, gen, cyls, cyl_stride, N, (float*)workspace);
  CUDA_TRY(cudaGetLastError());

  RenderKParams P{};
  fill_common(plan, o, P);
  P.packed[0] = (const uint8_t*)packed_coarse;
  P.packed[1] = (const uint8_t*)(Si > 0 ? packed_fine : packed_coarse);
  P.n_rays = N; P.Sc = Sc; P.Si = Si; P.Sf = Sc + Si;
  int R = choose_rays_per_item(Sc, Si > 0 ? Sc + Si : Sc, N);
  for (;; --R) {
    P.sl = make_smem_layout(plan->dims, plan->prog.sm.fixed_floats, R, Sc, Sc + Si);
    if (P.sl.total <= plan->max_smem || R == 1) break;
  }
  P.R = R;
  P.slotc = slot_chunks(R);
  P.tilesC = ceil_div(R * Sc, kTileM);
  P.tilesF = Si > 0 ? ceil_div(R * (Sc + Si), kTileM) : 0;
  P.n_items = ceil_div(N, R);
  P.rays = rays; P.gen = gen; P.skts = skts; P.skt_stride = skt_stride; P.cams = cams; P.cam_const = cam_const;
  if (draws) { P.t_rand = draws->t_rand; P.u_rand = draws->u_rand; P.noise0 = draws->noise0; P.noise1 = draws->noise1; P.pose_idx = draws->pose_idx; P.n_poses = draws->n_poses; }
  P.nearfar = (const float*)workspace;
  P.rgb_map = out->rgb_map; P.disp_map = out->disp_map; P.acc_map = out->acc_map; P.alpha = out->alpha;
  P.rgb0 = out->rgb0; P.disp0 = out->disp0; P.acc0 = out->acc0; P.alpha0 = out->alpha0;
  P.z_all_out = out->z_all; P.raw_out = out->raw;
  return launch_fused(plan, P, false, stream);
}

int anerf_render_fwd(const anerf_plan* plan, const void* packed_coarse, const void* packed_fine,
                     const anerf_render_opts* o, const anerf_render_inputs* in, const anerf_render_outputs* out,
                     void* workspace, size_t workspace_bytes, void* stream_) {
  ANERF_ENTRY();
  if (!plan || !packed_coarse || !o || !in || !out) return fail(ANERF_ERR_INVALID, "null argument");
  if (o->n_rays == 0) return ANERF_OK;
  if (!in->rays || !in->skts || !in->cyls) return fail(ANERF_ERR_INVALID, "rays/skts/cyls missing");
  if (plan->dims.fc_ch > 0 && !in->cams && !o->eval_mean_framecode) return fail(ANERF_ERR_INVALID, "cams missing (framecodes enabled)");
  return render_common(plan, packed_coarse, packed_fine, o, in->rays, RayGen{}, in->skts, (long long)plan->dims.J * 16, in->cyls, 5,
                       in->cams, 0.f, in, out, workspace, workspace_bytes, (cudaStream_t)stream_);
}

int anerf_render_frame(const anerf_plan* plan, const void* packed_coarse, const void* packed_fine,
                       const anerf_render_opts* o, const anerf_frame_inputs* fr, const anerf_render_outputs* out,
                       void* workspace, size_t workspace_bytes, void* stream_) {
  ANERF_ENTRY();
  if (!plan || !packed_coarse || !o || !fr || !out) return fail(ANERF_ERR_INVALID, "null argument");
  if (o->n_rays == 0) return ANERF_OK;
  if (!fr->skts || !fr->cyl) return fail(ANERF_ERR_INVALID, "skts/cyl missing");
  if (fr->width <= 0 || fr->height <= 0 || !(fr->focal_x != 0.f) || !(fr->focal_y != 0.f))
    return fail(ANERF_ERR_INVALID, "bad camera (width=%d height=%d)", fr->width, fr->height);
  if (!fr->pixels && (fr->pixel0 < 0 || (long long)fr->pixel0 + o->n_rays > (long long)fr->width * fr->height))
    return fail(ANERF_ERR_INVALID, "pixel range [%d, %d) outside the %d x %d image", fr->pixel0, fr->pixel0 + o->n_rays, fr->width, fr->height);
  RayGen g{};
  for (int i = 0; i < 12; ++i) g.c2w[i] = fr->c2w[i];
  g.fx = fr->focal_x; g.fy = fr->focal_y; g.cx = fr->center_x; g.cy = fr->center_y;
  g.near = fr->near; g.far = fr->far;
  g.W = fr->width; g.pixel0 = fr->pixel0; g.pixels = fr->pixels;
  return render_common(plan, packed_coarse, packed_fine, o, nullptr, g, fr->skts, 0, fr->cyl, 0, nullptr, fr->cam, nullptr, out,
                       workspace, workspace_bytes, (cudaStream_t)stream_);
}

int anerf_density_points(const anerf_plan* plan, const void* packed, const anerf_render_opts* o, const float* pts,
                         const float* skts, int64_t n_points, float* sigma, void* stream_) {
  ANERF_ENTRY();
  if (!plan || !packed || !o || !pts || !skts || !sigma) return fail(ANERF_ERR_INVALID, "null argument");
  if (n_points <= 0) return ANERF_OK;
  if (n_points > (int64_t)1 << 37) return fail(ANERF_ERR_INVALID, "too many points");
  RenderKParams P{};
  fill_common(plan, o, P);
  P.packed[0] = P.packed[1] = (const uint8_t*)packed;
  P.sl = make_smem_layout(plan->dims, plan->prog.sm.fixed_floats, 1, 4, 4);
  P.R = 1; P.Sc = 4; P.Sf = 4; P.slotc = slot_chunks(1);
  P.n_items = (int)((n_points + kTileM - 1) / kTileM);
  P.pts = pts; P.skts = skts; P.sigma = sigma; P.n_points = n_points;
  return launch_fused(plan, P, true, (cudaStream_t)stream_);
}

int anerf_density_grid(const anerf_plan* plan, const void* packed, const anerf_render_opts* o, const float* origin,
                       double radius, int32_t res, int64_t first, int64_t count, const float* skts, float* sigma, void* stream_) {
  ANERF_ENTRY();
  if (!plan || !packed || !o || !origin || !skts || !sigma) return fail(ANERF_ERR_INVALID, "null argument");
  if (res < 1 || res > 2047) return fail(ANERF_ERR_INVALID, "res must be 1..2047");
  const long long total = (long long)(res + 1) * (res + 1) * (res + 1);
  if (first < 0 || count < 0 || first + count > total) return fail(ANERF_ERR_INVALID, "voxel range outside the grid");
  if (count == 0) return ANERF_OK;
  RenderKParams P{};
  fill_common(plan, o, P);
  P.packed[0] = P.packed[1] = (const uint8_t*)packed;
  P.sl = make_smem_layout(plan->dims, plan->prog.sm.fixed_floats, 1, 4, 4);
  P.R = 1; P.Sc = 4; P.Sf = 4; P.slotc = slot_chunks(1);
  P.n_items = (int)((count + kTileM - 1) / kTileM);
  P.pts = nullptr; P.skts = skts; P.sigma = sigma; P.n_points = count;
  P.grid_first = first; P.grid_n1 = res + 1;
  // numpy.linspace(-r, r, res + 1) in fp64: step = (stop - start) / res; y[i] = i * step + start; y[-1] = stop
  P.grid_start = -radius; P.grid_stop = radius;
  P.grid_step = (P.grid_stop - P.grid_start) / (double)res;
  P.grid_origin = origin;
  return launch_fused(plan, P, true, (cudaStream_t)stream_);
}

int anerf_render_fwd_host(const anerf_plan* plan, const void* packed_coarse, const void* packed_fine,
                          const anerf_render_opts* o, const anerf_render_inputs* hin,
                          const anerf_render_outputs* hout, void* stream_) {
  ANERF_ENTRY();
  if (!plan || !o || !hin || !hout) return fail(ANERF_ERR_INVALID, "null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int N = o->n_rays, Sc = o->n_samples, Si = o->n_importance, Sf = Sc + Si, J = plan->dims.J;
  if (N <= 0) return ANERF_OK;
  // one device arena per call (cudaMallocAsync keeps it in the stream's pool, so repeated calls reuse memory)
  struct Seg { const void* h; size_t bytes; size_t off; };
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  auto seg = [&](const void* h, size_t bytes) { Seg s{h, bytes, off}; off += al(bytes); return s; };
  Seg s_rays = seg(hin->rays, (size_t)N * 8 * 4), s_skts = seg(hin->skts, (size_t)N * J * 16 * 4),
      s_cyls = seg(hin->cyls, (size_t)N * 5 * 4), s_cams = seg(hin->cams, hin->cams ? (size_t)N * 4 : 0),
      s_tr = seg(hin->t_rand, hin->t_rand ? (size_t)N * Sc * 4 : 0), s_ur = seg(hin->u_rand, hin->u_rand ? (size_t)N * Si * 4 : 0),
      s_n0 = seg(hin->noise0, hin->noise0 ? (size_t)N * Sc * 4 : 0), s_n1 = seg(hin->noise1, hin->noise1 ? (size_t)N * Sf * 4 : 0);
  Seg o_rgb = seg(hout->rgb_map, (size_t)N * 3 * 4), o_disp = seg(hout->disp_map, (size_t)N * 4), o_acc = seg(hout->acc_map, (size_t)N * 4),
      o_alpha = seg(hout->alpha, hout->alpha ? (size_t)N * (Si > 0 ? Sf : Sc) * 4 : 0), o_rgb0 = seg(hout->rgb0, hout->rgb0 ? (size_t)N * 3 * 4 : 0),
      o_disp0 = seg(hout->disp0, hout->disp0 ? (size_t)N * 4 : 0), o_acc0 = seg(hout->acc0, hout->acc0 ? (size_t)N * 4 : 0),
      o_alpha0 = seg(hout->alpha0, hout->alpha0 ? (size_t)N * Sc * 4 : 0), o_zall = seg(hout->z_all, hout->z_all ? (size_t)N * Sf * 4 : 0),
      o_raw = seg(hout->raw, hout->raw ? (size_t)N * (Si > 0 ? Sf : Sc) * 16 : 0);
  size_t ws_off = off;
  off += al(anerf_render_workspace_bytes(N));
  {   // keep the stream-ordered pool's memory across calls (the default releases it at every synchronisation)
    static bool pool_ready_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool& pool_ready = pool_ready_dev[(dev >= 0 && dev < 64) ? dev : 0];
    if (!pool_ready) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_ready = true;
    }
  }
  uint8_t* arena = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&arena, off, stream));
  auto up = [&](const Seg& s) -> const float* {
    if (!s.h || !s.bytes) return nullptr;
    cudaMemcpyAsync(arena + s.off, s.h, s.bytes, cudaMemcpyHostToDevice, stream);
    return (const float*)(arena + s.off);
  };
  auto dv = [&](const Seg& s) -> float* { return (s.h && s.bytes) ? (float*)(arena + s.off) : nullptr; };
  anerf_render_inputs din{up(s_rays), up(s_skts), up(s_cyls), up(s_cams), up(s_tr), up(s_ur), up(s_n0), up(s_n1)};
  anerf_render_outputs dout{dv(o_rgb), dv(o_disp), dv(o_acc), dv(o_alpha), dv(o_rgb0), dv(o_disp0), dv(o_acc0), dv(o_alpha0), dv(o_zall), dv(o_raw)};
  int rc = anerf_render_fwd(plan, packed_coarse, packed_fine, o, &din, &dout, arena + ws_off, anerf_render_workspace_bytes(N), stream);
  if (rc == ANERF_OK) {
    const Seg* outs[] = {&o_rgb, &o_disp, &o_acc, &o_alpha, &o_rgb0, &o_disp0, &o_acc0, &o_alpha0, &o_zall, &o_raw};
    for (const Seg* s : outs)
      if (s->h && s->bytes) cudaMemcpyAsync((void*)s->h, arena + s->off, s->bytes, cudaMemcpyDeviceToHost, stream);
  }
  cudaFreeAsync(arena, stream);
  cudaError_t e = cudaStreamSynchronize(stream);
  if (rc != ANERF_OK) return rc;
  if (e != cudaSuccess) {
    if (check_device_status() != 0) return ANERF_ERR_DEVICE;      // g_err holds the kernel's protocol error
    return fail(ANERF_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e));
  }
  return check_device_status();
}

namespace {
// per-device state of anerf_render_fwd_host_chunked: three streams, two arenas, events
struct HostPipe {
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  uint8_t* arena[2] = {nullptr, nullptr};
  size_t arena_bytes = 0;
  cudaEvent_t in_done[2] = {}, run_done[2] = {}, out_done[2] = {};
  bool ready = false;
};
HostPipe g_host_pipe[64];
std::mutex g_host_pipe_mu;
}  // namespace

int anerf_render_fwd_host_chunked(const anerf_plan* plan, const void* packed_coarse, const void* packed_fine,
                                  const anerf_render_opts* o, int32_t chunk, const anerf_render_inputs* hin,
                                  const anerf_render_outputs* hout) {
  ANERF_ENTRY();
  if (!plan || !o || !hin || !hout) return fail(ANERF_ERR_INVALID, "null argument");
  const int N = o->n_rays, Sc = o->n_samples, Si = o->n_importance, Sf = Sc + Si, J = plan->dims.J;
  if (N <= 0) return ANERF_OK;
  if (chunk <= 0) return fail(ANERF_ERR_INVALID, "chunk must be positive");
  if (!hin->rays || !hin->skts || !hin->cyls) return fail(ANERF_ERR_INVALID, "rays/skts/cyls missing");
  if (!hout->rgb_map || !hout->disp_map || !hout->acc_map) return fail(ANERF_ERR_INVALID, "rgb/disp/acc outputs missing");
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return fail(ANERF_ERR_INVALID, "device index %d not supported", dev);
  std::lock_guard<std::mutex> lk(g_host_pipe_mu);     // one frame at a time per process: the arenas are shared
  HostPipe& hp = g_host_pipe[dev];
  if (!hp.ready) {
    CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_run, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
      CUDA_TRY(cudaEventCreateWithFlags(&hp.in_done[b], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&hp.run_done[b], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&hp.out_done[b], cudaEventDisableTiming));
    }
    hp.ready = true;
  }
  // arena layout for one chunk of `c` rays (offsets independent of the actual chunk length: sized for `chunk`)
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t C = (size_t)(chunk < N ? chunk : N);
  const int Sa = Si > 0 ? Sf : Sc;
  struct Seg { size_t off; size_t per_ray; const void* h; };
  size_t off = 0;
  auto seg = [&](const void* h, size_t per_ray) { Seg s{off, per_ray, h}; if (h) off += al(C * per_ray); return s; };
  Seg i_rays = seg(hin->rays, 8 * 4), i_skts = seg(hin->skts, (size_t)J * 64), i_cyls = seg(hin->cyls, 5 * 4), i_cams = seg(hin->cams, 4),
      i_tr = seg(hin->t_rand, (size_t)Sc * 4), i_ur = seg(hin->u_rand, (size_t)Si * 4), i_n0 = seg(hin->noise0, (size_t)Sc * 4),
      i_n1 = seg(hin->noise1, (size_t)Sf * 4);
  Seg o_rgb = seg(hout->rgb_map, 12), o_disp = seg(hout->disp_map, 4), o_acc = seg(hout->acc_map, 4), o_alpha = seg(hout->alpha, (size_t)Sa * 4),
      o_rgb0 = seg(hout->rgb0, 12), o_disp0 = seg(hout->disp0, 4), o_acc0 = seg(hout->acc0, 4), o_alpha0 = seg(hout->alpha0, (size_t)Sc * 4),
      o_zall = seg(hout->z_all, (size_t)Sf * 4), o_raw = seg(hout->raw, (size_t)Sa * 16);
  const size_t ws_off = off;
  off += al(anerf_render_workspace_bytes((int)C));
  if (hp.arena_bytes < off) {
    for (int b = 0; b < 2; ++b) {
      if (hp.arena[b]) cudaFree(hp.arena[b]);
      hp.arena[b] = nullptr;
    }
    hp.arena_bytes = 0;
    for (int b = 0; b < 2; ++b) CUDA_TRY(cudaMalloc((void**)&hp.arena[b], off));
    hp.arena_bytes = off;
  }
  const Seg* ins[] = {&i_rays, &i_skts, &i_cyls, &i_cams, &i_tr, &i_ur, &i_n0, &i_n1};
  const Seg* outs[] = {&o_rgb, &o_disp, &o_acc, &o_alpha, &o_rgb0, &o_disp0, &o_acc0, &o_alpha0, &o_zall, &o_raw};
  int rc = ANERF_OK;
  int n_chunks = 0;
  for (long long r0 = 0; r0 < N && rc == ANERF_OK; r0 += chunk, ++n_chunks) {
    const int b = n_chunks & 1;
    const int n = (int)((N - r0) < chunk ? (N - r0) : chunk);
    uint8_t* A = hp.arena[b];
    // arena b is free once the device -> host copies of the chunk that used it two rounds ago are done
    if (n_chunks >= 2) CUDA_TRY(cudaStreamWaitEvent(hp.s_in, hp.out_done[b], 0));
    for (const Seg* s : ins)
      if (s->h) CUDA_TRY(cudaMemcpyAsync(A + s->off, (const uint8_t*)s->h + (size_t)r0 * s->per_ray, (size_t)n * s->per_ray, cudaMemcpyHostToDevice, hp.s_in));
    CUDA_TRY(cudaEventRecord(hp.in_done[b], hp.s_in));
    CUDA_TRY(cudaStreamWaitEvent(hp.s_run, hp.in_done[b], 0));
    if (n_chunks >= 2) CUDA_TRY(cudaStreamWaitEvent(hp.s_run, hp.out_done[b], 0));
    auto dp = [&](const Seg& s) -> float* { return s.h ? (float*)(A + s.off) : nullptr; };
    anerf_render_inputs din{dp(i_rays), dp(i_skts), dp(i_cyls), dp(i_cams), dp(i_tr), dp(i_ur), dp(i_n0), dp(i_n1)};
    anerf_render_outputs dout{dp(o_rgb), dp(o_disp), dp(o_acc), dp(o_alpha), dp(o_rgb0), dp(o_disp0), dp(o_acc0), dp(o_alpha0), dp(o_zall), dp(o_raw)};
    anerf_render_opts oc = *o;
    oc.n_rays = n;
    rc = anerf_render_fwd(plan, packed_coarse, packed_fine, &oc, &din, &dout, A + ws_off, anerf_render_workspace_bytes(n), hp.s_run);
    if (rc != ANERF_OK) break;
    CUDA_TRY(cudaEventRecord(hp.run_done[b], hp.s_run));
    CUDA_TRY(cudaStreamWaitEvent(hp.s_out, hp.run_done[b], 0));
    for (const Seg* s : outs)
      if (s->h) CUDA_TRY(cudaMemcpyAsync((uint8_t*)s->h + (size_t)r0 * s->per_ray, A + s->off, (size_t)n * s->per_ray, cudaMemcpyDeviceToHost, hp.s_out));
    CUDA_TRY(cudaEventRecord(hp.out_done[b], hp.s_out));
  }
  const std::string keep = g_err;                       // anerf_render_fwd's message, if it failed
  cudaError_t e = cudaStreamSynchronize(hp.s_in);
  cudaError_t e2 = cudaStreamSynchronize(hp.s_run);
  cudaError_t e3 = cudaStreamSynchronize(hp.s_out);
  if (rc != ANERF_OK) { g_err = keep; return rc; }
  if (e == cudaSuccess) e = e2;
  if (e == cudaSuccess) e = e3;
  if (e != cudaSuccess) {
    if (check_device_status() != 0) return ANERF_ERR_DEVICE;
    return fail(ANERF_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e));
  }
  return check_device_status();
}

static int fill_chain(pose::ChainArgs& a, int32_t n_poses, int32_t n_joints, const int32_t* parents, int32_t root_id,
                      const float* rots, const float* rest_pose, int32_t n_rest, const float* pelvis) {
  if (n_poses < 0 || n_joints < 1 || n_joints > pose::kMaxPoseJoints) return fail(ANERF_ERR_INVALID, "bad sizes (n_joints must be 1..%d)", pose::kMaxPoseJoints);
  if (!parents || !rots || !rest_pose || !pelvis) return fail(ANERF_ERR_INVALID, "null argument");
  if (root_id < 0 || root_id >= n_joints) return fail(ANERF_ERR_INVALID, "bad root_id");
  if (n_rest != 1 && n_rest != n_poses) return fail(ANERF_ERR_INVALID, "rest_pose must hold 1 or n_poses skeletons");
  for (int j = 0; j < n_joints; ++j) {
    // joints are visited root first, then in index order: every parent must come earlier in that order
    const int q = parents[j];
    if (j == root_id) continue;
    const bool earlier = q == root_id || (q != j && q >= 0 && q < j);
    if (!earlier) return fail(ANERF_ERR_INVALID, "joint %d: parent %d does not precede it (parents must come first, as in the SMPL tree)", j, q);
  }
  a.P = n_poses; a.J = n_joints; a.root = root_id;
  for (int j = 0; j < n_joints; ++j) a.parent[j] = parents[j];
  a.rots = rots; a.rest = rest_pose; a.rest_stride = n_rest == 1 ? 0 : (long long)n_joints * 3; a.pelvis = pelvis;
  return ANERF_OK;
}

int anerf_pose_chain_fwd(int32_t n_poses, int32_t n_joints, const int32_t* parents, int32_t root_id, const float* rots,
                         const float* rest_pose, int32_t n_rest, const float* pelvis, float* l2ws, float* skts, float* kps,
                         void* stream_) {
  ANERF_ENTRY();
  pose::ChainArgs a{};
  int rc = fill_chain(a, n_poses, n_joints, parents, root_id, rots, rest_pose, n_rest, pelvis);
  if (rc) return rc;
  if (!l2ws || !skts) return fail(ANERF_ERR_INVALID, "l2ws/skts outputs missing");
  if (n_poses == 0) return ANERF_OK;
  a.l2ws = l2ws; a.skts = skts; a.kps = kps;
  pose::pose_chain_fwd_kernel<<<(n_poses + 63) / 64, 64, 0, (cudaStream_t)stream_>>>(a);
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

size_t anerf_pose_chain_bwd_scratch_bytes(int32_t n_poses, int32_t n_joints) {
  return (size_t)(n_poses > 0 ? n_poses : 0) * (size_t)(n_joints > 0 ? n_joints : 0) * 12 * sizeof(float);
}

int anerf_pose_chain_bwd(int32_t n_poses, int32_t n_joints, const int32_t* parents, int32_t root_id, const float* rots,
                         const float* rest_pose, int32_t n_rest, const float* pelvis, const float* l2ws, const float* skts,
                         const float* g_skts, const float* g_l2ws, const float* g_kps, float* g_rots, float* g_pelvis,
                         void* scratch, size_t scratch_bytes, void* stream_) {
  ANERF_ENTRY();
  pose::ChainArgs a{};
  int rc = fill_chain(a, n_poses, n_joints, parents, root_id, rots, rest_pose, n_rest, pelvis);
  if (rc) return rc;
  if (!l2ws || !skts || !g_rots || !g_pelvis) return fail(ANERF_ERR_INVALID, "null argument");
  if (!scratch || scratch_bytes < anerf_pose_chain_bwd_scratch_bytes(n_poses, n_joints)) return fail(ANERF_ERR_INVALID, "scratch too small");
  if (n_poses == 0) return ANERF_OK;
  a.l2ws = const_cast<float*>(l2ws); a.skts = const_cast<float*>(skts);
  a.g_skts = g_skts; a.g_l2ws = g_l2ws; a.g_kps = g_kps; a.g_rots = g_rots; a.g_pelvis = g_pelvis; a.scratch = (float*)scratch;
  pose::pose_chain_bwd_kernel<<<(n_poses + 63) / 64, 64, 0, (cudaStream_t)stream_>>>(a);
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

int anerf_adam_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                    float* const* exp_avg_sq, const int64_t* sizes, int64_t step, double lr_d, double beta1_d, double beta2_d, double eps_d,
                    double weight_decay_d, double grad_scale_d, void* stream_) {
  const float lr = (float)lr_d, beta1 = (float)beta1_d, beta2 = (float)beta2_d, eps = (float)eps_d, weight_decay = (float)weight_decay_d, grad_scale = (float)grad_scale_d;
  ANERF_ENTRY();
  if (n_tensors < 0 || (n_tensors > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !sizes))) return fail(ANERF_ERR_INVALID, "null argument");
  if (step < 1) return fail(ANERF_ERR_INVALID, "step must be >= 1 (the count AFTER this update, as torch.optim.Adam keeps it)");
  const float bc1 = (float)(1.0 - pow(beta1_d, (double)step));
  const float bc2s = (float)sqrt(1.0 - pow(beta2_d, (double)step));
  for (int t0 = 0; t0 < n_tensors; t0 += optim::kMaxTensors) {
    optim::AdamArgs a{};
    a.n = n_tensors - t0 < optim::kMaxTensors ? n_tensors - t0 : optim::kMaxTensors;
    long long largest = 0;
    for (int i = 0; i < a.n; ++i) {
      if (!params[t0 + i] || !grads[t0 + i] || !exp_avg[t0 + i] || !exp_avg_sq[t0 + i] || sizes[t0 + i] < 0)
        return fail(ANERF_ERR_INVALID, "tensor %d: null pointer or negative size", t0 + i);
      a.p[i] = params[t0 + i]; a.g[i] = grads[t0 + i]; a.m[i] = exp_avg[t0 + i]; a.v[i] = exp_avg_sq[t0 + i];
      a.size[i] = sizes[t0 + i];
      if (sizes[t0 + i] > largest) largest = sizes[t0 + i];
    }
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.omb1 = (float)(1.0 - beta1_d); a.omb2 = (float)(1.0 - beta2_d);
    a.bc1 = bc1; a.bc2_sqrt = bc2s; a.grad_scale = grad_scale;
    if (largest == 0) continue;
    long long bx = (largest + 256 * 4 - 1) / (256 * 4);          // ~4 elements per thread for the largest tensor
    if (bx > 64) bx = 64;
    optim::adam_step_kernel<<<dim3((unsigned)bx, (unsigned)a.n), 256, 0, (cudaStream_t)stream_>>>(a);
    CUDA_TRY(cudaGetLastError());
  }
  return ANERF_OK;
}

static int mc_fill(mesh::McArgs& a, const float* vol, int32_t n0, int32_t n1, int32_t n2, int64_t s0, int64_t s1, int64_t s2, float iso) {
  if (!vol) return fail(ANERF_ERR_INVALID, "null argument");
  if (n0 < 2 || n1 < 2 || n2 < 2 || n0 > 4096 || n1 > 4096 || n2 > 4096) return fail(ANERF_ERR_INVALID, "volume must be 2..4096 voxels per axis");
  static bool tables[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !tables[dev]) {            // __constant__ memory is per device
    CUDA_TRY(cudaMemcpyToSymbol(mesh::c_tri_count, mesh::kMcTriCount, sizeof(mesh::kMcTriCount)));
    CUDA_TRY(cudaMemcpyToSymbol(mesh::c_tri_table, mesh::kMcTriTable, sizeof(mesh::kMcTriTable)));
    CUDA_TRY(cudaMemcpyToSymbol(mesh::c_edge_corner, mesh::kMcEdgeCorner, sizeof(mesh::kMcEdgeCorner)));
    CUDA_TRY(cudaMemcpyToSymbol(mesh::c_edge_axis, mesh::kMcEdgeAxis, sizeof(mesh::kMcEdgeAxis)));
    if (dev >= 0 && dev < 64) tables[dev] = true;
  }
  a.vol = vol; a.n0 = n0; a.n1 = n1; a.n2 = n2; a.s0 = s0; a.s1 = s1; a.s2 = s2; a.iso = iso;
  a.n_cells = (long long)(n0 - 1) * (n1 - 1) * (n2 - 1);
  return ANERF_OK;
}

int anerf_mc_count(const float* volume, int32_t n0, int32_t n1, int32_t n2, int64_t s0, int64_t s1, int64_t s2, float iso,
                   int32_t* counts, void* stream_) {
  ANERF_ENTRY();
  mesh::McArgs a{};
  int rc = mc_fill(a, volume, n0, n1, n2, s0, s1, s2, iso);
  if (rc) return rc;
  if (!counts) return fail(ANERF_ERR_INVALID, "null argument");
  a.counts = counts;
  mesh::mc_count_kernel<<<(unsigned)((a.n_cells + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(a);
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

int anerf_mc_emit(const float* volume, int32_t n0, int32_t n1, int32_t n2, int64_t s0, int64_t s1, int64_t s2, float iso,
                  const int64_t* offsets, float* verts, int64_t* keys, void* stream_) {
  ANERF_ENTRY();
  mesh::McArgs a{};
  int rc = mc_fill(a, volume, n0, n1, n2, s0, s1, s2, iso);
  if (rc) return rc;
  if (!offsets || !verts || !keys) return fail(ANERF_ERR_INVALID, "null argument");
  a.offsets = (const long long*)offsets; a.verts = verts; a.keys = (long long*)keys;
  mesh::mc_emit_kernel<<<(unsigned)((a.n_cells + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(a);
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

int anerf_sample_rays(const anerf_sampler_inputs* in, const int32_t* frames, int32_t n_images, int32_t rays_per_image,
                      uint64_t seed, const anerf_sampler_outputs* out, int32_t* n_valid, void* stream_) {
  ANERF_ENTRY();
  if (!in || !frames || !out || !n_valid) return fail(ANERF_ERR_INVALID, "null argument");
  if (!in->masks || !in->imgs || !in->c2ws || !in->focals) return fail(ANERF_ERR_INVALID, "masks / imgs / c2ws / focals missing");
  if (!out->rays || !out->target || !out->pixel_idx || !out->frame_of_ray) return fail(ANERF_ERR_INVALID, "rays / target / pixel_idx / frame_of_ray outputs missing");
  if (in->height < 1 || in->width < 1 || (long long)in->height * in->width > (1 << 26)) return fail(ANERF_ERR_INVALID, "bad image size");
  if (n_images < 0 || rays_per_image < 1) return fail(ANERF_ERR_INVALID, "bad sizes");
  if (n_images == 0) return ANERF_OK;
  sampler::SampleArgs a{};
  a.masks = in->masks; a.imgs = in->imgs; a.fgs = in->fgs; a.bgs = in->bgs; a.bg_idx = in->bg_idx;
  a.c2ws = in->c2ws; a.focals = in->focals; a.centers = in->centers;
  a.frames = frames; a.n_img = n_images; a.k = rays_per_image; a.H = in->height; a.W = in->width; a.n_frames = in->n_frames;
  a.seed = seed; a.fg_scale = in->fg_is_255 ? 1.0f / 255.0f : 1.0f; a.mask_img = in->mask_img;
  a.rays = out->rays; a.target = out->target; a.fg_out = out->fg; a.bg_out = out->bg; a.pixel_idx = out->pixel_idx;
  a.frame_of_ray = out->frame_of_ray; a.status = n_valid;
  sampler::sample_rays_kernel<<<n_images, 1024, 0, (cudaStream_t)stream_>>>(a);
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

int anerf_loss_seed(const float* rgb, const float* acc, const float* target, const float* bg, float bg_const, int32_t use_background,
                    int32_t mse, int32_t n_rays, float weight, float* g_rgb, float* g_acc, float* sums, void* stream_) {
  ANERF_ENTRY();
  if (!rgb || !acc || !target || !g_rgb || !g_acc || !sums) return fail(ANERF_ERR_INVALID, "null argument");
  if (n_rays <= 0) return ANERF_OK;
  optim::LossArgs a{};
  a.rgb = rgb; a.acc = acc; a.target = target; a.bg = bg; a.bg_const = bg_const; a.use_bg = use_background; a.mse = mse; a.N = n_rays;
  a.weight = weight; a.g_rgb = g_rgb; a.g_acc = g_acc; a.sums = sums;
  int blocks = (n_rays + 255) / 256;
  if (blocks > 296) blocks = 296;
  optim::loss_seed_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(a);
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

size_t anerf_render_bwd_workspace_bytes(const anerf_plan* plan, int32_t n_rays, int32_t n_samples, int32_t n_importance) {
  if (!plan || n_rays <= 0 || n_samples <= 0 || n_importance < 0) return 0;
  return train::train_workspace_bytes(plan->dims, n_rays, n_samples, n_importance);
}

int anerf_render_bwd(const anerf_plan* plan, const anerf_net_params* coarse, const anerf_net_params* fine,
                     const anerf_render_opts* o, const anerf_render_inputs* in, const float* nearfar, const float* z_all,
                     const anerf_render_grads* gout, const anerf_net_grads* g_coarse, const anerf_net_grads* g_fine,
                     float* g_skts, void* workspace, size_t workspace_bytes, void* stream_) {
  return anerf_render_bwd_pass(plan, coarse, fine, o, in, nearfar, z_all, gout, g_coarse, g_fine, g_skts, workspace, workspace_bytes, 3, stream_);
}

// argument checks shared by the training entry points
static int check_train_args(const anerf_plan* plan, const anerf_net_params* coarse, const anerf_net_params* fine,
                            const anerf_render_opts* o, const anerf_render_inputs* in) {
  if (!plan || !coarse || !o || !in) return fail(ANERF_ERR_INVALID, "null argument");
  const int N = o->n_rays, Sc = o->n_samples, Si = o->n_importance;
  if (N < 0 || Sc < 4 || Si < 0 || Sc + Si > 512) return fail(ANERF_ERR_INVALID, "bad sizes n_rays=%d n_samples=%d n_importance=%d", N, Sc, Si);
  if (Si > 0 && !fine) return fail(ANERF_ERR_INVALID, "fine parameters missing");
  if (N > 0 && (!in->rays || !in->skts)) return fail(ANERF_ERR_INVALID, "rays/skts missing");
  if (plan->dims.fc_ch > 0 && (!in->cams || o->eval_mean_framecode))
    return fail(ANERF_ERR_INVALID, "the training path needs per-ray cams (the eval-time mean framecode has no training path)");
  if (!(o->density_scale != 0.f)) return fail(ANERF_ERR_INVALID, "density_scale must be non-zero");
  const anerf_net_params* nets[2] = {coarse, Si > 0 ? fine : coarse};
  for (int n = 0; n < (Si > 0 ? 2 : 1); ++n) {
    const anerf_net_params* q = nets[n];
    for (int l = 0; l < plan->dims.D; ++l)
      if (!q->pts_w[l] || !q->pts_b[l]) return fail(ANERF_ERR_INVALID, "missing parameter pointer for layer %d", l);
    if (!q->alpha_w || !q->alpha_b || !q->feature_w || !q->feature_b || !q->views_w || !q->views_b || !q->rgb_w || !q->rgb_b)
      return fail(ANERF_ERR_INVALID, "missing head parameters");
    if (plan->dims.fc_ch > 0 && !q->framecodes) return fail(ANERF_ERR_INVALID, "framecodes pointer missing");
  }
  return ANERF_OK;
}

// GEMM engine of the training path bound to one workspace: tensor cores (tc_gemm.cuh) with fp16 hi/lo operands and
// per-matrix scales; ANERF_TRAIN_GEMM=bf16 selects round 1's bf16 hi/lo operands, ANERF_TRAIN_GEMM=simt the fp32 SIMT
// kernels (debug knobs).  Returns false for the SIMT engine (tc untouched).
static bool bind_engine(train::TcEngine& tc, const anerf_plan* plan, float* workspace, const train::Workspace& w) {
  const char* eng = getenv("ANERF_TRAIN_GEMM");
  if (eng && strcmp(eng, "simt") == 0) return false;
  tc.fmt = (eng && strcmp(eng, "bf16") == 0) ? 1 : 0;
  tc.n_sm = plan->n_sm;
  tc.status = g_status_dev;
  tc.wpack = (uint8_t*)(workspace + w.tc_w); tc.wpack_bytes = (size_t)w.tc_w_floats * 4; tc.wpack_used = 0;
  tc.gpack = (uint8_t*)(workspace + w.tc_g); tc.gpack_bytes = (size_t)w.tc_g_floats * 4;
  tc.slots = workspace + w.tc_slots; tc.w_used = 0; tc.b_used = train::TcEngine::kWeightSlots;
  tc.error = 0;
  tc.dry = false;
  tc.trace = g_trace;
  tc.wgrad_slice_chunks = 32;
  if (const char* sl = getenv("ANERF_WGRAD_SLICE")) { int v = atoi(sl); if (v >= 4 && v % 4 == 0 && v <= 256) tc.wgrad_slice_chunks = v; }
  return true;
}

int anerf_render_bwd_pass(const anerf_plan* plan, const anerf_net_params* coarse, const anerf_net_params* fine,
                          const anerf_render_opts* o, const anerf_render_inputs* in, const float* nearfar, const float* z_all,
                          const anerf_render_grads* gout, const anerf_net_grads* g_coarse, const anerf_net_grads* g_fine,
                          float* g_skts, void* workspace, size_t workspace_bytes, int32_t pass_mask, void* stream_) {
  ANERF_ENTRY();
  if (pass_mask < 1 || pass_mask > 3) return fail(ANERF_ERR_INVALID, "pass_mask must be 1 (coarse), 2 (fine) or 3 (both)");
  if (!gout || !nearfar) return fail(ANERF_ERR_INVALID, "null argument");
  int rc = check_train_args(plan, coarse, fine, o, in);
  if (rc) return rc;
  const int N = o->n_rays, Sc = o->n_samples, Si = o->n_importance;
  if (N == 0) return ANERF_OK;
  if (Si > 0 && !z_all) return fail(ANERF_ERR_INVALID, "z_all missing");
  const size_t need = train::train_workspace_bytes(plan->dims, N, Sc, Si);
  if (!workspace || workspace_bytes < need) return fail(ANERF_ERR_INVALID, "workspace too small (%zu < %zu)", workspace_bytes, need);
  train::TrainCall c{};
  c.dims = plan->dims;
  c.n_rays = N; c.Sc = Sc; c.Si = Si;
  c.opts = o; c.in = in; c.nearfar = nearfar; c.z_all = z_all; c.gout = gout;
  c.net[0] = coarse; c.net[1] = Si > 0 ? fine : coarse;
  c.grad[0] = g_coarse; c.grad[1] = g_fine;
  c.g_skts = g_skts;
  c.pass_mask = pass_mask;
  c.workspace = (float*)workspace;
  c.workspace_floats = workspace_bytes / sizeof(float);
  rc = ensure_status();
  if (rc) return rc;
  train::TcEngine tc{};
  if (bind_engine(tc, plan, (float*)workspace, train::make_workspace(plan->dims, N, Sc, Si))) c.tc = &tc;
  if (train::train_backward(c, (cudaStream_t)stream_) != 0) return fail(ANERF_ERR_INVALID, "internal: workspace layout");
  if (tc.error) return fail(ANERF_ERR_INVALID, "internal: tensor-core GEMM engine error %d (scratch size / launch)", tc.error);
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

size_t anerf_train_state_bytes(const anerf_plan* plan, int32_t n_rays, int32_t n_samples, int32_t n_importance) {
  if (!plan || n_rays <= 0 || n_samples <= 0 || n_importance < 0) return 0;
  const train::TrainState t = train::make_train_state(plan->dims, n_rays, n_samples, n_importance);
  return t.fits ? (size_t)t.total * sizeof(float) : 0;
}

int anerf_render_fwd_train(const anerf_plan* plan, const anerf_net_params* coarse, const anerf_net_params* fine,
                           const anerf_render_opts* o, const anerf_render_inputs* in, const anerf_render_outputs* out,
                           float* nearfar_out, void* state, size_t state_bytes, void* stream_) {
  ANERF_ENTRY();
  if (!out) return fail(ANERF_ERR_INVALID, "null argument");
  int rc = check_train_args(plan, coarse, fine, o, in);
  if (rc) return rc;
  const int N = o->n_rays, Sc = o->n_samples, Si = o->n_importance;
  if (N == 0) return ANERF_OK;
  if (!in->cyls) return fail(ANERF_ERR_INVALID, "cyls missing");
  if (!out->rgb_map || !out->disp_map || !out->acc_map) return fail(ANERF_ERR_INVALID, "rgb/disp/acc outputs missing");
  if (Si > 0 && (!out->rgb0 || !out->disp0 || !out->acc0)) return fail(ANERF_ERR_INVALID, "rgb0/disp0/acc0 outputs missing");
  const train::TrainState t = train::make_train_state(plan->dims, N, Sc, Si);
  if (!t.fits) return fail(ANERF_ERR_INVALID, "batch too large to keep its activations (anerf_train_state_bytes == 0): use anerf_render_fwd + anerf_render_bwd");
  if (!state || state_bytes < (size_t)t.total * sizeof(float)) return fail(ANERF_ERR_INVALID, "state too small (%zu < %zu)", state_bytes, (size_t)t.total * sizeof(float));
  rc = ensure_status();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream_;
  float* base = (float*)state;
  anerf_nearfar_kernel<<<1, 1024, 0, st>>>(in->rays, RayGen{}, in->cyls, 5, N, base + t.nearfar);
  CUDA_TRY(cudaGetLastError());
  train::TrainCall c{};
  c.dims = plan->dims;
  c.n_rays = N; c.Sc = Sc; c.Si = Si;
  c.opts = o; c.in = in; c.nearfar = base + t.nearfar;
  c.net[0] = coarse; c.net[1] = Si > 0 ? fine : coarse;
  c.workspace = base;
  c.workspace_floats = state_bytes / sizeof(float);
  train::TcEngine tc0{}, tc1{};
  const bool use_tc = bind_engine(tc0, plan, base, t.w);
  if (use_tc && Si > 0) bind_engine(tc1, plan, base + t.ws1, t.w);
  train::TrainFwdOut fo{out->rgb_map, out->disp_map, out->acc_map, out->alpha, out->rgb0, out->disp0, out->acc0, out->alpha0};
  if (train::train_forward(c, t, use_tc ? &tc0 : nullptr, use_tc ? &tc1 : nullptr, fo, st) != 0)
    return fail(ANERF_ERR_INVALID, "internal: per-ray stage launch");
  if (tc0.error || tc1.error) return fail(ANERF_ERR_INVALID, "internal: tensor-core GEMM engine error %d (scratch size / launch)", tc0.error ? tc0.error : tc1.error);
  if (nearfar_out) CUDA_TRY(cudaMemcpyAsync(nearfar_out, base + t.nearfar, (size_t)N * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (out->z_all && Si > 0) CUDA_TRY(cudaMemcpyAsync(out->z_all, base + t.z_all, (size_t)N * (Sc + Si) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

int anerf_render_bwd_saved(const anerf_plan* plan, const anerf_net_params* coarse, const anerf_net_params* fine,
                           const anerf_render_opts* o, const anerf_render_inputs* in, const anerf_render_grads* gout,
                           const anerf_net_grads* g_coarse, const anerf_net_grads* g_fine, float* g_skts, void* state,
                           size_t state_bytes, int32_t pass_mask, void* stream_) {
  ANERF_ENTRY();
  if (pass_mask < 1 || pass_mask > 3) return fail(ANERF_ERR_INVALID, "pass_mask must be 1 (coarse), 2 (fine) or 3 (both)");
  if (!gout) return fail(ANERF_ERR_INVALID, "null argument");
  int rc = check_train_args(plan, coarse, fine, o, in);
  if (rc) return rc;
  const int N = o->n_rays, Sc = o->n_samples, Si = o->n_importance;
  if (N == 0) return ANERF_OK;
  const train::TrainState t = train::make_train_state(plan->dims, N, Sc, Si);
  if (!t.fits) return fail(ANERF_ERR_INVALID, "batch too large for a saved-activation state");
  if (!state || state_bytes < (size_t)t.total * sizeof(float)) return fail(ANERF_ERR_INVALID, "state too small (%zu < %zu)", state_bytes, (size_t)t.total * sizeof(float));
  rc = ensure_status();
  if (rc) return rc;
  float* base = (float*)state;
  train::TrainCall c{};
  c.dims = plan->dims;
  c.n_rays = N; c.Sc = Sc; c.Si = Si;
  c.opts = o; c.in = in; c.nearfar = base + t.nearfar; c.z_all = base + t.z_all; c.gout = gout;
  c.net[0] = coarse; c.net[1] = Si > 0 ? fine : coarse;
  c.grad[0] = g_coarse; c.grad[1] = g_fine;
  c.g_skts = g_skts;
  c.pass_mask = pass_mask;
  c.workspace = base;
  c.workspace_floats = state_bytes / sizeof(float);
  train::TcEngine tc0{}, tc1{};
  const bool use_tc = bind_engine(tc0, plan, base, t.w);
  if (use_tc && Si > 0) bind_engine(tc1, plan, base + t.ws1, t.w);
  train::train_backward_saved(c, t, use_tc ? &tc0 : nullptr, use_tc ? &tc1 : nullptr, (cudaStream_t)stream_);
  if (tc0.error || tc1.error) return fail(ANERF_ERR_INVALID, "internal: tensor-core GEMM engine error %d (scratch size / launch)", tc0.error ? tc0.error : tc1.error);
  CUDA_TRY(cudaGetLastError());
  return ANERF_OK;
}

int anerf_selftest_tc_gemm(const float* A, int64_t a_ms, int64_t a_ks, int32_t M, int32_t K, const float* B, int64_t b_ns,
                           int64_t b_ks, int32_t N, float* C, int64_t c_ms, int64_t c_ns, const float* bias,
                           const float* mask, int64_t mask_ms, int32_t relu, int32_t mode, int32_t slice_chunks,
                           void* stream_) {
  ANERF_ENTRY();
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return fail(ANERF_ERR_INVALID, "bad argument");
  if (slice_chunks < 0 || slice_chunks % 4 != 0) return fail(ANERF_ERR_INVALID, "slice_chunks must be a multiple of 4");
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = ensure_status();
  if (rc) return rc;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  uint8_t* d_pack = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d_pack, tc_packed_bytes(N, K)));
  train::TcEngine tc{};
  tc.n_sm = n_sm; tc.status = g_status_dev; tc.error = 0; tc.trace = g_trace;
  // self test of the GEMM alone: fp16 operands with scales from explicit reductions over A and B
  const char* eng = getenv("ANERF_TRAIN_GEMM");
  tc.fmt = (eng && strcmp(eng, "bf16") == 0) ? 1 : 0;
  float* d_slots = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d_slots, 4 * sizeof(float)));
  CUDA_TRY(cudaMemsetAsync(d_slots, 0, 4 * sizeof(float), stream));
  AmaxRef ra{nullptr, nullptr}, rb{nullptr, nullptr};
  if (tc.fmt == 0) {
    tc_absmax_kernel<<<148, 256, 0, stream>>>(A, a_ms, a_ks, M, K, d_slots);
    tc_absmax_kernel<<<148, 256, 0, stream>>>(B, b_ns, b_ks, N, K, d_slots + 1);
    ra.p0 = d_slots; rb.p0 = d_slots + 1;
  }
  tc.pack(stream, B, b_ns, b_ks, N, K, d_pack, nullptr, rb);
  tc.run(stream, A, a_ms, a_ks, M, K, d_pack, N, C, c_ms, c_ns, bias, relu, mask, mask_ms, mode, slice_chunks, ra, rb, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(d_pack);
  cudaFree(d_slots);
  if (tc.error) return fail(ANERF_ERR_INVALID, "tensor-core GEMM launch failed (%d)", tc.error);
  if (e != cudaSuccess) { check_device_status(); return fail(ANERF_ERR_CUDA, "tc gemm failed: %s [%s]", cudaGetErrorString(e), g_err.c_str()); }
  return check_device_status();
}

int anerf_selftest_gemm(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t format, void* stream_) {
  ANERF_ENTRY();
  if (!A || !B || !D) return fail(ANERF_ERR_INVALID, "null argument");
  if (K <= 0 || K % (kGroups * kKC) != 0) return fail(ANERF_ERR_INVALID, "K must be a positive multiple of 128");
  if (N != 64 && N != 128 && N != 256) return fail(ANERF_ERR_INVALID, "N must be 64, 128 or 256");
  if (format != 0 && format != 1) return fail(ANERF_ERR_INVALID, "format must be 0 (fp16) or 1 (bf16)");
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = ensure_status();
  if (rc) return rc;
  const int chunks = K / kKC;
  std::vector<int> km(K);
  for (int k = 0; k < K; ++k) km[k] = k;
  int* d_km = nullptr;
  uint8_t* d_pack = nullptr;
  float* d_one = nullptr;
  const float one = 1.0f;
  CUDA_TRY(cudaMalloc((void**)&d_one, sizeof(float)));
  CUDA_TRY(cudaMemcpy(d_one, &one, sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc((void**)&d_km, K * sizeof(int)));
  CUDA_TRY(cudaMalloc((void**)&d_pack, (size_t)chunks * N * 128));
  CUDA_TRY(cudaMemcpy(d_km, km.data(), K * sizeof(int), cudaMemcpyHostToDevice));
  int blocks = (chunks * N * 4 + 255) / 256;
  int smem = kAStages * kAStageBytes + kBStages * kBStageBytes + 8 * kNumBars + 16 + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define ANERF_ST(FMT)                                                                                       \
  do {                                                                                                      \
    anerf_pack_layer_kernel<FMT><<<blocks, 256, 0, stream>>>(B, K, d_km, N, chunks, d_one, d_pack);          \
    cudaFuncSetAttribute(anerf_selftest_gemm_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    cudaLaunchKernelEx(&cfg, anerf_selftest_gemm_kernel<FMT>, A, (const uint8_t*)d_pack, D, (int)N, (int)K, g_status_dev); \
  } while (0)
  if (format == 0) ANERF_ST(0); else ANERF_ST(1);
#undef ANERF_ST
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(d_km);
  cudaFree(d_pack);
  cudaFree(d_one);
  if (e != cudaSuccess) { check_device_status(); return fail(ANERF_ERR_CUDA, "selftest failed: %s [%s]", cudaGetErrorString(e), g_err.c_str()); }
  return check_device_status();
}

}  // extern "C"
