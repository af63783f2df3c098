// TEST INFRASTRUCTURE.  Runs the training backward pass (anerf_b200/csrc/train_path.cuh + train_kernels.cuh)
// on the CPU through the SIMT emulation in simt_emu.h, on inputs written by tests/test_host_train.py, so that
// the kernels' indexing, the buffer carve-up and the gradient math are checked against the oracle's autograd
// in the build container (no GPU there).  The library never contains this code path.
//
//   train_harness <dir>     reads <dir>/meta.txt and <dir>/*.bin (float32), writes <dir>/out_*.bin
#include "simt_emu.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../anerf_b200/csrc/train_path.cuh"

using namespace anerf;

static std::string g_dir;
static std::map<std::string, std::vector<float>> g_store;

static float* load(const std::string& name, size_t expect = 0) {
  std::ifstream f(g_dir + "/" + name + ".bin", std::ios::binary | std::ios::ate);
  if (!f) return nullptr;
  size_t bytes = (size_t)f.tellg();
  f.seekg(0);
  std::vector<float>& v = g_store[name];
  v.resize(bytes / 4);
  f.read(reinterpret_cast<char*>(v.data()), bytes);
  if (expect && v.size() != expect) { fprintf(stderr, "%s: %zu floats, expected %zu\n", name.c_str(), v.size(), expect); exit(2); }
  return v.data();
}
static float* zeros(const std::string& name, size_t n) {
  std::vector<float>& v = g_store[name];
  v.assign(n, 0.f);
  return v.data();
}
static void save(const std::string& name) {
  std::ofstream f(g_dir + "/" + name + ".bin", std::ios::binary);
  const std::vector<float>& v = g_store[name];
  f.write(reinterpret_cast<const char*>(v.data()), v.size() * 4);
}

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  g_dir = argv[1];
  std::map<std::string, double> m;
  {
    std::ifstream f(g_dir + "/meta.txt");
    std::string k;
    double v;
    while (f >> k >> v) m[k] = v;
  }
  NetDims d{};
  d.J = (int)m["J"]; d.D = (int)m["D"]; d.W = (int)m["W"]; d.skip = (int)m["skip"]; d.fc_ch = (int)m["fc_ch"]; d.n_fc = (int)m["n_fc"]; d.fv = m.count("fv") ? (int)m["fv"] : kFv;
  const int N = (int)m["N"], Sc = (int)m["Sc"], Si = (int)m["Si"], Sf = Sc + Si, J = d.J, W = d.W, H = W / 2;
  const int P = in_pts_ref(d), LV = W + in_views_ref(d) + d.fc_ch;
  anerf_render_opts o{};
  o.n_rays = N; o.n_samples = Sc; o.n_importance = Si; o.lindisp = (int)m["lindisp"]; o.softplus = (int)m["softplus"];
  o.density_scale = (float)m["B"]; o.softplus_shift = (float)m["shift"]; o.tau_pts = (float)m["tau_p"]; o.tau_views = (float)m["tau_v"];
  for (int j = 0; j < 24; ++j) { o.cutoff_pts[j] = (float)m["cut"]; o.cutoff_views[j] = (float)m["cut"]; }
  anerf_render_inputs in{};
  in.rays = load("rays", (size_t)N * 8);
  in.skts = load("skts", (size_t)N * J * 16);
  in.cams = load("cams");
  in.t_rand = load("t_rand");
  in.noise0 = load("noise0");
  in.noise1 = load("noise1");
  const float* nearfar = load("nearfar", (size_t)N * 2);
  const float* z_all = Si > 0 ? load("z_all", (size_t)N * Sf) : nullptr;
  anerf_render_grads go{};
  go.rgb_map = load("g_rgb_map"); go.disp_map = load("g_disp_map"); go.acc_map = load("g_acc_map"); go.alpha = load("g_alpha");
  go.rgb0 = load("g_rgb0"); go.disp0 = load("g_disp0"); go.acc0 = load("g_acc0"); go.alpha0 = load("g_alpha0");

  anerf_net_params prm[2]{};
  anerf_net_grads grd[2]{};
  std::vector<std::string> outs;
  const int n_nets = Si > 0 ? 2 : 1;
  for (int n = 0; n < n_nets; ++n) {
    auto nm = [&](const std::string& s) { return "net" + std::to_string(n) + "_" + s; };
    auto both = [&](const std::string& s, size_t cnt, const float*& p, float*& g) {
      p = load(nm(s), cnt);
      if (!p) { fprintf(stderr, "missing %s\n", nm(s).c_str()); exit(2); }
      g = zeros("out_" + nm(s), cnt);
      outs.push_back("out_" + nm(s));
    };
    for (int l = 0; l < d.D; ++l) {
      const int K = l == 0 ? P : ((l - 1) == d.skip ? P + W : W);
      both("pts_w" + std::to_string(l), (size_t)W * K, prm[n].pts_w[l], grd[n].pts_w[l]);
      both("pts_b" + std::to_string(l), (size_t)W, prm[n].pts_b[l], grd[n].pts_b[l]);
    }
    both("alpha_w", W, prm[n].alpha_w, grd[n].alpha_w);
    both("alpha_b", 1, prm[n].alpha_b, grd[n].alpha_b);
    both("feature_w", (size_t)W * W, prm[n].feature_w, grd[n].feature_w);
    both("feature_b", W, prm[n].feature_b, grd[n].feature_b);
    both("views_w", (size_t)H * LV, prm[n].views_w, grd[n].views_w);
    both("views_b", H, prm[n].views_b, grd[n].views_b);
    both("rgb_w", (size_t)3 * H, prm[n].rgb_w, grd[n].rgb_w);
    both("rgb_b", 3, prm[n].rgb_b, grd[n].rgb_b);
    if (d.fc_ch > 0) both("framecodes", (size_t)d.n_fc * d.fc_ch, prm[n].framecodes, grd[n].framecodes);
  }
  float* g_skts = (int)m["need_pose"] ? zeros("out_g_skts", (size_t)N * J * 16) : nullptr;

  train::TrainCall c{};
  c.dims = d; c.n_rays = N; c.Sc = Sc; c.Si = Si; c.opts = &o; c.in = &in; c.nearfar = nearfar; c.z_all = z_all; c.gout = &go;
  c.net[0] = &prm[0]; c.net[1] = &prm[n_nets - 1];
  c.grad[0] = &grd[0]; c.grad[1] = &grd[n_nets - 1];
  c.g_skts = g_skts;
  std::vector<float> ws(train::train_workspace_bytes(d, N, Sc, Si) / 4 + 16, 0.f);
  // poison the workspace: every buffer must be written before it is read
  for (auto& x : ws) x = 1e30f;
  c.workspace = ws.data();
  c.workspace_floats = ws.size();
  if (train::train_backward(c, nullptr) != 0) { fprintf(stderr, "train_backward failed\n"); return 3; }
  for (const auto& s : outs) save(s);
  if (g_skts) save("out_g_skts");
  return 0;
}
