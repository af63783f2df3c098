"""Shared helpers for the test-suite: fixture loading and error metrics."""
import ast
import os

import numpy as np
import torch

from anerf_b200 import synthetic
from oracle import anerf_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RENDER_CASES = ["cfg1_j1_s16_i0", "cfg1_j1_s16_i16", "bench_j24_s64_i128", "surreal_j24_s64_i16_tau200",
                "mixamo_j24_s64_i16_fc", "train_j24_s64_i32_perturb", "single_j24_s64_i48",
                # round 2: the exact flags of configs/surreal/surreal_single.txt, the options no shipped config sets, other shapes
                "single_j24_s96_i48_mv0", "lindisp_j24_s64_i16", "softplus_j24_s64_i16_b2", "fcmean_j24_s64_i16",
                "tau2000_j24_s64_i16", "w128_d6_j17_s32_i16", "w64_d8_j5_s32_i16"]
BIG_CASE = "bench4096_j24_s64_i128"


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = ast.literal_eval(str(z["case"]))
    return case, {k: z[k] for k in z.files if k != "case"}


def bench_frame_scene(n_rays, H=512, W=512, focal=500., n_joints=24):
    """`n_rays` pixels of frame 0 of bench.py: the seeded subset bench.py's parity leg uses (oracle/make_golden.py)."""
    sc = synthetic.make_scene(seed=0, n_rays=None, H=H, W=W, focal=focal, n_joints=n_joints, cam_angle=0.0)
    idx = np.sort(np.random.RandomState(0).choice(H * W, n_rays, replace=False))
    out = {k: (np.ascontiguousarray(v[idx]) if isinstance(v, np.ndarray) and v.shape[:1] == (H * W,) else v)
           for k, v in sc.items()}
    out["pixel_idx"] = idx
    return out


def build_case(c):
    """Regenerates a fixture's inputs from its recorded config (mirrors oracle/make_golden.py:case_inputs)."""
    J = c["n_joints"]
    if c.get("bench_frame"):
        scene = bench_frame_scene(c["n_rays"], c.get("H", 512), c.get("W", 512), c.get("focal", 500.), J)
    else:
        scene = synthetic.make_scene(seed=11, n_rays=c.get("n_rays"), H=c.get("H", 64), W=c.get("W", 64),
                                     focal=c.get("focal", 60.), n_joints=J)
    fc = c.get("framecode_ch", 0)
    mv = c.get("multires_views", 4)
    wk = dict(n_joints=J, D=c["D"], W=c["W_net"], skips=c["skips"], framecode_ch=fc,
              n_framecodes=c.get("n_framecodes", 0), multires_views=mv)
    sd0 = synthetic.make_net_weights(101, **wk)
    sd1 = synthetic.make_net_weights(202, **wk) if c.get("N_importance", 1) > 0 else None
    if c.get("single_net") and sd1 is not None:
        sd1 = sd0                                    # --single_net: the fine pass re-uses the coarse network
    cfg = orc.PathConfig(n_joints=J, D=c["D"], W=c["W_net"], skips=c["skips"], N_samples=c.get("N_samples", 64),
                         N_importance=c.get("N_importance", 0), tau=c.get("tau", 20.), framecode_ch=fc,
                         single_net=bool(c.get("single_net", False)), multires_views=mv,
                         lindisp=bool(c.get("lindisp", False)), density_type=c.get("density_type", "relu"),
                         softplus_shift=c.get("softplus_shift", 0.), density_scale=c.get("density_scale", 1.0))
    N = scene["rays_o"].shape[0]
    if fc > 0:
        scene["cams"] = (np.full(N, -1) if c.get("eval_mean_fc") else np.arange(N) % c["n_framecodes"]).astype(np.int64)
    draws = None
    if c.get("perturb"):
        rng = np.random.RandomState(5)
        draws = dict(t_rand=rng.rand(N, cfg.N_samples).astype(np.float32),
                     u_rand=rng.rand(N, cfg.N_importance).astype(np.float32),
                     noise0=rng.randn(N, cfg.N_samples).astype(np.float32),
                     noise1=rng.randn(N, cfg.N_samples + cfg.N_importance).astype(np.float32))
    return scene, sd0, sd1, cfg, draws


def run_oracle(scene, sd0, sd1, cfg, draws=None, dtype=torch.float32, z_all_override=None):
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)
    d = {k: t(v) for k, v in (draws or {}).items()}
    cams = torch.as_tensor(scene["cams"]) if cfg.framecode_ch > 0 else None
    taps = {}
    with torch.no_grad():
        out = orc.render_rays(orc.to_torch(sd0, dtype), None if sd1 is None else orc.to_torch(sd1, dtype), cfg,
                              t(scene["rays_o"]), t(scene["rays_d"]), t(scene["skts"]), t(scene["cyls"]),
                              cams=cams, t_rand=d.get("t_rand"), u_rand=d.get("u_rand"),
                              noise0=d.get("noise0"), noise1=d.get("noise1"), taps=taps,
                              z_all_override=None if z_all_override is None else t(z_all_override))
    return {k: v.numpy() for k, v in out.items()}, {k: v.numpy() for k, v in taps.items()}


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'rel fp32' metric of BASELINE.json's parity bar)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
