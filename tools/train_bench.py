"""Training-step timing on one GPU: forward (fused kernel) + backward (anerf_render_bwd) + Adam step through the
Python boundary, next to the oracle port of the reference's PyTorch path (fp32 eager autograd) on the same GPU.

    python tools/train_bench.py [n_rays] [N_importance] [pose_grad 0/1] [steps]

Defaults: 3072 rays, 16 importance samples (configs/mixamo/mixamo.txt: N_rand 3072, N_samples 64, N_importance 16),
pose gradient on (pose refinement), 5 timed steps.  Prints one JSON line."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anerf_b200 import synthetic  # noqa: E402
from anerf_b200.raycasters import create_raycaster  # noqa: E402
import bench  # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
Si = int(sys.argv[2]) if len(sys.argv) > 2 else 16
pose = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
with_ref = not (len(sys.argv) > 5 and sys.argv[5] == "noref")
Sc = 64
dev = torch.device("cuda")

import collections
Skel = collections.namedtuple("Skel", ["joint_names", "joint_trees", "root_id"])
data_attrs = dict(skel_type=Skel(synthetic.SMPL_JOINT_NAMES, synthetic.SMPL_PARENTS, 0), near=0., far=1., n_views=1,
                  joint_coords=np.tile(np.eye(3, dtype=np.float32), (1, 24, 1, 1)))
args = bench.make_args(N_importance=Si, perturb=1.0)
with contextlib.redirect_stdout(io.StringIO()):
    rk_train, rk_test, _, grad_vars, optimizer, _ = create_raycaster(args, data_attrs, device=dev)
rc = rk_test["ray_caster"]
sd0, sd1 = synthetic.make_net_weights(101), synthetic.make_net_weights(202)
rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in sd0.items()})
rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in sd1.items()})
holder = rk_train["ray_caster"].train()
sc = synthetic.make_scene(seed=0, n_rays=n_rays, H=512, W=512, focal=500., n_joints=24)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
N = sc["rays_o"].shape[0]
rays = torch.cat([t(sc["rays_o"]), t(sc["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev),
                  torch.nn.functional.normalize(t(sc["rays_d"]), dim=-1)], 1)
skts0 = t(sc["skts"])
kw = {k: v for k, v in rk_train.items() if k not in ("ray_caster", "use_viewdirs")}
target = torch.rand(N, 3, device=dev)


def step():
    optimizer.zero_grad(set_to_none=True)
    skts = skts0.clone().requires_grad_(True) if pose else skts0
    out = holder(rays, kp_batch=t(sc["kps"]), skts=skts, cyls=t(sc["cyls"]), bones=t(sc["bones"]), cams=None, subject_idxs=None, **kw)
    loss = ((out["rgb_map"] - target) ** 2).mean() + ((out["rgb0"] - target) ** 2).mean()
    loss.backward()
    optimizer.step()
    return loss


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timed(step, steps)

# gradient agreement of the two GEMM engines on this batch (tensor cores, bf16 hi/lo split, vs fp32 SIMT kernels)
def grads_with(engine):
    os.environ["ANERF_TRAIN_GEMM"] = engine
    for p in grad_vars:
        p.grad = None
    out = holder(rays, kp_batch=t(sc["kps"]), skts=skts0, cyls=t(sc["cyls"]), bones=t(sc["bones"]), cams=None, subject_idxs=None,
                 **dict(kw, perturb=0., raw_noise_std=0.))
    (((out["rgb_map"] - target) ** 2).mean() + ((out["rgb0"] - target) ** 2).mean()).backward()
    return [p.grad.double().clone() for p in grad_vars]
g_tc, g_simt = grads_with("tc"), grads_with("simt")
os.environ["ANERF_TRAIN_GEMM"] = "tc"
names = [n for n, p in rc.named_parameters() if any(p is q for q in grad_vars)]
per = {n: float((a - b).norm() / b.norm().clamp_min(1e-30)) for n, a, b in zip(names, g_tc, g_simt)}
worst = max(per, key=per.get)
flat_tc, flat_simt = torch.cat([a.reshape(-1) for a in g_tc]), torch.cat([b.reshape(-1) for b in g_simt])
engine_err = {"l2_rel_all_gradients": float((flat_tc - flat_simt).norm() / flat_simt.norm()),
              "cosine": float(torch.dot(flat_tc, flat_simt) / (flat_tc.norm() * flat_simt.norm())),
              "l2_rel_worst_tensor": per[worst], "worst_tensor": f"{worst} ({g_simt[names.index(worst)].numel()} elements)",
              "l2_rel_weight_matrices_worst": max(v for n, v in per.items() if n.endswith("weight") and "alpha" not in n)}

# forward only / backward only split (CUDA events around the two halves)
def fwd_only():
    with torch.no_grad():
        holder(rays, kp_batch=t(sc["kps"]), skts=skts0, cyls=t(sc["cyls"]), bones=t(sc["bones"]), cams=None, subject_idxs=None, **kw)
ms_fwd = timed(fwd_only, steps)

# the reference's PyTorch path (oracle port) with autograd on the same GPU
ref_ms = None
try:
    if not with_ref:
        raise RuntimeError("skipped (noref)")
    from oracle import anerf_oracle as orc
    torch.backends.cuda.matmul.allow_tf32 = False
    p0 = {k: t(v).requires_grad_(True) for k, v in sd0.items()}
    p1 = {k: t(v).requires_grad_(True) for k, v in sd1.items()}
    cfg = orc.PathConfig(N_importance=Si)
    opt_ref = torch.optim.Adam(list(p0.values()) + list(p1.values()), lr=5e-4)

    def ref_step():
        opt_ref.zero_grad(set_to_none=True)
        skts = skts0.clone().requires_grad_(True) if pose else skts0
        out = orc.render_rays(p0, p1, cfg, rays[:, 0:3], rays[:, 3:6], skts, t(sc["cyls"]),
                              t_rand=torch.rand(N, Sc, device=dev), u_rand=torch.rand(N, Si, device=dev), training=True)
        loss = ((out["rgb_map"] - target) ** 2).mean() + ((out["rgb0"] - target) ** 2).mean()
        loss.backward()
        opt_ref.step()
    ref_ms = timed(ref_step, max(2, steps // 2))
except Exception as e:  # noqa: BLE001
    ref_ms = repr(e)[:200]

rows = N * (Sc + Sc + Si)
flop = 3 * rows * 1723648            # forward + dgrad + wgrad, MLP only
print(json.dumps({"n_rays": N, "N_samples": Sc, "N_importance": Si, "pose_grad": pose, "ms_per_step": ms, "ms_forward_only": ms_fwd,
                  "rays_per_s": N / (ms * 1e-3), "algorithmic_tflops": flop / (ms * 1e-3) / 1e12,
                  "tc_vs_simt_grad_err": engine_err,
                  "reference_port_ms_per_step": ref_ms,
                  "speedup_vs_reference_port": (ref_ms / ms) if isinstance(ref_ms, float) else None}))
