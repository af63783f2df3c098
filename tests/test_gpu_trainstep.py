"""SURVEY.md 8(f) row 3 on the GPU: `anerf_b200.train.FusedTrainStep` (forward -> loss seed -> backward per pass ->
FusedAdam as a fixed launch sequence) against the autograd route the reference's trainer takes through the boundary
(RayCaster in .train() mode + the reference's loss arithmetic in torch + loss.backward() + optimizer.step())."""
import contextlib
import io

import numpy as np
import pytest
import torch

from anerf_b200 import synthetic
from anerf_b200.pose_opt import PoseOptLayer
from anerf_b200.raycasters import create_raycaster
from anerf_b200.train import FusedTrainStep
from tests.test_gpu_api import data_attrs, make_args

pytestmark = pytest.mark.gpu


def _setup(dev, **over):
    with contextlib.redirect_stdout(io.StringIO()):
        rk_train, rk_test, _, grad_vars, optimizer, _ = create_raycaster(make_args(N_importance=16, N_samples=32, no_reload=True, **over), data_attrs(24, n_views=3))
    rc = rk_test["ray_caster"]
    wk = dict(framecode_ch=16, n_framecodes=3) if over.get("opt_framecode") else {}
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(101, **wk).items()})
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202, **wk).items()})
    return rk_train, rc, grad_vars, optimizer


def _batch(dev, N=160):
    sc = synthetic.make_scene(seed=4, n_rays=N, H=128, W=128, focal=120., n_joints=24)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    rays = torch.cat([t(sc["rays_o"]), t(sc["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev),
                      torch.nn.functional.normalize(t(sc["rays_d"]), dim=-1)], 1)
    rng = np.random.RandomState(8)
    return sc, rays, t(rng.rand(N, 3).astype(np.float32)), t(rng.rand(N, 3).astype(np.float32))


@pytest.mark.parametrize("loss_fn", ["L1", "MSE"])
def test_fused_step_equals_the_autograd_route(loss_fn):
    dev = torch.device("cuda")
    sc, rays, target, bgs = _batch(dev)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    cams = (torch.arange(rays.shape[0], device=dev) % 3).float()
    runs = {}
    for mode in ("autograd", "fused"):
        rk_train, rc, grad_vars, optimizer = _setup(dev, opt_framecode=True)
        kw = {k: v for k, v in rk_train.items() if k not in ("ray_caster", "use_viewdirs")}
        kw.update(perturb=0., raw_noise_std=0.)
        holder = rk_train["ray_caster"].train()
        step = FusedTrainStep(rc, optimizer, loss_fn=loss_fn, coarse_weight=0.7, use_background=True)
        losses = []
        for it in range(3):
            if mode == "fused":
                out, stats = step(rays, target, kp_batch=t(sc["kps"]), skts=t(sc["skts"]), cyls=t(sc["cyls"]), bones=t(sc["bones"]), cams=cams,
                                  bgs=bgs, **kw)
                losses.append(FusedTrainStep.losses(stats, mse=loss_fn == "MSE", coarse_weight=0.7)["total_loss"])
            else:
                optimizer.zero_grad()
                out = holder(rays, kp_batch=t(sc["kps"]), skts=t(sc["skts"]), cyls=t(sc["cyls"]), bones=t(sc["bones"]), cams=cams, subject_idxs=None, **kw)
                f = (lambda a, b: ((a - b) ** 2).mean()) if loss_fn == "MSE" else (lambda a, b: (a - b).abs().mean())
                pred = out["rgb_map"] + (1. - out["acc_map"])[..., None] * bgs            # core/trainer.py:362-365
                pred0 = out["rgb0"] + (1. - out["acc0"])[..., None] * bgs
                loss = f(pred, target) + f(pred0, target) * 0.7
                loss.backward()
                optimizer.step()
                losses.append(float(loss))
        runs[mode] = (losses, {k: v.detach().clone() for k, v in rc.state_dict()["network_fine_state_dict"].items()},
                      {k: v.detach().clone() for k, v in rc.state_dict()["network_fn_state_dict"].items()})
    (la, fa, ca), (lb, fb, cb) = runs["autograd"], runs["fused"]
    assert lb[2] < lb[0]                                        # it trains
    for x, y in zip(la, lb):
        assert abs(x - y) < 2e-5 * max(1., abs(x)), (la, lb)
    for wa, wb in ((fa, fb), (ca, cb)):
        for k in wa:
            assert float((wa[k] - wb[k]).abs().max()) < 2e-5 * max(1., float(wa[k].abs().max())), k


def test_fused_step_with_pose_refinement():
    """skts from PoseOptLayer.forward_poses (requires grad): the step hands d/d skts (per pose) back to the pose chain."""
    dev = torch.device("cuda")
    sc, rays, target, bgs = _batch(dev, N=96)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    P = 3
    poses = [synthetic.make_pose(70 + p, 24, pose_std=0.15) for p in range(P)]
    kp_idx = np.random.RandomState(1).randint(0, P, size=rays.shape[0])
    got = {}
    for mode in ("autograd", "fused"):
        rk_train, rc, grad_vars, optimizer = _setup(dev)
        layer = PoseOptLayer(torch.as_tensor(np.stack([p["kps"] for p in poses])), torch.as_tensor(np.stack([p["bones"] for p in poses])),
                             torch.as_tensor(synthetic.humanoid_rest_pose()[None]), use_rot6d=True, parents=synthetic.SMPL_PARENTS, root_id=0).to(dev)
        kw = {k: v for k, v in rk_train.items() if k not in ("ray_caster", "use_viewdirs")}
        kw.update(perturb=0., raw_noise_std=0.)
        (kps, bones, skts, _, _), pose_idx = layer.forward_poses(kp_idx)
        if mode == "fused":
            FusedTrainStep(rc, optimizer, loss_fn="L1", use_background=False)(rays, target, kp_batch=kps, skts=skts, cyls=t(sc["cyls"]), bones=bones,
                                                                               pose_idx=pose_idx, **kw)
        else:
            out = rk_train["ray_caster"].train()(rays, kp_batch=kps, skts=skts, cyls=t(sc["cyls"]), bones=bones, cams=None, subject_idxs=None,
                                                 pose_idx=pose_idx, **kw)
            ((out["rgb_map"] - target).abs().mean() + (out["rgb0"] - target).abs().mean()).backward()
        got[mode] = (layer.bones.grad.clone(), layer.pelvis.grad.clone())
    for a, b in zip(got["autograd"], got["fused"]):
        assert float(a.abs().max()) > 0
        assert float((a - b).abs().max()) < 2e-5 * float(a.abs().max())
