"""SURVEY.md 8(f) row 2 on the GPU: the fused pose chain (anerf_pose_chain_fwd/bwd behind anerf_b200.pose_opt) against
the outputs and autograd gradients of the unmodified reference's PoseOptLayer (tests/golden/pose_chain.npz, written by
oracle/make_golden_pose.py), and the renderer's per-pose `skts` path (pose_idx) against its per-ray form."""
import numpy as np
import pytest
import torch

from anerf_b200 import _lib, synthetic
from anerf_b200.pose_opt import PoseOptLayer
from oracle.make_golden_pose import inputs as pose_inputs
from tests.common import GOLDEN_DIR, build_case, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _layer(z, dev):
    P, J = z["bones6"].shape[:2]
    layer = PoseOptLayer(torch.zeros(P, J, 3), torch.zeros(P, J, 3) + 0.1, torch.as_tensor(z["rest"]), use_rot6d=True,
                         parents=z["parents"], root_id=int(z["root_id"])).to(dev)
    with torch.no_grad():
        layer.bones.copy_(torch.as_tensor(z["bones6"]))
        layer.pelvis.copy_(torch.as_tensor(z["pelvis"]))
    return layer


def test_pose_chain_matches_reference_poseoptlayer():
    import os
    z = np.load(os.path.join(GOLDEN_DIR, "pose_chain.npz"))
    dev = torch.device("cuda")
    layer = _layer(z, dev)
    assert set(layer.state_dict().keys()) == {"rest_pose", "pelvis", "bones"}          # the reference's checkpoint keys
    _, _, _, idxs, cot = pose_inputs()
    assert np.array_equal(idxs, z["idxs"])
    kps, bone, skts, l2ws, rots = layer(idxs)                                            # the reference's call convention
    for name, got in (("kps", kps), ("skts", skts), ("l2ws", l2ws), ("rots", rots)):
        assert got.shape == z["ref_" + name].shape
        assert rel_err(got.detach().cpu().numpy(), z["ref_" + name]) < 1e-5, name
    t = lambda a: torch.as_tensor(a).to(dev)
    loss = (kps * t(cot["kps"])).sum() + (skts * t(cot["skts"])).sum() + (l2ws * t(cot["l2ws"])).sum()
    g_bones, g_pelvis = torch.autograd.grad(loss, [layer.bones, layer.pelvis])
    assert rel_err(g_bones.cpu().numpy(), z["ref_g_bones"]) < 2e-5
    assert rel_err(g_pelvis.cpu().numpy(), z["ref_g_pelvis"]) < 2e-5
    # the fused form returns each pose once + the ray -> pose index; gathered, it is the same thing
    (kps_p, _, skts_p, l2ws_p, _), pose_idx = layer.forward_poses(idxs)
    assert skts_p.shape[0] == len(np.unique(idxs)) and pose_idx.dtype == torch.int32
    assert torch.equal(skts_p[pose_idx.long()], skts) and torch.equal(kps_p[pose_idx.long()], kps)


def test_bad_kinematic_trees_are_rejected():
    dev = torch.device("cuda")
    f = lambda *s: torch.zeros(*s, device=dev)
    with pytest.raises(RuntimeError, match="parent"):
        _lib.pose_chain_fwd(f(2, 3, 3, 3), f(1, 3, 3), f(2, 3), parents=[0, 2, 1])       # joint 1's parent comes after it
    with pytest.raises(RuntimeError, match="rest_pose"):
        _lib.pose_chain_fwd(f(2, 3, 3, 3), f(3, 3, 3), f(2, 3), parents=[0, 0, 1])


def _train_setup(dev, n_poses=5, N=96):
    case, _ = load_golden("grad_j24_s16_i8_fc_perturb")
    scene, sd0, sd1, cfg, draws = build_case(case)
    N = min(N, scene["rays_o"].shape[0])
    rng = np.random.RandomState(4)
    pose_idx = rng.randint(0, n_poses, size=N).astype(np.int32)
    poses = [synthetic.make_pose(100 + p, cfg.n_joints, pose_std=0.15) for p in range(n_poses)]
    skts_pose = np.stack([p["skts"] for p in poses]).astype(np.float32)
    return scene, sd0, sd1, cfg, draws, N, pose_idx, skts_pose


def test_render_with_pose_index_equals_per_ray_transforms():
    """anerf_render_fwd / anerf_render_bwd with `pose_idx` + per-POSE skts against the same rays with the transforms
    replicated per ray: identical outputs and parameter gradients; d/d skts comes back per pose and equals the segment sum
    of the per-ray gradient (what the reference's `skts[inverse_idxs]` backward computes)."""
    from oracle import grad_tools as gt
    dev = torch.device("cuda")
    scene, sd0, sd1, cfg, draws, N, pose_idx, skts_pose = _train_setup(dev)
    t = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(dev)
    fc = cfg.framecode_ch > 0
    plan = _lib.Plan(cfg.n_joints, cfg.D, cfg.W, cfg.skips, cfg.framecode_ch, 0 if not fc else sd0['framecodes.codes.weight'].shape[0], 0)
    names = _lib.param_names(cfg.D, fc)
    d0, d1 = {k: t(v) for k, v in sd0.items()}, {k: t(v) for k, v in sd1.items()}
    p0, p1 = plan.pack(d0), plan.pack(d1)
    rays = t(np.concatenate([scene["rays_o"], scene["rays_d"], np.zeros_like(scene["rays_o"][:, :1]), np.ones_like(scene["rays_o"][:, :1])], 1)[:N])
    opts = _lib.make_opts(N, cfg.N_samples, cfg.N_importance, tau_pts=cfg.tau, tau_views=cfg.tau_views)
    cams = t(scene["cams"].astype(np.float32)[:N]) if fc else None
    dr = {k: t(v[:N]) for k, v in draws.items()}
    cot = {k: t(v) for k, v in gt.cotangents(N, cfg.N_samples, cfg.N_importance).items()}
    pidx = torch.as_tensor(pose_idx).to(dev)
    res = {}
    for mode in ("per_ray", "per_pose"):
        skts = t(skts_pose[pose_idx]) if mode == "per_ray" else t(skts_pose)
        kw = {} if mode == "per_ray" else dict(pose_idx=pidx)
        out = _lib.render_fwd(plan, p0, p1, opts, rays, skts, t(scene["cyls"][:N]), cams, dr["t_rand"], dr["u_rand"], dr["noise0"],
                              dr["noise1"], keep_nearfar=True, want_z_all=True, **kw)
        g0, g1, g_skts = _lib.render_bwd(plan, opts, [d0[k] for k in names], [d1[k] for k in names], rays, skts, cams, dr["t_rand"],
                                         dr["noise0"], dr["noise1"], out["nearfar"].contiguous(), out["z_all"], cot,
                                         [True] * len(names), [True] * len(names), True, **kw)
        torch.cuda.synchronize()
        res[mode] = (out, g0, g1, g_skts)
    (oa, g0a, g1a, gsa), (ob, g0b, g1b, gsb) = res["per_ray"], res["per_pose"]
    for k in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "alpha0", "z_all"):
        assert torch.equal(oa[k], ob[k]), k
    for a, b in zip(g0a + g1a, g0b + g1b):
        assert rel_err(b.cpu().numpy(), a.cpu().numpy()) < 1e-5
    assert gsb.shape == (skts_pose.shape[0], cfg.n_joints, 4, 4)
    seg = torch.zeros_like(gsb).index_add_(0, pidx.long(), gsa)
    assert rel_err(gsb.cpu().numpy(), seg.cpu().numpy()) < 1e-5
    # an out-of-range pose index is clamped on the device (never an out-of-bounds access)
    bad = pidx.clone()
    bad[0], bad[1] = 10 ** 6, -5
    out = _lib.render_fwd(plan, p0, p1, opts, rays, t(skts_pose), t(scene["cyls"][:N]), cams, dr["t_rand"], dr["u_rand"], dr["noise0"],
                          dr["noise1"], pose_idx=bad)
    torch.cuda.synchronize()
    assert torch.isfinite(out["rgb_map"]).all()
    _lib.check_status()


def test_pose_refinement_step_through_the_boundary():
    """PoseOptLayer -> RayCaster -> loss -> backward, fused (forward_poses + pose_idx) against the reference-style call
    (per-ray gathered transforms): same loss, same gradients on the pose parameters and on the network."""
    import collections
    import contextlib
    import io
    import os
    from anerf_b200.raycasters import create_raycaster
    from tests.test_gpu_api import make_args, data_attrs
    dev = torch.device("cuda")
    z = np.load(os.path.join(GOLDEN_DIR, "pose_chain.npz"))
    with contextlib.redirect_stdout(io.StringIO()):
        rk_train, rk_test, _, grad_vars, _, _ = create_raycaster(make_args(N_importance=8, N_samples=16, no_reload=True), data_attrs(24))
    rc = rk_test["ray_caster"]
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(101).items()})
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202).items()})
    holder = rk_train["ray_caster"].train()
    P, N = 4, 64
    poses = [synthetic.make_pose(50 + p, 24, pose_std=0.15) for p in range(P)]
    rest = synthetic.humanoid_rest_pose()[None]
    kps0 = np.stack([p["kps"] for p in poses]); bones0 = np.stack([p["bones"] for p in poses])
    layer = PoseOptLayer(torch.as_tensor(kps0), torch.as_tensor(bones0), torch.as_tensor(rest), use_rot6d=True,
                         parents=synthetic.SMPL_PARENTS, root_id=0).to(dev)
    sc = synthetic.make_scene(seed=50, n_rays=N, H=256, W=256, focal=250., n_joints=24)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    rays = torch.cat([t(sc["rays_o"]), t(sc["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev),
                      torch.nn.functional.normalize(t(sc["rays_d"]), dim=-1)], 1)
    kp_idx = np.random.RandomState(1).randint(0, P, size=N)
    target = t(np.random.RandomState(2).rand(N, 3).astype(np.float32))
    kw = {k: v for k, v in rk_train.items() if k not in ("ray_caster", "use_viewdirs")}
    kw.update(perturb=0., raw_noise_std=0.)
    got = {}
    for mode in ("reference_style", "fused"):
        for p in list(layer.parameters()) + grad_vars:
            p.grad = None
        if mode == "fused":
            (kps, bones, skts, _, _), pose_idx = layer.forward_poses(kp_idx)
            out = holder(rays, kp_batch=kps[pose_idx.long()], skts=skts, cyls=t(sc["cyls"]), bones=bones[pose_idx.long()], cams=None,
                         subject_idxs=None, pose_idx=pose_idx, **kw)
        else:
            kps, bones, skts, _, _ = layer(kp_idx)
            out = holder(rays, kp_batch=kps, skts=skts, cyls=t(sc["cyls"]), bones=bones, cams=None, subject_idxs=None, **kw)
        loss = ((out["rgb_map"] - target) ** 2).mean() + ((out["rgb0"] - target) ** 2).mean()
        loss.backward()
        got[mode] = (float(loss), layer.bones.grad.clone(), layer.pelvis.grad.clone(), rc.network_fine.pts_linears[3].weight.grad.clone())
    a, b = got["reference_style"], got["fused"]
    assert a[0] == b[0]
    assert float(a[1].abs().max()) > 0 and float(a[2].abs().max()) > 0
    for x, y in zip(a[1:], b[1:]):
        assert rel_err(y.cpu().numpy(), x.cpu().numpy()) < 2e-5
