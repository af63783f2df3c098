"""The reference-facing Python boundary (create_raycaster / RayCaster) on the GPU."""
import argparse
import collections
import os
import tempfile

import numpy as np
import pytest
import torch

from anerf_b200 import synthetic
from anerf_b200.raycasters import batchify_rays, create_raycaster
from tests.common import build_case, load_golden, rel_err, run_oracle

pytestmark = pytest.mark.gpu


def make_args(**over):
    tmp = tempfile.mkdtemp(prefix="anerf_t_")
    os.makedirs(os.path.join(tmp, "exp"), exist_ok=True)
    d = dict(n_framecodes=None, use_cutoff=True, normalize_cutoff=False, cutoff_mm=500., ext_scale=0.001,
             cutoff_inputs=True, opt_cutoff=False, freq_schedule=False, init_freq=0., cut_to_dist=False,
             cutoff_shift=False, multires=7, i_embed=0, cutoff_bones=False, multires_bones=0, use_viewdirs=True,
             cutoff_viewdir=True, multires_views=4, N_importance=16, netdepth=8, netwidth=256, opt_framecode=False,
             framecode_size=16, density_scale=1.0, single_net=False, lrate=5e-4, ft_path=None, basedir=tmp,
             expname="exp", no_reload=False, finetune=False, fix_layer=0, weight_decay=None, density_type="relu",
             softplus_shift=0., pts_tr_type="local", kp_dist_type="reldist", view_type="relray", bone_type="reldir",
             debug=True, perturb=0., N_samples=64, raw_noise_std=0., ray_noise_std=0., lindisp=False, nerf_type="nerf",
             cutoff_step=250, cutoff_rate=10., freq_schedule_step=50)
    d.update(over)
    return argparse.Namespace(**d)


def data_attrs(J=24, n_views=1):
    Skel = collections.namedtuple("Skel", ["joint_names", "joint_trees", "root_id"])
    return dict(skel_type=Skel([f"j{i}" for i in range(J)], synthetic.SMPL_PARENTS[:J], 0), near=0., far=1.,
                n_views=n_views, joint_coords=np.tile(np.eye(3, dtype=np.float32), (1, J, 1, 1)))


def test_create_raycaster_contract_and_render():
    case, gold = load_golden("surreal_j24_s64_i16_tau200")
    scene, sd0, sd1, cfg, _ = build_case(case)
    args = make_args()
    rk_train, rk_test, start, grad_vars, optimizer, loaded = create_raycaster(args, data_attrs())
    assert start == 0 and loaded is None and len(grad_vars) == 2 * 24
    assert set(rk_test) == {'ray_caster', 'perturb', 'N_importance', 'N_samples', 'use_viewdirs', 'raw_noise_std',
                            'ray_noise_std', 'ext_scale', 'preproc_kwargs', 'lindisp', 'nerf_type'}
    rc = rk_test['ray_caster']
    assert rk_train['ray_caster'].module is rc
    assert sum(p.numel() for p in rc.network.parameters()) == 864260          # reference prints this number
    assert set(rc.state_dict()) == {'network_fn_state_dict', 'network_fine_state_dict', 'embed_state_dict',
                                    'embedbones_state_dict', 'embeddirs_state_dict'}
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in sd0.items()})
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in sd1.items()})
    rc.embed_fn.tau.fill_(200.)
    rc.embeddirs_fn.tau.fill_(200.)
    rc.eval()
    dev = torch.device("cuda")
    t = lambda a: torch.as_tensor(a).to(dev)
    N = scene["rays_o"].shape[0]
    rays = torch.cat([t(scene["rays_o"]), t(scene["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev),
                      torch.nn.functional.normalize(t(scene["rays_d"]), dim=-1)], 1)        # [N,11] like trainer.render
    kw = {k: v for k, v in rk_test.items() if k not in ('ray_caster', 'use_viewdirs')}
    out = batchify_rays(rays, 24, ray_caster=rc, kp_batch=t(scene["kps"]), skts=t(scene["skts"]), cyls=t(scene["cyls"]),
                        bones=t(scene["bones"]), cams=None, subject_idxs=None, **kw)
    one = rc(rays, kp_batch=t(scene["kps"]), skts=t(scene["skts"]), cyls=t(scene["cyls"]), bones=t(scene["bones"]),
             cams=None, subject_idxs=None, **kw)
    for k in ("rgb0", "disp0", "acc0", "alpha0"):
        assert rel_err(one[k].cpu().numpy(), gold["ref_" + k]) < 1e-4, k
        assert out[k].shape == one[k].shape
    # checkpoint round trip through the reference's nested layout
    path = os.path.join(args.basedir, args.expname, "000100.tar")
    torch.save({'global_step': 100, 'optimizer_state_dict': optimizer.state_dict(), **rc.state_dict()}, path)
    _, rk2, start2, _, _, loaded2 = create_raycaster(make_args(basedir=args.basedir), data_attrs())
    assert start2 == 100 and loaded2 is not None
    rc2 = rk2['ray_caster'].eval()
    assert float(rc2.embed_fn.tau) == 200.
    two = rc2(rays, kp_batch=t(scene["kps"]), skts=t(scene["skts"]), cyls=t(scene["cyls"]), bones=t(scene["bones"]),
              cams=None, subject_idxs=None, **kw)
    assert torch.equal(two["rgb_map"], one["rgb_map"])
    # in-place parameter updates are picked up (re-pack on version change)
    with torch.no_grad():
        rc2.network_fine.rgb_linear.bias.add_(0.5)
    three = rc2(rays, kp_batch=t(scene["kps"]), skts=t(scene["skts"]), cyls=t(scene["cyls"]), bones=t(scene["bones"]),
                cams=None, subject_idxs=None, **kw)
    assert not torch.equal(three["rgb_map"], one["rgb_map"]) and torch.equal(three["rgb0"], one["rgb0"])


def test_unsupported_flags_raise():
    for bad in (dict(kp_dist_type='relpos'), dict(view_type='world'), dict(use_viewdirs=False),
                dict(multires=10), dict(cutoff_bones=True)):
        with pytest.raises(NotImplementedError):
            create_raycaster(make_args(**bad), data_attrs())
    with pytest.raises(NotImplementedError):
        create_raycaster(make_args(density_type='gelu'), data_attrs())


def test_mesh_density_entry_point():
    case, gold = load_golden("mesh_j24_res15")
    J = case["n_joints"]
    pose = synthetic.make_pose(11, J)
    _, rk, _, _, _, _ = create_raycaster(make_args(N_importance=16, no_reload=True), data_attrs(J))
    rc = rk['ray_caster'].eval()
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202).items()})
    dev = torch.device("cuda")
    t = lambda a: torch.as_tensor(a).to(dev)
    sig = rc(kps=t(pose["kps"])[None], skts=t(pose["skts"])[None], bones=t(pose["bones"])[None], radius=case["radius"],
             render_kwargs=rk['preproc_kwargs'], res=case["res"], netchunk=4096, fwd_type='mesh')
    assert tuple(sig.shape) == (16, 16, 16)
    assert rel_err(sig.cpu().numpy(), gold["ref_sigma"]) < 1e-4


def test_sharded_density_grid_and_ragged_point_counts():
    from anerf_b200 import mesh
    case, gold = load_golden("mesh_j24_res15")
    J = case["n_joints"]
    pose = synthetic.make_pose(11, J)
    _, rk, _, _, _, _ = create_raycaster(make_args(N_importance=16, no_reload=True), data_attrs(J))
    rc = rk['ray_caster'].eval()
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202).items()})
    dev = torch.device("cuda")
    t = lambda a: torch.as_tensor(a).to(dev)
    kps, skts = t(pose["kps"])[None], t(pose["skts"])[None]
    sig = mesh.density_grid_sharded(rc, kps, skts, radius=case["radius"], res=case["res"])
    assert rel_err(sig.cpu().numpy(), gold["ref_sigma"]) < 1e-4
    # the generated slab points equal the reference's meshgrid, and point counts that are not a multiple of the
    # 128-row tile (or of the CTA pair) work
    t_ = np.linspace(-case["radius"], case["radius"], case["res"] + 1)
    ref_pts = np.stack(np.meshgrid(t_, t_, t_), axis=-1).astype(np.float32).reshape(-1, 3) + pose["kps"][0]
    pts = mesh.grid_points(kps, case["radius"], case["res"])
    assert np.abs(pts.cpu().numpy() - ref_pts).max() < 1e-6
    full = rc.render_pts_density(pts.reshape(-1, 1, 3), kps, skts, None).reshape(-1)
    for n in (1, 127, 129, 1000):
        part = rc.render_pts_density(pts[:n].reshape(-1, 1, 3), kps, skts, None).reshape(-1)
        assert torch.equal(part, full[:n])
    # points generated inside the kernel (anerf_density_grid) == the explicit points, bit for bit, whole grid and slabs;
    # a radius whose linspace steps are not exact in fp32 exercises the fp64 coordinate arithmetic
    n1 = case["res"] + 1
    grid = rc.render_mesh_density(kps, skts, None, radius=case["radius"], res=case["res"])
    assert torch.equal(grid, full.reshape(n1, n1, n1).transpose(1, 0))
    for first, count in ((0, 1), (5, 300), (n1 ** 3 - 129, 129)):
        slab = rc.render_mesh_density(kps, skts, None, radius=case["radius"], res=case["res"], first=first, count=count)
        assert torch.equal(slab, full[first:first + count])
    for radius, res in ((1.3, 9), (0.77, 12)):
        t_ = np.linspace(-radius, radius, res + 1)
        p_ref = torch.as_tensor(np.stack(np.meshgrid(t_, t_, t_), axis=-1).astype(np.float32).reshape(-1, 3)).to(dev) + kps[0, 0]
        a = rc.render_pts_density(p_ref.reshape(-1, 1, 3), kps, skts, None).reshape(-1)
        b = rc.render_mesh_density(kps, skts, None, radius=radius, res=res, first=0, count=(res + 1) ** 3)
        assert torch.equal(a, b)


def test_chunked_host_entry_point_equals_per_chunk_calls():
    """anerf_render_fwd_host_chunked (a frame of host buffers, copies overlapped with the kernels, optional outputs)
    against one anerf_render_fwd_host call per chunk: identical results, ragged last chunk, chunk-wise near/far repair."""
    from anerf_b200 import _lib
    from tests.common import build_case
    case, _ = load_golden("surreal_j24_s64_i16_tau200")
    scene, sd0, sd1, cfg, _ = build_case(case)
    scene["rays_d"] = scene["rays_d"].copy()
    scene["rays_d"][3:6, 0] += 2.0                      # rays that miss the cylinder: repaired with their CHUNK's mean
    dev = torch.device("cuda")
    N = scene["rays_o"].shape[0]
    plan = _lib.Plan(cfg.n_joints, cfg.D, cfg.W, cfg.skips, 0, 0, 0)
    p0 = plan.pack({k: torch.as_tensor(v).to(dev) for k, v in sd0.items()})
    p1 = plan.pack({k: torch.as_tensor(v).to(dev) for k, v in sd1.items()})
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).pin_memory()
    rays = pin(np.concatenate([scene["rays_o"], scene["rays_d"], np.zeros((N, 1), np.float32), np.ones((N, 1), np.float32)], 1))
    skts, cyls = pin(scene["skts"]), pin(scene["cyls"])
    mk = lambda n: _lib.make_opts(n, cfg.N_samples, cfg.N_importance, tau_pts=cfg.tau, tau_views=cfg.tau_views)
    chunk = 24                                           # 64 rays -> 24, 24, 16
    ref = {}
    for i in range(0, N, chunk):
        o = _lib.render_fwd_host(plan, p0, p1, mk(min(chunk, N - i)), rays[i:i + chunk], skts[i:i + chunk], cyls[i:i + chunk])
        for k, v in o.items():
            ref.setdefault(k, []).append(v.clone())
    ref = {k: torch.cat(v) for k, v in ref.items()}
    for rep in range(2):                                 # second call re-uses the cached arenas / streams
        out = _lib.render_fwd_host_chunked(plan, p0, p1, mk(N), chunk, rays, skts, cyls)
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
    slim = _lib.render_fwd_host_chunked(plan, p0, p1, mk(N), chunk, rays, skts, cyls, keys=("rgb_map", "disp_map", "acc_map"))
    assert set(slim) == {"rgb_map", "disp_map", "acc_map"} and all(torch.equal(slim[k], ref[k]) for k in slim)
    _lib.check_status()


def test_render_frame_equals_explicit_rays():
    """anerf_render_frame (rays generated in the kernels, one pose per frame) against the reference-style call with
    explicit get_rays rays and the pose replicated per ray: identical pixels, including ragged chunks, chunk-wise
    near/far repair, a pixel subset and framecodes."""
    H, W, focal = 40, 56, 55.0
    J = 24
    pose = synthetic.make_pose(11, J)
    c2w = synthetic.orbit_c2w(0.4, 3.0, centre=pose['kps'][0] * np.array([1., 0., 1.])).astype(np.float32)
    _, rk, _, _, _, _ = create_raycaster(make_args(N_importance=16, no_reload=True, opt_framecode=True), data_attrs(J, n_views=3))
    rc = rk['ray_caster'].eval()
    wk = dict(framecode_ch=16, n_framecodes=3)
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(101, **wk).items()})
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202, **wk).items()})
    dev = torch.device("cuda")
    t = lambda a: torch.as_tensor(a).to(dev)
    # the reference's get_rays (core/utils/ray_utils.py:6-28)
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='ij')
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * torch.as_tensor(c2w[:3, :3]), -1).reshape(-1, 3)
    rays_o = torch.as_tensor(c2w[:3, 3]).expand(rays_d.shape)
    n = H * W
    rays = torch.cat([rays_o, rays_d, torch.zeros(n, 1), torch.ones(n, 1), torch.nn.functional.normalize(rays_d, dim=-1)], 1).to(dev)
    skts, cyl, kps, bones = t(pose["skts"]), t(pose["cyl"]), t(pose["kps"]), t(pose["bones"])
    kw = {k: v for k, v in rk.items() if k not in ('ray_caster', 'use_viewdirs')}
    chunk = 1000                                          # ragged last chunk; several chunks with missing rays
    for pix in (None, torch.arange(n)[::3][:777]):
        sel = slice(None) if pix is None else pix.to(dev)
        m = n if pix is None else len(pix)
        ref = batchify_rays(rays[sel], chunk, ray_caster=rc, kp_batch=kps.expand(m, J, 3), skts=skts.expand(m, J, 4, 4),
                            cyls=cyl.expand(m, 5), bones=bones.expand(m, J, 3), cams=torch.full((m,), 2., device=dev),
                            subject_idxs=None, **kw)
        out = rc.render_frame(H, W, focal, c2w, skts[None], cyl[None], cams=torch.tensor([2.]), pixel_idx=pix, chunk=chunk, **kw)
        for k in ref:
            assert torch.equal(out[k], ref[k]), (k, float((out[k] - ref[k]).abs().max()))
    assert float(ref['acc_map'].max()) > 0.5 and float(ref['acc_map'].min()) < 0.01


def test_render_path_frame_composites_like_render_path():
    """frames.render_path_frame (box of the cylinder -> render_frame on those pixels -> composite over the background,
    run_nerf.py:77-136) against the same steps done with explicit rays through the reference-style call."""
    from anerf_b200 import frames
    H, W, focal, J = 96, 128, 110.0, 24
    pose = synthetic.make_pose(11, J)
    c2w = synthetic.orbit_c2w(5.1, 2.6, centre=pose['kps'][0] * np.array([1., 0., 1.])).astype(np.float32)
    _, rk, _, _, _, _ = create_raycaster(make_args(N_importance=16, no_reload=True), data_attrs(J))
    rc = rk['ray_caster'].eval()
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(101).items()})
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202).items()})
    dev = torch.device("cuda")
    t = lambda a: torch.as_tensor(a).to(dev)
    skts, cyl, kps, bones = t(pose["skts"]), t(pose["cyl"]), t(pose["kps"]), t(pose["bones"])
    bg = torch.rand(H, W, 3, device=dev)
    rgb, disp, acc = frames.render_path_frame(rc, c2w, H, W, focal, skts[None], cyl[None], rk, bg=bg, chunk=4096)
    assert rgb.shape == (H, W, 3) and disp.shape == (H, W, 1) and acc.shape == (H, W, 1)
    # the same with explicit rays
    idx, (tl, br) = frames.valid_pixels(pose["cyl"], H, W, focal, c2w, device=dev)
    assert 0 < len(idx) < H * W
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='ij')
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * torch.as_tensor(c2w[:3, :3]), -1).reshape(-1, 3)[idx.cpu().long()]
    m = len(idx)
    rays = torch.cat([torch.as_tensor(c2w[:3, 3]).expand(m, 3), rays_d, torch.zeros(m, 1), torch.ones(m, 1),
                      torch.nn.functional.normalize(rays_d, dim=-1)], 1).to(dev)
    kw = {k: v for k, v in rk.items() if k not in ('ray_caster', 'use_viewdirs')}
    ref = batchify_rays(rays, 4096, ray_caster=rc, kp_batch=kps.expand(m, J, 3), skts=skts.expand(m, J, 4, 4),
                        cyls=cyl.expand(m, 5), bones=bones.expand(m, J, 3), cams=None, subject_idxs=None, **kw)
    img = bg.reshape(-1, 3).clone()
    img[idx.long()] = ref['rgb_map'] + (1. - ref['acc_map'][..., None]) * img[idx.long()]
    assert torch.equal(rgb.reshape(-1, 3), img)
    outside = torch.ones(H * W, dtype=torch.bool, device=dev)
    outside[idx.long()] = False
    assert torch.equal(rgb.reshape(-1, 3)[outside], bg.reshape(-1, 3)[outside]) and float(acc.reshape(-1)[outside].abs().max()) == 0.


def test_single_net_through_the_boundary():
    """--single_net (configs/surreal/surreal_single.txt): one network for both passes, blurred importance pdf; forward
    against the reference's golden outputs, and loss.backward() accumulates both passes into the one network."""
    case, gold = load_golden("single_j24_s64_i48")
    scene, sd0, sd1, cfg, _ = build_case(case)
    rk_train, rk, _, grad_vars, _, _ = create_raycaster(make_args(N_importance=48, single_net=True, no_reload=True), data_attrs())
    rc = rk['ray_caster']
    assert rc.network_fine is rc.network and len(grad_vars) == 24
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in sd0.items()})
    rc.eval()
    dev = torch.device("cuda")
    t = lambda a: torch.as_tensor(a).to(dev)
    N = scene["rays_o"].shape[0]
    rays = torch.cat([t(scene["rays_o"]), t(scene["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev),
                      torch.nn.functional.normalize(t(scene["rays_d"]), dim=-1)], 1)
    kw = {k: v for k, v in rk.items() if k not in ('ray_caster', 'use_viewdirs')}
    batch = dict(kp_batch=t(scene["kps"]), skts=t(scene["skts"]), cyls=t(scene["cyls"]), bones=t(scene["bones"]), cams=None,
                 subject_idxs=None)
    out = rc(rays, **batch, **kw)
    for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "alpha0"):
        assert rel_err(out[k].cpu().numpy(), gold["ref_" + k]) < 1e-4, k
    holder = rk_train['ray_caster'].train()
    o = holder(rays, **batch, **dict(kw, perturb=0.))
    (o['rgb_map'].sum() + o['rgb0'].sum()).backward()
    g = rc.network.pts_linears[3].weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0
