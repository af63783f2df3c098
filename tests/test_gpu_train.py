"""Training path on the GPU: gradients of the CUDA backward (C ABI anerf_render_bwd, and the same through the
Python boundary's autograd node) against the oracle's autograd on identical rays, weights, draws and output
cotangents, and against the digests of the reference's own autograd in tests/golden/grad_*.npz.

Tolerances (max-norm relative, per tensor): 2e-4 against the oracle evaluated at the kernel's own fine sample
positions -- for the DEFAULT tensor-core engine (fp16 hi/lo operands with per-matrix power-of-two scales) as well as for
the fp32 SIMT engine; 2e-3 against the reference's stored gradients (its sample positions differ by the fp32 conditioning
of the inverse-CDF step).  Round 1's bf16 hi/lo operands (16 mantissa bits) remain selectable
(ANERF_TRAIN_GEMM=bf16) and are held to a norm-wise bound."""
import numpy as np
import pytest
import torch

from anerf_b200 import _lib, synthetic
from oracle import grad_tools as gt
from tests.common import RENDER_CASES, build_case, load_golden, run_oracle

pytestmark = pytest.mark.gpu

GRAD_CASES = ["grad_cfg1_j1_s16_i16", "grad_j24_s24_i0", "grad_j24_s16_i8_fc_perturb", "grad_single_j24_s16_i8"]
TC_L2_TOL = {"grad_cfg1_j1_s16_i16": 3e-4, "grad_j24_s24_i0": 1e-3, "grad_j24_s16_i8_fc_perturb": 3e-2, "grad_single_j24_s16_i8": 1e-3}


def gpu_grads(scene, sd0, sd1, cfg, draws, cot, need_pose=True, saved=False):
    """saved: the forward that keeps its activations (anerf_render_fwd_train) + the backward that starts from them
    (anerf_render_bwd_saved) instead of the fused forward kernel + the recomputing backward."""
    dev = torch.device("cuda")
    t = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(dev)
    N = scene["rays_o"].shape[0]
    fc = cfg.framecode_ch > 0
    plan = _lib.Plan(cfg.n_joints, cfg.D, cfg.W, cfg.skips, cfg.framecode_ch,
                     0 if not fc else sd0['framecodes.codes.weight'].shape[0], 0, view_freqs=cfg.multires_views)
    names = _lib.param_names(cfg.D, fc)
    d0 = {k: t(v) for k, v in sd0.items()}
    d1 = None if sd1 is None else {k: t(v) for k, v in sd1.items()}
    p0 = plan.pack(d0)
    p1 = None if d1 is None else plan.pack(d1)
    rays = t(np.concatenate([scene["rays_o"], scene["rays_d"], np.zeros((N, 1), np.float32), np.ones((N, 1), np.float32)], 1))
    opts = _lib.make_opts(N, cfg.N_samples, cfg.N_importance, tau_pts=cfg.tau, tau_views=cfg.tau_views,
                          cutoff_pts=cfg.cutoff_dist, cutoff_views=cfg.cutoff_dist, n_joints=cfg.n_joints,
                          single_net=getattr(cfg, "single_net", False), lindisp=cfg.lindisp,
                          softplus=cfg.density_type == "softplus", softplus_shift=cfg.softplus_shift, density_scale=cfg.density_scale)
    d = draws or {}
    cams = t(scene["cams"].astype(np.float32)) if fc else None
    skts = t(scene["skts"])
    params0 = [d0[k] for k in names]
    params1 = None if d1 is None else [d1[k] for k in names]
    state = None
    if saved:
        state = torch.empty(_lib.train_state_bytes(plan, opts), dtype=torch.uint8, device=dev)
        assert state.numel() > 0
        out = _lib.render_fwd_train(plan, opts, params0, params1, rays, skts, t(scene["cyls"]), state, cams, t(d.get("t_rand")),
                                    t(d.get("u_rand")), t(d.get("noise0")), t(d.get("noise1")))
    else:
        out = _lib.render_fwd(plan, p0, p1, opts, rays, skts, t(scene["cyls"]), cams, t(d.get("t_rand")), t(d.get("u_rand")),
                              t(d.get("noise0")), t(d.get("noise1")), keep_nearfar=True, want_z_all=True)
    g0, g1, g_skts = _lib.render_bwd(plan, opts, params0, params1, rays, skts, cams, t(d.get("t_rand")), t(d.get("noise0")),
                                     t(d.get("noise1")), out['nearfar'].contiguous(), out.get('z_all'),
                                     {k: t(v) for k, v in cot.items()}, [True] * len(names), [True] * len(names), need_pose,
                                     state=state)
    torch.cuda.synchronize()
    if getattr(cfg, "single_net", False) and g1 is not None:     # one network: the two passes' gradients add up (as autograd does)
        g0, g1 = [a + b for a, b in zip(g0, g1)], None
    grads = {f"net0.{k}": g.cpu().numpy() for k, g in zip(names, g0)}
    if g1 is not None:
        grads.update({f"net1.{k}": g.cpu().numpy() for k, g in zip(names, g1)})
    if need_pose:
        grads["skts"] = g_skts.cpu().numpy()
    return grads, {k: v.cpu().numpy() for k, v in out.items()}


def train_forward(scene, sd0, sd1, cfg, draws):
    """anerf_render_fwd_train with every option of the configuration (the gradient fixtures above use the defaults)."""
    dev = torch.device("cuda")
    t = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(dev)
    N = scene["rays_o"].shape[0]
    fc = cfg.framecode_ch > 0
    plan = _lib.Plan(cfg.n_joints, cfg.D, cfg.W, cfg.skips, cfg.framecode_ch,
                     0 if not fc else sd0['framecodes.codes.weight'].shape[0], 0, view_freqs=cfg.multires_views)
    names = _lib.param_names(cfg.D, fc)
    d0 = {k: t(v) for k, v in sd0.items()}
    d1 = None if sd1 is None else {k: t(v) for k, v in sd1.items()}
    rays = t(np.concatenate([scene["rays_o"], scene["rays_d"], np.zeros((N, 1), np.float32), np.ones((N, 1), np.float32)], 1))
    opts = _lib.make_opts(N, cfg.N_samples, cfg.N_importance, tau_pts=cfg.tau, tau_views=cfg.tau_views, cutoff_pts=cfg.cutoff_dist,
                          cutoff_views=cfg.cutoff_dist, n_joints=cfg.n_joints, single_net=getattr(cfg, "single_net", False),
                          lindisp=cfg.lindisp, softplus=cfg.density_type == "softplus", softplus_shift=cfg.softplus_shift,
                          density_scale=cfg.density_scale)
    state = torch.empty(_lib.train_state_bytes(plan, opts), dtype=torch.uint8, device=dev)
    d = draws or {}
    cams = t(scene["cams"].astype(np.float32)) if fc else None
    out = _lib.render_fwd_train(plan, opts, [d0[k] for k in names], None if d1 is None else [d1[k] for k in names], rays,
                                t(scene["skts"]), t(scene["cyls"]), state, cams, t(d.get("t_rand")), t(d.get("u_rand")),
                                t(d.get("noise0")), t(d.get("noise1")))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("name", [n for n in RENDER_CASES if "fcmean" not in n])      # (the eval-time mean framecode has no training path)
def test_training_forward_matches_reference_golden(name):
    """The forward that keeps its activations (the training forward by default), on every forward fixture of the unmodified
    reference -- lindisp, softplus + shift + density_scale, tau 2000, raw view directions, --single_net, widths 64 / 128,
    depth 6, 1 / 5 / 17 / 24 joints, framecodes, jitter + noise: same bar as the fused kernel (1e-4; the fine pass at its
    own sample positions)."""
    case, gold = load_golden(name)
    scene, sd0, sd1, cfg, draws = build_case(case)
    out = train_forward(scene, sd0, sd1, cfg, draws)
    keys = ["rgb0", "disp0", "acc0", "alpha0"] if cfg.N_importance > 0 else ["rgb_map", "disp_map", "acc_map", "alpha"]
    for k in keys:
        assert gt.rel_err(out[k], gold["ref_" + k]) < 1e-4, (k, gt.rel_err(out[k], gold["ref_" + k]))
    if cfg.N_importance > 0:
        orc_out, _ = run_oracle(scene, sd0, sd1, cfg, draws, z_all_override=out["z_all"])
        for k in ("rgb_map", "disp_map", "acc_map", "alpha"):
            assert gt.rel_err(out[k], orc_out[k]) < 1e-4, (k, gt.rel_err(out[k], orc_out[k]))
        _, taps = run_oracle(scene, sd0, sd1, cfg, draws)
        assert gt.rel_err(out["z_all"], taps["z_all"]) < 2e-3


OPTION_CASES = ["lindisp_j24_s64_i16", "softplus_j24_s64_i16_b2", "tau2000_j24_s64_i16", "single_j24_s96_i48_mv0",
                "w128_d6_j17_s32_i16", "w64_d8_j5_s32_i16", "mixamo_j24_s64_i16_fc"]


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("saved", [False, True])
@pytest.mark.parametrize("name", OPTION_CASES)
def test_gradients_of_the_option_fixtures(name, saved, engine, monkeypatch):
    """Gradients under the options the gradient fixtures do not set -- inverse-depth sampling, softplus density with shift and
    density_scale 2, tau 2000, raw view directions + --single_net + 96/48 samples, width 128 / depth 6 / 17 joints, width 64 /
    5 joints, framecodes at 24 joints -- on 24 rays of the forward fixtures' scenes, both routes, against the fp64 autograd of
    the oracle (pinned to the unmodified reference on exactly these configurations by the forward fixtures).

    fp32 SIMT engine: every tensor within 2e-4 (max-norm) -- the backward LOGIC of every option.  Default tensor-core engine
    (22-bit operands): 24-ray batches put single samples in charge of whole columns, so a hidden unit whose pre-activation
    is within rounding of zero takes its ReLU derivative the other way than fp64 and moves the weight / bias gradients of
    its layer and the one below by up to a few percent (tools/probes/engine_option_probe.py: softplus fixture, fine
    network, layers 0-1 at 6e-3 / 4e-2, every other tensor at 1e-5; the bf16 engine flips the same unit).  Bar per tensor:
    2e-4, or 4x what the oracle's own fp32 autograd deviates from fp64 there (inverse-depth sampling: 1.6e-3); tensor-core
    engine: all gradients together within 2e-3 (L2), at most 6 of the ~45 tensors above the bar, none above 1e-1."""
    monkeypatch.setenv("ANERF_TRAIN_GEMM", engine)
    case, _ = load_golden(name)
    scene, sd0, sd1, cfg, draws = build_case(case)
    keep = slice(0, 24)                                # a few rays are enough; keeps the CPU autograd short
    scene = {k: (v[keep] if isinstance(v, np.ndarray) and v.shape[:1] == scene["rays_o"].shape[:1] else v) for k, v in scene.items()}
    draws = None if draws is None else {k: v[keep] for k, v in draws.items()}
    N = scene["rays_o"].shape[0]
    cot = gt.cotangents(N, cfg.N_samples, cfg.N_importance)
    g, out = gpu_grads(scene, sd0, sd1, cfg, draws, cot, saved=saved)
    _, g64, _ = gt.oracle_grads(scene, sd0, sd1, cfg, draws, cot, dtype=torch.float64, z_all_override=out.get("z_all"))
    _, g32, _ = gt.oracle_grads(scene, sd0, sd1, cfg, draws, cot, z_all_override=out.get("z_all"))
    assert set(g) == set(g64)
    errs = {k: gt.rel_err(g[k], g64[k]) for k in g}
    yard = {k: gt.rel_err(g32[k], g64[k]) for k in g}      # what the oracle's own fp32 autograd deviates by (1.6e-3 with lindisp)
    print(name, engine, "saved" if saved else "recompute", "worst max-norm", max(errs, key=errs.get), max(errs.values()), "oracle fp32:", max(yard.values()))
    above = {k: e for k, e in errs.items() if not (e < max(2e-4, 4 * yard[k]))}
    if engine == "simt":
        assert not above, above
    else:
        fa = np.concatenate([g[k].astype(np.float64).ravel() for k in sorted(g64)])
        fb = np.concatenate([g64[k].ravel() for k in sorted(g64)])
        assert np.linalg.norm(fa - fb) / np.linalg.norm(fb) < 2e-3
        assert len(above) <= 6 and all(e < 1e-1 for e in above.values()), above


@pytest.mark.parametrize("engine", ["tc", "simt", "bf16"])
@pytest.mark.parametrize("name", GRAD_CASES)
def test_backward_matches_oracle_and_reference_autograd(name, engine, monkeypatch):
    """engine: the GEMMs of the backward pass on tensor cores with fp16 hi/lo operands and per-matrix scales ("tc", the
    default), as fp32 SIMT kernels (ANERF_TRAIN_GEMM=simt, the bring-up path that the host emulation also runs), or on
    tensor cores with round 1's bf16 hi/lo operands (ANERF_TRAIN_GEMM=bf16)."""
    monkeypatch.setenv("ANERF_TRAIN_GEMM", engine)
    c, gold = load_golden(name)
    scene, sd0, sd1, cfg, draws = build_case(c)
    N = scene["rays_o"].shape[0]
    cot = gt.cotangents(N, cfg.N_samples, cfg.N_importance)
    g, out = gpu_grads(scene, sd0, sd1, cfg, draws, cot)
    _, g_orc, _ = gt.oracle_grads(scene, sd0, sd1, cfg, draws, cot, z_all_override=out.get("z_all"))
    assert set(g) == set(g_orc)
    if engine in ("simt", "tc"):
        # fp32-grade arithmetic: exactness of the hand-written backward, max-norm per tensor
        errs = {k: gt.rel_err(g[k], g_orc[k]) for k in g}
        print(name, engine, "worst max-norm", max(errs, key=errs.get), max(errs.values()))
        bad = {k: e for k, e in errs.items() if not (e < 2e-4)}
        assert not bad, bad
        errs_ref = {k: gt.digest_err(g[k], {f: gold[f"g|{k}|{f}"] for f in ("sum", "norm", "amax", "idx", "val")}) for k in g}
        bad = {k: e for k, e in errs_ref.items() if not (e < 2e-3)}
        assert not bad, bad
    else:
        # bf16 hi/lo operands carry 16 of fp32's 24 mantissa bits, so the rounding noise of this engine is ~30x
        # fp32's, and these tiny batches amplify rounding strongly (the reference's own fp32 gradients deviate from
        # its fp64 evaluation by 7e-6 / 2e-5 / 6e-5 on the three fixtures; a CPU emulation of the same arithmetic,
        # tests/host -DANERF_EMU_BF16X3, gives L2 errors of 5e-5 / 1.2e-4 / 2.9e-3).  Norm-wise bound per tensor:
        l2 = lambda a, b: float(np.linalg.norm((a.astype(np.float64) - b).ravel()) / max(np.linalg.norm(b.astype(np.float64).ravel()), 1e-30))
        errs = {k: l2(g[k], g_orc[k]) for k in g}
        print(name, engine, "worst L2", max(errs, key=errs.get), max(errs.values()))
        tol = TC_L2_TOL[name]
        bad = {k: e for k, e in errs.items() if not (e < tol)}
        assert not bad, bad


@pytest.mark.parametrize("engine", ["tc", "simt"])
@pytest.mark.parametrize("name", GRAD_CASES)
def test_saved_activation_route(name, engine, monkeypatch):
    """anerf_render_fwd_train + anerf_render_bwd_saved: the forward outputs equal the fused kernel's (both are fp32-grade
    evaluations of the same path: 1e-4; the fine pass at ITS OWN sample positions against the oracle), the gradients hold
    the same 2e-4 against the oracle's autograd as the recomputing route."""
    monkeypatch.setenv("ANERF_TRAIN_GEMM", engine)
    c, gold = load_golden(name)
    scene, sd0, sd1, cfg, draws = build_case(c)
    N = scene["rays_o"].shape[0]
    cot = gt.cotangents(N, cfg.N_samples, cfg.N_importance)
    g, out = gpu_grads(scene, sd0, sd1, cfg, draws, cot, saved=True)
    _, out_fused = gpu_grads(scene, sd0, sd1, cfg, draws, cot, need_pose=False)
    coarse_keys = ("rgb0", "disp0", "acc0", "alpha0") if cfg.N_importance > 0 else ("rgb_map", "disp_map", "acc_map", "alpha")
    for k in coarse_keys + ("nearfar",):
        assert gt.rel_err(out[k], out_fused[k]) < 1e-4, k
    out_orc, g_orc, _ = gt.oracle_grads(scene, sd0, sd1, cfg, draws, cot, z_all_override=out.get("z_all"))
    for k in out_orc:
        assert gt.rel_err(out[k], out_orc[k]) < 1e-4, k
    if cfg.N_importance > 0:
        assert gt.rel_err(out["z_all"], out_fused["z_all"]) < 2e-3      # conditioning of the inverse-CDF step
    assert set(g) == set(g_orc)
    errs = {k: gt.rel_err(g[k], g_orc[k]) for k in g}
    print(name, engine, "saved route, worst max-norm", max(errs, key=errs.get), max(errs.values()))
    bad = {k: e for k, e in errs.items() if not (e < 2e-4)}
    assert not bad, bad


def test_saved_activation_route_at_the_sample_limit():
    """384 + 128 samples per ray (the 512-sample limit of the entry points): the per-ray stage kernel then needs more than
    48 KB of shared memory per CTA (opt-in attribute path).  Outputs against the fused kernel (coarse: 1e-4) and the oracle
    at the route's own fine sample positions."""
    scene = synthetic.make_scene(seed=5, n_rays=24, H=64, W=64, focal=60., n_joints=24)
    sd0, sd1 = synthetic.make_net_weights(101), synthetic.make_net_weights(202)
    from oracle import anerf_oracle as orc
    cfg = orc.PathConfig(n_joints=24, N_samples=384, N_importance=128)
    N = scene["rays_o"].shape[0]
    cot = gt.cotangents(N, cfg.N_samples, cfg.N_importance)
    g, out = gpu_grads(scene, sd0, sd1, cfg, None, cot, saved=True)
    g_rec, out_fused = gpu_grads(scene, sd0, sd1, cfg, None, cot)
    for k in ("rgb0", "disp0", "acc0", "alpha0", "nearfar"):
        assert gt.rel_err(out[k], out_fused[k]) < 1e-4, k
    out_orc, g_orc, _ = gt.oracle_grads(scene, sd0, sd1, cfg, None, cot, z_all_override=out["z_all"])
    for k in out_orc:
        assert gt.rel_err(out[k], out_orc[k]) < 1e-4, k
    # the coarse pass of both routes is the same arithmetic on the same depths: equal up to the order of the atomic adds.
    # (Against the oracle's fp32 autograd this 384-sample configuration is conditioned at ~1e-3 for the coarse network --
    # sample spacings of 1e-3 of the ray -- so that comparison is only a sanity bound here.)
    for k in g:
        if k.startswith("net0."):
            assert gt.rel_err(g[k], g_rec[k]) < 2e-5, k
    errs = {k: gt.rel_err(g[k], g_orc[k]) for k in g}
    assert max(errs.values()) < 5e-3, errs
    assert max(v for k, v in errs.items() if k.startswith("net1.")) < 1e-3, errs          # 512 fine samples: measured 2.7e-4


def test_frozen_parameters_and_no_pose_gradient():
    c, _ = load_golden("grad_j24_s24_i0")
    scene, sd0, sd1, cfg, draws = build_case(c)
    cot = gt.cotangents(scene["rays_o"].shape[0], cfg.N_samples, cfg.N_importance)
    g_full, _ = gpu_grads(scene, sd0, sd1, cfg, draws, cot, need_pose=True)
    g_nop, _ = gpu_grads(scene, sd0, sd1, cfg, draws, cot, need_pose=False)
    assert "skts" not in g_nop
    for k in g_nop:        # weight gradients do not depend on whether the pose gradient was asked for
        assert gt.rel_err(g_nop[k], g_full[k]) < 1e-5, k


def _train_setup(N_importance=16, opt_framecode=False, n_rays=96):
    from tests.test_gpu_api import data_attrs, make_args
    from anerf_b200.raycasters import create_raycaster
    args = make_args(N_importance=N_importance, no_reload=True, perturb=1.0, raw_noise_std=0., opt_framecode=opt_framecode)
    rk_train, rk_test, _, grad_vars, optimizer, _ = create_raycaster(args, data_attrs(24, n_views=4))
    rc = rk_test['ray_caster']
    wk = dict(framecode_ch=16, n_framecodes=4) if opt_framecode else {}
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(101, **wk).items()})
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202, **wk).items()})
    scene = synthetic.make_scene(seed=3, n_rays=n_rays, H=512, W=512, focal=500., n_joints=24)
    dev = torch.device("cuda")
    t = lambda a: torch.as_tensor(a).to(dev)
    N = scene["rays_o"].shape[0]
    rays = torch.cat([t(scene["rays_o"]), t(scene["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev),
                      torch.nn.functional.normalize(t(scene["rays_d"]), dim=-1)], 1)
    kw = {k: v for k, v in rk_train.items() if k not in ('ray_caster', 'use_viewdirs')}
    batch = dict(kp_batch=t(scene["kps"]), skts=t(scene["skts"]), cyls=t(scene["cyls"]), bones=t(scene["bones"]),
                 cams=(torch.arange(N, device=dev) % 4).float() if opt_framecode else None, subject_idxs=None)
    return rk_train['ray_caster'], rc, optimizer, grad_vars, rays, batch, kw


def test_autograd_through_the_python_boundary():
    """loss.backward() through RayCaster in training mode fills .grad of every parameter and of a pose tensor that
    requires grad; an eval-mode call under no_grad (the fused kernel) agrees with the training forward (the layer-wise
    chain that keeps its activations) to the parity tolerance on the same draws (pytest), and bit for bit once the
    saved-activation route is switched off."""
    holder, rc, optimizer, grad_vars, rays, batch, kw = _train_setup(opt_framecode=True)
    holder.train()
    skts = batch['skts'].clone().requires_grad_(True)
    kw = dict(kw, pytest=True)                    # numpy-seeded draws: the same samples in both calls
    out = holder(rays, **dict(batch, skts=skts), **kw)
    target = torch.full_like(out['rgb_map'], 0.5)
    loss = ((out['rgb_map'] - target) ** 2).mean() + 0.5 * ((out['rgb0'] - target) ** 2).mean() + 0.01 * out['acc_map'].mean()
    loss.backward()
    assert skts.grad is not None and torch.isfinite(skts.grad).all() and float(skts.grad.abs().max()) > 0
    assert float(skts.grad[:, :, 3].abs().max()) == 0.          # bottom row of the transforms carries nothing
    for p in grad_vars:
        assert p.grad is not None and torch.isfinite(p.grad).all()
    assert float(rc.network.framecodes.codes.weight.grad.abs().max()) > 0
    assert float(rc.network_fine.pts_linears[0].weight.grad.abs().max()) > 0
    with torch.no_grad():
        again = holder(rays, **batch, **kw)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(out['rgb0'].detach(), again['rgb0']) < 1e-4
    assert rel(out['rgb_map'].detach(), again['rgb_map']) < 1e-3          # fine pass: conditioning of the inverse-CDF step
    rc.keep_activations = False
    try:
        out_fused = holder(rays, **dict(batch, skts=skts), **kw)
        assert torch.equal(again['rgb_map'], out_fused['rgb_map'].detach())
    finally:
        del rc.keep_activations
    # frozen layers get no gradient (--fix_layer semantics)
    optimizer.zero_grad(set_to_none=True)
    for p in rc.network.pts_linears[0].parameters():
        p.requires_grad_(False)
    out = holder(rays, **batch, **kw)
    out['rgb0'].sum().backward()
    assert rc.network.pts_linears[0].weight.grad is None and rc.network.pts_linears[1].weight.grad is not None
    assert rc.network_fine.pts_linears[1].weight.grad is None or float(rc.network_fine.pts_linears[1].weight.grad.abs().max()) == 0.


@pytest.mark.parametrize("single_net", [False, True])
def test_in_place_gradient_accumulation_equals_returned_gradients(single_net):
    """The autograd node adds parameter gradients straight into .grad (one flat allocation, AccumulateGrad bypassed);
    RayCaster.accumulate_param_grads_in_place = False returns them to autograd instead.  Same numbers either way, also with
    --single_net (one parameter set used by both passes), with existing .grad buffers (accumulation over two backward
    calls) and after zero_grad(set_to_none=False)."""
    from tests.test_gpu_api import data_attrs, make_args
    from anerf_b200.raycasters import RayCaster, create_raycaster
    dev = torch.device("cuda")
    args = make_args(N_importance=16, no_reload=True, single_net=single_net)
    got = {}
    for mode in (True, False):
        rk_train, rk_test, _, grad_vars, optimizer, _ = create_raycaster(args, data_attrs(24))
        rc = rk_test['ray_caster']
        rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(101).items()})
        if not single_net:
            rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202).items()})
        assert (rc.network_fine is rc.network) == single_net
        rc.accumulate_param_grads_in_place = mode
        holder = rk_train['ray_caster'].train()
        scene = synthetic.make_scene(seed=3, n_rays=64, H=512, W=512, focal=500., n_joints=24)
        t = lambda a: torch.as_tensor(a).to(dev)
        N = 64
        rays = torch.cat([t(scene["rays_o"]), t(scene["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev),
                          torch.nn.functional.normalize(t(scene["rays_d"]), dim=-1)], 1)
        kw = {k: v for k, v in rk_train.items() if k not in ('ray_caster', 'use_viewdirs')}
        kw.update(perturb=0., raw_noise_std=0.)
        batch = dict(kp_batch=t(scene["kps"]), skts=t(scene["skts"]), cyls=t(scene["cyls"]), bones=t(scene["bones"]), cams=None, subject_idxs=None)
        target = torch.full((N, 3), 0.4, device=dev)
        snaps = []
        for rep in range(3):                     # 1: fresh .grad; 2: accumulate into it; 3: after an in-place zero_grad
            if rep == 2:
                optimizer.zero_grad(set_to_none=False)
            out = holder(rays, **batch, **kw)
            (((out['rgb_map'] - target) ** 2).mean() + ((out['rgb0'] - target) ** 2).mean()).backward()
            snaps.append([p.grad.clone() for p in grad_vars])
        got[mode] = snaps
    RayCaster.accumulate_param_grads_in_place = True
    for a_rep, b_rep in zip(got[True], got[False]):
        for a, b in zip(a_rep, b_rep):
            assert float((a - b).abs().max()) <= 1e-5 * max(float(b.abs().max()), 1e-30)      # fp32 summation order of the two passes
    for g1, g2 in zip(got[True][0], got[True][1]):   # the second backward doubled the gradients
        assert float((g2 - 2 * g1).abs().max()) <= 1e-5 * max(float(g1.abs().max()), 1e-30)


def test_kept_activations_equal_recomputed_and_stale_state_falls_back():
    """Through the Python boundary: (1) gradients of the route that keeps the activations (default) equal those of the
    recomputing route (RayCaster.keep_activations = False) to the engines' tolerance; (2) two forwards before one
    backward: the first call's state has been overwritten, its backward must notice and recompute -- the summed gradient
    equals the sum of the two separate ones."""
    holder, rc, optimizer, grad_vars, rays, batch, kw = _train_setup(N_importance=16, n_rays=192)
    holder.train()
    kw = dict(kw, perturb=0., raw_noise_std=0.)
    N = rays.shape[0]
    tgt = torch.full((N, 3), 0.45, device=rays.device)
    loss_of = lambda out: ((out['rgb_map'] - tgt[:out['rgb_map'].shape[0]]) ** 2).mean() + ((out['rgb0'] - tgt[:out['rgb0'].shape[0]]) ** 2).mean()
    sub = lambda lo, hi: (rays[lo:hi], {k: (v[lo:hi] if torch.is_tensor(v) else v) for k, v in batch.items()})

    def grads(fn):
        optimizer.zero_grad(set_to_none=True)
        fn()
        return [p.grad.clone() for p in grad_vars]

    def one(lo, hi):
        r, b = sub(lo, hi)
        loss_of(holder(r, **b, **kw)).backward()

    def both_then_backward():
        r1, b1 = sub(0, 96)
        r2, b2 = sub(96, 192)
        o1 = holder(r1, **b1, **kw)
        o2 = holder(r2, **b2, **kw)             # same batch shape: re-uses (overwrites) the state buffer of the first call
        (loss_of(o1) + loss_of(o2)).backward()

    g_keep = grads(lambda: one(0, 192))
    assert rc._state_buf[1] is not None and rc._train_state_epoch >= 1
    rc.keep_activations = False
    try:
        g_rec = grads(lambda: one(0, 192))
        g_a, g_b = grads(lambda: one(0, 96)), grads(lambda: one(96, 192))
    finally:
        del rc.keep_activations
    for a, b in zip(g_keep, g_rec):
        assert float((a - b).abs().max()) <= 5e-4 * max(float(b.abs().max()), 1e-30)
    g_two = grads(both_then_backward)
    for t, a, b in zip(g_two, g_a, g_b):
        assert float((t - (a + b)).abs().max()) <= 5e-4 * max(float((a + b).abs().max()), 1e-30)


def test_adam_steps_reduce_the_loss():
    """A few optimizer steps on a fixed batch through the boundary: the photometric loss must go down, and the
    re-packed weights must be what the next forward uses."""
    holder, rc, optimizer, grad_vars, rays, batch, kw = _train_setup(N_importance=16, n_rays=256)
    holder.train()
    torch.manual_seed(0)
    target = torch.rand(rays.shape[0], 3, device=rays.device) * 0.5 + 0.25
    losses = []
    for step in range(12):
        optimizer.zero_grad()
        out = holder(rays, **batch, **dict(kw, perturb=0.))
        loss = ((out['rgb_map'] - target) ** 2).mean() + ((out['rgb0'] - target) ** 2).mean()
        loss.backward()
        optimizer.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses))
    assert losses[-1] < 0.9 * losses[0], losses
