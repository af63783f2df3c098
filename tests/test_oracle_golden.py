"""The oracle must reproduce the reference's own outputs stored in tests/golden (CPU, any box)."""
import numpy as np
import pytest
import torch

from tests.common import RENDER_CASES, build_case, load_golden, rel_err, run_oracle
from oracle import anerf_oracle as orc
from anerf_b200 import synthetic

# maps: 2e-5 (fp32 re-association only).  per-sample alpha after importance sampling: 2e-3, because the
# inverse-CDF step amplifies 1e-7 differences in the cdf by up to 1/1e-5 (measured: the reference differs
# from its own fp64 evaluation by 5e-4..7e-4 on `alpha`, see oracle/make_golden.py output).
TOL = dict(rgb_map=2e-5, disp_map=2e-5, acc_map=2e-5, rgb0=2e-5, disp0=2e-5, acc0=2e-5, alpha0=2e-5, alpha=2e-3)


@pytest.mark.parametrize("name", RENDER_CASES)
def test_oracle_reproduces_reference_outputs(name):
    case, gold = load_golden(name)
    scene, sd0, sd1, cfg, draws = build_case(case)
    out, taps = run_oracle(scene, sd0, sd1, cfg, draws)
    for k, tol in TOL.items():
        if "ref_" + k in gold:
            if k == "alpha" and cfg.N_importance == 0:
                tol = 2e-5
            assert rel_err(out[k], gold["ref_" + k]) < tol, k
    # the fixtures are not vacuous: densities are a mix of empty and occupied
    acc = gold["ref_acc_map"]
    if cfg.density_type == "softplus":      # softplus density is never exactly 0 and the last interval is 1e10 long: acc == 1
        assert gold["ref_alpha0"][:, :-1].mean() < 0.5 and gold["ref_alpha0"].max() > 0.5
    else:
        assert 0.01 < acc.mean() < 0.95 and acc.max() > 0.5


def test_headline_fixture_is_well_formed():
    """tests/golden/bench4096_*: reference outputs (fp32), the fp64 evaluation and the per-ray conditioning that
    tests/test_gpu_parity.py::test_headline_config_4096_rays relies on."""
    from tests.common import BIG_CASE
    case, gold = load_golden(BIG_CASE)
    assert case["n_rays"] == 4096 and case["N_samples"] == 64 and case["N_importance"] == 128
    for k in ("rgb_map", "disp_map", "acc_map"):
        ref, ref64 = gold["ref_" + k].astype(np.float64), gold["ref64_" + k].astype(np.float64)
        cond = np.abs(ref - ref64).reshape(4096, -1).max(1) / np.abs(ref).max()
        assert (cond > 5e-5).sum() <= 4            # the rays the reference's own fp32 arithmetic cannot resolve
        assert np.median(cond) < 1e-6
    assert gold["tap_z_all"].shape == (4096, 192) and (np.diff(gold["tap_z_all"], axis=1) >= 0).all()
    assert 0.2 < gold["ref_acc_map"].mean() < 0.8


def test_oracle_reproduces_reference_density_grid():
    case, gold = load_golden("mesh_j24_res15")
    J = case["n_joints"]
    pose = synthetic.make_pose(11, J)
    sd = synthetic.make_net_weights(202, n_joints=J, D=case["D"], W=case["W_net"], skips=case["skips"])
    cfg = orc.PathConfig(n_joints=J, D=case["D"], W=case["W_net"], skips=case["skips"])
    t = torch.as_tensor
    with torch.no_grad():
        sig = orc.density_grid(orc.to_torch(sd), cfg, t(pose["kps"]), t(pose["skts"]), case["radius"], case["res"])
    assert rel_err(sig.numpy(), gold["ref_sigma"]) < 2e-5
    assert 0.05 < float((gold["ref_sigma"] > 0).mean()) < 0.95


def test_flop_count_matches_baseline():
    # 441.25 MFLOP per ray at 24 joints, 64+128 samples, 8x256
    assert orc.algorithmic_flops_per_ray(orc.PathConfig()) == 256 * 1723648


@pytest.mark.parametrize("name", ["grad_cfg1_j1_s16_i16", "grad_j24_s24_i0", "grad_j24_s16_i8_fc_perturb", "grad_single_j24_s16_i8"])
def test_oracle_autograd_matches_reference_gradient_digests(name):
    """The oracle's gradients (torch autograd over the restatement) against the digests of the reference's own
    autograd on the same rays, weights, draws and output cotangents (oracle/make_golden_grad.py)."""
    from oracle import grad_tools as gt
    c, gold = load_golden(name)
    scene, sd0, sd1, cfg, draws = build_case(c)
    cot = gt.cotangents(scene["rays_o"].shape[0], cfg.N_samples, cfg.N_importance)
    out, g, _ = gt.oracle_grads(scene, sd0, sd1, cfg, draws, cot)
    for k in out:
        assert rel_err(out[k], gold["ref_" + k]) < (2e-3 if k == "alpha" else 2e-5), k
    for k in g:
        dg = {f: gold[f"g|{k}|{f}"] for f in ("sum", "norm", "amax", "idx", "val")}
        assert gt.digest_err(g[k], dg) < 2e-5, k


def test_frame_glue_matches_reference_boxes():
    """anerf_b200.frames.valid_pixels against the reference's cylinder_to_box_2d / kp_to_valid_rays results stored by
    oracle/make_golden_frames.py (integer boxes and pixel index lists: exact)."""
    import ast
    import os
    from anerf_b200 import frames
    from tests.common import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "frames_box2d.npz"))
    cases = ast.literal_eval(str(z["cases"]))
    for i, c in enumerate(cases):
        idx, (tl, br) = frames.valid_pixels(z["cyl"], c["H"], c["W"], c["focal"], z[f"{i}|c2w"])
        assert np.array_equal(tl, z[f"{i}|tl"]) and np.array_equal(br, z[f"{i}|br"]), i
        assert len(idx) == int(z[f"{i}|n_valid"])
        assert np.array_equal(idx[:16].numpy(), z[f"{i}|valid_head"]) and np.array_equal(idx[-16:].numpy(), z[f"{i}|valid_tail"])
