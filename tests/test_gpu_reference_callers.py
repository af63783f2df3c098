""""run_nerf.py and run_render.py call it unchanged", exercised: the reference's OWN callers -- `core.trainer.render`
/ `batchify_rays` (core/trainer.py:64-145) and `Trainer.train_batch` -> `compute_loss` -> `optimize`
(core/trainer.py:237-483) -- are imported unmodified (oracle/ref_import.py; on the GPU box from oracle/_ref) and
handed the render kwargs of `anerf_b200.create_raycaster`.  Results are compared with the golden outputs of the
reference and with the same callers driving the reference's own ray caster on the same GPU."""
import collections
import contextlib
import io

import numpy as np
import pytest
import torch

from anerf_b200 import synthetic
from oracle import ref_import
from tests.common import build_case, load_golden, rel_err
from tests.test_configs import reference_args, CONFIGS

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_import.reference_available(), reason="reference sources not present")]


def _data_attrs(J=24, n_views=1):
    Skel = collections.namedtuple("Skel", ["joint_names", "joint_trees", "root_id"])
    return dict(skel_type=Skel(synthetic.SMPL_JOINT_NAMES[:J], synthetic.SMPL_PARENTS[:J], 0), near=0., far=1., n_views=n_views,
                joint_coords=np.tile(np.eye(3, dtype=np.float32), (1, J, 1, 1)), hwf=(512, 512, 500.))


def _surreal_args(**over):
    path = [p for p in CONFIGS if p.endswith("surreal.txt")][0]
    return reference_args(path, **over)


def _load(rc, sd0, sd1, dev):
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in sd0.items()})
    if sd1 is not None and rc.network_fine is not rc.network:
        rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in sd1.items()})
    return rc.to(dev)


def test_reference_render_drives_our_caster():
    """core.trainer.render (unmodified) -> batchify_rays -> anerf_b200.RayCaster, against the reference's golden outputs."""
    ref_import.import_reference()
    from core.trainer import render
    from anerf_b200.raycasters import create_raycaster
    dev = torch.device("cuda")
    case, gold = load_golden("bench_j24_s64_i128")
    scene, sd0, sd1, cfg, _ = build_case(case)
    args = _surreal_args(N_importance=cfg.N_importance)
    with contextlib.redirect_stdout(io.StringIO()):
        _, rk_test, _, _, _, _ = create_raycaster(args, _data_attrs(), device=dev)
    _load(rk_test["ray_caster"], sd0, sd1, dev).eval()
    t = lambda a: torch.as_tensor(a).to(dev)
    call = lambda rk, chunk: render(scene["H"], scene["W"], scene["focal"], chunk=chunk, rays=(t(scene["rays_o"]), t(scene["rays_d"])),
                                    kp_batch=t(scene["kps"]), skts=t(scene["skts"]), cyls=t(scene["cyls"]), bones=t(scene["bones"]),
                                    cams=None, subject_idxs=None, **rk)
    out = call(rk_test, 4096)                  # the fixture was rendered as one chunk
    for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "alpha0"):
        assert out[k].shape == gold["ref_" + k].shape
        assert rel_err(out[k].cpu().numpy(), gold["ref_" + k]) < 1e-4, k
    # 96 rays in chunks of 40 (ragged last chunk; the near/far repair of rays that miss the cylinder is a CHUNK mean,
    # so the result depends on the chunking): against the reference's own caster on this GPU with the same chunking
    from core.raycasters import create_raycaster as ref_create
    with ref_import.reference_on_cuda():
        with contextlib.redirect_stdout(io.StringIO()):
            _, ref_test, _, _, _, _ = ref_create(args, _data_attrs())
        _load(ref_test["ray_caster"], sd0, sd1, dev).eval()
        with torch.no_grad():
            want = call(ref_test, 40)
    got = call(rk_test, 40)
    for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "alpha0"):
        assert rel_err(got[k].cpu().numpy(), want[k].cpu().numpy()) < 1e-4, k


def _train_batch(N, dev, seed=3):
    sc = synthetic.make_scene(seed=0, n_rays=N, H=512, W=512, focal=500., n_joints=24)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    rng = np.random.RandomState(seed)
    return dict(rays=torch.stack([t(sc["rays_o"]), t(sc["rays_d"])]), cam_idxs=torch.zeros(N, device=dev),
                kp3d=t(sc["kps"]), skts=t(sc["skts"]), bones=t(sc["bones"]), cyls=t(sc["cyls"]),
                target_s=t(rng.rand(N, 3).astype(np.float32)), bgs=t(rng.rand(N, 3).astype(np.float32)))


def test_reference_trainer_step_drives_our_caster():
    """Trainer.train_batch (unmodified: render -> compute_loss -> loss.backward -> Adam step -> lrate decay ->
    update_embed_fns) once with our render kwargs and once with the reference's own, same weights and batch."""
    ref_import.import_reference()
    from core.trainer import Trainer
    from core.raycasters import create_raycaster as ref_create
    from anerf_b200.raycasters import create_raycaster
    dev = torch.device("cuda")
    sd0, sd1 = synthetic.make_net_weights(101), synthetic.make_net_weights(202)
    N = 256
    # deterministic sampling so that both runs see the same samples (the random draws come from different generators)
    args = _surreal_args(perturb=0., raw_noise_std=0., debug=True)
    results = {}
    for who in ("ours", "reference"):
        batch = _train_batch(N, dev)
        # the reference's trainer module itself builds tensors with the legacy constructors (core/trainer.py:8), so it
        # runs the way run_nerf.py runs it: CUDA as the default tensor type -- for both casters
        with contextlib.redirect_stdout(io.StringIO()), ref_import.reference_on_cuda():
            if who == "ours":
                rk_train, rk_test, _, grad_vars, optimizer, _ = create_raycaster(args, _data_attrs(), device=dev)
            else:
                rk_train, rk_test, _, grad_vars, optimizer, _ = ref_create(args, _data_attrs())
            _load(rk_test["ray_caster"], sd0, sd1, dev)
            rk_train["ray_caster"].train()
            trainer = Trainer(args, _data_attrs(), optimizer, None, rk_train, rk_test, popt_kwargs=None, device=dev)
            loss, stats = trainer.train_batch(batch, i=1, global_step=1)
        rc = rk_test["ray_caster"]
        results[who] = dict(loss={k: float(v) for k, v in loss.items()}, stats=stats,
                            w={k: v.detach().float().cpu().numpy() for k, v in rc.network_fine.state_dict().items()},
                            tau=rc.embed_fn.get_tau())
    a, b = results["ours"], results["reference"]
    for k in b["loss"]:
        assert abs(a["loss"][k] - b["loss"][k]) < 1e-5 * max(1.0, abs(b["loss"][k])), (k, a["loss"][k], b["loss"][k])
    for k in ("psnr", "psnr0", "alpha", "lrate", "cutoff"):
        assert abs(a["stats"][k] - b["stats"][k]) < 1e-4 * max(1.0, abs(b["stats"][k])), (k, a["stats"][k], b["stats"][k])
    # gradient norm over all parameters as Trainer.optimize reports it (get_gradnorm)
    assert abs(a["stats"]["total_norm"] - b["stats"]["total_norm"]) < 2e-3 * b["stats"]["total_norm"], (a["stats"]["total_norm"], b["stats"]["total_norm"])
    assert a["tau"] == pytest.approx(b["tau"], rel=1e-6)
    # one Adam step moved the weights the same way (step size = lrate per element on the first step)
    moved = same = 0
    for k in b["w"]:
        d_ref = b["w"][k] - np.asarray(sd1[k])
        d_our = a["w"][k] - np.asarray(sd1[k])
        big = np.abs(d_ref) > 0.5 * args.lrate           # elements whose gradient is well above Adam's eps
        moved += int(big.sum())
        same += int((np.sign(d_our[big]) == np.sign(d_ref[big])).sum())
        assert (np.sign(d_our[big]) == np.sign(d_ref[big])).mean() > 0.98, k
    assert moved > 100000 and same / moved > 0.999, (moved, same)
