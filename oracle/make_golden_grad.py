"""Gradient fixtures: run the UNMODIFIED reference (from /root/reference) in TRAINING mode with autograd
on synthetic rays, fixed random draws and fixed output cotangents, and pin the oracle's gradients
(oracle/grad_tools.py) against it.  TEST INFRASTRUCTURE; build container only.

    python oracle/make_golden_grad.py            # rewrite tests/golden/grad_*.npz
    python oracle/make_golden_grad.py --check    # compare only

A fixture stores the case config (inputs are regenerated from seeds), the reference's outputs and a digest
(sum, norm, max, 1024 sampled entries) of the reference's gradient of every parameter and of `skts`.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import grad_tools as gt  # noqa: E402
from oracle import make_golden as mg  # noqa: E402
from oracle import ref_import  # noqa: E402

GRAD_CASES = {
    # 24 joints, 8x256, framecodes, stratified jitter + random importance draws + density noise (training sampling)
    "grad_j24_s16_i8_fc_perturb": dict(n_joints=24, n_rays=12, H=512, W=512, focal=500., D=8, W_net=256, skips=(4,),
                                       N_samples=16, N_importance=8, framecode_ch=16, n_framecodes=3, perturb=True),
    # BASELINE.json configs[0] shape: 1 joint, 4x64 (no skip), deterministic sampling
    "grad_cfg1_j1_s16_i16": dict(n_joints=1, n_rays=32, H=32, W=32, focal=40., D=4, W_net=64, skips=(4,),
                                 N_samples=16, N_importance=16),
    # --single_net: both passes through one network (gradients of the two passes add up), blurred importance pdf
    "grad_single_j24_s16_i8": dict(n_joints=24, n_rays=10, H=512, W=512, focal=500., D=8, W_net=256, skips=(4,),
                                   N_samples=16, N_importance=8, single_net=True),
    # coarse pass only
    "grad_j24_s24_i0": dict(n_joints=24, n_rays=8, H=512, W=512, focal=500., D=8, W_net=256, skips=(4,),
                            N_samples=24, N_importance=0),
}


def build(c):
    scene, sd0, sd1, cfg = mg.build_case(c)
    N = scene["rays_o"].shape[0]
    if cfg.framecode_ch > 0:
        scene["cams"] = (np.arange(N) % c["n_framecodes"]).astype(np.int64)
    draws = mg.make_draws(c, cfg, N)
    cot = gt.cotangents(N, cfg.N_samples, cfg.N_importance)
    return scene, sd0, sd1, cfg, draws, cot


def reference_grads(core, c, scene, sd0, sd1, cfg, draws, cot):
    """The reference's own training-mode forward (core/trainer.py:82 render -> batchify_rays -> RayCaster.forward
    -> render_rays) and torch autograd."""
    from core.trainer import render
    rc, rk = mg.make_reference_caster(core, c, cfg, sd0, sd1)
    rc.train()
    t = lambda a: torch.as_tensor(a)
    kw = {k: v for k, v in rk.items()}
    cams = torch.as_tensor(scene["cams"]).float() if cfg.framecode_ch > 0 else None
    skts = t(scene["skts"]).clone().requires_grad_(True)
    orig_rand, orig_randn = torch.rand, torch.randn
    if draws is not None:
        kw["perturb"], kw["raw_noise_std"] = 1.0, 1.0
        seq_rand = [draws["t_rand"], draws["u_rand"]]
        seq_randn = [draws["noise0"], draws["noise1"]]
        torch.rand = lambda *a, **k: torch.as_tensor(seq_rand.pop(0))
        torch.randn = lambda *a, **k: torch.as_tensor(seq_randn.pop(0))
    try:
        out = render(scene["H"], scene["W"], scene["focal"], chunk=4096, rays=(t(scene["rays_o"]), t(scene["rays_d"])),
                     kp_batch=t(scene["kps"]), skts=skts, cyls=t(scene["cyls"]), bones=t(scene["bones"]), cams=cams,
                     subject_idxs=None, **kw)
    finally:
        torch.rand, torch.randn = orig_rand, orig_randn
    out = {k: v for k, v in out.items() if k in gt.OUT_KEYS}
    loss = sum((out[k] * t(cot[k])).sum() for k in out)
    loss.backward()
    grads = {}
    for tag, net in (("net0", rc.network), ("net1", rc.network_fine)):
        if net is None or (tag == "net1" and net is rc.network):
            continue
        for k, p in net.named_parameters():
            grads[f"{tag}.{k}"] = (torch.zeros(1) if p.grad is None else p.grad).numpy()
    grads["skts"] = skts.grad.numpy()
    return {k: v.detach().numpy() for k, v in out.items()}, grads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    opt = ap.parse_args()
    core = ref_import.import_reference()
    worst = 0.0
    for name, c in GRAD_CASES.items():
        scene, sd0, sd1, cfg, draws, cot = build(c)
        ref_out, ref_g = reference_grads(core, c, scene, sd0, sd1, cfg, draws, cot)
        out, g, _ = gt.oracle_grads(scene, sd0, sd1, cfg, draws, cot)
        _, g64, _ = gt.oracle_grads(scene, sd0, sd1, cfg, draws, cot, torch.float64)
        assert set(g) == set(ref_g), (sorted(set(g) ^ set(ref_g)))
        errs = {k: gt.rel_err(g[k], ref_g[k]) for k in ref_g}
        errs64 = {k: gt.rel_err(ref_g[k], g64[k]) for k in ref_g}
        kmax = max(errs, key=errs.get)
        print(f"[{name}] N={scene['rays_o'].shape[0]} acc_mean={ref_out['acc_map'].mean():.3f} params={len(ref_g)} "
              f"worst oracle-vs-ref {errs[kmax]:.1e} ({kmax}); ref-vs-fp64 worst {max(errs64.values()):.1e}; "
              f"|g skts|={np.abs(ref_g['skts']).max():.2e}")
        worst = max(worst, errs[kmax])
        if not opt.check:
            save = {f"ref_{k}": v for k, v in ref_out.items()}
            for k, v in ref_g.items():
                for f, x in gt.digest(v).items():
                    save[f"g|{k}|{f}"] = np.asarray(x)
            save["case"] = np.array(repr(c))
            np.savez_compressed(os.path.join(mg.GOLDEN_DIR, name + ".npz"), **save)
    print(f"worst oracle-vs-reference gradient error: {worst:.2e}")
    assert worst < 2e-4, "oracle gradients drifted from the reference's autograd"


if __name__ == "__main__":
    main()
