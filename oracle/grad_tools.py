"""Gradient side of the oracle (TEST INFRASTRUCTURE, same rules as anerf_oracle.py).

The oracle is plain torch code, so its gradients are torch autograd over the restated forward; what is
pinned against the reference is therefore the reference's own autograd on identical rays, weights,
random draws and output cotangents (`oracle/make_golden_grad.py`, `tests/golden/grad_*.npz`).

    loss = sum_k <out[k], cot[k]>     for every entry of the dict render_rays returns
    grads: every parameter of both networks (incl. framecodes) and the per-ray bone transforms `skts`.
"""
import numpy as np
import torch

from oracle import anerf_oracle as orc

OUT_KEYS = ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0", "alpha0")


def cotangents(N, Sc, Si, seed=9):
    """Fixed dL/d(outputs): O(1) on the colour / opacity maps, smaller on disparity and per-sample alpha."""
    rng = np.random.RandomState(seed)
    S = Sc + Si if Si > 0 else Sc
    cot = dict(rgb_map=rng.randn(N, 3), disp_map=0.05 * rng.randn(N), acc_map=rng.randn(N), alpha=0.1 * rng.randn(N, S))
    if Si > 0:
        cot.update(rgb0=rng.randn(N, 3), disp0=0.05 * rng.randn(N), acc0=rng.randn(N), alpha0=0.1 * rng.randn(N, Sc))
    return {k: v.astype(np.float32) for k, v in cot.items()}


def oracle_grads(scene, sd0, sd1, cfg, draws, cot, dtype=torch.float32, z_all_override=None):
    """-> (outputs, grads) as numpy; grads keys: 'net0.<param>', 'net1.<param>', 'skts'."""
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)
    d = {k: t(v) for k, v in (draws or {}).items()}
    p0 = {k: v.requires_grad_(True) for k, v in orc.to_torch(sd0, dtype).items()}
    p1 = None if (sd1 is None or cfg.single_net) else {k: v.requires_grad_(True) for k, v in orc.to_torch(sd1, dtype).items()}
    skts = t(scene["skts"]).requires_grad_(True)
    cams = torch.as_tensor(scene["cams"]) if cfg.framecode_ch > 0 else None
    taps = {}
    out = orc.render_rays(p0, p1, cfg, t(scene["rays_o"]), t(scene["rays_d"]), skts, t(scene["cyls"]), cams=cams,
                          t_rand=d.get("t_rand"), u_rand=d.get("u_rand"), noise0=d.get("noise0"), noise1=d.get("noise1"),
                          training=True, taps=taps,
                          z_all_override=None if z_all_override is None else t(z_all_override))
    loss = sum((out[k] * t(cot[k])).sum() for k in out)
    loss.backward()
    grads = {f"net0.{k}": v.grad for k, v in p0.items()}
    if p1 is not None:
        grads.update({f"net1.{k}": v.grad for k, v in p1.items()})
    grads["skts"] = skts.grad
    grads = {k: (torch.zeros(1) if v is None else v).detach().numpy() for k, v in grads.items()}
    return ({k: v.detach().numpy() for k, v in out.items()}, grads,
            {k: v.detach().numpy() for k, v in taps.items()})


def digest(a, n_sample=1024):
    """Compact fingerprint of a gradient tensor for the committed fixtures: sum, L2 norm, max |.| and the values
    at fixed pseudo-random flat positions (full tensor when it is small)."""
    a = np.asarray(a, np.float64).reshape(-1)
    if a.size <= 4 * n_sample:
        idx = np.arange(a.size)
    else:
        idx = np.sort(np.random.RandomState(a.size % 9973).choice(a.size, n_sample, replace=False))
    return dict(sum=float(a.sum()), norm=float(np.sqrt((a * a).sum())), amax=float(np.abs(a).max()),
                idx=idx.astype(np.int64), val=a[idx].astype(np.float32))


def digest_err(a, dg):
    """Relative deviation of tensor `a` from a stored digest: sampled values against the tensor's max, and the norm."""
    a = np.asarray(a, np.float64).reshape(-1)
    scale = max(float(dg["amax"]), 1e-30)
    e_val = float(np.abs(a[dg["idx"]] - dg["val"].astype(np.float64)).max() / scale)
    e_norm = abs(float(np.sqrt((a * a).sum())) - float(dg["norm"])) / max(float(dg["norm"]), 1e-30)
    return max(e_val, e_norm)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
