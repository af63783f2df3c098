"""tests/golden/pose_chain.npz: outputs and autograd gradients of the UNMODIFIED reference's PoseOptLayer
(core/pose_opt.py:240-445: calculate_kinematic + unrolled_kinematic_chain + torch.inverse) on seeded inputs, for the
fused pose chain of SURVEY.md 8(f) row 2.  TEST INFRASTRUCTURE; needs the reference sources (build container).
    python oracle/make_golden_pose.py [--check]
pytorch3d is not installed: PoseOptLayer.__init__ converts the initial axis-angle bones through
pytorch3d.axis_angle_to_matrix, which is stubbed with a Rodrigues formula here; the parameters are then OVERWRITTEN with
seeded 6-D rotations, so nothing of that third-party function reaches the fixture (the chain itself is pure torch)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "pose_chain.npz")


def rodrigues(v):
    theta = v.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    k = v / theta
    K = torch.zeros(*v.shape[:-1], 3, 3, dtype=v.dtype)
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = torch.sin(theta)[..., None], torch.cos(theta)[..., None]
    return torch.eye(3, dtype=v.dtype) + s * K + (1 - c) * (K @ K)


def inputs(P=6, N=40, J=24):
    rng = np.random.RandomState(21)
    bones6 = rng.randn(P, J, 6).astype(np.float32)            # any 6-D vector is a valid rotation parameter
    pelvis = (rng.randn(P, 3) * 0.3).astype(np.float32)
    rest = (rng.randn(1, J, 3) * 0.3).astype(np.float32)
    idxs = rng.randint(0, P, size=N)
    idxs[:3] = [3, 0, 3]                                       # repeated poses: the per-ray -> per-pose reduction
    cot = dict(kps=rng.randn(N, J, 3).astype(np.float32), skts=rng.randn(N, J, 4, 4).astype(np.float32),
               l2ws=rng.randn(N, J, 4, 4).astype(np.float32))
    return bones6, pelvis, rest, idxs, cot


def run_reference():
    core = ref_import.import_reference()
    import pytorch3d.transforms.rotation_conversions as p3dr
    p3dr.axis_angle_to_matrix = rodrigues
    import core.pose_opt as po
    from core.utils.skeleton_utils import SMPLSkeleton
    bones6, pelvis, rest, idxs, cot = inputs()
    P, J = bones6.shape[:2]
    layer = po.PoseOptLayer(torch.zeros(P, J, 3), torch.zeros(P, J, 3) + 0.1, torch.as_tensor(rest), use_rot6d=True)
    with torch.no_grad():
        layer.bones.copy_(torch.as_tensor(bones6))
        layer.pelvis.copy_(torch.as_tensor(pelvis))
    kps, bone, skts, l2ws, rots = layer(idxs)
    loss = (kps * torch.as_tensor(cot["kps"])).sum() + (skts * torch.as_tensor(cot["skts"])).sum() + (l2ws * torch.as_tensor(cot["l2ws"])).sum()
    g_bones, g_pelvis = torch.autograd.grad(loss, [layer.bones, layer.pelvis])
    return dict(parents=np.asarray(SMPLSkeleton.joint_trees, np.int32), root_id=np.int32(SMPLSkeleton.root_id),
                bones6=bones6, pelvis=pelvis, rest=rest, idxs=idxs.astype(np.int64),
                ref_kps=kps.detach().numpy(), ref_skts=skts.detach().numpy(), ref_l2ws=l2ws.detach().numpy(),
                ref_rots=rots.detach().numpy(), ref_g_bones=g_bones.numpy(), ref_g_pelvis=g_pelvis.numpy())


if __name__ == "__main__":
    out = run_reference()
    if "--check" in sys.argv:
        z = np.load(OUT)
        for k in out:
            assert np.allclose(z[k], out[k], rtol=1e-6, atol=1e-7), k
        print("fixture matches the live reference")
    else:
        np.savez_compressed(OUT, **out)
        print("wrote", OUT, {k: np.asarray(v).shape for k, v in out.items()})
