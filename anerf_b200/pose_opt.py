"""the pose chain of pose-refinement training behind the reference's `PoseOptLayer`
interface (core/pose_opt.py:240-445), fused with the reduction of the renderer's d/d skts.

* `pose_chain(rots, rest_pose, pelvis, parents)` -- one autograd node over the C ABI
  (`anerf_pose_chain_fwd` / `anerf_pose_chain_bwd`): the 24-link chain of 4x4 products, the pelvis shift and the
  inverse (closed form for a rigid transform; the reference calls LU `torch.inverse`), forward and backward in one
  kernel each instead of ~60 small batched-matmul / cat / inverse launches.
* `PoseOptLayer` -- same constructor, parameters (`pelvis`, `bones` as axis-angle or 6-D rotations), state_dict and
  `forward(idxs) -> (kps, bones, skts, l2ws, rots)` as the reference, so `core.trainer.Trainer.get_kp_args` calls it
  unchanged.  `forward_poses(idxs)` is the fused form: it returns the transforms once per UNIQUE pose together with the
  ray -> pose index; `RayCaster(..., skts=skts_pose, pose_idx=idx)` then reads the pose through the index and its backward
  adds every ray's d/d skt straight into the per-pose gradient (atomics), so neither the [N,24,4,4] per-ray copy of the
  transforms nor its gradient is ever materialised (the reference: `skts[inverse_idxs]` + index backward,
  core/pose_opt.py:438-441).

Rotation parametrisations are plain torch: `rot6d_to_rotmat` follows core/utils/skeleton_utils.py:420-436; axis-angle
uses the Rodrigues formula (the reference delegates to pytorch3d's axis_angle_to_matrix, which is not installed here --
parity at that third-party boundary is unpinned; every shipped pose-refinement config uses opt_rot6d).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def rot6d_to_rotmat(x):
    """[...,6] -> [...,3,3]  (Zhou et al. 2019; core/utils/skeleton_utils.py:420-436: columns b1, b2, b1 x b2)."""
    sh = x.shape[:-1]
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-1).reshape(*sh, 3, 3)


def axisang_to_rot(v):
    """[...,3] axis-angle -> [...,3,3] (Rodrigues)."""
    theta = v.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    k = v / theta
    K = torch.zeros(*v.shape[:-1], 3, 3, dtype=v.dtype, device=v.device)
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = torch.sin(theta)[..., None], torch.cos(theta)[..., None]
    return torch.eye(3, dtype=v.dtype, device=v.device) + s * K + (1 - c) * (K @ K)


class _PoseChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rots, rest_pose, pelvis, parents, root_id):
        rots_c, rest_c, pel_c = rots.float().contiguous(), rest_pose.float().contiguous(), pelvis.float().contiguous()
        with torch.cuda.device(rots.device):
            l2ws, skts, kps = _lib.pose_chain_fwd(rots_c, rest_c, pel_c, parents, root_id)
        ctx.save_for_backward(rots_c, rest_c, pel_c, l2ws, skts)
        ctx.parents, ctx.root_id = parents, root_id
        return l2ws, skts, kps

    @staticmethod
    def backward(ctx, g_l2ws, g_skts, g_kps):
        rots, rest, pelvis, l2ws, skts = ctx.saved_tensors
        c = lambda g: None if g is None else g.float().contiguous()
        with torch.cuda.device(rots.device):
            g_rots, g_pelvis = _lib.pose_chain_bwd(rots, rest, pelvis, ctx.parents, ctx.root_id, l2ws, skts,
                                                   g_skts=c(g_skts), g_l2ws=c(g_l2ws), g_kps=c(g_kps))
        return g_rots, None, g_pelvis, None, None


def pose_chain(rots, rest_pose, pelvis, parents, root_id=0):
    """rots [P,J,3,3], rest_pose [1|P,J,3], pelvis [P,3] (CUDA) -> (l2ws [P,J,4,4], skts [P,J,4,4], kps [P,J,3]);
    differentiable w.r.t. rots and pelvis (core/pose_opt.py:395-445)."""
    if rots.device.type != 'cuda':
        raise RuntimeError("anerf_b200.pose_chain: tensors must be on a CUDA device (no CPU path)")
    return _PoseChainFn.apply(rots, rest_pose, pelvis, tuple(int(p) for p in parents), int(root_id))


class PoseOptLayer(nn.Module):
    """Per-frame pose parameters + the kinematic chain (reference: core/pose_opt.py:240-445; single-view layout, i.e.
    `kp_map is None`; the multi-view parameter sharing and the cache are not implemented and raise)."""

    def __init__(self, kps, bones, rest_pose, skel_type=None, kp_map=None, kp_uidxs=None, use_cache=False, use_rot6d=False,
                 beta=None, rest_pose_idxs=None, parents=None, root_id=0):
        super().__init__()
        if kp_map is not None or use_cache:
            raise NotImplementedError("anerf_b200.PoseOptLayer: multi-view parameter sharing (kp_map) / use_cache are not implemented")
        if parents is None:
            parents, root_id = skel_type.joint_trees, skel_type.root_id
        self.parents, self.root_id = tuple(int(p) for p in parents), int(root_id)
        self.skel_type, self.use_rot6d, self.use_cache = skel_type, use_rot6d, False
        self.kp_map = self.kp_uidxs = None
        self.rest_pose_idxs = rest_pose_idxs
        self.beta = torch.tensor(beta, requires_grad=False) if beta is not None else None
        kps, bones = torch.as_tensor(kps).float(), torch.as_tensor(bones).float()
        self.register_buffer('rest_pose', torch.as_tensor(rest_pose).float().clone())
        self.register_parameter('pelvis', nn.Parameter(kps[:, self.root_id].clone(), requires_grad=True))
        if use_rot6d:
            NJ = bones.shape[1]
            bones = axisang_to_rot(bones.reshape(-1, 3)).reshape(-1, NJ, 3, 3)[..., :3, :2].reshape(-1, NJ, 6)
        self.register_parameter('bones', nn.Parameter(bones.clone(), requires_grad=True))
        self.N_kps = self.pelvis.shape[0]

    def idx_to_params(self, idx):
        return self.pelvis[idx], self.bones[idx]

    def get_pelvis(self, idx=None):
        return self.pelvis if idx is None else self.pelvis[idx]

    def get_rest_pose(self, kp_idxs=None, rest_pose_idxs=None):
        if len(self.rest_pose) == 1:
            return self.rest_pose
        if rest_pose_idxs is not None:
            return self.rest_pose[rest_pose_idxs]
        return self.rest_pose[self.rest_pose_idxs[kp_idxs]]

    def _rots(self, bone):
        N, NJ = bone.shape[:2]
        if self.use_rot6d:
            return rot6d_to_rotmat(bone.reshape(-1, 6)).reshape(N, NJ, 3, 3)
        return axisang_to_rot(bone.reshape(-1, 3)).reshape(N, NJ, 3, 3)

    def forward_poses(self, idxs, rest_pose_idxs=None):
        """Fused form: (kps, bones, skts, l2ws, rots) once per UNIQUE pose + `pose_idx` int32 [len(idxs)] mapping every
        requested row to its pose."""
        if idxs is None:
            idxs = np.arange(len(self.pelvis))
        if torch.is_tensor(idxs):
            idxs = idxs.cpu().numpy()
        idxs = np.atleast_1d(np.asarray(idxs))
        unique_idxs, inverse_idxs = np.unique(idxs, return_inverse=True)
        dev = self.pelvis.device
        # index uploads through pinned memory, non-blocking: a pageable host->device copy would stall the host until the
        # previous step's kernels have drained and put the launch overhead of this step on the critical path
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).pin_memory().to(dev, non_blocking=True)
        uidx = up(unique_idxs, np.int64) if dev.type == 'cuda' else torch.from_numpy(np.ascontiguousarray(unique_idxs, dtype=np.int64))
        pelvis, bone = self.pelvis.index_select(0, uidx), self.bones.index_select(0, uidx)
        rots = self._rots(bone)
        rest = self.get_rest_pose(unique_idxs, rest_pose_idxs)
        l2ws, skts, kps = pose_chain(rots, rest, pelvis, self.parents, self.root_id)
        pose_idx = up(inverse_idxs.reshape(-1), np.int32)
        return (kps, bone, skts, l2ws, rots), pose_idx

    def calculate_kinematic(self, idxs, rest_pose_idxs=None):
        """The reference's return convention: everything gathered back to the requested (possibly repeated) rows."""
        (kps, bone, skts, l2ws, rots), pose_idx = self.forward_poses(idxs, rest_pose_idxs)
        i = pose_idx.long()
        return kps[i], bone[i], skts[i], l2ws[i], rots[i]

    def forward(self, idxs, rest_pose_idxs=None):
        return self.calculate_kinematic(idxs, rest_pose_idxs)
