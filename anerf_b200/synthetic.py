"""Synthetic inputs for the hot path: skeleton poses, cameras, rays and network weights.

There is no dataset or checkpoint in this environment, so tests, `bench.py` and
`__graft_entry__.smoke()` draw their inputs from here.  Everything is numpy with
`np.random.RandomState`, so the same seed gives the same bytes on every box (torch's CPU
generator is not used on purpose).

What the generated tensors mean follows the reference's data contract:
  * skts  [J,4,4]  world -> bone-local transforms = inverse of the local-to-world chain
                   (reference: core/utils/skeleton_utils.py:334-376 builds l2w, run_render.py:762-765 inverts)
  * kps   [J,3]    joint locations = l2w[:, :3, 3]
  * cyls  [5]      (cx, cz, radius, top, bottom) bounding cylinder, ground plane x-z
                   (reference: core/utils/skeleton_utils.py:542-592, head='-y')
  * rays  o,d      pinhole camera looking down -z (reference: core/utils/ray_utils.py:6-28)
"""
import numpy as np

# SMPL kinematic tree (parent of each of the 24 joints); joint 0 (pelvis) is the root.
SMPL_PARENTS = np.array([0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21])

SMPL_JOINT_NAMES = [
    'pelvis', 'left_hip', 'right_hip', 'spine1', 'left_knee', 'right_knee', 'spine2', 'left_ankle',
    'right_ankle', 'spine3', 'left_foot', 'right_foot', 'neck', 'left_collar', 'right_collar', 'head',
    'left_shoulder', 'right_shoulder', 'left_elbow', 'right_elbow', 'left_wrist', 'right_wrist',
    'left_hand', 'right_hand']


def humanoid_rest_pose():
    """A 24-joint T-pose in metres (y up, x to the figure's left), pelvis at the origin.

    Proportions of a ~1.7 m adult; these are our own round numbers, not the SMPL template."""
    P = np.zeros((24, 3), dtype=np.float32)
    P[0] = (0.00, 0.00, 0.00)
    P[1] = (0.09, -0.09, 0.00); P[2] = (-0.09, -0.09, 0.00)      # hips
    P[3] = (0.00, 0.12, -0.02)                                   # spine1
    P[4] = (0.10, -0.48, 0.00); P[5] = (-0.10, -0.48, 0.00)      # knees
    P[6] = (0.00, 0.26, -0.01)                                   # spine2
    P[7] = (0.10, -0.90, -0.03); P[8] = (-0.10, -0.90, -0.03)    # ankles
    P[9] = (0.00, 0.32, 0.01)                                    # spine3
    P[10] = (0.11, -0.96, 0.09); P[11] = (-0.11, -0.96, 0.09)    # feet
    P[12] = (0.00, 0.53, -0.03)                                  # neck
    P[13] = (0.08, 0.44, -0.02); P[14] = (-0.08, 0.44, -0.02)    # collars
    P[15] = (0.00, 0.62, 0.02)                                   # head
    P[16] = (0.18, 0.47, -0.03); P[17] = (-0.18, 0.47, -0.03)    # shoulders
    P[18] = (0.44, 0.46, -0.05); P[19] = (-0.44, 0.46, -0.05)    # elbows
    P[20] = (0.69, 0.47, -0.05); P[21] = (-0.69, 0.47, -0.05)    # wrists
    P[22] = (0.78, 0.46, -0.06); P[23] = (-0.78, 0.46, -0.06)    # hands
    return P


def rodrigues(rotvec):
    """Axis-angle [...,3] -> rotation matrices [...,3,3] (float64)."""
    rv = np.asarray(rotvec, dtype=np.float64)
    theta = np.linalg.norm(rv, axis=-1, keepdims=True)
    k = rv / np.maximum(theta, 1e-12)
    K = np.zeros(rv.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = np.sin(theta)[..., None], np.cos(theta)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def local_to_world_chain(bones, rest_pose, parents=SMPL_PARENTS):
    """bones [J,3] axis-angle -> l2w [J,4,4]: each joint's frame composed down the kinematic tree."""
    J = rest_pose.shape[0]
    R = rodrigues(bones)
    l2w = np.zeros((J, 4, 4))
    for j in range(J):
        T = np.eye(4)
        T[:3, :3] = R[j]
        if j == 0:
            T[:3, 3] = rest_pose[0]
            l2w[0] = T
        else:
            p = parents[j]
            T[:3, 3] = rest_pose[j] - rest_pose[p]
            l2w[j] = l2w[p] @ T
    return l2w


def rigid_inverse(T):
    """Closed-form inverse of rigid 4x4 transforms [...,4,4]."""
    R = T[..., :3, :3]
    t = T[..., :3, 3]
    out = np.zeros_like(T)
    Rt = np.swapaxes(R, -1, -2)
    out[..., :3, :3] = Rt
    out[..., :3, 3] = -(Rt @ t[..., None])[..., 0]
    out[..., 3, 3] = 1.0
    return out


def make_pose(seed=0, n_joints=24, pose_std=0.2):
    """One random pose. Returns dict of float32 arrays: bones[J,3], kps[J,3], skts[J,4,4], l2ws[J,4,4], cyl[5]."""
    rng = np.random.RandomState(seed)
    if n_joints == 24:
        rest, parents = humanoid_rest_pose(), SMPL_PARENTS
    else:  # small test skeletons: a chain of joints going up the y axis
        rest = np.stack([np.zeros(n_joints), 0.3 * np.arange(n_joints), np.zeros(n_joints)], -1).astype(np.float32)
        parents = np.maximum(np.arange(n_joints) - 1, 0)
    bones = (rng.randn(n_joints, 3) * pose_std).astype(np.float32)
    l2w = local_to_world_chain(bones, rest, parents)
    skts = rigid_inverse(l2w)
    kps = l2w[:, :3, 3]
    return dict(bones=bones, kps=kps.astype(np.float32), skts=skts.astype(np.float32),
                l2ws=l2w.astype(np.float32), cyl=bounding_cylinder(kps).astype(np.float32))


def bounding_cylinder(kps, extend=0.25, top_ratio=1.0, bot_ratio=0.25):
    """(cx, cz, radius, top, bottom) around the joints; height axis is -y (image-up convention)."""
    root = kps[0]
    dist = np.linalg.norm(kps[:, [0, 2]] - root[[0, 2]], axis=-1)
    hi, lo = (-kps[:, 1]).max(), (-kps[:, 1]).min()
    return np.array([root[0], root[2], dist.max() + extend,
                     -(hi + extend * top_ratio), -(lo - extend * bot_ratio)], dtype=np.float64)


def camera_rays(H, W, focal, c2w):
    """Pinhole rays for every pixel, row-major [H*W,3] origins and (un-normalised) directions."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -np.ones_like(i)], -1)
    rays_d = (dirs[..., None, :] * c2w[:3, :3]).sum(-1)
    rays_o = np.broadcast_to(c2w[:3, 3], rays_d.shape)
    return rays_o.reshape(-1, 3).astype(np.float32), rays_d.reshape(-1, 3).astype(np.float32)


def orbit_c2w(angle, dist=3.0, centre=(0., 0., 0.)):
    """Camera on a circle of radius `dist` around `centre` in the x-z plane, looking at the centre."""
    c, s = np.cos(angle), np.sin(angle)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)   # rotation about +y
    c2w = np.eye(4)
    c2w[:3, :3] = R
    c2w[:3, 3] = np.asarray(centre) + R @ np.array([0., 0., dist])
    return c2w.astype(np.float32)


def linear_init(rng, out_f, in_f):
    b = 1.0 / np.sqrt(in_f)
    return (rng.uniform(-b, b, size=(out_f, in_f)).astype(np.float32),
            rng.uniform(-b, b, size=(out_f,)).astype(np.float32))


def make_net_weights(seed, n_joints=24, multires=7, multires_views=4, D=8, W=256, skips=(4,),
                     framecode_ch=0, n_framecodes=0, frac_positive=0.3, mean_positive=6.0):
    """Weights of one density/radiance MLP with the reference's state_dict key names
    (core/networks/nerf.py:57-88).  alpha_linear is then calibrated so that densities are a mix of
    zero and positive values (default init gives raw sigma < 0 everywhere)."""
    rng = np.random.RandomState(seed)
    in_pts = n_joints * (1 + 2 * multires) + n_joints * 3
    in_views = n_joints * 3 * (1 + 2 * multires_views)
    sd = {}
    fan_in = in_pts
    for i in range(D):
        w, b = linear_init(rng, W, fan_in)
        sd[f'pts_linears.{i}.weight'], sd[f'pts_linears.{i}.bias'] = w, b
        fan_in = W + in_pts if i in skips else W
    w, b = linear_init(rng, 1, W)
    sd['alpha_linear.weight'], sd['alpha_linear.bias'] = w, b
    w, b = linear_init(rng, W // 2, W + in_views + framecode_ch)
    sd['views_linears.0.weight'], sd['views_linears.0.bias'] = w, b
    w, b = linear_init(rng, W, W)
    sd['feature_linear.weight'], sd['feature_linear.bias'] = w, b
    w, b = linear_init(rng, 3, W // 2)
    sd['rgb_linear.weight'], sd['rgb_linear.bias'] = w, b
    if framecode_ch > 0:
        std = np.sqrt(2.0 / (n_framecodes + framecode_ch))
        sd['framecodes.codes.weight'] = (rng.randn(n_framecodes, framecode_ch) * std).astype(np.float32)
    calibrate_density(sd, n_joints, multires, D, skips, frac_positive, mean_positive)
    return sd


def make_scene(seed=0, n_rays=None, H=512, W=512, focal=500.0, n_joints=24, cam_angle=0.0,
               cam_dist=3.0, pixel_offset=0):
    """One posed skeleton seen by one camera; per-ray tensors replicated as `render_path` does
    (reference: run_nerf.py:84-90).  `n_rays=None` -> the full H*W frame."""
    pose = make_pose(seed, n_joints)
    c2w = orbit_c2w(cam_angle, cam_dist, centre=pose['kps'][0] * np.array([1., 0., 1.]))
    ro, rd = camera_rays(H, W, focal, c2w)
    if n_rays is not None:
        # a deterministic spread of pixels across the frame
        rng = np.random.RandomState(seed + 7919 + pixel_offset)
        idx = np.sort(rng.choice(H * W, size=n_rays, replace=False))
        ro, rd = ro[idx], rd[idx]
    N = ro.shape[0]
    out = dict(rays_o=ro, rays_d=rd, c2w=c2w, H=H, W=W, focal=focal)
    out['skts'] = np.ascontiguousarray(np.broadcast_to(pose['skts'], (N,) + pose['skts'].shape))
    out['kps'] = np.ascontiguousarray(np.broadcast_to(pose['kps'], (N,) + pose['kps'].shape))
    out['bones'] = np.ascontiguousarray(np.broadcast_to(pose['bones'], (N,) + pose['bones'].shape))
    out['cyls'] = np.ascontiguousarray(np.broadcast_to(pose['cyl'], (N, 5)))
    out['pose'] = pose
    return out


# ------------------------------------------------------------------------------------------------
# density calibration: make sigma a mix of empty space and matter, like a trained model
# ------------------------------------------------------------------------------------------------
def _encode_points_np(pts, skts, multires, tau=20.0, cutoff=0.5):
    """float64 numpy encoding of world points [P,3] for one pose -> [P, J*(1+2F) + 3J]
    (distance embedding, channel k*J+j, then unit bone-local positions).  Used only to calibrate
    synthetic weights; the product's encoder is the CUDA kernel."""
    pts = np.asarray(pts, np.float64)
    skts = np.asarray(skts, np.float64)
    pt = np.einsum('jab,pb->pja', skts[:, :3, :3], pts) + skts[None, :, :3, 3]
    v = np.linalg.norm(pt, axis=-1)
    r = pt / np.maximum(v, 1e-12)[..., None]
    w = 1.0 - 1.0 / (1.0 + np.exp(-tau * (v - cutoff)))
    feats = [v]
    for f in range(multires):
        feats += [np.sin(v * 2.0 ** f), np.cos(v * 2.0 ** f)]
    return np.concatenate([x * w for x in feats] + [r.reshape(len(pts), -1)], -1)


def _trunk_np(sd, x, D, skips):
    h = x
    for i in range(D):
        h = np.maximum(h @ sd[f'pts_linears.{i}.weight'].astype(np.float64).T + sd[f'pts_linears.{i}.bias'], 0.0)
        if i in skips:
            h = np.concatenate([x, h], -1)
    return h


def _round_sig(x, digits=3):
    if x == 0:
        return 0.0
    return float(np.round(x, digits - 1 - int(np.floor(np.log10(abs(x))))))


def calibrate_density(sd, n_joints=24, multires=7, D=8, skips=(4,), frac_positive=0.3, mean_positive=6.0):
    """Rescale/shift alpha_linear (in place) so that about `frac_positive` of the points inside the
    skeleton's bounding cylinder have sigma > 0, with mean positive sigma ~ `mean_positive`.
    Gain and bias are rounded to 3 significant digits so every box derives identical weights."""
    pose = make_pose(0, n_joints)
    rng = np.random.RandomState(12345)
    cx, cz, R, top, bot = pose['cyl']
    P = 4096
    ang, rad = rng.uniform(0, 2 * np.pi, P), R * np.sqrt(rng.uniform(0, 1, P))
    pts = np.stack([cx + rad * np.cos(ang), rng.uniform(min(top, bot), max(top, bot), P), cz + rad * np.sin(ang)], -1)
    h = _trunk_np(sd, _encode_points_np(pts, pose['skts'], multires), D, skips)
    a = h @ sd['alpha_linear.weight'].astype(np.float64)[0]
    thr = np.quantile(a, 1.0 - frac_positive)
    pos = a[a > thr] - thr
    gain = _round_sig(mean_positive / max(pos.mean(), 1e-9))
    bias = _round_sig(-gain * thr)
    sd['alpha_linear.weight'] = (sd['alpha_linear.weight'].astype(np.float64) * gain).astype(np.float32)
    sd['alpha_linear.bias'] = np.array([bias], dtype=np.float32)
    return gain, bias
