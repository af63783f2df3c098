"""Frame loop glue around the ray caster: which pixels of a frame can see the
skeleton, rendering them with in-kernel ray generation, and compositing over the background.

Reference this stands in for (paths relative to the reference root):
  cylinder_to_box_2d        core/utils/skeleton_utils.py:607-694   -> cylinder_box_2d
  kp_to_valid_rays          core/utils/ray_utils.py:83-136         -> valid_pixels   (no rays are materialised here)
  render_path (per frame)   run_nerf.py:77-136                     -> render_path_frame

The box is a few dozen floating-point operations per frame and stays on the host (numpy, same operations and dtypes
as the reference so that the integer box is the same); everything per pixel runs on the device.
"""
import numpy as np
import torch


def _swap_mat(mat):
    """[right, -up, -forward] (skeleton_utils.py:1308-1317)."""
    return np.concatenate([mat[..., 0:1], -mat[..., 1:2], -mat[..., 2:3], mat[..., 3:]], axis=-1)


def cylinder_box_2d(cyl, H, W, focal, c2w, center=None):
    """2-D box (top-left, bottom-right; x then y, integer pixels, clipped to the image) of the bounding cylinder
    `cyl` = (cx, cz, radius, top, bottom) seen from camera `c2w` [4,4] (skeleton_utils.py:607-694 with scale = 1)."""
    cyl = np.asarray(cyl)
    c2w = np.asarray(c2w)
    if c2w.shape[0] == 3:
        c2w = np.concatenate([c2w, np.array([[0, 0, 0, 1]], c2w.dtype)], 0)
    w2c = np.linalg.inv(_swap_mat(c2w))                                  # nerf_c2w_to_extrinsic (skeleton_utils.py:442)
    root_loc, radius = cyl[None, :2], cyl[None, 2:3]
    top, bot = cyl[None, 3:4], cyl[None, 4:5]
    rads = np.linspace(0., 2 * np.pi, 50)
    x = root_loc[..., 0:1] + np.cos(rads)[None] * radius
    z = root_loc[..., 1:2] + np.sin(rads)[None] * radius
    ones = np.ones_like(x)
    caps = np.concatenate([np.stack([x, top * ones, z, ones], -1), np.stack([x, bot * ones, z, ones], -1)], -2).reshape(-1, 4)
    fx, fy = (focal, focal) if np.ndim(focal) == 0 or np.size(focal) < 2 else (focal[0], focal[1])
    intrinsic = np.array([[fx, 0, 0, 0], [0, fy, 0, 0], [0, 0, 1, 0]], dtype=np.float32)
    pts = (caps @ w2c.T) @ intrinsic.T
    pts_2d = pts[..., :2] / pts[..., 2:3]
    max_x, min_x = np.ceil(pts_2d[..., 0].max(-1)).astype(np.int32), np.floor(pts_2d[..., 0].min(-1)).astype(np.int32)
    max_y, min_y = np.ceil(pts_2d[..., 1].max(-1)).astype(np.int32), np.floor(pts_2d[..., 1].min(-1)).astype(np.int32)
    off_x, off_y = (int(W * .5), int(H * .5)) if center is None else (int(center[0]), int(center[1]))
    tl = np.array([np.clip(min_x + off_x, 0, W - 1), np.clip(min_y + off_y, 0, H - 1)], np.int32)
    br = np.array([np.clip(max_x + off_x, 0, W - 1), np.clip(max_y + off_y, 0, H - 1)], np.int32)
    return tl, br


def valid_pixels(cyl, H, W, focal, c2w, center=None, device=None):
    """Flat indices (row-major j*W + i, int32) of the pixels inside the cylinder's 2-D box: the `valid_idx` of
    kp_to_valid_rays (ray_utils.py:127-131)."""
    tl, br = cylinder_box_2d(cyl, H, W, focal, c2w, center)
    h_range = torch.arange(int(tl[1]), int(br[1]), device=device)
    w_range = torch.arange(int(tl[0]), int(br[0]), device=device)
    return (h_range[:, None] * W + w_range[None, :]).reshape(-1).to(torch.int32), (tl, br)


@torch.no_grad()
def render_path_frame(ray_caster, c2w, H, W, focal, skts, cyl, render_kwargs, bg=None, white_bkgd=False, center=None,
                      chunk=4096, cams=None):
    """One frame of run_nerf.render_path (run_nerf.py:77-136): the pixels inside the skeleton's box are rendered
    (RayCaster.render_frame: rays generated in the kernels, pose read once) and composited over the background,
    everything else is background.  skts [J,4,4] / [1,J,4,4], cyl [5] / [1,5] CUDA tensors; bg: optional [H,W,3] CUDA
    tensor.  Returns rgb [H,W,3], disp [H,W,1], acc [H,W,1] on the device."""
    dev = skts.device
    cyl_t = cyl.reshape(-1, cyl.shape[-1])[0]
    valid_idx, _ = valid_pixels(cyl_t.detach().cpu().numpy(), H, W, focal, np.asarray(torch.as_tensor(c2w).cpu()), center, device=dev)
    if bg is not None and not white_bkgd:
        rgb_img = bg.reshape(H * W, 3).to(dev, torch.float32).clone()
    else:
        rgb_img = torch.ones(H * W, 3, device=dev) if white_bkgd else torch.zeros(H * W, 3, device=dev)
    disp_img = torch.zeros(H * W, device=dev)
    acc_img = torch.zeros(H * W, device=dev)
    if valid_idx.numel() > 0:
        kw = {k: v for k, v in render_kwargs.items() if k not in ('ray_caster', 'use_viewdirs')}
        out = ray_caster.render_frame(H, W, focal, c2w, skts, cyl, cams=cams, pixel_idx=valid_idx, chunk=chunk, center=center, **kw)
        idx = valid_idx.long()
        rgb_img[idx] = out['rgb_map'] + (1. - out['acc_map'][..., None]) * rgb_img[idx]
        disp_img[idx] = out['disp_map']
        acc_img[idx] = out['acc_map']
    return rgb_img.view(H, W, 3), disp_img.view(H, W, 1), acc_img.view(H, W, 1)
