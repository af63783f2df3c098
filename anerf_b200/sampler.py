"""SURVEY.md 8(f) row 4 (second half): the training-ray sampler on the device.

`RaySampler` keeps what the reference's `BaseH5Dataset` reads from its .h5 file per item (images, sampling masks,
foreground masks, backgrounds, cameras; core/dataset.py:57-160) resident in HBM as uint8 / fp32 tensors and draws a
whole training batch in ONE kernel launch (C ABI `anerf_sample_rays`): `n_images` images, `N_rand // n_images` distinct
pixels each, uniformly from the image's sampling mask, in increasing pixel order -- the reference's image_batching
layout (`RayImageSampler`, core/dataset.py), where a DataLoader worker does this on the host one image at a time.

The returned dict has the reference's batch keys (`rays_o`, `rays_d`, `target_s`, `fgs`, `bgs`, `cam_idxs`, `kp_idx`)
plus `pixel_idx`; pose tensors are looked up by `kp_idx` (e.g. anerf_b200.pose_opt.PoseOptLayer.forward_poses).
Not implemented (raise): patch sampling (patch_size > 1) and box-constrained samples (N_nms > 0).
The random stream is a counter-based hash of (seed, image, pixel) -- not numpy's generator, so draws differ from the
reference's for the same seed while the distribution is the same (uniform k-subsets of the valid pixels).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class _SamplerInputs(C.Structure):
    _fields_ = [("masks", C.c_void_p), ("imgs", C.c_void_p), ("fgs", C.c_void_p), ("bgs", C.c_void_p), ("bg_idx", C.c_void_p),
                ("c2ws", C.c_void_p), ("focals", C.c_void_p), ("centers", C.c_void_p), ("height", C.c_int32), ("width", C.c_int32),
                ("n_frames", C.c_int32), ("fg_is_255", C.c_int32), ("mask_img", C.c_int32)]


class _SamplerOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rays", "target", "fg", "bg", "pixel_idx", "frame_of_ray")]


class RaySampler:
    def __init__(self, imgs, sampling_masks, c2ws, focals, H, W, fgs=None, bgs=None, bg_idxs=None, centers=None, cam_idxs=None,
                 kp_idxs=None, mask_img=False, fg_is_255=False, device=None, seed=0):
        dev = torch.device(device if device is not None else "cuda")
        if dev.type != "cuda":
            raise RuntimeError("anerf_b200.RaySampler: CUDA only (no CPU path)")
        u8 = lambda a, *shape: None if a is None else torch.as_tensor(np.ascontiguousarray(a)).to(torch.uint8).reshape(*shape).contiguous().to(dev)
        F = len(imgs)
        self.F, self.H, self.W, self.device = F, int(H), int(W), dev
        self.imgs = u8(imgs, F, H * W, 3)
        self.masks = u8(sampling_masks, F, H * W)
        self.fgs = u8(fgs, F, H * W)
        self.bgs = None if bgs is None else u8(bgs, len(bgs), H * W, 3)
        f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).contiguous().to(dev)
        i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(torch.int32).contiguous().to(dev)
        self.bg_idx = None if bg_idxs is None else i32(bg_idxs)
        self.c2ws = f32(np.asarray(c2ws, np.float32)[:, :3, :4])
        fo = np.asarray(focals, np.float32)
        self.focals = f32(np.stack([fo, fo], -1) if fo.ndim == 1 else fo.reshape(F, 2))
        self.centers = None if centers is None else f32(np.asarray(centers, np.float32).reshape(F, 2))
        self.cam_idxs = torch.arange(F, device=dev) if cam_idxs is None else torch.as_tensor(cam_idxs).to(dev)
        self.kp_idxs = torch.arange(F, device=dev) if kp_idxs is None else torch.as_tensor(kp_idxs).to(dev)
        self.mask_img, self.fg_is_255 = bool(mask_img), bool(fg_is_255)
        self._seed, self._calls = int(seed), 0
        self._frame_rng = np.random.RandomState(seed)

    def sample(self, N_rand, n_images, frames=None):
        """One training batch: `n_images` images (random without replacement unless `frames` lists them) x N_rand //
        n_images rays.  Returns the reference's batch dict with CUDA tensors."""
        k = N_rand // n_images
        if k < 1:
            raise ValueError("N_rand must be at least n_images")
        if frames is None:
            frames = self._frame_rng.choice(self.F, size=n_images, replace=n_images > self.F)
        frames_t = torch.as_tensor(np.asarray(frames)).to(torch.int32).contiguous().to(self.device)
        n_images = int(frames_t.shape[0])
        N = n_images * k
        dev = self.device
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        rays, target, pix, fr = f(N, 8), f(N, 3), torch.empty(N, dtype=torch.int32, device=dev), torch.empty(N, dtype=torch.int32, device=dev)
        fg = f(N) if self.fgs is not None else None
        bg = f(N, 3) if self.bgs is not None else None
        n_valid = torch.empty(n_images, dtype=torch.int32, device=dev)
        p = _lib._ptr
        sin = _SamplerInputs(p(self.masks), p(self.imgs), p(self.fgs), p(self.bgs), p(self.bg_idx), p(self.c2ws), p(self.focals),
                             p(self.centers), self.H, self.W, self.F, int(self.fg_is_255), int(self.mask_img))
        sout = _SamplerOutputs(p(rays), p(target), p(fg), p(bg), p(pix), p(fr))
        self._calls += 1
        seed = (self._seed * 0x9E3779B97F4A7C15 + self._calls * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
        lib = _lib.load()
        lib.anerf_sample_rays.argtypes = [C.POINTER(_SamplerInputs), C.c_void_p, C.c_int32, C.c_int32, C.c_uint64,
                                          C.POINTER(_SamplerOutputs), C.c_void_p, C.c_void_p]
        with torch.cuda.device(dev):
            _lib.check(lib.anerf_sample_rays(C.byref(sin), p(frames_t), n_images, k, C.c_uint64(seed), C.byref(sout), p(n_valid), _lib._stream()))
        if bool((n_valid < 0).any()):
            raise IndexError(f"image index outside [0, {self.F}): {[int(x) for x in frames_t[n_valid < 0][:8]]}")
        if bool((n_valid < k).any()):
            bad = int(torch.nonzero(n_valid < k)[0])
            raise ValueError(f"image {int(frames_t[bad])} has {int(n_valid[bad])} valid pixels, fewer than the {k} rays asked for")
        fl = fr.long()
        return {"rays_o": rays[:, 0:3], "rays_d": rays[:, 3:6], "rays": rays, "target_s": target, "fgs": None if fg is None else fg[:, None],
                "bgs": bg, "cam_idxs": self.cam_idxs[fl], "kp_idx": self.kp_idxs[fl], "pixel_idx": pix, "frame": fr}
