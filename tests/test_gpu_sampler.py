"""SURVEY.md 8(f) row 4 (second half) on the GPU: the training-ray sampler (anerf_sample_rays behind
anerf_b200.sampler.RaySampler) against what BaseH5Dataset.__getitem__ produces for the same pixels
(core/dataset.py:57-105, 258-275, 277-322, 346-362): valid, distinct, sorted pixels; rays of get_rays_np; colours of the
images; uniformity of the draws.  numpy's generator is not reproduced (different draws, same distribution)."""
import numpy as np
import pytest
import torch

from anerf_b200.sampler import RaySampler
from anerf_b200 import synthetic

pytestmark = pytest.mark.gpu


def _dataset(F=6, H=40, W=56, seed=0):
    rng = np.random.RandomState(seed)
    imgs = rng.randint(0, 256, size=(F, H, W, 3)).astype(np.uint8)
    bgs = rng.randint(0, 256, size=(2, H, W, 3)).astype(np.uint8)
    masks = np.zeros((F, H, W), np.uint8)
    for f in range(F):
        masks[f, 5 + f:30, 8:40 + f] = 1
    if F > 1:
        masks[1, 12, 20] = 0
    fgs = (rng.rand(F, H, W) > 0.5).astype(np.uint8)
    c2ws = np.stack([synthetic.orbit_c2w(0.3 * f, 3.0) for f in range(F)]).astype(np.float32)
    focals = (50. + 3 * np.arange(F)).astype(np.float32)
    centers = np.stack([np.full(F, W * 0.5 + 1.5), np.full(F, H * 0.5 - 0.5)], -1).astype(np.float32)
    return dict(imgs=imgs, bgs=bgs, masks=masks, fgs=fgs, c2ws=c2ws, focals=focals, centers=centers, bg_idxs=np.arange(F) % 2, H=H, W=W)


def _rays_np(H, W, focal, c2w, center):
    """get_rays_np (core/utils/ray_utils.py:31-60) for the whole image."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i - center[0]) / focal, -(j - center[1]) / focal, -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    return np.broadcast_to(c2w[:3, -1], rays_d.shape).reshape(-1, 3), rays_d.reshape(-1, 3)


def test_batch_is_what_the_dataset_returns_for_those_pixels():
    d = _dataset()
    s = RaySampler(d["imgs"], d["masks"], d["c2ws"], d["focals"], d["H"], d["W"], fgs=d["fgs"], bgs=d["bgs"], bg_idxs=d["bg_idxs"],
                   centers=d["centers"], cam_idxs=np.arange(6) + 10, kp_idxs=np.arange(6) * 2, seed=3)
    b = s.sample(N_rand=4 * 64, n_images=4, frames=[4, 1, 5, 1])
    pix, fr = b["pixel_idx"].cpu().numpy(), b["frame"].cpu().numpy()
    assert pix.shape == (256,) and list(fr[::64]) == [4, 1, 5, 1]
    for g in range(4):
        p, f = pix[g * 64:(g + 1) * 64], fr[g * 64]
        assert (np.diff(p) > 0).all()                                       # distinct and increasing (np.sort in the reference)
        assert d["masks"][f].reshape(-1)[p].all()                           # only pixels of the sampling mask
        ro, rd = _rays_np(d["H"], d["W"], d["focals"][f], d["c2ws"][f], d["centers"][f])
        assert np.abs(b["rays_o"].cpu().numpy()[g * 64:(g + 1) * 64] - ro[p]).max() < 1e-6
        assert np.abs(b["rays_d"].cpu().numpy()[g * 64:(g + 1) * 64] - rd[p]).max() < 1e-5
        assert np.array_equal(b["target_s"].cpu().numpy()[g * 64:(g + 1) * 64], d["imgs"][f].reshape(-1, 3)[p].astype(np.float32) / 255.)
        assert np.array_equal(b["fgs"].cpu().numpy()[g * 64:(g + 1) * 64, 0], d["fgs"][f].reshape(-1)[p].astype(np.float32))
        assert np.array_equal(b["bgs"].cpu().numpy()[g * 64:(g + 1) * 64], d["bgs"][d["bg_idxs"][f]].reshape(-1, 3)[p].astype(np.float32) / 255.)
    assert list(b["cam_idxs"].cpu().numpy()[::64]) == [14, 11, 15, 11] and list(b["kp_idx"].cpu().numpy()[::64]) == [8, 2, 10, 2]
    assert float(b["rays"][:, 6].abs().max()) == 0. and float((b["rays"][:, 7] - 1).abs().max()) == 0.
    # the two draws from image 1 inside one batch share a seed and an image: same pixels; the next batch differs
    assert np.array_equal(pix[64:128], pix[192:256])
    b2 = s.sample(N_rand=4 * 64, n_images=4, frames=[4, 1, 5, 1])
    assert not np.array_equal(b2["pixel_idx"].cpu().numpy(), pix)
    # mask_img: target = img * fg + (1 - fg) * bg
    sm = RaySampler(d["imgs"], d["masks"], d["c2ws"], d["focals"], d["H"], d["W"], fgs=d["fgs"], bgs=d["bgs"], bg_idxs=d["bg_idxs"],
                    mask_img=True, seed=3)
    bm = sm.sample(N_rand=64, n_images=1, frames=[2])
    p = bm["pixel_idx"].cpu().numpy()
    fg = d["fgs"][2].reshape(-1)[p].astype(np.float32)[:, None]
    want = d["imgs"][2].reshape(-1, 3)[p] / 255. * fg + (1 - fg) * d["bgs"][0].reshape(-1, 3)[p] / 255.
    assert np.abs(bm["target_s"].cpu().numpy() - want).max() < 1e-6


def test_draws_are_uniform_and_exhaustive():
    d = _dataset(F=1)
    s = RaySampler(d["imgs"], d["masks"], d["c2ws"], d["focals"], d["H"], d["W"], seed=7)
    valid = np.flatnonzero(d["masks"][0].reshape(-1))
    # k = all valid pixels: a permutation-free take of everything
    b = s.sample(N_rand=len(valid), n_images=1, frames=[0])
    assert np.array_equal(b["pixel_idx"].cpu().numpy(), valid)
    with pytest.raises(ValueError, match="valid pixels"):
        s.sample(N_rand=len(valid) + 1, n_images=1, frames=[0])
    with pytest.raises(IndexError, match="image index"):
        s.sample(N_rand=8, n_images=1, frames=[3])
    # many batches: every valid pixel is drawn about equally often
    hits = np.zeros(d["H"] * d["W"])
    n_batches, k = 400, 50
    for _ in range(n_batches):
        hits[s.sample(N_rand=k, n_images=1, frames=[0])["pixel_idx"].cpu().numpy()] += 1
    assert hits[np.setdiff1d(np.arange(hits.size), valid)].sum() == 0
    expect = n_batches * k / len(valid)
    chi2 = ((hits[valid] - expect) ** 2 / expect).sum() / (len(valid) - 1)
    assert 0.8 < chi2 < 1.2, chi2                                  # reduced chi-square of a uniform draw is ~1 (hypergeometric: slightly below)
