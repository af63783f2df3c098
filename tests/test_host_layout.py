"""CPU checks of the host-side logic of the CUDA path: the K permutation that the weight packer
applies must be the inverse of the order in which the kernel's producers emit encoding values.
Compiles anerf_b200/csrc/path_math.cuh (the __host__ __device__ scalar stages) with g++."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from anerf_b200 import synthetic
from oracle import anerf_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("harness") / "layout_harness.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(HERE, "host", "layout_harness.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.h_linspace01.restype = C.c_float
    return lib


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.mark.parametrize("J,fc", [(24, 0), (24, 16), (1, 0), (5, 0)])
def test_packed_k_order_inverts_producer_order(harness, J, fc):
    rng = np.random.RandomState(J)
    pose = synthetic.make_pose(3, J)
    skt12 = np.ascontiguousarray(pose["skts"][:, :3, :].reshape(J, 12))
    p = (pose["kps"][0] + rng.randn(3) * 0.3).astype(np.float32)
    dirv = rng.randn(3).astype(np.float32)
    cut = np.full(J, 0.5, np.float32)
    tau = 20.0
    cfg = orc.PathConfig(n_joints=J, framecode_ch=fc)
    t = torch.as_tensor
    xp, xv = orc.encode_samples(t(p)[None, None], t(dirv)[None], t(pose["skts"])[None], cfg)
    xp, xv = xp[0, 0].numpy().astype(np.float64), xv[0, 0].numpy().astype(np.float64)
    fcode = rng.randn(16).astype(np.float32)
    D, W, skip = 8, 256, 4
    nP = harness.h_layer_chunks(J, D, W, skip, fc, 0) * 32
    ep = np.zeros(nP, np.float32)
    harness.h_emit_pts(fptr(skt12), fptr(p), C.c_float(tau), fptr(cut), J, fptr(ep))
    col = lambda l, k: harness.h_layer_ref_col(J, D, W, skip, fc, l, k)
    # layer 0: emitted value k must be reference input column col(0,k)
    for k in range(nP):
        c = col(0, k)
        assert (ep[k] == 0.0) if c < 0 else abs(ep[k] - xp[c]) < 2e-6, (k, c)
    assert sorted(col(0, k) for k in range(nP) if col(0, k) >= 0) == list(range(cfg.in_pts))
    # skip layer (l = skip + 1): [encoding | h]
    l = skip + 1
    cols = [col(l, k) for k in range(nP + W)]
    assert cols[:nP] == [col(0, k) for k in range(nP)]
    assert cols[nP:] == [cfg.in_pts + n for n in range(W)]
    # plain trunk layer and the streamed (h) part of the views layer: identity
    assert [col(2, k) for k in range(W)] == list(range(W))
    assert [col(D, k) for k in range(W)] == list(range(W))
    # view branch contracted per ray: sum_j w_j sum_q Wv[n, col(j,q)] table[j][q] == Wv[n, W:] . x_view
    wv = rng.randn(5, W + cfg.in_views + fc)
    tab = np.zeros((J, 27), np.float32)
    harness.h_view_table(fptr(skt12), fptr(dirv), J, fptr(tab))
    harness.h_cutoff_w.restype = C.c_float
    wj = np.array([harness.h_cutoff_w(fptr(np.ascontiguousarray(skt12[j])), fptr(p), C.c_float(tau), C.c_float(0.5)) for j in range(J)])
    vcol = lambda j, q: harness.h_view_weight_col(J, D, W, skip, fc, j, q)
    got = np.zeros(5)
    seen = []
    for j in range(J + 1):
        for q in range(27):
            c = vcol(j, q)
            if c < 0:
                continue
            seen.append(c)
            got += wv[:, c] * ((tab[j, q] * wj[j]) if j < J else fcode[q])
    ref_in = np.concatenate([np.zeros(W), xv, fcode[:fc].astype(np.float64)])
    assert np.abs(got - wv @ ref_in).max() < 1e-5
    assert sorted(seen) == list(range(W, W + cfg.in_views + fc))


def test_linspace_matches_torch(harness):
    # ours is the scalar formula of ATen's CUDA linspace kernel (start + step*i below the middle,
    # end - step*(n-1-i) above); the CPU kernel evaluates the same thing per SIMD vector and can
    # differ in the last bit, hence 1 ulp.
    for n in (2, 16, 63, 64, 128, 129):
        ours = np.array([harness.h_linspace01(i, n) for i in range(n)], np.float32)
        ref = torch.linspace(0., 1., n).numpy()
        assert np.abs(ours - ref).max() <= 6e-8, n
        assert ours[0] == 0. and ours[-1] == 1.


def test_near_far_matches_oracle(harness):
    scene = synthetic.make_scene(seed=5, n_rays=64)
    scene["rays_d"][:4, 0] += 2.0
    t = torch.as_tensor
    N = 64
    near, far = torch.zeros(N, 1), torch.ones(N, 1)
    # oracle without the repair step = raw intersection
    o, d, cyl = scene["rays_o"], scene["rays_d"], scene["cyls"]
    out = np.zeros((N, 3), np.float32)
    for i in range(N):
        harness.h_near_far(fptr(np.ascontiguousarray(o[i])), fptr(np.ascontiguousarray(d[i])),
                           fptr(np.ascontiguousarray(cyl[i])), C.c_float(0.), C.c_float(1.), fptr(out[i]))
    nn, ff = orc.near_far_in_cylinder(t(o), t(d), t(cyl), near, far)
    hit = out[:, 2] == 0
    assert hit.sum() >= 32 and (~hit).sum() >= 4 and not hit[:4].any()
    assert np.allclose(out[hit, 0], nn[:, 0].numpy()[hit], rtol=2e-6, atol=1e-6)
    assert np.allclose(out[hit, 1], ff[:, 0].numpy()[hit], rtol=2e-6, atol=1e-6)
    assert np.isnan(out[~hit, 0]).all()


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only box and exports what include/anerf_b200.h declares."""
    import re
    from anerf_b200 import _lib, build
    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(os.path.dirname(HERE), "include", "anerf_b200.h")).read()
    declared = set(re.findall(r"\b(anerf_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.anerf_version() >= 100


def test_pixel_rays_match_reference_get_rays(harness):
    """The kernels' in-place ray generation (anerf_render_frame) against the reference's get_rays
    (core/utils/ray_utils.py:6-28), restated here with the same torch ops: bit-exact on the CPU."""
    H, W, focal = 48, 64, 57.5
    c2w = synthetic.orbit_c2w(0.7, 3.0)[:3, :4].astype(np.float32)
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='ij')
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * torch.as_tensor(c2w[:3, :3]), -1).reshape(-1, 3).numpy()
    n, p0 = 777, 1234
    out = np.zeros((n, 8), np.float32)
    harness.h_pixel_rays(fptr(np.ascontiguousarray(c2w.reshape(-1))), C.c_float(focal), C.c_float(focal), C.c_float(W * 0.5),
                         C.c_float(H * 0.5), W, p0, n, fptr(out))
    assert np.array_equal(out[:, 3:6], rays_d[p0:p0 + n])
    assert np.array_equal(out[:, 0:3], np.broadcast_to(c2w[:, 3], (n, 3)))
    assert np.all(out[:, 6] == 0.) and np.all(out[:, 7] == 1.)


@pytest.mark.parametrize("lindisp", [0, 1])
@pytest.mark.parametrize("jitter", [False, True])
def test_coarse_depths_match_oracle(harness, lindisp, jitter):
    """coarse_depth (path_math.cuh; what the fused kernel and the backward pass sample at) against the oracle's
    restatement of sample_from_lineseg (ray_utils.py:204-251), incl. stratified jitter and lindisp."""
    rng = np.random.RandomState(3)
    for Sc in (16, 64, 65):
        near, far = np.float32(1.2345), np.float32(4.321)
        tr = rng.rand(Sc).astype(np.float32) if jitter else None
        out = np.zeros(Sc, np.float32)
        harness.h_coarse_depths(C.c_float(near), C.c_float(far), Sc, lindisp, fptr(tr) if jitter else None, fptr(out))
        ref = orc.coarse_depths(torch.tensor([[near]]), torch.tensor([[far]]), Sc, None if tr is None else torch.as_tensor(tr)[None],
                                bool(lindisp))[0].numpy()
        assert np.abs(out - ref).max() <= 5e-7 * far, (Sc, np.abs(out - ref).max())
