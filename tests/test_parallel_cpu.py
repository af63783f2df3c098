"""Multi-process sharding logic on CPU (gloo, world_size 2): frame dealing, pixel gather, slab split."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from anerf_b200 import parallel


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_distributed(backend="gloo")
    n_frames = 5
    mine = {f: torch.full((7, 5), float(f)) + torch.arange(5.) * 0.1 for f in parallel.frames_for_rank(n_frames, r, w)}
    frames = parallel.gather_pixels(mine, n_frames, r, w)
    # ragged frames (per-frame valid-pixel crops) and fewer frames than ranks: rank 1 owns nothing in the 1-frame case
    rag = {f: torch.arange((3 + 2 * f) * 5, dtype=torch.float32).reshape(3 + 2 * f, 5) + f for f in parallel.frames_for_rank(3, r, w)}
    rag_out = parallel.gather_pixels(rag, 3, r, w)
    one = {f: torch.ones(4, 5) * 9 for f in parallel.frames_for_rank(1, r, w)}
    one_out = parallel.gather_pixels(one, 1, r, w)
    n = 11
    a, b = parallel.slab_for_rank(n, r, w)
    full = parallel.gather_slabs(torch.arange(a, b, dtype=torch.float32)[:, None] * torch.ones(1, 3), n, r, w)
    tmax = parallel.max_over_ranks(10.0 + r, torch.device("cpu"))
    # training exchange step: one flat all-reduce averages the gradients (a parameter without .grad counts as zero)
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(2))]
    ps[0].grad = torch.full((3, 4), float(r + 1))
    if r == 0:
        ps[1].grad = torch.arange(5.)
    ps[2].requires_grad_(False)
    parallel.allreduce_gradients(ps, w)
    g_ok = torch.allclose(ps[0].grad, torch.full((3, 4), 1.5)) and torch.allclose(ps[1].grad, torch.arange(5.) / 2) \
        and ps[2].grad is None
    # equal contiguous shares + the averaged gradient == the full-batch gradient of a mean-over-rays loss
    torch.manual_seed(1)
    lin = torch.nn.Linear(6, 3)
    xs, ys = torch.randn(64, 6), torch.randn(64, 3)
    full_g = torch.autograd.grad(((lin(xs) - ys) ** 2).mean(), list(lin.parameters()))
    lo, hi = parallel.rays_for_rank(64, r, w)
    lin.zero_grad()
    ((lin(xs[lo:hi]) - ys[lo:hi]) ** 2).mean().backward()
    parallel.allreduce_gradients(list(lin.parameters()), w)
    g_ok = g_ok and all(torch.allclose(p.grad, f, atol=1e-6) for p, f in zip(lin.parameters(), full_g))
    a0, b0 = parallel.rays_for_rank(3072, r, w)
    g_ok = g_ok and (b0 - a0) == 1536 and a0 == r * 1536
    if r == 0:
        ok = all(torch.equal(frames[f], torch.full((7, 5), float(f)) + torch.arange(5.) * 0.1) for f in range(n_frames))
        ok = ok and torch.equal(full, torch.arange(n, dtype=torch.float32)[:, None] * torch.ones(1, 3))
        ok = ok and all(torch.equal(rag_out[f], torch.arange((3 + 2 * f) * 5, dtype=torch.float32).reshape(3 + 2 * f, 5) + f) for f in range(3))
        ok = ok and len(one_out) == 1 and torch.equal(one_out[0], torch.ones(4, 5) * 9)
        q.put((ok and g_ok, tmax))
    else:
        assert frames is None and full is None and g_ok
    dist.barrier()
    dist.destroy_process_group()


def test_frame_and_slab_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and tmax == 11.0


def test_slabs_cover_everything():
    for n in (1, 7, 256, 257):
        for w in (1, 2, 3, 8):
            s = [parallel.slab_for_rank(n, r, w) for r in range(w)]
            assert s[0][0] == 0 and s[-1][1] == n and all(s[i][1] == s[i + 1][0] for i in range(w - 1))
            assert sorted(f for r in range(w) for f in parallel.frames_for_rank(n, r, w)) == list(range(n))


def test_grid_points_match_reference_meshgrid():
    import numpy as np
    from anerf_b200 import mesh
    res, radius = 6, 0.7
    kps = torch.tensor([[[0.1, -0.2, 0.3]]])
    t = np.linspace(-radius, radius, res + 1)
    ref = np.stack(np.meshgrid(t, t, t), axis=-1).astype(np.float32).reshape(-1, 3) + kps[0, 0].numpy()
    pts = mesh.grid_points(kps, radius, res).numpy()
    assert np.abs(pts - ref).max() < 1e-6
    a, b = parallel.slab_for_rank((res + 1) ** 3, 1, 3)
    assert np.abs(mesh.grid_points(kps, radius, res, a, b).numpy() - ref[a:b]).max() < 1e-6
