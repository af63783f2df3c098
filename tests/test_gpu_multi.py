"""Two ranks over NCCL on one box (skipped with fewer than 2 GPUs; run with `gpurun --gpus 2`): the three ways the path
shards -- frame-parallel rendering + gather of pixels, voxel slabs + gather, training with equal ray
shares + ONE gradient all-reduce -- each against the single-GPU result."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import contextlib
    import io
    from anerf_b200 import mesh, parallel, synthetic
    from anerf_b200.raycasters import batchify_rays, create_raycaster
    from tests.test_gpu_api import data_attrs, make_args
    r, w, local = parallel.init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    with contextlib.redirect_stdout(io.StringIO()):
        rk_train, rk_test, _, grad_vars, optimizer, _ = create_raycaster(make_args(N_importance=16, no_reload=True), data_attrs(24), device=dev)
    rc = rk_test["ray_caster"]
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(101).items()})
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202).items()})
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    kw = {k: v for k, v in rk_test.items() if k not in ("ray_caster", "use_viewdirs")}
    ok = {}

    # ---- rendering: 3 frames of different sizes (valid-pixel crops), frame f on rank f mod 2, pixels gathered on rank 0
    def frame(f):
        sc = synthetic.make_scene(seed=0, n_rays=200 + 64 * f, H=128, W=128, focal=120., n_joints=24, cam_angle=0.3 * f)
        rays = torch.cat([t(sc["rays_o"]), t(sc["rays_d"]), torch.zeros(len(sc["rays_o"]), 1, device=dev), torch.ones(len(sc["rays_o"]), 1, device=dev)], 1)
        o = batchify_rays(rays, 128, ray_caster=rc.eval(), kp_batch=t(sc["kps"]), skts=t(sc["skts"]), cyls=t(sc["cyls"]), bones=t(sc["bones"]),
                          cams=None, subject_idxs=None, **kw)
        return torch.cat([o["rgb_map"], o["disp_map"][:, None], o["acc_map"][:, None]], 1)
    n_frames = 3
    mine = {f: frame(f) for f in parallel.frames_for_rank(n_frames, r, w)}
    allf = parallel.gather_pixels(mine, n_frames, r, w)
    if r == 0:
        ok["gather_pixels"] = all(torch.equal(allf[f], frame(f)) for f in range(n_frames))

    # ---- mesh: voxel slabs + gather == the whole grid on one GPU
    pose = synthetic.make_pose(11, 24)
    kps, skts = t(pose["kps"])[None], t(pose["skts"])[None]
    grid = mesh.density_grid_sharded(rc, kps, skts, radius=1.0, res=23, rank=r, world=w)
    if r == 0:
        ok["mesh"] = torch.equal(grid, rc.render_mesh_density(kps, skts, None, radius=1.0, res=23))

    # ---- training: equal contiguous ray shares + one all-reduce == the full-batch gradient
    holder = rk_train["ray_caster"].train()
    N = 128
    # a narrow field of view: every ray crosses the bounding cylinder, so the near/far repair (a CHUNK-wide mean that
    # would differ between the full batch and its halves) is not involved
    sc = synthetic.make_scene(seed=5, n_rays=N, H=64, W=64, focal=250., n_joints=24)
    rays = torch.cat([t(sc["rays_o"]), t(sc["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev),
                      torch.nn.functional.normalize(t(sc["rays_d"]), dim=-1)], 1)
    target = t(np.random.RandomState(9).rand(N, 3).astype(np.float32))
    tk = {k: v for k, v in rk_train.items() if k not in ("ray_caster", "use_viewdirs")}
    tk.update(perturb=0., raw_noise_std=0.)

    def grads(lo, hi):
        for p in grad_vars:
            p.grad = None
        o = holder(rays[lo:hi], kp_batch=t(sc["kps"])[lo:hi], skts=t(sc["skts"])[lo:hi], cyls=t(sc["cyls"])[lo:hi], bones=t(sc["bones"])[lo:hi],
                   cams=None, subject_idxs=None, **tk)
        (((o["rgb_map"] - target[lo:hi]) ** 2).mean() + ((o["rgb0"] - target[lo:hi]) ** 2).mean()).backward()
    grads(0, N)
    full = [p.grad.clone() for p in grad_vars]
    lo, hi = parallel.rays_for_rank(N, r, w)
    grads(lo, hi)
    parallel.allreduce_gradients(grad_vars, w)
    worst = max(float((p.grad - f).abs().max() / f.abs().max().clamp_min(1e-20)) for p, f in zip(grad_vars, full))
    ok["training"] = worst < 2e-4
    ok["training_worst"] = worst
    # the un-averaged sum + FusedAdam(grad_scale) is the same update
    grads(lo, hi)
    parallel.allreduce_gradients(grad_vars, w, average=False)
    ok["sum"] = max(float((p.grad / w - f).abs().max() / f.abs().max().clamp_min(1e-20)) for p, f in zip(grad_vars, full)) < 2e-4
    torch.cuda.synchronize()
    if r == 0:
        q.put(ok)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_ranks_nccl():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    assert ok["gather_pixels"] and ok["mesh"] and ok["training"] and ok["sum"], ok
