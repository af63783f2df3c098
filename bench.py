#!/usr/bin/env python
"""Benchmark of the A-NeRF ray-marching hot path (BASELINE.json metric: rays/sec at 512^2,
64 coarse + 128 fine samples, 24 joints, 8x256 coarse + fine nets, 4096-ray chunks).

    python bench.py --gpus 1 --steps 3 --warmup 3                # our CUDA path
    python bench.py --impl reference --steps 1 --warmup 1        # the reference algorithm on the host CPU
    torchrun --nproc-per-node N ... bench.py --gpus N            # frame-parallel, one rank per GPU

One "step" = one full synthetic 512x512 frame per rank (262,144 rays = 64 chunks of 4096) driven the way
the reference's render() -> batchify_rays() drives the boundary.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from anerf_b200 import synthetic  # noqa: E402

H = W = 512
FOCAL = 500.0
CHUNK = 4096
N_SAMPLES, N_IMPORTANCE, N_JOINTS = 64, 128, 24
FLOP_PER_RAY = 256 * 1723648           # 441.25 MFLOP forward
METRIC = "rays/sec at 512^2 (64c+128f samples, 24 joints)"
WORKLOAD = "SURREAL-style bullet-time frame 512x512, 4096-ray chunks, 64+128 samples, 24 joints, 8x256 coarse+fine (BASELINE.json configs[1])"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '', 1).isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '', 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 3 + i and r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_args(**over):
    """The flags create_raycaster reads, SURREAL-style (reference: configs/surreal/surreal.txt + --N_importance 128)."""
    import tempfile
    tmp = tempfile.mkdtemp(prefix="anerf_bench_")
    os.makedirs(os.path.join(tmp, "exp"), exist_ok=True)
    d = dict(n_framecodes=None, use_cutoff=True, normalize_cutoff=False, cutoff_mm=500., ext_scale=0.001,
             cutoff_inputs=True, opt_cutoff=False, freq_schedule=False, init_freq=0., cut_to_dist=False,
             cutoff_shift=False, multires=7, i_embed=0, cutoff_bones=False, multires_bones=0, use_viewdirs=True,
             cutoff_viewdir=True, multires_views=4, N_importance=N_IMPORTANCE, netdepth=8, netwidth=256,
             opt_framecode=False, framecode_size=16, density_scale=1.0, single_net=False, lrate=5e-4, ft_path=None,
             basedir=tmp, expname="exp", no_reload=True, finetune=False, fix_layer=0, weight_decay=None,
             density_type="relu", softplus_shift=0., pts_tr_type="local", kp_dist_type="reldist", view_type="relray",
             bone_type="reldir", debug=True, perturb=0., N_samples=N_SAMPLES, raw_noise_std=0., ray_noise_std=0.,
             lindisp=False, nerf_type="nerf", cutoff_step=250, cutoff_rate=10., freq_schedule_step=50)
    d.update(over)
    return argparse.Namespace(**d)


def frame_inputs(frame_idx, n_frames_total=8):
    """Synthetic bullet-time frame: same pose, camera orbiting; all 262,144 pixels; per-ray replicated
    pose tensors as run_nerf.render_path builds them (run_nerf.py:84-90)."""
    ang = 2 * np.pi * frame_idx / n_frames_total
    sc = synthetic.make_scene(seed=0, n_rays=None, H=H, W=W, focal=FOCAL, n_joints=N_JOINTS, cam_angle=ang)
    N = sc["rays_o"].shape[0]
    rays = np.concatenate([sc["rays_o"], sc["rays_d"], np.zeros((N, 1), np.float32), np.ones((N, 1), np.float32)], 1)
    return dict(rays=rays, skts=sc["skts"], cyls=sc["cyls"], kps=sc["kps"], bones=sc["bones"], c2w=sc["c2w"], pose=sc["pose"])


def pick_cpu_threads(fn):
    """torch's CPU ops stop scaling long before 128 threads on this workload; probe a few thread counts on a
    small sample and keep the fastest (the count used is reported as `cores`)."""
    ncpu = os.cpu_count() or 1
    best, best_t = None, None
    for n in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(n)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def _data_attrs():
    import collections
    Skel = collections.namedtuple("Skel", ["joint_names", "joint_trees", "root_id"])
    return dict(skel_type=Skel(synthetic.SMPL_JOINT_NAMES, synthetic.SMPL_PARENTS, 0), near=0., far=1., n_views=1,
                joint_coords=np.tile(np.eye(3, dtype=np.float32), (1, N_JOINTS, 1, 1)))


def reference_renderer(device):
    """-> (fn(rays [n,8], skts, cyls, kps, bones) -> output dict, kind).  kind "reference": the UNMODIFIED reference
    (core.raycasters.create_raycaster + core.trainer.render, imported from /root/reference or from the copy under
    oracle/_ref that oracle/build_ref.py ships with the snapshot), its stock code path, none of our code on it.
    kind "port": the oracle restatement -- only when the reference sources are not on this machine."""
    import contextlib
    import io
    from oracle import ref_import
    sd0, sd1 = synthetic.make_net_weights(101), synthetic.make_net_weights(202)
    if ref_import.reference_available():
        ref_import.import_reference()
        from core.raycasters import create_raycaster as ref_create
        from core.trainer import render as ref_render
        with contextlib.redirect_stdout(io.StringIO()):
            _, rk, _, _, _, _ = ref_create(make_args(), _data_attrs())
        rc = rk["ray_caster"]
        rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in sd0.items()})
        rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in sd1.items()})
        rc.to(device).eval()

        def fn(rays, skts, cyls, kps, bones):
            with torch.no_grad():
                return ref_render(H, W, FOCAL, chunk=CHUNK, rays=(rays[:, 0:3], rays[:, 3:6]), kp_batch=kps, skts=skts, cyls=cyls,
                                  bones=bones, cams=None, subject_idxs=None, **rk)
        return fn, "reference"
    from oracle import anerf_oracle as orc
    s0 = {k: v.to(device) for k, v in orc.to_torch(sd0).items()}
    s1 = {k: v.to(device) for k, v in orc.to_torch(sd1).items()}
    cfg = orc.PathConfig()

    def fn(rays, skts, cyls, kps, bones):
        with torch.no_grad():
            return orc.render_rays(s0, s1, cfg, rays[:, 0:3], rays[:, 3:6], skts, cyls)
    return fn, "port"


def frame_sample(n_sample, device="cpu"):
    fr = frame_inputs(0)
    idx = np.sort(np.random.RandomState(0).choice(H * W, n_sample, replace=False))
    return [torch.as_tensor(np.ascontiguousarray(fr[k][idx])).to(device) for k in ("rays", "skts", "cyls", "kps", "bones")]


def run_reference_arm(opt):
    """The reference's own implementation of the path on the host cores: `core.trainer.render` of the unmodified
    reference (kind "reference"; the oracle port only if its sources are missing), all the threads torch can use well,
    each step a bounded sample of the frame sized so that the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn, kind = reference_renderer("cpu")
    probe_in = frame_sample(128)
    pick_cpu_threads(lambda: fn(*probe_in))
    t0 = time.perf_counter()
    fn(*frame_sample(256))
    rate = 256 / (time.perf_counter() - t0)
    budget_s = 150.0                                   # CPU seconds for warm-up + timed steps
    n_sample = int(min(opt.ref_rays, max(256, rate * budget_s / (opt.steps + opt.warmup))))
    if n_sample >= 256:
        n_sample = n_sample // 256 * 256
    args = frame_sample(n_sample)
    for _ in range(opt.warmup):
        fn(*args)
    t0 = time.perf_counter()
    for _ in range(opt.steps):
        fn(*args)
    dt = time.perf_counter() - t0
    v = n_sample * opt.steps / dt
    cores = torch.get_num_threads()
    what = ("core.trainer.render of the unmodified reference (PyTorch CPU, fp32)" if kind == "reference"
            else "oracle port of the reference's PyTorch path (reference sources not on this machine)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": opt.gpus, "steps": opt.steps,
            "warmup": opt.warmup, "ms_per_step": 1e3 * dt / opt.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{n_sample} rays of the frame per step"},
            "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": kind,
                             "sample": f"{n_sample} rays x {opt.steps} steps; {what}; host has {os.cpu_count()} cpus"},
            "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(n_sample=1024):
    """The reference's CPU path timed beside ours (rank 0, N = 1): one pass over `n_sample` rays after a warm-up."""
    fn, kind = reference_renderer("cpu")
    probe_in = frame_sample(128)
    pick_cpu_threads(lambda: fn(*probe_in))
    args = frame_sample(n_sample)
    fn(*args)                                        # warm-up
    t0 = time.perf_counter()
    fn(*args)
    dt = time.perf_counter() - t0
    what = "core.trainer.render of the unmodified reference" if kind == "reference" else "oracle port of the reference"
    return {"value": n_sample / dt, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{n_sample} rays of frame 0, one timed pass after one warm-up ({what}, PyTorch CPU fp32); thread count "
                      f"picked from {{8,16,32,64,{os.cpu_count()}}} by a 128-ray probe"}


def reference_gpu_sample(dev, n_chunks=6):
    """the reference's PyTorch path on the same B200 (fp32, eager, TF32 off), 4096-ray chunks of frame 0:
    2 warm-up chunks, then `n_chunks` timed ones."""
    from oracle import ref_import
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    fr = frame_inputs(0)
    with ref_import.reference_on_cuda() if ref_import.reference_available() else contextlib_null():
        fn, kind = reference_renderer(dev)
        chunks = []
        for c in range(2 + n_chunks):
            sl = slice((100 + 7 * c) * W, (100 + 7 * c) * W + CHUNK)        # rows spread over the figure
            chunks.append([torch.as_tensor(np.ascontiguousarray(fr[k][sl])).to(dev) for k in ("rays", "skts", "cyls", "kps", "bones")])
        for a in chunks[:2]:
            fn(*a)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a in chunks[2:]:
            fn(*a)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_chunks
    return {"value": CHUNK / dt, "unit": "rays/s", "kind": kind,
            "what": ("core.trainer.render of the unmodified reference" if kind == "reference" else "oracle port of the reference")
                    + ", CUDA tensors, fp32 eager, TF32 off",
            "sample": f"{n_chunks} timed 4096-ray chunks after 2 warm-up chunks", "ms_per_chunk": dt * 1e3}


class contextlib_null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def headline_parity(rc, kw, dev):
    """Errors of the benchmarked configuration on the 4096 rays of frame 0 stored in tests/golden/bench4096_*.npz with the
    outputs of the UNMODIFIED reference (fp32) and an fp64 evaluation of the same algorithm (oracle/make_golden.py)."""
    import ast
    z = np.load(os.path.join(ROOT, "tests", "golden", "bench4096_j24_s64_i128.npz"))
    c = ast.literal_eval(str(z["case"]))
    idx = np.sort(np.random.RandomState(0).choice(H * W, c["n_rays"], replace=False))
    fr = frame_inputs(0)
    sub = {k: torch.as_tensor(np.ascontiguousarray(fr[k][idx])).to(dev) for k in ("rays", "skts", "cyls", "kps", "bones")}
    o = rc(sub["rays"], kp_batch=sub["kps"], skts=sub["skts"], cyls=sub["cyls"], bones=sub["bones"], cams=None, subject_idxs=None, **kw)
    o = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in o.items()}
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    par = {"rays": int(c["n_rays"]), "reference": "tests/golden/bench4096_j24_s64_i128.npz (unmodified reference, fp32)"}
    for k in ("rgb0", "disp0", "acc0", "alpha0", "rgb_map", "disp_map", "acc_map"):
        par[k + "_rel"] = rel(o[k], z["ref_" + k].astype(np.float64))
    ref, ref64 = z["ref_rgb_map"].astype(np.float64), z["ref64_rgb_map"].astype(np.float64)
    scale = np.abs(ref).max()
    err = np.abs(o["rgb_map"] - ref).max(1) / scale
    cond = np.abs(ref - ref64).max(1) / scale
    par.update({
        "ref_fp32_vs_fp64_rgb_map_rel": float(cond.max()), "ref_fp32_vs_fp64_rgb0_rel": rel(z["ref_rgb0"].astype(np.float64), z["ref64_rgb0"].astype(np.float64)),
        "rays_above_1e-4": int((err > 1e-4).sum()), "ref_fp32_vs_fp64_rays_above_1e-4": int((cond > 1e-4).sum()),
        "rgb_map_rel_p50": float(np.percentile(err, 50)), "rgb_map_rel_p99": float(np.percentile(err, 99)),
        "rgb_map_rel_p99.9": float(np.percentile(err, 99.9)),
        "rgb_map_rel_excluding_rays_the_reference_cannot_resolve": float(err[cond <= 5e-5].max()),
        "rays_the_reference_cannot_resolve": int((cond > 5e-5).sum()),
        "rgb_map_psnr_db": float(-10 * np.log10(max(float(((o["rgb_map"] - ref) ** 2).mean()), 1e-30))),
        "note": "max-norm relative errors, NOT conditioned on the kernel's sample positions; a ray whose coarse weights sum to "
                "~1e-4 has an importance pdf dominated by its 1e-5 floor, so the reference's own fp32 arithmetic does not resolve "
                "it to 1e-4 either (ref_fp32_vs_fp64_*)"})
    return par


def training_probe(dev, rank, world, n_rays=3072, n_importance=16, steps=5, n_poses=256):
    """Training step through the boundary, weak scaling (n_rays per rank), two routes:
    `fused_step` (primary): anerf_b200.train.FusedTrainStep -- forward that keeps its activations (anerf_render_fwd_train),
        loss seed, per-pass backward from the kept activations (anerf_render_bwd_saved), one gradient all-reduce,
        FusedAdam; pose refinement ON with the transforms coming
        from anerf_b200.pose_opt.PoseOptLayer once per POSE (n_poses poses, rays -> pose index; the reference's
        image_batching: N_sample_images 256, configs/mixamo/mixamo.txt) and d/d skts reduced per pose in the backward;
    `autograd_route`: create_raycaster's train kwargs in .train() mode with per-RAY transforms that require grad, torch loss,
        loss.backward(), gradient all-reduce, optimizer.step() -- what the reference's Trainer does through the boundary."""
    import contextlib
    import io
    from anerf_b200 import parallel
    from anerf_b200.pose_opt import PoseOptLayer
    from anerf_b200.raycasters import create_raycaster
    from anerf_b200.train import FusedTrainStep
    data_attrs = _data_attrs()
    args = make_args(N_importance=n_importance, perturb=1.0, raw_noise_std=1.0)
    with contextlib.redirect_stdout(io.StringIO()):
        rk_train, rk_test, _, grad_vars, optimizer, _ = create_raycaster(args, data_attrs, device=dev)
    rc = rk_test["ray_caster"]
    rc.network.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(101).items()})
    rc.network_fine.load_state_dict({k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202).items()})
    holder = rk_train["ray_caster"].train()
    sc = synthetic.make_scene(seed=0, n_rays=n_rays, H=H, W=W, focal=FOCAL, n_joints=N_JOINTS, pixel_offset=rank)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    N = n_rays
    rays = torch.cat([t(sc["rays_o"]), t(sc["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev),
                      torch.nn.functional.normalize(t(sc["rays_d"]), dim=-1)], 1)
    skts0, kps, cyls, bones = t(sc["skts"]), t(sc["kps"]), t(sc["cyls"]), t(sc["bones"])
    kw = {k: v for k, v in rk_train.items() if k not in ("ray_caster", "use_viewdirs")}
    target = torch.rand(N, 3, device=dev)
    flat = [None]
    # pose layer: n_poses copies of the scene's pose (slightly perturbed), every ray assigned to one of them
    pose = sc["pose"]
    rng = np.random.RandomState(1)
    layer = PoseOptLayer(torch.as_tensor(np.tile(pose["kps"][None], (n_poses, 1, 1))),
                         torch.as_tensor(np.tile(pose["bones"][None], (n_poses, 1, 1)) + 0.01 * rng.randn(n_poses, N_JOINTS, 3).astype(np.float32)),
                         torch.as_tensor(synthetic.humanoid_rest_pose()[None]), use_rot6d=True, parents=synthetic.SMPL_PARENTS, root_id=0).to(dev)
    pose_optimizer = torch.optim.Adam(layer.parameters(), lr=5e-4)
    kp_idx = np.sort(rng.randint(0, n_poses, size=N))
    fused = FusedTrainStep(rc, optimizer, loss_fn="L1", use_background=True, world=world)
    it = [0]

    def step_fused():
        (kps_p, bones_p, skts_p, _, _), pose_idx = layer.forward_poses(kp_idx)
        fused(rays, target, kp_batch=kps_p, skts=skts_p, cyls=cyls, bones=bones_p, cams=None, pose_idx=pose_idx, **kw)
        it[0] += 1
        if it[0] % 20 == 0:                                  # opt_pose_step (configs/mixamo/mixamo.txt): gradients accumulate in between
            if world > 1:
                parallel.allreduce_gradients(list(layer.parameters()), world)
            pose_optimizer.step()
            pose_optimizer.zero_grad()

    def step_autograd():
        optimizer.zero_grad(set_to_none=True)
        skts = skts0.clone().requires_grad_(True)           # pose refinement: the bone transforms carry gradient
        out = holder(rays, kp_batch=kps, skts=skts, cyls=cyls, bones=bones, cams=None, subject_idxs=None, **kw)
        loss = ((out["rgb_map"] - target) ** 2).mean() + ((out["rgb0"] - target) ** 2).mean()
        loss.backward()
        flat[0] = parallel.allreduce_gradients(grad_vars, world, flat[0])
        optimizer.step()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return parallel.max_over_ranks(e0.elapsed_time(e1) / steps, dev)
    ms_auto = timed(step_autograd)
    for p in grad_vars:
        p.grad = None
    ms = timed(step_fused)
    rows = N * (N_SAMPLES + N_SAMPLES + n_importance)
    return {"metric": "training rays/sec (fwd + loss + bwd + grad all-reduce + Adam, pose refinement on)", "value": N * world / (ms * 1e-3),
            "unit": "rays/s", "ms_per_step": ms, "route": "FusedTrainStep (forward keeps its activations: 3 GEMM passes per step), per-pose transforms from PoseOptLayer (%d poses), d/dskts reduced per pose" % n_poses,
            "activations_kept": getattr(rc, "_state_buf", None) is not None and rc._state_buf[1] is not None,
            "rays_per_rank": N, "samples": f"{N_SAMPLES}+{n_importance}", "pose_grad": True,
            "allreduce_bytes": int(sum(p.numel() for p in grad_vars) * 4) if world > 1 else 0,
            "algorithmic_tflops": 3 * rows * world * 1723648 / (ms * 1e-3) / 1e12,
            "autograd_route": {"ms_per_step": ms_auto, "value": N * world / (ms_auto * 1e-3),
                               "what": "RayCaster.train() + torch loss + loss.backward() + all-reduce + FusedAdam.step(), per-ray transforms [N,24,4,4] requiring grad"},
            "gemm_engine": os.environ.get("ANERF_TRAIN_GEMM", "tc") + " (default tc: fp16 hi/lo operands with per-matrix power-of-two scales on tcgen05)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-rays", type=int, default=2048, help="rays per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--format", type=int, default=None, help="operand format override: 0 fp16x3, 1 bf16x3")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement")
    opt = ap.parse_args()
    if opt.impl == "reference":
        return run_reference_arm(opt)

    from anerf_b200 import _lib, build, parallel
    from anerf_b200.raycasters import create_raycaster, batchify_rays
    build.build()
    rank, world, local = parallel.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (anerf_b200 has no CPU path)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if opt.format is not None:
        os.environ["ANERF_OPERAND_FORMAT"] = str(opt.format)

    data_attrs = _data_attrs()
    args = make_args()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        _, rk, _, _, _, _ = create_raycaster(args, data_attrs, device=dev)
    rc = rk["ray_caster"]
    sd0 = {k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(101).items()}
    sd1 = {k: torch.as_tensor(v) for k, v in synthetic.make_net_weights(202).items()}
    rc.network.load_state_dict(sd0)
    rc.network_fine.load_state_dict(sd1)
    rc.eval()
    kw = {k: v for k, v in rk.items() if k not in ("ray_caster", "use_viewdirs")}

    # two frames per rank, alternated, resident in HBM (402 MB of per-ray skts each: larger than the 126 MB L2)
    my_frames = [rank * 2, rank * 2 + 1]
    host = [frame_inputs(f, n_frames_total=2 * world * 4) for f in my_frames]
    devf = [{k: torch.as_tensor(v).to(dev) for k, v in fr.items() if k not in ("c2w", "pose")} for fr in host]
    n_rays = H * W
    n_chunks = (n_rays + CHUNK - 1) // CHUNK

    def render_frame(fr):
        return batchify_rays(fr["rays"], CHUNK, ray_caster=rc, kp_batch=fr["kps"], skts=fr["skts"], cyls=fr["cyls"],
                             bones=fr["bones"], cams=None, subject_idxs=None, **kw)

    def step(i):
        out = render_frame(devf[i % 2])
        pix = torch.cat([out["rgb_map"], out["disp_map"][:, None], out["acc_map"][:, None]], 1)
        if world > 1:                      # pixels of every rank's frame to rank 0 (bullet-time video)
            bufs = [torch.empty_like(pix) for _ in range(world)] if rank == 0 else None
            torch.distributed.gather(pix, bufs, dst=0)
        return out

    for i in range(max(opt.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(opt.steps):
        out = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        torch.distributed.barrier()
    clocks = sampler.stop()
    ms = parallel.max_over_ranks(ms, dev)
    rays_total = n_rays * opt.steps * world
    value = rays_total / (ms * 1e-3)

    # ---- dominant kernel: the fused render kernel, timed per launch with CUDA events on its stream -----
    fr = devf[0]
    sl = slice(0, CHUNK)
    call = lambda: rc(fr["rays"][sl], kp_batch=fr["kps"][sl], skts=fr["skts"][sl], cyls=fr["cyls"][sl], bones=fr["bones"][sl],
                      cams=None, subject_idxs=None, **kw)
    evs = []
    for c in range(n_chunks):
        sl = slice(c * CHUNK, (c + 1) * CHUNK)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        call()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    launch_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    peak, peak_src = load_peaks()
    achieved = FLOP_PER_RAY * CHUNK / (launch_ms * 1e-3) / 1e12
    # DRAM traffic of one launch cannot be measured outside a profiler: it is the figure of the committed ncu capture
    # (profiles/fused_kernel_traffic.json says which build and command it comes from)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), f"{tj.get('build')}; {tj.get('source')}"
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "anerf_fused_kernel",
                "launch_ms": launch_ms, "algorithmic_flop_per_launch": FLOP_PER_RAY * CHUNK,
                "note": "achieved = algorithmic fp32-equivalent FLOPs (441.25 MFLOP/ray, BASELINE.md). Each product is issued as "
                        "3 fp16 MMAs (lo*hi+hi*lo+hi*hi); with feature_linear folded into the views layer, the view branch "
                        "contracted per ray (views layer K 904 -> 384) and K padding the kernel executes 770,048 MACs x 3 "
                        "per sample against 861,824 algorithmic, so tensor-pipe work is 2.68x `achieved`: issued_frac below",
                "issued_frac": 2.6805 * achieved / peak}

    # ---- e2e: the C-ABI call with HOST buffers, H2D and D2H inside the timed region ------------------------
    e2e = None
    if rank == 0 or world > 1:
        plan = rc._get_plan()
        p0, p1 = rc._packed_image('network'), rc._packed_image('network_fine')
        hp = host[0]
        pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()
        h_rays, h_skts, h_cyls = pin(hp["rays"]), pin(hp["skts"]), pin(hp["cyls"])
        opts_full = _lib.make_opts(CHUNK, N_SAMPLES, N_IMPORTANCE, tau_pts=20., tau_views=20., cutoff_pts=0.5, cutoff_views=0.5)
        f = lambda *s: torch.empty(*s, dtype=torch.float32).pin_memory()
        Sf = N_SAMPLES + N_IMPORTANCE
        h_out = dict(rgb_map=f(n_rays, 3), disp_map=f(n_rays), acc_map=f(n_rays), alpha=f(n_rays, Sf), rgb0=f(n_rays, 3),
                     disp0=f(n_rays), acc0=f(n_rays), alpha0=f(n_rays, N_SAMPLES))

        opts_frame = _lib.make_opts(n_rays, N_SAMPLES, N_IMPORTANCE, tau_pts=20., tau_views=20., cutoff_pts=0.5, cutoff_views=0.5)

        def e2e_frame(out=h_out):
            # ONE C-ABI call per frame: 64 chunks of 4096 rays, each chunk's host->device and device->host copies overlapping
            # the kernels of its neighbours (three streams, two arenas inside the library)
            _lib.render_fwd_host_chunked(plan, p0, p1, opts_frame, CHUNK, h_rays, h_skts, h_cyls, None, out=out)
        e2e_frame()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(opt.steps, 3))
        for _ in range(n_e2e):
            e2e_frame()
        torch.cuda.synchronize()
        dt_ms = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
        h2d = n_rays * (8 + N_JOINTS * 16 + 5) * 4
        d2h = n_rays * (3 + 1 + 1 + Sf + 3 + 1 + 1 + N_SAMPLES) * 4
        e2e = {"value": n_rays * n_e2e * world / (dt_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h,
               "api": "anerf_render_fwd_host_chunked (C ABI): one call per frame with pinned host buffers, 4096-ray chunks, "
                      "copies overlapped with the kernels; all eight outputs copied back"}
        # what run_nerf.render_path actually reads back (rgb, disp, acc): per-sample alpha not copied
        slim = {k: h_out[k] for k in ("rgb_map", "disp_map", "acc_map")}
        e2e_frame(slim)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_frame(slim)
        torch.cuda.synchronize()
        dt2 = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
        e2e["render_path_outputs_only"] = {"value": n_rays * n_e2e * world / (dt2 * 1e-3), "d2h_bytes_per_step": n_rays * 5 * 4}

    # ---- the frame API (rays generated in the kernels from the camera, pose passed once per frame):
    # host -> device per frame = camera + one pose (1.6 KB instead of 416 MB), all eight outputs back to pinned host memory
    e2e_frame = None
    if e2e is not None:
        hp = host[0]
        c2w = np.asarray(hp["c2w"], np.float32)
        h_skts, h_cyl = torch.as_tensor(hp["pose"]["skts"]).pin_memory(), torch.as_tensor(hp["pose"]["cyl"]).pin_memory()
        d_out = {k: torch.empty(v.shape, dtype=torch.float32, device=dev) for k, v in h_out.items()}

        def frame_api():
            o = rc.render_frame(H, W, FOCAL, c2w, h_skts.to(dev, non_blocking=True)[None], h_cyl.to(dev, non_blocking=True)[None],
                                chunk=CHUNK, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, preproc_kwargs=kw["preproc_kwargs"], out=d_out)
            for k, v in h_out.items():
                v.copy_(o[k], non_blocking=True)
        frame_api()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            frame_api()
        torch.cuda.synchronize()
        dt_ms = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
        ref_px = render_frame(devf[0])
        e2e_frame = {"value": n_rays * n_e2e * world / (dt_ms * 1e-3), "unit": "rays/s",
                     "h2d_bytes_per_step": int(12 * 4 + h_skts.numel() * 4 + h_cyl.numel() * 4), "d2h_bytes_per_step": d2h,
                     "api": "RayCaster.render_frame -> anerf_render_frame (C ABI), camera + one pose in, all outputs to pinned host memory",
                     "max_abs_diff_vs_explicit_rays": {k: float((h_out[k].to(dev) - ref_px[k]).abs().max()) for k in ("rgb_map", "acc_map")}}

    # ---- secondary: one training step per rank (Mixamo-style N_rand 3072, 64 + 16 samples, pose
    # gradient on): fused forward + CUDA backward + the single flat gradient all-reduce + Adam, through the boundary
    training = None
    if not opt.no_train:
        training = training_probe(dev, rank, world)

    # ---- secondary: 256^3 density grid for mesh extraction (BASELINE.json configs[4]), voxel slabs over the ranks
    mesh_grid = None
    if not opt.no_train:
        from anerf_b200 import mesh
        pose = host[0]["pose"]
        kps_m, skts_m = torch.as_tensor(pose["kps"]).to(dev)[None], torch.as_tensor(pose["skts"]).to(dev)[None]
        for _ in range(2):
            mesh.density_grid_sharded(rc, kps_m, skts_m, 1.8, 255, rank, world)
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        mesh.density_grid_sharded(rc, kps_m, skts_m, 1.8, 255, rank, world)
        m1.record()
        torch.cuda.synchronize()
        mms = parallel.max_over_ranks(m0.elapsed_time(m1), dev)
        mesh_grid = {"metric": "voxels/s, 256^3 density grid (24 joints, 8x256 trunk, fine network)", "value": 256 ** 3 / (mms * 1e-3),
                     "ms": mms, "algorithmic_tflops": 256 ** 3 * 1.3604e6 / (mms * 1e-3) / 1e12}

    if rank != 0:
        return
    cpu_base, parity = None, None
    if world == 1:
        try:
            parity = headline_parity(rc, kw, dev)
        except Exception as e:  # noqa: BLE001
            parity = {"error": repr(e)[:300]}
    if not opt.no_cpu_baseline and world == 1:
        cpu_base = cpu_baseline_sample(1024)
    ref_gpu = None
    if not opt.no_cpu_baseline and world == 1:
        try:
            ref_gpu = reference_gpu_sample(dev)
        except Exception as e:  # noqa: BLE001
            ref_gpu = {"error": repr(e)[:300]}
    line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": opt.steps, "warmup": max(opt.warmup, 3),
            "ms_per_step": ms / opt.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 via fp16 hi/lo split on tcgen05 (3 MMAs per product, fp32 accumulate)" if rc._operand_format == 0
            else "f32 via bf16 hi/lo split on tcgen05 (3 MMAs per product, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": n_rays, "chunks_per_step": n_chunks,
                       "parallelism": f"frame-parallel x{world}, gather of [rays,5] pixels to rank 0" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2: 402 MB of per-ray skts per frame, two frames alternated"},
            "clocks": clocks, "gpu_launches": 2 * n_chunks * opt.steps, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu_base, "reference_gpu": ref_gpu, "parity": parity, "training": training,
            "e2e_frame_api": e2e_frame, "mesh_grid": mesh_grid}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    try:
        main()
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
