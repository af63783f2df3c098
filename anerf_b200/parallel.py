"""One process per GPU: how the ray-marching path shards.

Rays are independent given the weights and the pose, so nothing on the data path needs a collective:
  * rendering  -- frames are dealt round-robin to ranks (`frames_for_rank`); inside a frame the unit of
                  work stays the reference's 4096-ray chunk, because the near/far repair of rays that
                  miss the bounding cylinder is a chunk-wide mean (core/utils/ray_utils.py:328-342);
                  finished pixels [rays, 5] (rgb, disp, acc) are gathered on rank 0 (`gather_pixels`);
  * mesh grids -- voxels are split into contiguous slabs per rank (`slab_for_rank`) and the densities
                  gathered on rank 0 (`gather_slabs`).
  * training   -- each rank takes an equal, contiguous share of the step's N_rand rays (`rays_for_rank`), runs
                  forward + backward locally, and the gradients of all trainable tensors are averaged with ONE
                  all-reduce of a flat fp32 buffer (`allreduce_gradients`; ~6.9 MB for two 8x256 networks), the
                  only exchange step of the path.  Gradients of pose tensors that live outside the caster
                  (PoseOptLayer) go into the same buffer when their parameters are passed along.
This replaces the reference's single-process nn.DataParallel (core/raycasters.py:157).  The backend is
NCCL on GPUs; the same functions run on gloo/CPU tensors in the tests (the gather is plain data movement).
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                **({"device_id": torch.device("cuda", local)} if backend == "nccl" else {}))
    return rank, world, local


def frames_for_rank(n_frames, rank, world):
    """Frame f is rendered by rank f mod world."""
    return list(range(rank, n_frames, world))


def slab_for_rank(n, rank, world):
    """Contiguous [start, stop) slab of n voxels (or chunks) for this rank; sizes differ by at most one."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _comm_device():
    """Where collectives take their tensors: the current CUDA device under NCCL, the host under gloo."""
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_pixels(local_frames, n_frames, rank, world, dst=0):
    """local_frames: {frame index: tensor [rays_f, C]} rendered by this rank (fp32).  Returns the list of all
    n_frames tensors on rank `dst` (None elsewhere).  One gather per round of `world` frames.

    Frames may have DIFFERENT ray counts (per-frame valid-pixel crops: frames.valid_pixels / render_frame(pixel_idx=...))
    and a rank may own no frame at all (n_frames < world): every round first exchanges the [rows, channels] of each
    rank's frame (one tiny all_gather), pads the payloads to the round's largest frame and allocates every buffer --
    placeholders included -- on the communication device with that padded shape."""
    if world == 1:
        return [local_frames[f] for f in range(n_frames)]
    out = [None] * n_frames if rank == dst else None
    dev = _comm_device()
    rounds = (n_frames + world - 1) // world
    for r in range(rounds):
        f = r * world + rank
        mine = local_frames.get(f)
        shape = torch.tensor([0, 0] if mine is None else [mine.shape[0], mine.reshape(mine.shape[0], -1).shape[1]],
                             dtype=torch.int64, device=dev)
        shapes = [torch.empty_like(shape) for _ in range(world)]
        dist.all_gather(shapes, shape)
        shapes = [tuple(int(x) for x in s.tolist()) for s in shapes]
        rows, ch = max(s[0] for s in shapes), max(s[1] for s in shapes)
        pad = torch.zeros((rows, ch), dtype=torch.float32, device=dev)
        if mine is not None:
            pad[:mine.shape[0]] = mine.reshape(mine.shape[0], -1).to(dev, torch.float32)
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst)
        if rank == dst:
            for src in range(world):
                g = r * world + src
                if g < n_frames:
                    out[g] = bufs[src][:shapes[src][0], :shapes[src][1]]
    return out


def gather_slabs(local, n, rank, world, dst=0):
    """local: this rank's slab [stop-start, ...] of a length-n array split by `slab_for_rank`.
    Returns the full array on rank `dst` (None elsewhere)."""
    if world == 1:
        return local
    sizes = [slab_for_rank(n, r, world) for r in range(world)]
    longest = max(b - a for a, b in sizes)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:sizes[r][1] - sizes[r][0]] for r in range(world)], 0)


def rays_for_rank(n_rays, rank, world):
    """Contiguous, equally sized share of a training batch: every rank gets n_rays // world rays (the loss is a
    mean over rays, so equal shares make mean-of-means == the global mean; a remainder is dropped like a short
    last batch)."""
    per = n_rays // world
    return rank * per, (rank + 1) * per


def allreduce_gradients(params, world=None, flat=None, average=True):
    """Sum (average=False) or average the .grad of `params` over the ranks in ONE collective launch; a missing .grad
    counts as zero, so every rank reduces the same layout.

    NCCL: the per-tensor all-reduces are issued inside one coalescing group (ncclGroupStart/End), i.e. one fused NCCL
    kernel working on the gradient tensors in place -- no flat staging buffer, none of the ~100 small copy kernels
    that packing and unpacking one would take.  Other backends (gloo in the CPU tests): one all-reduce of a flat fp32
    buffer (returned, reusable as `flat`).  With average=False the 1/world can be folded into the optimizer step
    (FusedAdam.step(grad_scale=1/world))."""
    params = [p for p in params if p.requires_grad]
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1 or not params:
        return flat
    if dist.get_backend() == "nccl":
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in params]
        try:
            from torch.distributed.distributed_c10d import _coalescing_manager
            with _coalescing_manager(device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM)
            done = True
        except Exception:  # noqa: BLE001  (private API: fall back to the flat buffer below)
            done = False
        if done:
            if average:
                torch._foreach_mul_(grads, 1.0 / world)
            return flat
    n = sum(p.numel() for p in params)
    dev = params[0].device
    if flat is None or flat.numel() != n or flat.device != dev:
        flat = torch.empty(n, dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        k = p.numel()
        if p.grad is None:
            flat[off:off + k].zero_()
        else:
            flat[off:off + k].copy_(p.grad.reshape(-1))
        off += k
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat.mul_(1.0 / world)
    off = 0
    for p in params:
        k = p.numel()
        g = flat[off:off + k].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += k
    return flat


def max_over_ranks(value, device):
    """Device-timed milliseconds -> max over ranks (the job's time)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
