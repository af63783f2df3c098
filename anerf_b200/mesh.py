"""Density grids for mesh extraction, sharded over ranks.

`RayCaster.render_mesh_density` (the reference's entry point, core/raycasters.py:579-595) queries the whole
[res+1]^3 grid on one GPU.  `density_grid_sharded` gives every rank a contiguous slab of the flattened grid,
runs the same fused density kernel on it and gathers the slabs on rank 0, which returns the grid in the
reference's layout (meshgrid 'xy' then transpose(1, 0)) ready for marching cubes.
"""
import numpy as np
import torch

from . import _lib, parallel


def grid_points(kps, radius, res, start=0, stop=None):
    """World points [start, stop) of the flattened reference grid (np.meshgrid(t, t, t) order, 'xy' indexing)
    around kps[0, 0] as explicit tensors.  The product path no longer uses this (the kernel generates the points,
    `RayCaster.render_mesh_density`); tests compare the two."""
    n1 = res + 1
    total = n1 ** 3
    stop = total if stop is None else stop
    idx = torch.arange(start, stop, device=kps.device, dtype=torch.int64)
    # flat index -> (a, b, c) of np.meshgrid(t, t, t)[..., :] with shape [n1(y), n1(x), n1(z), 3]
    a, rem = idx // (n1 * n1), idx % (n1 * n1)
    b, c = rem // n1, rem % n1
    t = torch.linspace(-radius, radius, n1, device=kps.device, dtype=torch.float64)
    pts = torch.stack([t[b], t[a], t[c]], -1).float()        # 'xy' indexing: x varies along axis 1, y along axis 0
    return pts + kps[0, 0].float()


@torch.no_grad()
def density_grid_sharded(ray_caster, kps, skts, radius=1.0, res=255, rank=0, world=1, dst=0):
    """-> [res+1]^3 raw densities on rank `dst` (None elsewhere).  kps [1,J,3], skts [1,J,4,4] on this rank's GPU."""
    n1 = res + 1
    total = n1 ** 3
    a, b = parallel.slab_for_rank(total, rank, world)
    # the slab's points are generated inside the density kernel (anerf_density_grid): algorithmic HBM traffic only
    sig = ray_caster.render_mesh_density(kps, skts, None, radius=radius, res=res, first=a, count=b - a).reshape(-1, 1)
    full = parallel.gather_slabs(sig, total, rank, world, dst=dst)
    if full is None:
        return None
    return full.reshape(n1, n1, n1).transpose(1, 0)


# ------------------------------------------------------------------------------------------------
# density grid -> mesh on the device (reference: run_render.render_mesh, run_render.py:969-986)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def marching_cubes(volume, threshold):
    """Isosurface of a CUDA fp32 volume [n0,n1,n2] (any strides: the transposed view render_mesh_density returns is
    taken as it is) -> (vertices [V,3] fp32 in index coordinates like `mcubes.marching_cubes`, triangles [F,3] int64).
    Two HBM-bound kernels through the C ABI (anerf_mc_count / anerf_mc_emit) around a scan; duplicate vertices are welded
    by the id of the volume edge they lie on.  A cell corner is inside when value > threshold; triangle normals point
    towards lower values (out of the body for a density).  Parity with PyMCubes is unpinned (mc_table.py)."""
    if volume.device.type != 'cuda' or volume.dtype != torch.float32 or volume.dim() != 3:
        raise RuntimeError("anerf_b200.marching_cubes: fp32 CUDA volume [n0,n1,n2] expected (no CPU path)")
    n0, n1, n2 = volume.shape
    s0, s1, s2 = volume.stride()
    dev = volume.device
    cells = (n0 - 1) * (n1 - 1) * (n2 - 1)
    with torch.cuda.device(dev):
        counts = torch.empty(cells, dtype=torch.int32, device=dev)
        _lib.check(_lib.load().anerf_mc_count(_lib._ptr(volume), n0, n1, n2, s0, s1, s2, float(threshold), _lib._ptr(counts), _lib._stream()))
        incl = torch.cumsum(counts, 0, dtype=torch.int64)
        total = int(incl[-1])                                   # the one host sync: the mesh size
        if total == 0:
            return torch.zeros(0, 3, device=dev), torch.zeros(0, 3, dtype=torch.int64, device=dev)
        offsets = (incl - counts).contiguous()
        verts = torch.empty(total, 3, 3, dtype=torch.float32, device=dev)
        keys = torch.empty(total, 3, dtype=torch.int64, device=dev)
        _lib.check(_lib.load().anerf_mc_emit(_lib._ptr(volume), n0, n1, n2, s0, s1, s2, float(threshold), _lib._ptr(offsets),
                                             _lib._ptr(verts), _lib._ptr(keys), _lib._stream()))
        uniq, inv = torch.unique(keys.reshape(-1), return_inverse=True)
        vertices = torch.empty(uniq.shape[0], 3, dtype=torch.float32, device=dev)
        vertices[inv] = verts.reshape(-1, 3)                   # duplicates carry identical coordinates
    return vertices, inv.reshape(-1, 3)


def export_ply(path, vertices, triangles):
    """Binary little-endian PLY (what trimesh's `mesh.export('*.ply')` writes: float x/y/z, uchar-counted int faces)."""
    v = np.ascontiguousarray(torch.as_tensor(vertices).detach().cpu().numpy(), dtype='<f4')
    f = np.ascontiguousarray(torch.as_tensor(triangles).detach().cpu().numpy(), dtype='<i4')
    header = ("ply\nformat binary_little_endian 1.0\ncomment anerf_b200\n"
              f"element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n"
              f"element face {len(f)}\nproperty list uchar int vertex_indices\nend_header\n")
    rec = np.empty(len(f), dtype=[('n', 'u1'), ('idx', '<i4', (3,))])
    rec['n'] = 3
    rec['idx'] = f
    with open(path, 'wb') as fh:
        fh.write(header.encode('ascii'))
        fh.write(v.tobytes())
        fh.write(rec.tobytes())


@torch.no_grad()
def render_mesh(basedir, render_kwargs, tensor_data, chunk=1024, radius=1.80, res=255, threshold=10., rank=0, world=1):
    """run_render.render_mesh (run_render.py:969-986) with everything after the network on the device too: density grid
    (voxel slabs over the ranks when world > 1) -> max(raw, 0) -> marching cubes -> vertices / res - 0.5 -> .ply per pose.
    Returns the list of (vertices, triangles) of the poses this rank wrote (rank 0 writes)."""
    import os
    ray_caster = render_kwargs['ray_caster']
    os.makedirs(os.path.join(basedir, 'meshes'), exist_ok=True)
    kps, skts = tensor_data['kp'], tensor_data['skts']
    out = []
    for i in range(len(kps)):
        if world > 1:
            raw = density_grid_sharded(ray_caster, kps[i:i + 1], skts[i:i + 1], radius, res, rank, world)
        else:
            raw = ray_caster(kps=kps[i:i + 1], skts=skts[i:i + 1], bones=None, radius=radius, render_kwargs=render_kwargs.get('preproc_kwargs'),
                             res=res, netchunk=chunk, fwd_type='mesh')
        if raw is None:
            continue
        vertices, triangles = marching_cubes(raw.clamp_min(0.), threshold)
        vertices = vertices / res - .5
        export_ply(os.path.join(basedir, 'meshes', f'{i:03d}.ply'), vertices, triangles)
        out.append((vertices, triangles))
    return out
