"""Density grids for mesh extraction, sharded over ranks.

`RayCaster.render_mesh_density` (the reference's entry point, core/raycasters.py:579-595) queries the whole
[res+1]^3 grid on one GPU.  `density_grid_sharded` gives every rank a contiguous slab of the flattened grid,
runs the same fused density kernel on it and gathers the slabs on rank 0, which returns the grid in the
reference's layout (meshgrid 'xy' then transpose(1, 0)) ready for marching cubes.
"""
import numpy as np
import torch

from . import parallel


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
