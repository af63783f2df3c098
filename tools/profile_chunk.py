"""A short run of the benchmark configuration for ncu: N chunk launches of 4096 rays (64+128, 24 joints)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anerf_b200 import _lib, synthetic  # noqa: E402

n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
fmt = int(os.environ.get("ANERF_OPERAND_FORMAT", "0"))
dev = torch.device("cuda")
sc = synthetic.make_scene(seed=0, n_rays=None, H=512, W=512, focal=500., n_joints=24)
N = 4096
off = 128 * 512
t = lambda a: torch.as_tensor(np.ascontiguousarray(a[off:off + N])).to(dev)
rays = torch.cat([t(sc["rays_o"]), t(sc["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev)], 1).contiguous()
plan = _lib.Plan(24, 8, 256, (4,), 0, 0, fmt)
p0 = plan.pack({k: torch.as_tensor(v).to(dev) for k, v in synthetic.make_net_weights(101).items()})
p1 = plan.pack({k: torch.as_tensor(v).to(dev) for k, v in synthetic.make_net_weights(202).items()})
opts = _lib.make_opts(N, 64, 128)
skts, cyls = t(sc["skts"]), t(sc["cyls"])
for i in range(n_launch):
    out = _lib.render_fwd(plan, p0, p1, opts, rays, skts, cyls)
torch.cuda.synchronize()
ts = []
for i in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = _lib.render_fwd(plan, p0, p1, opts, rays, skts, cyls)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("chunk ms", min(ts), "all", " ".join(f"{t:.3f}" for t in ts), "acc mean", float(out["acc_map"].mean()))
