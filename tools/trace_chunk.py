"""Timeline of CTA 0 for one benchmark chunk (debug aid): prints per-layer gaps of the MMA thread and worker groups."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anerf_b200 import _lib, synthetic
dev = torch.device("cuda")
sc = synthetic.make_scene(seed=0, n_rays=None, H=512, W=512, focal=500., n_joints=24)
N = 4096; off = 128 * 512
t = lambda a: torch.as_tensor(np.ascontiguousarray(a[off:off + N])).to(dev)
rays = torch.cat([t(sc["rays_o"]), t(sc["rays_d"]), torch.zeros(N, 1, device=dev), torch.ones(N, 1, device=dev)], 1).contiguous()
plan = _lib.Plan(24, 8, 256, (4,), 0, 0, 0)
p0 = plan.pack({k: torch.as_tensor(v).to(dev) for k, v in synthetic.make_net_weights(101).items()})
p1 = plan.pack({k: torch.as_tensor(v).to(dev) for k, v in synthetic.make_net_weights(202).items()})
opts = _lib.make_opts(N, 64, 128)
skts, cyls = t(sc["skts"]), t(sc["cyls"])
for _ in range(2):
    _lib.render_fwd(plan, p0, p1, opts, rays, skts, cyls)
buf = torch.zeros(3 * 1024, dtype=torch.int64, device=dev)
_lib.load().anerf_debug_set_trace(buf.data_ptr())
_lib.render_fwd(plan, p0, p1, opts, rays, skts, cyls)
torch.cuda.synchronize()
_lib.load().anerf_debug_set_trace(None)
b = buf.cpu().numpy().reshape(3, 1024)
t0 = None
for s in range(3):
    n = int(b[s, 1023]); ev = [(int(x >> 48), int(x & 0xFFFFFFFFFFFF)) for x in b[s, :n]]
    if t0 is None: t0 = ev[0][1]
    print(f"--- stream {s}: {n} events")
    prev = None; line = []
    for tag, c in ev[:260]:
        line.append(f"{tag}@{c - t0}" + (f"(+{c - prev})" if prev is not None else ""))
        prev = c
    print(" ".join(line))
