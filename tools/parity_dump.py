"""Dump the kernel's outputs (with taps) on N rays of the benchmark frame, for offline error analysis against the oracle:
    python tools/parity_dump.py [n_rays] [out_prefix]       (needs a B200; writes <prefix>_fmt{0,1}.npz)
The rays are a seeded random subset of frame 0 of bench.py (one chunk, so the near/far repair sees the same chunk)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from anerf_b200 import _lib, build, synthetic  # noqa: E402


def bench_rays(n):
    fr = bench.frame_inputs(0)
    idx = np.sort(np.random.RandomState(0).choice(bench.H * bench.W, n, replace=False))
    return idx, {k: np.ascontiguousarray(fr[k][idx]) for k in ("rays", "skts", "cyls")}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    prefix = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "parity_dump")
    build.build()
    dev = torch.device("cuda")
    idx, sub = bench_rays(n)
    t = lambda a: torch.as_tensor(a, dtype=torch.float32).to(dev)
    sd0, sd1 = synthetic.make_net_weights(101), synthetic.make_net_weights(202)
    sweep = [(0, None), (1, None)]
    if len(sys.argv) > 3:          # calibration sweep of the truncation compensation: "fmt:kappa,fmt:kappa,..."
        sweep = [(int(a.split(":")[0]), a.split(":")[1]) for a in sys.argv[3].split(",")]
    for fmt, kappa in sweep:
        if kappa is not None:
            os.environ["ANERF_TRUNC_KAPPA"] = kappa
        plan = _lib.Plan(24, 8, 256, (4,), 0, 0, fmt)
        p0 = plan.pack({k: t(v) for k, v in sd0.items()})
        p1 = plan.pack({k: t(v) for k, v in sd1.items()})
        res = {"idx": idx}
        for Si, tag in ((128, ""), (0, "c_")):
            opts = _lib.make_opts(n, 64, Si, tau_pts=20., tau_views=20., cutoff_pts=0.5, cutoff_views=0.5)
            out = _lib.render_fwd(plan, p0, p1 if Si else None, opts, t(sub["rays"]), t(sub["skts"]), t(sub["cyls"]), None,
                                  None, None, None, None, want_taps=True, keep_nearfar=True)
            torch.cuda.synchronize()
            res.update({tag + k: v.cpu().numpy() for k, v in out.items()})
        name = f"{prefix}_fmt{fmt}" + ("" if kappa is None else f"_k{kappa}") + ".npz"
        if kappa is not None:      # sweep: keep the dumps small
            res = {k: v for k, v in res.items() if k in ("idx", "rgb_map", "acc_map", "disp_map", "rgb0", "acc0", "z_all", "c_raw")}
        np.savez_compressed(name, **res)
        print("wrote", name, {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
