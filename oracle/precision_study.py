"""Which tensor-core operand format meets the 1e-4 parity bar?  (TEST INFRASTRUCTURE, CPU emulation.)

Runs the oracle on the benchmark-shaped fixture with every nn.Linear of both nets replaced by an
emulation of a split-precision tensor-core GEMM (operands rounded to bf16/fp16/tf32, optionally as
hi+lo pairs, products accumulated in fp32) and reports max relative error of each output against
the fp32 oracle and against an fp64 run.  Results are recorded in DESIGN.md.
    python oracle/precision_study.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import anerf_oracle as orc  # noqa: E402
from oracle.make_golden import CASES, build_case, rel_err  # noqa: E402


def rnd(x, fmt):
    if fmt == 'bf16':
        return x.bfloat16().float()
    if fmt == 'fp16':
        return x.half().float()
    if fmt == 'tf32':   # round-to-nearest-even to 10 explicit mantissa bits
        i = x.contiguous().view(torch.int32)
        i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(fmt)


def split_linear(fmt, terms):
    def lin(x, w, b=None):
        sh = x.shape
        x = x.reshape(-1, sh[-1]).float()
        xh = rnd(x, fmt); wh = rnd(w, fmt)
        y = xh @ wh.t()
        if terms >= 2:
            xl = rnd(x - xh, fmt)
            y = y + xl @ wh.t()
        if terms >= 3:
            wl = rnd(w - wh, fmt)
            y = y + xh @ wl.t()
        if terms >= 4:
            y = y + xl @ wl.t()
        if b is not None:
            y = y + b
        return y.reshape(*sh[:-1], -1)
    return lin


def main():
    c = CASES["bench_j24_s64_i128"]
    scene, sd0, sd1, cfg = build_case(c)
    t = lambda a, dt=torch.float32: torch.as_tensor(np.asarray(a)).to(dt)

    def run(dt=torch.float32):
        with torch.no_grad():
            return orc.render_rays(orc.to_torch(sd0, dt), orc.to_torch(sd1, dt), cfg, t(scene["rays_o"], dt),
                                   t(scene["rays_d"], dt), t(scene["skts"], dt), t(scene["cyls"], dt))
    ref32 = run()
    ref64 = run(torch.float64)
    orig = F.linear
    keys = ['rgb_map', 'disp_map', 'acc_map', 'rgb0', 'acc0', 'alpha0']
    print("mode            " + " ".join(f"{k:>10s}" for k in keys) + "   (max rel err vs fp32 oracle | vs fp64)")
    print("fp32-vs-fp64    " + " ".join(f"{rel_err(ref32[k].numpy(), ref64[k].numpy()):10.1e}" for k in keys))
    for fmt, terms in [('bf16', 1), ('tf32', 1), ('fp16', 1), ('bf16', 2), ('bf16', 3), ('bf16', 4), ('fp16', 3), ('tf32', 3)]:
        F.linear = split_linear(fmt, terms)
        orc.F.linear = F.linear
        try:
            out = run()
        finally:
            F.linear = orig
            orc.F.linear = orig
        print(f"{fmt}x{terms:<10d}  " + " ".join(f"{rel_err(out[k].numpy(), ref32[k].numpy()):10.1e}" for k in keys)
              + "  | " + " ".join(f"{rel_err(out[k].numpy(), ref64[k].numpy()):8.1e}" for k in keys[:3]))


if __name__ == "__main__":
    main()
