"""How does tcgen05.mma (kind::f16, fp32 accumulate) round?  Needs a B200.
   python tools/probes/trunc_probe.py
Uses the library's GEMM self test (D[256,N] = A[256,K] B[N,K]^T through the fused kernel's pipeline, bf16 operands so
that tiny powers of two are representable).  Each row of A is one experiment: a[0] = +-1 (first K=16 slab -> the
accumulator holds +-1 after the first MMA), then a tiny addend t either in the SAME slab (a[1]) or in a LATER slab
(a[16], a separate MMA instruction).  Compares the result with round-to-nearest, truncation toward zero and floor.
Also: statistics of D(A,B) + D(-A,B) on random operands (zero for any sign-symmetric rounding)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from anerf_b200 import _lib, build  # noqa: E402

build.build()
dev = torch.device("cuda")
K, N = 128, 64
ts = []
for e in (-24, -25, -26, -27, -30):
    for m in (1.0, 1.5, 1.75):
        ts += [m * 2.0 ** e, -m * 2.0 ** e]
rows = []
A = np.zeros((256, K), np.float32)
for s in (1.0, -1.0):
    for where in (1, 16, 32):
        for t in ts:
            i = len(rows)
            A[i, 0] = s
            A[i, where] = t
            rows.append((s, where, t))
B = np.zeros((N, K), np.float32)
B[0, :] = 1.0
D = _lib.selftest_gemm(torch.as_tensor(A).to(dev), torch.as_tensor(B).to(dev), 1)
torch.cuda.synchronize()
d = D[0, :, 0].cpu().numpy().astype(np.float64)


def rn(x):
    return float(np.float32(x))


def rz(x):
    f = np.float32(x)
    if abs(float(f)) > abs(x):
        f = np.nextafter(f, np.float32(0))
    return float(f)


def fl(x):
    f = np.float32(x)
    if float(f) > x:
        f = np.nextafter(f, np.float32(-np.inf))
    return float(f)


cnt = {"rn": 0, "rz": 0, "floor": 0}
print(" s   slab  t(ulp of 1)    result-s (ulp)   rn   rz  floor")
for i, (s, where, t) in enumerate(rows):
    x = s + t
    u = 2.0 ** -23
    tag = [abs(d[i] - f(x)) == 0 for f in (rn, rz, fl)]
    for k, ok in zip(cnt, tag):
        cnt[k] += ok
    print(f"{s:+.0f}  k={where:2d}  {t / u:+10.5f}   {(d[i] - s) / u:+10.5f}      {int(tag[0])}    {int(tag[1])}    {int(tag[2])}")
print("matches:", cnt, "of", len(rows))

g = torch.Generator().manual_seed(3)
for Kb in (256, 1024):
    Ar = torch.randn(256, Kb, generator=g)
    Br = torch.randn(256, Kb, generator=g) / Kb ** 0.5
    for fmt in (1, 0):
        Dp = _lib.selftest_gemm(Ar.to(dev), Br.to(dev), fmt)[0].double()
        Dn = _lib.selftest_gemm((-Ar).to(dev), Br.to(dev), fmt)[0].double()
        ref = (Ar.double() @ Br.double().t()).to(dev)
        sym = (Dp + Dn)
        print(f"K={Kb} fmt={fmt}: mean(D(A)+D(-A)) = {float(sym.mean()):+.3e}  rms {float(sym.pow(2).mean().sqrt()):.3e};  "
              f"mean err D(A) {float((Dp - ref).mean()):+.3e}  rms err {float((Dp - ref).pow(2).mean().sqrt()):.3e};  "
              f"mean err (D(A)-D(-A))/2 {float(((Dp - Dn) / 2 - ref).mean()):+.3e} rms {float(((Dp - Dn) / 2 - ref).pow(2).mean().sqrt()):.3e}; "
              f"fp32 matmul rms err {float(((Ar.to(dev) @ Br.to(dev).t()).double() - ref).pow(2).mean().sqrt()):.3e}")
