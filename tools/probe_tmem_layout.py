"""Round-2 probe: TMEM layout of cta_group::2 accumulators for M = 256 / 128 and for a lane-offset destination
(tools/probes/tmem_layout_probe.cu).  Prints, per configuration and CTA, which output rows m land in which TMEM lanes and
whether column j holds n = j.   python tools/probe_tmem_layout.py      (needs a B200)"""
import ctypes as C
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
src = os.path.join(HERE, "probes", "tmem_layout_probe.cu")
so = os.path.join(HERE, "probes", "tmem_layout_probe.so")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared",
                           "-Xcompiler", "-fPIC", "-I" + os.path.join(HERE, "..", "anerf_b200", "csrc"), "-o", so, src])
lib = C.CDLL(so)
lib.tmem_layout_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
for (M, N, lane_off, col_off) in [(256, 64, 0, 0), (128, 64, 0, 0), (128, 64, 64, 0), (128, 64, 0, 64), (128, 64, 64, 64)]:
    out = torch.full((2, 128, 128), -2.0, device="cuda")
    rc = lib.tmem_layout_probe(M, N, lane_off, col_off, C.c_void_p(out.data_ptr()))
    print(f"=== M={M} N={N} lane_off={lane_off} col_off={col_off}: rc={rc}")
    if rc != 0:
        continue
    o = out.cpu().numpy()
    for cta in range(2):
        written = o[cta] >= 0
        lanes = np.where(written.any(1))[0]
        cols = np.where(written.any(0))[0]
        desc = []
        for l in lanes:
            row = o[cta, l][written[l]]
            ms = np.unique((row // 256).astype(int))
            ns = (row % 256).astype(int)
            desc.append((int(l), ms.tolist(), bool(np.array_equal(ns, np.arange(ns[0], ns[0] + len(ns))))))
        runs = []
        for l, ms, ok in desc:          # compress consecutive lanes holding consecutive single rows
            if runs and len(ms) == 1 and runs[-1][3] and ok and l == runs[-1][1] + 1 and ms[0] == runs[-1][2] + (l - runs[-1][0]):
                runs[-1][1] = l
            else:
                runs.append([l, l, ms[0] if len(ms) == 1 else ms, ok and len(ms) == 1])
        print(f"  cta {cta}: columns {cols.min() if len(cols) else None}..{cols.max() if len(cols) else None} ({len(cols)} written); "
              f"lane runs [first lane, last lane, first m, clean]: {runs[:12]}")
