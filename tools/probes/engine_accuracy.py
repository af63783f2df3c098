"""Accuracy of the three GEMM engines of the training path on a batch of the BENCHMARK shape (rays of bench.py's frame 0,
64 + 16 samples, 24 joints, MSE loss on rgb_map + rgb0), against the oracle's autograd evaluated in fp64 at the kernel's
own fine sample positions; the oracle's fp32 autograd is the yardstick (what a re-ordered fp32 evaluation differs by).

    python tools/probes/engine_accuracy.py [n_rays=768] [seed=0]

A relu density gate (`relu(raw)`) is a discrete event: a sample whose raw density is within rounding of zero is "on" in one
evaluation and "off" in another, and a single such sample on a surface moves the whole gradient by ~1e-3 (seen as
IDENTICAL deviations of two unrelated fp32 evaluations from fp64); `l2_rel_all_vs_oracle_fp32` tells that apart from
arithmetic noise.

Test infrastructure (imports oracle/): run by hand on a GPU box, one JSON line per engine."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import anerf_oracle as orc          # noqa: E402
from oracle import grad_tools as gt             # noqa: E402
from tests.common import bench_frame_scene      # noqa: E402
from tests.test_gpu_train import gpu_grads      # noqa: E402
from anerf_b200 import synthetic                # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 768
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
Sc, Si = 64, 16
scene = bench_frame_scene(n_rays)
sd0, sd1 = synthetic.make_net_weights(101), synthetic.make_net_weights(202)
cfg = orc.PathConfig(n_joints=24, N_samples=Sc, N_importance=Si)

# dL/d(outputs) of the training loss at these weights: forward once through the library, MSE against a seeded target
os.environ["ANERF_TRAIN_GEMM"] = "simt"
zero = {k: np.zeros_like(v) for k, v in gt.cotangents(n_rays, Sc, Si).items()}
_, out = gpu_grads(scene, sd0, sd1, cfg, None, zero, need_pose=False)
target = np.random.RandomState(seed).rand(n_rays, 3).astype(np.float32)
cot = dict(zero)
cot["rgb_map"] = (2.0 * (out["rgb_map"] - target) / (3 * n_rays)).astype(np.float32)
cot["rgb0"] = (2.0 * (out["rgb0"] - target) / (3 * n_rays)).astype(np.float32)

_, g64, taps64 = gt.oracle_grads(scene, sd0, sd1, cfg, None, cot, dtype=torch.float64, z_all_override=out["z_all"])
_, g32, taps32 = gt.oracle_grads(scene, sd0, sd1, cfg, None, cot, dtype=torch.float32, z_all_override=out["z_all"])


# density gates that the fp32 and fp64 evaluations of the oracle resolve differently
for key, wkey in (("raw0", "weights0"), ("raw1", "weights1")):
    r64, r32 = taps64[key][..., 3], taps32[key][..., 3]
    flips = np.argwhere((r64 > 0) != (r32 > 0))
    print(json.dumps({"gate": key, "samples": int(r64.size), "resolved_differently": int(len(flips)),
                      "cases": [{"ray": int(i), "sample": int(j), "raw_fp64": float(r64[i, j]), "raw_fp32": float(r32[i, j]),
                                 "transmittance_weight_of_neighbours": float(taps64[wkey][i, max(j - 2, 0):j + 3].sum())}
                                for i, j in flips[:6]]}))


def report(tag, g):
    l2 = {k: float(np.linalg.norm((g[k].astype(np.float64) - g64[k]).ravel()) / max(np.linalg.norm(g64[k].ravel()), 1e-300)) for k in g64}
    mx = {k: float(np.abs(g[k].astype(np.float64) - g64[k]).max() / max(np.abs(g64[k]).max(), 1e-300)) for k in g64}
    fa = np.concatenate([g[k].astype(np.float64).ravel() for k in sorted(g64)])
    fb = np.concatenate([g64[k].ravel() for k in sorted(g64)])
    fc = np.concatenate([g32[k].astype(np.float64).ravel() for k in sorted(g64)])
    w = max(l2, key=l2.get)
    wm = max(mx, key=mx.get)
    print(json.dumps({"engine": tag, "l2_rel_all": float(np.linalg.norm(fa - fb) / np.linalg.norm(fb)),
                      "l2_rel_all_vs_oracle_fp32": float(np.linalg.norm(fa - fc) / np.linalg.norm(fc)),
                      "l2_rel_worst_tensor": l2[w], "worst_tensor": w, "norm_of_worst": float(np.linalg.norm(g64[w].ravel())),
                      "maxnorm_rel_worst_tensor": mx[wm], "worst_tensor_maxnorm": wm,
                      "l2_rel_median_tensor": float(np.median(list(l2.values())))}))


report("oracle fp32 autograd (yardstick)", g32)
for engine in ("simt", "tc", "bf16"):
    os.environ["ANERF_TRAIN_GEMM"] = engine
    g, _ = gpu_grads(scene, sd0, sd1, cfg, None, cot, need_pose=True)
    report(engine, g)
