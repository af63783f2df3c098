"""Per-tensor gradient error of the three GEMM engines against the oracle fp64 autograd on 24 rays of a forward fixture:
    python tools/probes/engine_option_probe.py <fixture name>   (test infrastructure: imports oracle/ and tests/)"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from tests.common import build_case, load_golden
from tests.test_gpu_train import gpu_grads
from oracle import grad_tools as gt
name = sys.argv[1]
case, _ = load_golden(name)
scene, sd0, sd1, cfg, draws = build_case(case)
keep = slice(0, 24)
scene = {k: (v[keep] if isinstance(v, np.ndarray) and v.shape[:1] == scene["rays_o"].shape[:1] else v) for k, v in scene.items()}
draws = None if draws is None else {k: v[keep] for k, v in draws.items()}
N = scene["rays_o"].shape[0]
cot = gt.cotangents(N, cfg.N_samples, cfg.N_importance)
res = {}
for eng in ("simt", "tc", "bf16"):
    os.environ["ANERF_TRAIN_GEMM"] = eng
    res[eng], out = gpu_grads(scene, sd0, sd1, cfg, draws, cot)
_, g64, _ = gt.oracle_grads(scene, sd0, sd1, cfg, draws, cot, dtype=torch.float64, z_all_override=out.get("z_all"))
for k in sorted(g64):
    n = np.abs(g64[k]).max()
    print(k.ljust(34), "amax %.3e" % n, " ".join("%s %.2e" % (e, gt.rel_err(res[e][k], g64[k])) for e in res))
