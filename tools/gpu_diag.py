"""Run on the GPU box: prints error tables for the tensor-core self test and the fused path, and
writes them to gpurun_out/diag.json.  Never stops at the first failure (debug aid, not a test)."""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anerf_b200 import _lib  # noqa: E402
from tests.common import build_case, load_golden, rel_err, run_oracle  # noqa: E402
from tests.test_gpu_parity import gpu_render  # noqa: E402

res = {}


def section(name, fn):
    try:
        res[name] = fn()
    except Exception as e:  # noqa: BLE001
        res[name] = "EXC: " + repr(e)
        traceback.print_exc()
    print(name, json.dumps(res[name]), flush=True)


def gemm():
    out = {}
    for fmt in (1, 0):
        for N, K in [(256, 128), (256, 256), (128, 1024), (64, 128), (64, 384)]:
            g = torch.Generator().manual_seed(1)
            A = torch.randn(256, K, generator=g).cuda()
            B = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
            try:
                D = _lib.selftest_gemm(A, B, fmt)
                torch.cuda.synchronize()
                ref = A.double() @ B.double().t()
                out[f"fmt{fmt}_N{N}_K{K}"] = [float((D[r].double() - ref).abs().max() / ref.abs().max()) for r in range(2)]
            except Exception as e:  # noqa: BLE001
                out[f"fmt{fmt}_N{N}_K{K}"] = repr(e)
                return out
        break
    return out


def accum_probe():
    """operands exactly representable in bf16 -> lo parts vanish, products are exact; what remains is the
    tensor core's accumulation rounding.  Reports mean signed and max relative error vs fp64."""
    out = {}
    for K in (128, 256, 1024, 4096):
        g = torch.Generator().manual_seed(K)
        A = torch.rand(256, K, generator=g).bfloat16().float().cuda()          # positive: no cancellation
        B = torch.rand(256, K, generator=g).bfloat16().float().cuda()
        D = _lib.selftest_gemm(A, B, 1)[0].double()
        ref = A.double() @ B.double().t()
        ref32 = (A @ B.t()).double()                                           # cuBLAS fp32 for comparison
        rel = (D - ref) / ref
        out[f"K{K}"] = dict(mean_signed=float(rel.mean()), max_abs=float(rel.abs().max()),
                            cublas_mean_signed=float(((ref32 - ref) / ref).mean()),
                            cublas_max=float(((ref32 - ref) / ref).abs().max()))
    return out


def big(fmt):
    def f():
        from anerf_b200 import synthetic
        from tests.common import run_oracle
        from oracle import anerf_oracle as orc
        scene = synthetic.make_scene(seed=11, n_rays=2048, H=512, W=512, focal=500., n_joints=24)
        sd0 = synthetic.make_net_weights(101)
        sd1 = synthetic.make_net_weights(202)
        cfg = orc.PathConfig()
        out = gpu_render(scene, sd0, sd1, cfg, None, want_taps=True, fmt=fmt)
        ref, _ = run_oracle(scene, sd0, sd1, cfg)
        ref2, _ = run_oracle(scene, sd0, sd1, cfg, z_all_override=out["z_all"])
        r = {k: rel_err(out[k], ref[k]) for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "alpha0")}
        r["alpha_same_z"] = rel_err(out["alpha"], ref2["alpha"])
        r["rgb_same_z"] = rel_err(out["rgb_map"], ref2["rgb_map"])
        r["acc_mean"] = float(ref["acc_map"].mean())
        return r
    return f


def render(name, **kw):
    def f():
        case, gold = load_golden(name)
        scene, sd0, sd1, cfg, draws = build_case(case)
        t0 = time.time()
        out = gpu_render(scene, sd0, sd1, cfg, draws, want_taps=True, **kw)
        r = {"secs": time.time() - t0}
        for k in out:
            if "ref_" + k in gold and gold["ref_" + k].shape == out[k].shape:
                r[k] = rel_err(out[k], gold["ref_" + k])
        _, taps = run_oracle(scene, sd0, sd1, cfg, draws)
        if "z_all" in out:
            r["z_all"] = rel_err(out["z_all"], taps["z_all"])
            r["raw1_vs_oracle"] = rel_err(out["raw"], taps["raw1"])
        else:
            r["raw0_vs_oracle"] = rel_err(out["raw"], taps["raw0"])
        r["finite"] = bool(all(np.isfinite(v).all() for v in out.values()))
        return r
    return f


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    section("gemm", gemm)
    if isinstance(res["gemm"], dict) and not any(isinstance(v, str) for v in res["gemm"].values()):
        section("accum_probe", accum_probe)
        section("cfg1_coarse_only", render("cfg1_j1_s16_i0"))
        section("cfg1", render("cfg1_j1_s16_i16"))
        section("bench_coarse_only", render("bench_j24_s64_i128", n_importance=0))
        section("bench", render("bench_j24_s64_i128"))
        section("bench_fp16", render("bench_j24_s64_i128", fmt=0))
        section("surreal_tau200", render("surreal_j24_s64_i16_tau200"))
        section("mixamo_fc", render("mixamo_j24_s64_i16_fc"))
        section("train_perturb", render("train_j24_s64_i32_perturb"))
        section("big2048_bf16", big(1))
        section("big2048_fp16", big(0))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(res, f, indent=1)
