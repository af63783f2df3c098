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
    for fmt in (1, 0, 2):
        for N, K in [(256, 32), (256, 256), (128, 928), (64, 64), (32, 96)]:
            g = torch.Generator().manual_seed(1)
            A = torch.randn(128, K, generator=g).cuda()
            B = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
            try:
                D = _lib.selftest_gemm(A, B, fmt)
                torch.cuda.synchronize()
                ref = A.double() @ B.double().t()
                out[f"fmt{fmt}_N{N}_K{K}"] = [float((D[r].double() - ref).abs().max() / ref.abs().max()) for r in range(2)]
            except Exception as e:  # noqa: BLE001
                out[f"fmt{fmt}_N{N}_K{K}"] = repr(e)
                return out
    return out


def render(name, **kw):
    def f():
        case, gold = load_golden(name)
        scene, sd0, sd1, cfg, draws = build_case(case)
        t0 = time.time()
        out = gpu_render(scene, sd0, sd1, cfg, draws, want_taps=True, **kw)
        r = {"secs": time.time() - t0}
        for k in out:
            if "ref_" + k in gold:
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
        section("cfg1_coarse_only", render("cfg1_j1_s16_i0"))
        section("cfg1", render("cfg1_j1_s16_i16"))
        section("bench_coarse_only", render("bench_j24_s64_i128", n_importance=0))
        section("bench", render("bench_j24_s64_i128"))
        section("bench_fp16", render("bench_j24_s64_i128", fmt=0))
        section("surreal_tau200", render("surreal_j24_s64_i16_tau200"))
        section("mixamo_fc", render("mixamo_j24_s64_i16_fc"))
        section("train_perturb", render("train_j24_s64_i32_perturb"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(res, f, indent=1)
