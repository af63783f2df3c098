"""Kernel A/B runs on one GPU box: build variants of the library with different compile-time knobs
(-DANERF_...=...) here, then time the benchmark chunk with each of them on the GPU.

    python tools/ab_variants.py build  name:-DANERF_B_STAGES=4 name2:-DANERF_EARLY_COLS=64,-DANERF_TAIL_CHUNKS=1 ...
    python tools/ab_variants.py run    [repeats]        (on the GPU box: every variant found, interleaved)

Variants live in anerf_b200/csrc/variants/ (git-ignored .so files; they travel with the snapshot)."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "anerf_b200", "csrc", "variants")
sys.path.insert(0, ROOT)


def build(specs):
    from anerf_b200 import build as b
    os.makedirs(VAR, exist_ok=True)
    for f in glob.glob(os.path.join(VAR, "*.so")):
        os.remove(f)
    for spec in specs:
        name, _, flags = spec.partition(":")
        out = os.path.join(VAR, f"lib_{name}.so")
        cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + b.NVCC_FLAGS + [f for f in flags.split(",") if f] + \
              ["-o", out] + b.SOURCES
        r = subprocess.run(cmd, cwd=b.CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stderr)
        print("built", out)


def run(repeats):
    libs = sorted(glob.glob(os.path.join(VAR, "*.so")))
    res = {os.path.basename(l): [] for l in libs}
    for _ in range(repeats):
        for l in libs:
            env = dict(os.environ, ANERF_B200_LIB=l)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_chunk.py"), "8"], env=env,
                               capture_output=True, text=True, timeout=120)
            m = re.search(r"chunk ms ([0-9.]+)", r.stdout)
            res[os.path.basename(l)].append(float(m.group(1)) if m else float("nan"))
            if not m:
                print(r.stdout[-400:], r.stderr[-800:])
    for k, v in res.items():
        print(f"{k:40s} min {min(v):.4f}  all {' '.join(f'{x:.4f}' for x in v)}")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
