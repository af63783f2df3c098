"""Build the CUDA library in-tree: anerf_b200/csrc/libanerf_b200.so (sm_100a only).

    python -m anerf_b200.build [--force]

nvcc cross-compiles without a GPU, so this runs in the build container; the .so travels to the
GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored)."""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libanerf_b200.so")
SOURCES = ["anerf_api.cu"]
HEADERS = ["render_kernels.cuh", "tc_sm100.cuh", "path_math.cuh", "train_kernels.cuh", "train_path.cuh", "tc_gemm.cuh", "pose_kernels.cuh", "optim_kernels.cuh", "mesh_kernels.cuh", "mc_table.inc", "sampler_kernels.cuh", os.path.join("..", "..", "include", "anerf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


STAMP = LIB + ".srchash"


def _source_hash():
    """Content hash of everything the library is compiled from (+ the flags): the snapshot that carries the built .so
    to a GPU box need not preserve modification times, so staleness is decided by content, not by mtime."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _source_hash()


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tmp = f"{LIB}.{os.getpid()}.tmp"           # several ranks may build at once: write aside, then rename atomically
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB)
    with open(f"{STAMP}.{os.getpid()}.tmp", "w") as fh:
        fh.write(_source_hash())
    os.replace(f"{STAMP}.{os.getpid()}.tmp", STAMP)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
