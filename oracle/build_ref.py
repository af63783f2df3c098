"""Recipe for `oracle/_ref/`: the UNMODIFIED reference files of the hot path, copied from where they lie under
/root/reference so that the real reference can run beside the kernels on the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE.  The reference is pure Python, so "building" it is a copy; nothing is edited.
`oracle/_ref/` is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so it
travels to the GPU box with the snapshot exactly like the built .so files do.  `__graft_entry__.build()` runs this
whenever /root/reference is present (the build container); on the GPU box the prebuilt copy is used as it is.
Users: `oracle/ref_import.py` (falls back to oracle/_ref when /root/reference is absent), hence `bench.py --impl reference`,
bench.py's `cpu_baseline` / `reference_gpu` legs and the tests that drive our boundary through the reference's own callers.

    python oracle/build_ref.py [--force]
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("ANERF_REFERENCE_SRC", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")

# What `import core.raycasters, core.trainer, core.pose_opt` pulls in (module-top imports of
# plotly / matplotlib / pytorch3d / smplx / h5py are stubbed by oracle/ref_import.py, none is on the path).
FILES = [
    "core/__init__.py",
    "core/raycasters.py",             # a1, a13-a15: create_raycaster, RayCaster
    "core/encoders.py",               # a4-a7
    "core/cutoff_embedder.py",        # a8-a10, a16
    "core/trainer.py",                # the caller: render / batchify_rays / Trainer
    "core/networks/__init__.py",
    "core/networks/nerf.py",          # a11, a12
    "core/networks/embedding.py",     # a17
    "core/utils/__init__.py",
    "core/utils/ray_utils.py",        # a2, a3, a13
    "core/utils/run_nerf_helpers.py",
    "core/utils/skeleton_utils.py",   # skeletons, cylinders, kinematic helpers
    "core/pose_opt.py",               # PoseOptLayer (the pose chain our fused version is checked against)
    "core/process_spin.py",           # imported by pose_opt.py (SMPL_JOINT_MAPPER)
    "run_nerf.py",                    # the caller: config_parser() (flag defaults), train(), render_path()
    # the shipped configurations (tests/test_configs.py: create_raycaster must accept every one of them)
    "configs/h36m/h36m_prot2.txt", "configs/h36m/h36m_prot2_finetune.txt", "configs/mixamo/mixamo.txt",
    "configs/mixamo/mixamo_finetune.txt", "configs/perfcap/perfcap.txt", "configs/perfcap/perfcap_finetune.txt",
    "configs/surreal/surreal.txt", "configs/surreal/surreal_single.txt",
]


def _sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def source_available():
    return os.path.isfile(os.path.join(SRC_ROOT, "core", "raycasters.py"))


def build(force=False):
    """Copies FILES into oracle/_ref/ and writes MANIFEST.json (sha256 per file).  Returns the destination, or None
    when the reference sources are not on this machine (GPU box: use the copy that travelled with the snapshot)."""
    if not source_available():
        return DST_ROOT if os.path.isdir(os.path.join(DST_ROOT, "core")) else None
    manifest_path = os.path.join(DST_ROOT, "MANIFEST.json")
    want = {f: _sha(os.path.join(SRC_ROOT, f)) for f in FILES}
    if not force and os.path.exists(manifest_path):
        try:
            have = json.load(open(manifest_path))["files"]
            if have == want and all(os.path.exists(os.path.join(DST_ROOT, f)) and _sha(os.path.join(DST_ROOT, f)) == h
                                    for f, h in want.items()):
                return DST_ROOT
        except Exception:  # noqa: BLE001
            pass
    for f in FILES:
        dst = os.path.join(DST_ROOT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC_ROOT, f), dst)
    commit = None
    sub = os.path.join(SRC_ROOT, ".SUBMODULES.json")
    if os.path.exists(sub):
        try:
            commit = json.load(open(sub)).get("commit")
        except Exception:  # noqa: BLE001
            commit = None
    with open(manifest_path, "w") as fh:
        json.dump({"source": SRC_ROOT, "commit": commit, "files": want,
                   "note": "unmodified copies; see oracle/build_ref.py"}, fh, indent=1)
    return DST_ROOT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
