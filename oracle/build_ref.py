"""Recipe for `oracle/_ref/`: the UNMODIFIED reference files of the hot path, copied from where they lie under
/root/reference so that the real reference can run beside the kernels on the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE.  The reference is pure Python, so "building" it is packing its files, byte for byte,
into ONE archive `oracle/_ref/reference_src.tar.gz` (+ MANIFEST.json with the sha256 of every member); nothing is edited,
and no reference source file exists as a file of this repository -- the archive is a build artefact like a .so.
`oracle/_ref/` is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so it
travels to the GPU box with the snapshot exactly like the built .so files do; there `oracle/ref_import.py` unpacks it
into a temporary directory outside the repository and imports the reference from that.  `__graft_entry__.build()` runs this
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


ARCHIVE = os.path.join(DST_ROOT, "reference_src.tar.gz")
MANIFEST = os.path.join(DST_ROOT, "MANIFEST.json")


def build(force=False):
    """Packs FILES into oracle/_ref/reference_src.tar.gz and writes MANIFEST.json (sha256 per file).  Returns the
    archive path, or None when neither the reference sources nor a previously built archive are on this machine."""
    import io
    import tarfile
    if not source_available():
        return ARCHIVE if os.path.isfile(ARCHIVE) else None
    want = {f: _sha(os.path.join(SRC_ROOT, f)) for f in FILES}
    if not force and os.path.exists(MANIFEST) and os.path.isfile(ARCHIVE):
        try:
            if json.load(open(MANIFEST))["files"] == want:
                return ARCHIVE
        except Exception:  # noqa: BLE001
            pass
    os.makedirs(DST_ROOT, exist_ok=True)
    for stale in ("core", "configs", "run_nerf.py"):          # plain copies of an earlier recipe
        pth = os.path.join(DST_ROOT, stale)
        if os.path.isdir(pth):
            shutil.rmtree(pth)
        elif os.path.exists(pth):
            os.remove(pth)
    tmp = ARCHIVE + f".{os.getpid()}.tmp"
    with tarfile.open(tmp, "w:gz") as tar:
        for f in FILES:
            with open(os.path.join(SRC_ROOT, f), "rb") as fh:
                data = fh.read()
            info = tarfile.TarInfo(f)
            info.size, info.mtime, info.mode = len(data), 0, 0o644
            tar.addfile(info, io.BytesIO(data))
    os.replace(tmp, ARCHIVE)
    commit = None
    sub = os.path.join(SRC_ROOT, ".SUBMODULES.json")
    if os.path.exists(sub):
        try:
            commit = json.load(open(sub)).get("commit")
        except Exception:  # noqa: BLE001
            commit = None
    with open(MANIFEST, "w") as fh:
        json.dump({"source": SRC_ROOT, "commit": commit, "files": want,
                   "note": "unmodified files packed by oracle/build_ref.py"}, fh, indent=1)
    return ARCHIVE


def unpack(dst=None):
    """Extracts the archive (verifying every member against MANIFEST.json) into `dst` (default: a per-archive directory
    under the system temp dir, outside the repository) and returns that directory, or None without an archive."""
    import tarfile
    import tempfile
    if not os.path.isfile(ARCHIVE) or not os.path.isfile(MANIFEST):
        return None
    files = json.load(open(MANIFEST))["files"]
    tag = hashlib.sha256(json.dumps(files, sort_keys=True).encode()).hexdigest()[:16]
    dst = dst or os.path.join(tempfile.gettempdir(), f"anerf_reference_{tag}")
    ok = all(os.path.isfile(os.path.join(dst, f)) and _sha(os.path.join(dst, f)) == h for f, h in files.items())
    if not ok:
        stage = tempfile.mkdtemp(prefix="anerf_reference_stage_")
        with tarfile.open(ARCHIVE, "r:gz") as tar:
            for m in tar.getmembers():
                if m.name not in files or not m.isfile():
                    raise RuntimeError(f"unexpected member {m.name} in {ARCHIVE}")
            tar.extractall(stage, filter="data")
        for f, h in files.items():
            if _sha(os.path.join(stage, f)) != h:
                raise RuntimeError(f"{f}: checksum differs from MANIFEST.json")
        if os.path.isdir(dst):
            shutil.rmtree(dst, ignore_errors=True)
        try:
            os.replace(stage, dst)
        except OSError:                       # another rank got there first
            shutil.rmtree(stage, ignore_errors=True)
    return dst


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
