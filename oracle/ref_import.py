"""Import the UNMODIFIED reference hot path from /root/reference (this container only).

TEST / MEASUREMENT INFRASTRUCTURE -- not part of the product path.  Users: `oracle/make_golden*.py`,
`oracle/precision_study.py`, the pinning tests, the tests that drive our boundary through the reference's own callers,
and bench.py's reference legs.  /root/reference does not exist on the GPU box; there the archive of unmodified files under
`oracle/_ref/` (packed by `oracle/build_ref.py` in the build container, git-ignored, shipped with the snapshot) is unpacked
into a temporary directory and used.

The reference imports plotly / matplotlib / pytorch3d / smplx / h5py at module top
(core/utils/skeleton_utils.py:1-13, core/pose_opt.py:5, core/dataset.py:2); none of them is
on the hot path, so they are replaced by empty stub modules.
"""
import argparse
import os
import sys
import tempfile
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """/root/reference in the build container; on the GPU box the unmodified files packed by oracle/build_ref.py into
    oracle/_ref/reference_src.tar.gz (it travels with the snapshot like the built .so files), unpacked -- checksums
    verified -- into a temporary directory outside the repository."""
    for cand in (os.environ.get("ANERF_REFERENCE_ROOT"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "core", "raycasters.py")):
            return cand
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_anerf_build_ref", os.path.join(_HERE, "build_ref.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        d = mod.unpack()
        if d and os.path.isfile(os.path.join(d, "core", "raycasters.py")):
            return d
    except Exception:  # noqa: BLE001
        pass
    return "/root/reference"


REF_ROOT = _find_root()


def reference_available():
    if os.environ.get("ANERF_NO_REFERENCE"):       # tests: exercise the fall-back to the oracle port
        return False
    return os.path.isfile(os.path.join(REF_ROOT, "core", "raycasters.py"))


def reference_kind():
    """'reference' = the unmodified sources (either location)."""
    return "reference" if reference_available() else None


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference's `core` package (raycasters, trainer, ... importable)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    for name in ("plotly", "plotly.graph_objects", "matplotlib", "matplotlib.pyplot",
                 "pytorch3d", "pytorch3d.transforms", "pytorch3d.transforms.rotation_conversions",
                 "h5py", "smplx"):
        _stub(name)
    sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pytorch3d"].transforms = sys.modules["pytorch3d.transforms"]
    sys.modules["pytorch3d.transforms"].rotation_conversions = \
        sys.modules["pytorch3d.transforms.rotation_conversions"]
    if not hasattr(sys.modules["smplx"], "SMPL"):
        sys.modules["smplx"].SMPL = object
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import core.raycasters  # noqa: F401
    import core.trainer  # noqa: F401
    import core
    return core


def make_args(**over):
    """argparse.Namespace with every field `create_raycaster` reads (run_nerf.py:184-488 defaults,
    SURREAL-style flags from configs/surreal/surreal.txt)."""
    tmp = tempfile.mkdtemp(prefix="anerf_ref_")
    os.makedirs(os.path.join(tmp, "exp"), exist_ok=True)
    d = dict(
        n_framecodes=None, use_cutoff=True, normalize_cutoff=False, cutoff_mm=500.,
        ext_scale=0.001, cutoff_inputs=True, opt_cutoff=False, freq_schedule=False, init_freq=0.,
        cut_to_dist=False, cutoff_shift=False, multires=7, i_embed=0, cutoff_bones=False,
        multires_bones=0, use_viewdirs=True, cutoff_viewdir=True, multires_views=4,
        N_importance=128, netdepth=8, netwidth=256, opt_framecode=False, framecode_size=16,
        density_scale=1.0, single_net=False, lrate=5e-4, ft_path=None, basedir=tmp, expname="exp",
        no_reload=True, finetune=False, fix_layer=0, weight_decay=None, density_type="relu",
        softplus_shift=0., pts_tr_type="local", kp_dist_type="reldist", view_type="relray",
        bone_type="reldir", debug=True, perturb=0., N_samples=64, raw_noise_std=0.,
        ray_noise_std=0., lindisp=False, nerf_type="nerf", cutoff_step=250, cutoff_rate=10.,
        freq_schedule_step=50, chunk=4096, netchunk=65536,
    )
    d.update(over)
    return argparse.Namespace(**d)


class reference_on_cuda:
    """Run the unmodified reference on the GPU the way run_nerf.py does (`torch.set_default_tensor_type(
    'torch.cuda.FloatTensor')`, run_nerf.py:__main__): several of its tensors are created with the legacy constructors
    (`torch.Tensor([1e10])`, core/networks/nerf.py:168), which only follow the legacy default type."""

    def __enter__(self):
        import torch
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.set_default_tensor_type('torch.cuda.FloatTensor')
        return self

    def __exit__(self, *exc):
        import torch
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.set_default_tensor_type('torch.FloatTensor')
        return False
