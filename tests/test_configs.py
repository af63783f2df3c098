"""Every configuration the reference ships must be accepted by `anerf_b200.create_raycaster`, and the
networks it builds must have exactly the parameters the reference's factory builds for the same flags.

The flags are parsed by the reference's OWN `config_parser()` (run_nerf.py:184-488; its source is executed from the
unmodified file, with a 20-line stand-in for the `configargparse` package, which is not installed here)."""
import argparse
import ast
import collections
import contextlib
import glob
import io
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from anerf_b200 import synthetic
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="reference sources not present")

CONFIGS = sorted(glob.glob(os.path.join(ref_import.REF_ROOT, "configs", "*", "*.txt")))


class _ConfigArgParser(argparse.ArgumentParser):
    """What the reference needs of configargparse: `is_config_file` options and `key = value` files."""

    def add_argument(self, *a, **k):
        k.pop("is_config_file", None)
        return super().add_argument(*a, **k)


def reference_args(config_path, **over):
    src = open(os.path.join(ref_import.REF_ROOT, "run_nerf.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "config_parser")
    ns = {}
    shim = types.ModuleType("configargparse")
    shim.ArgumentParser = _ConfigArgParser
    import sys
    had = sys.modules.get("configargparse")
    sys.modules["configargparse"] = shim
    try:
        exec(compile(ast.Module(body=[fn], type_ignores=[]), "run_nerf.py", "exec"), ns)
        parser = ns["config_parser"]()
    finally:
        if had is None:
            del sys.modules["configargparse"]
        else:
            sys.modules["configargparse"] = had
    flags = {a.dest: a for a in parser._actions}
    argv = []
    for line in open(config_path):
        line = line.split("#")[0].strip()
        if "=" not in line:
            continue
        k, v = [x.strip() for x in line.split("=", 1)]
        act = flags[k]                                    # KeyError = a config key the parser does not know
        if isinstance(act, (argparse._StoreTrueAction, argparse._StoreFalseAction)):
            if v.lower() == "true":
                argv.append("--" + k)
        else:
            argv += ["--" + k, v]
    args = parser.parse_args(argv)
    tmp = tempfile.mkdtemp(prefix="anerf_cfg_")
    os.makedirs(os.path.join(tmp, "exp"), exist_ok=True)
    args.basedir, args.expname, args.no_reload = tmp, "exp", True
    for k, v in over.items():
        setattr(args, k, v)
    return args


def _data_attrs(J=24):
    Skel = collections.namedtuple("Skel", ["joint_names", "joint_trees", "root_id"])
    return dict(skel_type=Skel(synthetic.SMPL_JOINT_NAMES[:J], synthetic.SMPL_PARENTS[:J], 0), near=0., far=1., n_views=7,
                joint_coords=np.tile(np.eye(3, dtype=np.float32), (1, J, 1, 1)))


def test_configs_found():
    assert len(CONFIGS) == 8, CONFIGS


@pytest.mark.parametrize("path", CONFIGS, ids=[os.path.basename(p) for p in CONFIGS])
def test_create_raycaster_accepts_shipped_config(path):
    from anerf_b200.raycasters import create_raycaster
    args = reference_args(path)
    with contextlib.redirect_stdout(io.StringIO()):
        rk_train, rk_test, start, grad_vars, optimizer, ckpt = create_raycaster(args, _data_attrs(), device=torch.device("cpu"))
    ours = rk_test["ray_caster"]
    assert rk_train["N_samples"] == args.N_samples and rk_train["N_importance"] == args.N_importance
    assert rk_train["ray_caster"].module is ours
    # the same parameter names and shapes as the reference's factory builds for these flags
    ref_import.import_reference()
    from core.raycasters import create_raycaster as ref_create
    with contextlib.redirect_stdout(io.StringIO()):
        _, ref_test, _, ref_vars, _, _ = ref_create(args, _data_attrs())
    ref = ref_test["ray_caster"]
    for name in ("network", "network_fine"):
        a, b = getattr(ours, name), getattr(ref, name)
        assert (a is None) == (b is None)
        if a is not None:
            assert {k: tuple(v.shape) for k, v in a.state_dict().items()} == {k: tuple(v.shape) for k, v in b.state_dict().items()}
    assert (ours.network_fine is ours.network) == (ref.network_fine is ref.network)          # --single_net
    assert set(ours.state_dict().keys()) == set(ref.state_dict().keys())                     # checkpoint layout
    assert sum(p.numel() for p in grad_vars) == sum(p.numel() for p in ref_vars)


def test_unsupported_flags_raise_not_implemented():
    from anerf_b200.raycasters import create_raycaster
    base = CONFIGS[-2]
    for over in (dict(cutoff_bones=True), dict(opt_cutoff=True), dict(pts_tr_type="bone"), dict(multires=10), dict(multires_views=2)):
        with pytest.raises(NotImplementedError), contextlib.redirect_stdout(io.StringIO()):
            create_raycaster(reference_args(base, **over), _data_attrs(), device=torch.device("cpu"))
