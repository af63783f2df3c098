"""Host-side behaviour of the boundary that needs no GPU: the cutoff schedule (a16), the option cache, error handling."""
import contextlib
import io

import numpy as np
import pytest
import torch

from anerf_b200 import _lib
from oracle import ref_import
from tests.test_configs import CONFIGS, _data_attrs, reference_args


def _caster(**over):
    from anerf_b200.raycasters import create_raycaster
    args = reference_args([p for p in CONFIGS if p.endswith("surreal.txt")][0], **over)
    with contextlib.redirect_stdout(io.StringIO()):
        rk_train, rk_test, *_ = create_raycaster(args, _data_attrs(), device=torch.device("cpu"))
    return args, rk_test["ray_caster"]


needs_ref = pytest.mark.skipif(not ref_import.reference_available(), reason="reference sources not present")


@needs_ref
def test_update_embed_fns_follows_the_reference_schedule():
    """RayCaster.update_embed_fns / CutoffEmbedder.update_tau (core/raycasters.py:731-748, core/cutoff_embedder.py:176-183):
    tau = min(2000, 20 * rate^(step / (cutoff_step * 1000))) on both cutoff embedders, against the reference's own objects."""
    args, rc = _caster()
    ref_import.import_reference()
    from core.raycasters import create_raycaster as ref_create
    with contextlib.redirect_stdout(io.StringIO()):
        _, ref_test, *_ = ref_create(args, _data_attrs())
    ref = ref_test["ray_caster"]
    for step in (0, 1, 1000, 125000, 250000, 499999, 500000, 2000000):
        rc.update_embed_fns(step, args)
        ref.update_embed_fns(step, args)
        assert rc.embed_fn.get_tau() == pytest.approx(ref.embed_fn.get_tau(), rel=1e-6)
        assert rc.embeddirs_fn.get_tau() == pytest.approx(ref.embeddirs_fn.get_tau(), rel=1e-6)
    assert rc.embed_fn.get_tau() == 2000.0                       # the ceiling
    assert rc.embedbones_fn.get_tau() == ref.embedbones_fn.get_tau() == 0.0


@needs_ref
def test_option_cache_follows_every_tau_update():
    """The kernels take tau as a launch parameter read back from the device only when it changed; the cache key must not
    be fooled by update_tau rebinding the buffer (fresh tensor, version 0, recycled id) -- ADVICE r1."""
    args, rc = _caster()
    seen = []
    for step in (0, 200000, 400000, 200000, 0, 400000):
        rc.update_embed_fns(step, args)
        o = rc._opts(4, 64, 16, False, 1.0, None)
        seen.append(o.tau_pts)
        assert o.tau_pts == pytest.approx(rc.embed_fn.get_tau(), rel=1e-6) and o.tau_views == pytest.approx(rc.embeddirs_fn.get_tau(), rel=1e-6)
    assert seen[0] == seen[4] and seen[1] == seen[3] and seen[2] == seen[5] and len(set(seen)) == 3
    rc.embed_fn.tau.fill_(123.0)                                  # in-place edits are seen too
    assert rc._opts(4, 64, 16, False, 1.0, None).tau_pts == 123.0
    sd = {k: {n: t.clone() for n, t in v.items()} for k, v in rc.state_dict().items()}
    rc.embed_fn.tau.fill_(7.0)
    rc.load_state_dict(sd)                                        # and checkpoint loads
    assert rc._opts(4, 64, 16, False, 1.0, None).tau_pts == 123.0


def test_error_string_is_per_call():
    """A failed call leaves its message; the next successful call clears it (stale g_err)."""
    lib = _lib.load()
    with pytest.raises(RuntimeError, match="n_joints"):
        _lib.Plan(99, 8, 256)
    assert b"n_joints" in lib.anerf_last_error()
    assert lib.anerf_check_status() == 0
    assert lib.anerf_last_error() == b""


def test_cpu_tensors_are_rejected():
    args, rc = _caster()
    with pytest.raises(RuntimeError, match="CUDA"):
        rc(torch.zeros(4, 11), N_samples=64, kp_batch=torch.zeros(4, 24, 3), skts=torch.eye(4).expand(4, 24, 4, 4),
           cyls=torch.zeros(4, 5), bones=torch.zeros(4, 24, 3), N_importance=16)
