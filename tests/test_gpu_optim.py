"""SURVEY.md 8(f) row 3 on the GPU: FusedAdam (one launch per step, C ABI anerf_adam_step) against torch.optim.Adam."""
import copy

import numpy as np
import pytest
import torch

from anerf_b200.optim import FusedAdam

pytestmark = pytest.mark.gpu


def _params(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(256, 432), (256,), (1, 256), (1,), (128, 904), (3, 128), (3,), (5, 16), (7,)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]


def _grads(params, step):
    g = torch.Generator().manual_seed(100 + step)
    return [torch.randn(*p.shape, generator=g).to(p.device) * (0.1 + 0.3 * i) for i, p in enumerate(params)]


def test_fused_adam_matches_torch_adam():
    dev = torch.device("cuda")
    pa, pb = _params(dev), _params(dev)
    oa = torch.optim.Adam(pa, lr=5e-4, betas=(0.9, 0.999))
    ob = FusedAdam(pb, lr=5e-4, betas=(0.9, 0.999))
    for step in range(1, 8):
        for ps in (pa, pb):
            for p, g in zip(ps, _grads(ps, step)):
                p.grad = g if step != 3 or p.dim() != 1 else None      # step 3: the 1-D tensors have no gradient (frozen that step)
        if step == 5:                                                     # the trainer's learning-rate decay writes the group
            for o in (oa, ob):
                o.param_groups[0]['lr'] = 2e-4
        v0 = [p._version for p in pb]
        oa.step()
        ob.step()
        assert all(p._version > v for p, v in zip(pb, v0) if p.grad is not None)      # re-pack trigger of RayCaster
    for a, b in zip(pa, pb):
        a, b = a.detach(), b.detach()
        assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(a.abs().max()))
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa['state'].keys() == sb['state'].keys()
    for k in sa['state']:
        assert set(sa['state'][k].keys()) == set(sb['state'][k].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}
        assert float(sa['state'][k]['step']) == float(sb['state'][k]['step'])
        assert torch.allclose(sa['state'][k]['exp_avg'], sb['state'][k]['exp_avg'], rtol=1e-5, atol=1e-6)
        assert torch.allclose(sa['state'][k]['exp_avg_sq'], sb['state'][k]['exp_avg_sq'], rtol=1e-5, atol=1e-7)
    # the reference's decay_optimizer_lrate reads the step like this (core/trainer.py:178)
    assert int(ob.state[ob.param_groups[0]['params'][0]]['step'] // 1) == 7
    # checkpoints move between the two optimizers (reference checkpoints hold torch.optim.Adam state)
    oc = FusedAdam(_params(dev), lr=1e-3)
    oc.load_state_dict(copy.deepcopy(sa))
    pc = oc.param_groups[0]['params']
    with torch.no_grad():
        for c, a in zip(pc, pa):
            c.copy_(a)
    for ps in (pa, pc):
        for p, g in zip(ps, _grads(ps, 8)):
            p.grad = g
    oa.step()
    oc.step()
    for a, c in zip(pa, pc):
        assert float((a - c).abs().max()) <= 1e-6 * max(1.0, float(a.abs().max()))


def test_grad_scale_folds_the_allreduce_average():
    dev = torch.device("cuda")
    pa, pb = _params(dev, 1), _params(dev, 1)
    oa, ob = FusedAdam(pa, lr=1e-3), FusedAdam(pb, lr=1e-3)
    for p, q, g in zip(pa, pb, _grads(pa, 1)):
        p.grad, q.grad = g * 0.25, g.clone()
    oa.step()
    ob.step(grad_scale=0.25)
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)


def test_create_raycaster_returns_the_fused_optimizer():
    import contextlib
    import io
    from anerf_b200.raycasters import create_raycaster
    from tests.test_gpu_api import data_attrs, make_args
    with contextlib.redirect_stdout(io.StringIO()):
        _, _, _, grad_vars, optimizer, _ = create_raycaster(make_args(no_reload=True), data_attrs(24))
    assert isinstance(optimizer, FusedAdam) and isinstance(optimizer, torch.optim.Optimizer)
    assert sum(p.numel() for g in optimizer.param_groups for p in g['params']) == sum(p.numel() for p in grad_vars)
    assert optimizer.param_groups[0]['betas'] == (0.9, 0.999)


def test_rejects_cpu_parameters():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedAdam([p]).step()
