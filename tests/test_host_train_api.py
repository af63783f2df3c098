"""Host-side behaviour of the training API that needs no GPU: argument checks and defaults of FusedTrainStep, the switch
of the saved-activation route."""
import pytest
import torch

from anerf_b200 import raycasters
from anerf_b200.optim import FusedAdam
from anerf_b200.train import FusedTrainStep


def _adam():
    return FusedAdam([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)


def test_fused_train_step_argument_checks():
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(TypeError):
        FusedTrainStep(object(), torch.optim.Adam([p], lr=1e-3))
    with pytest.raises(NotImplementedError):
        FusedTrainStep(object(), _adam(), loss_fn="Huber")


def test_exchange_schedule_default_and_override(monkeypatch):
    """One gradient exchange per step unless asked otherwise (measured faster at 8 GPUs, profiles/r2_final.md)."""
    monkeypatch.delenv("ANERF_TRAIN_OVERLAP", raising=False)
    assert FusedTrainStep(object(), _adam(), world=8).overlap is False
    monkeypatch.setenv("ANERF_TRAIN_OVERLAP", "1")
    assert FusedTrainStep(object(), _adam(), world=8).overlap is True
    assert FusedTrainStep(object(), _adam(), world=8, overlap_exchange=False).overlap is False
    monkeypatch.setenv("ANERF_TRAIN_OVERLAP", "0")
    assert FusedTrainStep(object(), _adam(), world=8, overlap_exchange=True).overlap is True


def test_saved_activation_route_switch():
    """RayCaster.keep_activations: on by default, a per-instance attribute turns it off (the state buffer is then never asked for)."""
    assert raycasters.RayCaster.keep_activations is True

    class Probe(raycasters.RayCaster):
        def __init__(self):          # no networks: only the switch is exercised
            torch.nn.Module.__init__(self)

    rc = Probe()
    rc.keep_activations = False
    assert rc._train_state(None, torch.device("cpu")) is None
    assert rc._claim_train_state() == 1 and rc._claim_train_state() == 2
