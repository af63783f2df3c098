"""the optimizer step of the training loop as one CUDA launch.

`FusedAdam` is a `torch.optim.Optimizer` with torch.optim.Adam's hyper-parameters, state layout (`step`, `exp_avg`,
`exp_avg_sq` per parameter) and state_dict, so the reference's trainer keeps working on it unchanged
(`decay_optimizer_lrate` reads `optimizer.state[p]['step']` and writes `param_group['lr']`, core/trainer.py:172-185;
checkpoints store `optimizer.state_dict()`, :503) and checkpoints written with either optimizer load into the other.
`step()` launches `anerf_adam_step` (C ABI) once for all parameters that have a gradient.  `create_raycaster` returns
it in place of torch.optim.Adam when the parameters live on a CUDA device.

`grad_scale` folds the 1/world of the gradient all-reduce into the update (parallel.allreduce_gradients(average=False)).
"""
import ctypes as C

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.):
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameters")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            if group.get('amsgrad') or group.get('maximize'):
                raise NotImplementedError("FusedAdam: amsgrad / maximize are not implemented")
            by_step = {}
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse or p.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("FusedAdam: dense fp32 CUDA parameters only (no CPU path)")
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = torch.tensor(0.0, dtype=torch.float32)
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['step'] += 1
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters and gradients must be contiguous")
                by_step.setdefault((int(st['step']), p.device), []).append((p, p.grad, st['exp_avg'], st['exp_avg_sq']))
            beta1, beta2 = group['betas']
            lr = group['lr']
            lr = float(lr) if not torch.is_tensor(lr) else float(lr.item())
            for (step, dev), items in by_step.items():
                n = len(items)
                arr = lambda i: (C.c_void_p * n)(*[t[i].data_ptr() for t in items])
                sizes = (C.c_int64 * n)(*[t[0].numel() for t in items])
                with torch.cuda.device(dev):
                    _lib.check(_lib.load().anerf_adam_step(n, arr(0), arr(1), arr(2), arr(3), sizes, step, lr, beta1, beta2,
                                                           group['eps'], group['weight_decay'], float(grad_scale), _lib._stream()))
                # the kernel wrote the parameters behind autograd's back: bump their version counters (what an in-place
                # torch op would have done) so that RayCaster re-packs the tensor-core images before the next forward
                ps = [t[0] for t in items]
                torch._C._autograd._unsafe_set_version_counter(ps, [p._version + 1 for p in ps])
        return loss
