"""one training step of the reference's loop without the autograd graph in between.

`Trainer.train_batch` (core/trainer.py:237-287) = render (forward) -> `compute_loss` (:319-381: photometric L1 / MSE of
`rgb + (1 - acc) * bg` against the targets, for the fine and the coarse pair) -> `loss.backward()` -> `optimizer.step()`.
`FusedTrainStep` runs the same arithmetic as a fixed sequence of launches through the C ABI:

    anerf_render_fwd_train      forward as the layer-wise GEMM chain, activations of both passes kept in a state buffer
                                (batches too large to keep: anerf_render_fwd, the fused kernel, and a recomputing backward)
    anerf_loss_seed  x2         loss values + d loss / d (rgb, acc) for the fine and the coarse outputs
    anerf_render_bwd_saved(1)   backward of the coarse network's pass from the kept activations
      [overlap_exchange=True: NCCL all-reduce of the coarse network's gradients on NCCL's stream, overlapping ...]
    anerf_render_bwd_saved(2)   ... the backward of the fine network's pass
      [NCCL all-reduce of the fine network's gradients -- by default of ALL gradients, one exchange per step]
    anerf_adam_step             FusedAdam over all parameters (grad_scale = 1 / world: the all-reduces sum)

Gradients are accumulated into persistent `.grad` buffers (zeroed by one memset per step), the backward workspace is kept
across steps.  A `skts` tensor that requires grad (pose refinement: `PoseOptLayer.forward_poses`) gets its gradient by
`skts.backward(g_skts)`, which continues into the pose chain's own autograd node; everything else bypasses autograd.

The result equals the autograd route (RayCaster in .train() mode + torch loss + backward + FusedAdam.step) -- checked
in tests/test_gpu_trainstep.py -- at fewer launches.
"""
import math

import torch
import torch.distributed as dist

from . import _lib
from .optim import FusedAdam


class FusedTrainStep:
    def __init__(self, ray_caster, optimizer, loss_fn="L1", coarse_weight=1.0, use_background=True, base_bg=1.0, world=None,
                 overlap_exchange=None):
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("FusedTrainStep drives anerf_b200.optim.FusedAdam (what create_raycaster returns on a GPU)")
        if loss_fn not in ("L1", "MSE"):
            raise NotImplementedError(f"loss_fn {loss_fn} (the shipped configs use L1 / MSE)")
        self.rc, self.opt = ray_caster, optimizer
        self.mse, self.coarse_weight, self.use_bg, self.base_bg = loss_fn == "MSE", float(coarse_weight), bool(use_background), float(base_bg)
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        # False (default): one all-reduce of all gradients after both passes.  True: the coarse network's gradients are
        # exchanged while the fine pass's backward runs (two all-reduces).  Measured (profiles/r2_final.md, 3072 rays per
        # rank): 2 GPUs 16.81 / 16.80 ms, 8 GPUs 16.76 / 16.91 ms against 16.62 / 16.68 on one -- at 8 ranks NCCL's CTAs
        # displace CTAs of the persistent 148-CTA GEMM grids they overlap with, which costs more than the exposed
        # 6.9 MB exchange.  ANERF_TRAIN_OVERLAP=0/1 sets the default.
        if overlap_exchange is None:
            import os
            overlap_exchange = os.environ.get("ANERF_TRAIN_OVERLAP", "0") != "0"
        self.overlap = bool(overlap_exchange)
        self._flat = None          # one buffer behind every parameter's .grad: coarse network first, then the fine one
        self._ws = None
        self._side = None

    # ---- persistent gradient storage: [coarse params | fine params] as views of one flat tensor ---------------
    def _grad_views(self, nets, names):
        params = [[dict(n.named_parameters())[k] for k in names] for n in nets]
        if self._flat is None:
            uniq, seen = [], set()
            for ps in params:
                for p in ps:
                    if id(p) not in seen:
                        seen.add(id(p))
                        uniq.append(p)
            sizes = [-(-p.numel() // 4) * 4 for p in uniq]
            self._flat = torch.zeros(sum(sizes), dtype=torch.float32, device=uniq[0].device)
            self._views, off = {}, 0
            for p, s in zip(uniq, sizes):
                self._views[id(p)] = self._flat[off:off + p.numel()].view_as(p)
                off += s
            # element ranges of the two networks inside the flat buffer (for the two all-reduces)
            n0 = sum(-(-p.numel() // 4) * 4 for p in params[0])
            self._ranges = [(0, n0), (n0, self._flat.numel())] if len(params) > 1 and params[1][0] is not params[0][0] else [(0, self._flat.numel())]
        for ps in params:
            for p in ps:
                if p.requires_grad:
                    p.grad = self._views[id(p)]
        return params

    def __call__(self, ray_batch, target, N_samples, kp_batch=None, skts=None, cyls=None, bones=None, cams=None, bgs=None,
                 N_importance=0, perturb=0., raw_noise_std=0., lindisp=False, preproc_kwargs={}, pose_idx=None, **unused):
        rc = self.rc
        dev = ray_batch.device
        N = ray_batch.shape[0]
        J = rc._n_joints()
        Sc, Si = int(N_samples), int(N_importance)
        rays = ray_batch[:, :8].float().contiguous()
        pidx = None if pose_idx is None else pose_idx.to(device=dev, dtype=torch.int32).contiguous()
        # built with autograd ON when the pose is being refined: the expand / reshape in front of the kernels then routes
        # the kernels' d/d skts back to whatever produced `skts` (PoseOptLayer) when skts_c.backward() is called below
        skts_c = (skts.float().reshape(-1, J, 4, 4) if pidx is not None else skts.float().expand(N, J, 4, 4)).contiguous()
        if skts.requires_grad and not skts_c.requires_grad:
            raise RuntimeError("FusedTrainStep must be called with grad mode enabled when skts requires grad")
        cyls_c = cyls.float().expand(N, cyls.shape[-1]).contiguous()
        density_scale = preproc_kwargs.get('density_scale', 1.0)
        use_fc = rc.network.use_framecode
        cams_c = cams.float().reshape(-1).expand(N).contiguous() if use_fc else None
        t_rand = u_rand = noise0 = noise1 = None
        if perturb > 0.:
            t_rand = torch.rand(N, Sc, device=dev)
            u_rand = torch.rand(N, Si, device=dev) if Si > 0 else None
        if raw_noise_std > 0.:
            noise0 = torch.randn(N, Sc, device=dev) * (raw_noise_std * density_scale)
            noise1 = torch.randn(N, Sc + Si, device=dev) * (raw_noise_std * density_scale) if Si > 0 else None
        opts = rc._opts(N, Sc, Si, lindisp, density_scale, preproc_kwargs.get('density_fn', None), False)
        names = _lib.param_names(rc.network.D, use_fc)
        nets = [rc.network] + ([rc.network_fine] if Si > 0 else [])
        with torch.cuda.device(dev), torch.no_grad():
            plan = rc._get_plan()
            params = self._grad_views(nets, names)
            self._flat.zero_()
            det = [[p.detach() for p in ps] for ps in params]
            state = rc._train_state(opts, dev)
            if state is not None:       # the forward keeps its activations: nothing is recomputed in the backward
                rc._claim_train_state()
                out = _lib.render_fwd_train(plan, opts, det[0], det[1] if Si > 0 else None, rays, skts_c.detach(), cyls_c, state,
                                            cams_c, t_rand, u_rand, noise0, noise1, pose_idx=pidx)
            else:
                p0 = rc._packed_image('network')
                p1 = rc._packed_image('network_fine') if Si > 0 else None
                out = _lib.render_fwd(plan, p0, p1, opts, rays, skts_c.detach(), cyls_c, cams_c, t_rand, u_rand, noise0, noise1,
                                      keep_nearfar=True, want_z_all=True, pose_idx=pidx)
            # ---- loss + gradient seed
            sums = torch.zeros(4, dtype=torch.float32, device=dev)
            tgt = target.float().contiguous()
            bg = bgs.float().contiguous() if bgs is not None else self.base_bg
            gout = {}
            g_rgb, g_acc = _lib.loss_seed(out['rgb_map'], out['acc_map'], tgt, bg, self.use_bg, self.mse, 1.0, sums[0:2])
            gout.update(rgb_map=g_rgb, acc_map=g_acc)
            if Si > 0:
                g_rgb0, g_acc0 = _lib.loss_seed(out['rgb0'], out['acc0'], tgt, bg, self.use_bg, self.mse, self.coarse_weight, sums[2:4])
                gout.update(rgb0=g_rgb0, acc0=g_acc0)
            # ---- backward, one network pass at a time; the coarse gradients are exchanged while the fine pass runs
            want = [[bool(p.requires_grad) for p in ps] for ps in params]
            want_skts = bool(skts_c.requires_grad)
            g_skts = torch.zeros_like(skts_c) if want_skts else None
            if state is None and (self._ws is None or self._ws.numel() < _lib.load().anerf_render_bwd_workspace_bytes(plan.handle, N, Sc, Si)):
                self._ws = _lib.bwd_workspace(plan, opts, dev)
            into = [[p.grad if w else None for p, w in zip(ps, ws)] for ps, ws in zip(params, want)]
            common = dict(pose_idx=pidx, into0=into[0], into1=into[1] if Si > 0 else None, g_skts=g_skts, workspace=self._ws, state=state)
            args = (plan, opts, det[0], None if Si == 0 else det[1], rays, skts_c.detach(),
                    cams_c, t_rand, noise0, noise1, out['nearfar'].contiguous(), out.get('z_all'), gout, want[0], want[1] if Si > 0 else None, want_skts)
            handles = []
            if Si > 0:
                split = self.overlap and len(self._ranges) == 2
                _lib.render_bwd(*args, pass_mask=1, **common)
                if self.world > 1 and split:
                    handles.append(self._exchange(0))
                _lib.render_bwd(*args, pass_mask=2, **common)
                if self.world > 1:
                    handles.append(self._exchange(1) if split else self._exchange(None))
            else:
                _lib.render_bwd(*args, pass_mask=3, **common)
                if self.world > 1:
                    handles.append(self._exchange(None))
            for h in handles:
                h.wait()
            self.opt.step(grad_scale=1.0 / self.world)
        if want_skts:
            if skts_c.grad_fn is None:          # `skts` itself was already float / contiguous / full size: a leaf
                skts.grad = g_skts if skts.grad is None else skts.grad + g_skts
            else:
                skts_c.backward(g_skts)
        n_el = 3.0 * N
        stats = {'sums': sums, 'n': n_el}
        return out, stats

    def _exchange(self, which):
        a, b = (0, self._flat.numel()) if which is None else self._ranges[which]
        return dist.all_reduce(self._flat[a:b], op=dist.ReduceOp.SUM, async_op=True)

    @staticmethod
    def losses(stats, mse=False, coarse_weight=1.0):
        """Host-side view of the step's statistics (one device->host read): the loss terms and PSNRs the reference logs."""
        s = stats['sums'].tolist()
        n = stats['n']
        out = {'rgb_loss': s[0] / n, 'psnr': -10. * math.log10(max(s[1] / n, 1e-30))}
        if s[3] > 0 or s[2] > 0:
            out.update(rgb_loss0=coarse_weight * s[2] / n, psnr0=-10. * math.log10(max(s[3] / n, 1e-30)))
        out['total_loss'] = out['rgb_loss'] + out.get('rgb_loss0', 0.)
        return out
