"""Host side of the drop-in boundary: `create_raycaster` / `RayCaster` with the reference's names,
argument meaning and error behaviour (reference: core/raycasters.py:17-184, 326-794), so that the
reference's run_nerf.py / run_render.py call it unchanged.

The modules below only *hold* parameters and buffers under the reference's state_dict layout
(core/networks/nerf.py:57-88, core/cutoff_embedder.py:91-95, core/networks/embedding.py:9);
all arithmetic happens in libanerf_b200.so through the C ABI (`anerf_b200._lib`).  There is no
PyTorch or CPU fallback: if the library is missing, or a tensor is not on a CUDA device, calls raise.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib

MULTIRES = 7                       # distance frequencies compiled into the kernels (every shipped config)
MULTIRES_VIEWS = (4, 0)            # view-direction frequencies: 4, or 0 = raw directions (configs/surreal/surreal_single.txt:32)


# ------------------------------------------------------------------------------------------------
# parameter containers
# ------------------------------------------------------------------------------------------------
class Embedder(nn.Module):
    """Plain positional encoding descriptor (core/cutoff_embedder.py:9-58).  With num_freqs == 0 it is the
    identity (`embedbones_fn` of every shipped config)."""

    def __init__(self, input_dims, num_freqs):
        super().__init__()
        self.input_dims, self.num_freqs = input_dims, num_freqs
        self.out_dim = input_dims * (1 + 2 * num_freqs)

    def update_threshold(self, *args, **kwargs):
        pass

    def update_tau(self, *args, **kwargs):
        pass

    def update_alpha(self, *args, **kwargs):
        pass

    def get_tau(self):
        return 0.0

    def forward(self, *args, **kwargs):
        raise RuntimeError("anerf_b200 embedders hold parameters only; encoding runs inside the fused CUDA kernel")


class CutoffEmbedder(Embedder):
    """Cutoff positional encoding parameters (core/cutoff_embedder.py:61-197): `cutoff_dist` per joint
    (Parameter without grad) and the sigmoid sharpness `tau` (buffer), with the reference's schedule."""

    def __init__(self, input_dims, num_freqs, cutoff_dist, cutoff_dim, dist_inputs):
        super().__init__(input_dims, num_freqs)
        self.dist_inputs = dist_inputs
        self.cutoff_dim = cutoff_dim
        self.cutoff_dist = nn.Parameter(torch.ones(cutoff_dim) * cutoff_dist, requires_grad=False)
        self.init_tau = 20.
        self.register_buffer('tau', torch.tensor(self.init_tau))
        self._epoch = 0          # bumped whenever tau / cutoff_dist may have changed (cache key of RayCaster._opts)

    def get_tau(self):
        return self.tau.item()

    def _load_from_state_dict(self, *args, **kwargs):
        self._epoch += 1
        return super()._load_from_state_dict(*args, **kwargs)

    def get_cutoff_dist(self):
        return self.cutoff_dist

    def update_threshold(self, global_step, tau_step, tau_rate, alpha_step, alpha_target):
        self.update_tau(global_step, tau_step, tau_rate)

    def update_tau(self, global_step, step, rate):
        # tau = min(2000, 20 * rate^(step / (cutoff_step * 1000)))   (cutoff_embedder.py:181-183)
        self.tau = (self.init_tau * torch.ones_like(self.tau) * rate ** (global_step / float(step * 1000))).clamp(max=2000.)
        self._epoch += 1


class Optcodes(nn.Module):
    """Per-frame appearance codes (core/networks/embedding.py:4-44)."""

    def __init__(self, n_codes, code_ch):
        super().__init__()
        self.n_codes, self.code_ch = n_codes, code_ch
        self.codes = nn.Embedding(n_codes, code_ch)
        nn.init.xavier_normal_(self.codes.weight)


class NeRF(nn.Module):
    """Parameters of the density/radiance MLP under the reference's names (core/networks/nerf.py:12-88)."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_bones=0, input_ch_views=3, output_ch=4, skips=(4,),
                 use_viewdirs=True, use_framecode=False, framecode_ch=16, n_framecodes=0, skel_type=None,
                 density_scale=1.0):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_bones, self.input_ch_views = input_ch, input_ch_bones, input_ch_views
        self.skips = list(skips)
        self.use_viewdirs, self.use_framecode = use_viewdirs, use_framecode
        self.framecode_ch, self.n_framecodes = framecode_ch, n_framecodes
        self.output_ch, self.skel_type, self.density_scale = output_ch, skel_type, density_scale
        dnet = input_ch + input_ch_bones
        layers = [nn.Linear(dnet, W)]
        for i in range(D - 1):
            layers.append(nn.Linear(W + dnet if i in self.skips else W, W))
        self.pts_linears = nn.ModuleList(layers)
        self.alpha_linear = nn.Linear(W, 1)
        vnet = input_ch_views + (framecode_ch if use_framecode else 0) + W
        self.views_linears = nn.ModuleList([nn.Linear(vnet, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.rgb_linear = nn.Linear(W // 2, 3)
        if use_framecode:
            self.framecodes = Optcodes(n_framecodes, framecode_ch)

    @property
    def dnet_input(self):
        return self.input_ch + self.input_ch_bones

    @property
    def vnet_input(self):
        return self.input_ch_views + (self.framecode_ch if self.use_framecode else 0) + self.W

    def forward(self, *args, **kwargs):
        raise RuntimeError("anerf_b200.NeRF holds parameters only; evaluation runs inside the fused CUDA kernel")


class _EncoderTag:
    """Stand-in for the reference's encoder objects inside `preproc_kwargs` (core/encoders.py); only the
    default trio is compiled into the kernels, so these carry the name and nothing else."""

    def __init__(self, name, dims):
        self.encoder_name, self.dims = name, dims


# ------------------------------------------------------------------------------------------------
# training: the chunk as one autograd node
# ------------------------------------------------------------------------------------------------
OUT_KEYS = ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0", "alpha0")


class _RenderRaysFn(torch.autograd.Function):
    """render_rays as a single autograd node (reference graph: core/raycasters.py:361-474).  When something needs a
    gradient: forward = anerf_render_fwd_train (layer-wise chain, activations kept in the caster's state buffer), backward
    = anerf_render_bwd_saved.  Otherwise, or when the batch is too large to keep (RayCaster.keep_activations = False turns
    it off): forward = the fused kernel (anerf_render_fwd), keeping only the repaired near/far and the sorted fine depths;
    backward = anerf_render_bwd, which recomputes the activations layer by layer.  Differentiable inputs: `skts` (pose
    refinement) and the parameters of both networks; sample positions carry no gradient (ray_utils.py:285)."""

    @staticmethod
    def forward(ctx, caster, opts, aux, skts, *params):
        plan = caster._get_plan()
        ctx.caster, ctx.opts, ctx.aux = caster, opts, aux
        ctx.state = None
        # A step that will be differentiated runs the layer-wise chain and KEEPS its activations (anerf_render_fwd_train):
        # the backward then has nothing to recompute.  Everything else is the fused kernel.
        state = caster._train_state(opts, skts.device) if any(ctx.needs_input_grad[3:]) else None
        if state is not None:
            n0 = aux['n_params0']
            det = [p.detach() for p in params]
            out = _lib.render_fwd_train(plan, opts, det[:n0], det[n0:] if len(det) > n0 else None, aux['rays'], skts.detach(),
                                        aux['cyls'], state, aux['cams'], aux['t_rand'], aux['u_rand'], aux['noise0'], aux['noise1'],
                                        pose_idx=aux.get('pose_idx'))
            ctx.state, ctx.state_epoch = state, caster._claim_train_state()
        else:
            p0 = caster._packed_image('network')
            p1 = caster._packed_image('network_fine') if opts.n_importance > 0 else None
            out = _lib.render_fwd(plan, p0, p1, opts, aux['rays'], skts, aux['cyls'], aux['cams'], aux['t_rand'], aux['u_rand'],
                                  aux['noise0'], aux['noise1'], keep_nearfar=True, want_z_all=True, pose_idx=aux.get('pose_idx'))
        ctx.nearfar, ctx.z_all = out['nearfar'], out.get('z_all')
        ctx.save_for_backward(skts, *params)
        ctx.keys = [k for k in OUT_KEYS if k in out]
        return tuple(out[k] for k in ctx.keys)

    @staticmethod
    def backward(ctx, *gouts):
        skts, *params = ctx.saved_tensors
        aux, opts = ctx.aux, ctx.opts
        n0 = aux['n_params0']
        params0, params1 = params[:n0], (params[n0:] if len(params) > n0 else None)
        need = ctx.needs_input_grad
        want_skts = bool(need[3])
        want0 = [bool(x) for x in need[4:4 + n0]]
        want1 = [bool(x) for x in need[4 + n0:]]
        gout = {k: (None if g is None else g.float().contiguous()) for k, g in zip(ctx.keys, gouts)}
        # Parameter gradients: the kernels ADD into buffers, so an existing contiguous fp32 .grad is used as it is and
        # missing ones are carved out of one zero-filled allocation and installed as .grad -- autograd's AccumulateGrad
        # (one clone or add per tensor, ~90 tiny launches per step for two networks) is bypassed by returning None for
        # those inputs.  RayCaster.accumulate_param_grads_in_place = False restores the returned-gradient behaviour
        # (needed only by torch.autograd.grad(..., network_parameters)).
        direct = ctx.caster.accumulate_param_grads_in_place
        usable = lambda p, w: direct and w and p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32 and p.grad.device == p.device
        into0 = [p.grad if usable(p, w) else None for p, w in zip(params0, want0)]
        into1 = None if params1 is None else [p.grad if usable(p, w) else None for p, w in zip(params1, want1)]
        # the kept activations are this call's only while no later forward has re-used the state buffer
        state = ctx.state if (ctx.state is not None and ctx.caster._train_state_epoch == ctx.state_epoch) else None
        with torch.cuda.device(skts.device):
            g0, g1, g_skts = _lib.render_bwd(ctx.caster._get_plan(), opts, [p.detach() for p in params0],
                                             None if params1 is None else [p.detach() for p in params1],
                                             aux['rays'], skts.detach(), aux['cams'], aux['t_rand'], aux['noise0'], aux['noise1'],
                                             ctx.nearfar, ctx.z_all, gout, want0, want1, want_skts, pose_idx=aux.get('pose_idx'),
                                             into0=into0, into1=into1, state=state)
        if not direct:
            return (None, None, None, g_skts, *g0, *(g1 or []))
        seen = {}
        for p, g, w in list(zip(params0, g0, want0)) + list(zip(params1 or [], g1 or [], want1)):
            if not w:
                continue
            if p.grad is None:
                if id(p) in seen:                       # --single_net: the same parameter in both passes
                    seen[id(p)].add_(g)
                else:
                    p.grad = g
                    seen[id(p)] = g
            elif p.grad is not g:                       # an unusable .grad (other dtype / layout): add the usual way
                p.grad.add_(g.to(p.grad.dtype))
        return (None, None, None, g_skts) + (None,) * (len(params0) + len(params1 or []))


# ------------------------------------------------------------------------------------------------
# the ray caster
# ------------------------------------------------------------------------------------------------
class RayCaster(nn.Module):
    accumulate_param_grads_in_place = True       # see _RenderRaysFn.backward
    keep_activations = os.environ.get("ANERF_KEEP_ACTIVATIONS", "1") != "0"     # see _RenderRaysFn.forward

    def __init__(self, network, embed_fn, embedbones_fn, embeddirs_fn, network_fine=None, joint_coords=None,
                 single_net=False, operand_format=None):
        super().__init__()
        self.network = network
        self.network_fine = network_fine
        self.embed_fn = embed_fn
        self.embedbones_fn = embedbones_fn
        self.embeddirs_fn = embeddirs_fn
        if joint_coords is not None:
            n_j = joint_coords.shape[-3]
            self.register_buffer('joint_coords', joint_coords.reshape(-1, n_j, 3, 3))
        self.single_net = single_net
        if operand_format is None:
            operand_format = int(os.environ.get("ANERF_OPERAND_FORMAT", "0"))
        self._operand_format = operand_format
        self._plan = None
        self._packed = {}        # net name -> (packed image tensor, version signature)

    # ---- engine plumbing ---------------------------------------------------------------------
    def _n_joints(self):
        return self.embed_fn.cutoff_dim

    def _get_plan(self):
        """The plan owns device buffers: it is created (and must be used) under the device of the parameters."""
        dev = next(self.network.parameters()).device
        if self._plan is not None and self._plan_device != dev:
            self._plan, self._packed = None, {}
        if self._plan is None:
            self._plan_device = dev
            net = self.network
            with torch.cuda.device(dev):
                self._plan = _lib.Plan(self._n_joints(), net.D, net.W, net.skips,
                                       net.framecode_ch if net.use_framecode else 0,
                                       net.n_framecodes if net.use_framecode else 0, self._operand_format,
                                       view_freqs=self.embeddirs_fn.num_freqs)
        return self._plan

    def _train_state(self, opts, dev):
        """The saved-activation state buffer for this batch shape (one per caster, re-used every step), or None when
        the route is off or the batch is too large to keep resident."""
        if not self.keep_activations:
            return None
        key = (opts.n_rays, opts.n_samples, opts.n_importance, dev)
        cached = getattr(self, '_state_buf', None)
        if cached is not None and cached[0] == key:
            return cached[1]
        with torch.cuda.device(dev):
            nb = _lib.train_state_bytes(self._get_plan(), opts)
            self._state_buf = None                  # release the old one first
            try:
                buf = torch.empty(nb, dtype=torch.uint8, device=dev) if nb > 0 else None
            except torch.cuda.OutOfMemoryError:     # no room to keep the activations: the recomputing route needs ~half
                buf = None
        self._state_buf = (key, buf)
        return buf

    _train_state_epoch = 0

    def _claim_train_state(self):
        """Every forward that fills the state buffer takes a new epoch; a backward whose epoch is no longer current
        (two forwards before one backward) falls back to recomputing."""
        self._train_state_epoch += 1
        return self._train_state_epoch

    def _packed_image(self, which):
        """Packed tensor-core image of a network, re-packed whenever a parameter changed in place
        (optimizer step, load_state_dict) -- tracked through the tensors' version counters."""
        if which == 'network_fine' and self.network_fine is self.network:      # --single_net: one image serves both passes
            which = 'network'
        net = self.network if which == 'network' else self.network_fine
        params = list(net.parameters())
        sig = tuple((p.data_ptr(), p._version) for p in params)
        cached = self._packed.get(which)
        if cached is not None and cached[1] == sig:
            return cached[0]
        sd = {k: v for k, v in net.state_dict().items()}
        dev = params[0].device
        if dev.type != 'cuda':
            raise RuntimeError("anerf_b200.RayCaster: parameters must live on a CUDA device (no CPU path)")
        out = cached[0] if cached is not None and cached[0].device == dev else None
        img = self._get_plan().pack(sd, out)
        self._packed[which] = (img, sig)
        return img

    def _opts(self, n_rays, n_samples, n_importance, lindisp, density_scale, density_fn, eval_mean=False):
        softplus, shift = False, 0.
        if density_fn is not None and getattr(density_fn, 'anerf_softplus', False):
            softplus, shift = True, float(density_fn.anerf_shift)
        # embedder scalars live on the device; read them back only when they change (no per-chunk sync)
        e0, e1 = self.embed_fn, self.embeddirs_fn
        # key: an explicit epoch (update_tau rebinds `tau` to a fresh tensor whose id / version can repeat) plus the
        # in-place version counters (tau.fill_(...), cutoff_dist.data edits)
        sig = (e0._epoch, e1._epoch, e0.tau._version, e1.tau._version, e0.cutoff_dist._version, e1.cutoff_dist._version,
               e0.tau.data_ptr(), e1.tau.data_ptr(), e0.cutoff_dist.data_ptr(), e1.cutoff_dist.data_ptr())
        if getattr(self, '_embed_cache', (None,))[0] != sig:
            self._embed_cache = (sig, float(e0.tau), float(e1.tau), e0.cutoff_dist.detach().float().cpu().tolist(),
                                 e1.cutoff_dist.detach().float().cpu().tolist())
        _, tau_p, tau_v, cp, cv = self._embed_cache
        return _lib.make_opts(n_rays, n_samples, n_importance, tau_pts=tau_p,
                              tau_views=tau_v, cutoff_pts=cp, cutoff_views=cv,
                              lindisp=bool(lindisp), softplus=softplus, softplus_shift=shift,
                              density_scale=float(density_scale), eval_mean_framecode=eval_mean,
                              single_net=self.single_net)

    # ---- reference API -------------------------------------------------------------------------
    @torch.no_grad()
    def forward_eval(self, *args, **kwargs):
        return self.render_rays(*args, **kwargs)

    def forward(self, *args, fwd_type='', **kwargs):
        if fwd_type == 'density':
            return self.render_pts_density(*args, **kwargs)
        elif fwd_type == 'density_color':
            raise NotImplementedError("fwd_type='density_color' needs texture layers the reference never defines")
        elif fwd_type == 'mesh':
            return self.render_mesh_density(*args, **kwargs)
        if not self.training:
            return self.forward_eval(*args, **kwargs)
        return self.render_rays(*args, **kwargs)

    def render_rays(self, ray_batch, N_samples, kp_batch, skts=None, cyls=None, bones=None, cams=None,
                    subject_idxs=None, retraw=False, lindisp=False, perturb=0., N_importance=0, network_fine=None,
                    raw_noise_std=0., ray_noise_std=0., verbose=False, ext_scale=0.001, pytest=False,
                    preproc_kwargs={}, nerf_type="nerf", use_viewdirs=True, pose_idx=None, **unused):
        """One chunk of rays -> {'rgb_map','disp_map','acc_map','alpha'[, 'rgb0','disp0','acc0','alpha0']}
        (reference: core/raycasters.py:361-474, 711-724).

        `pose_idx` (ours, optional int tensor [N]): `skts` then holds one set of bone transforms per POSE ([P,J,4,4], e.g.
        from anerf_b200.pose_opt.PoseOptLayer.forward_poses) and ray n uses skts[pose_idx[n]]; in training the gradient
        w.r.t. `skts` comes back per pose, already summed over each pose's rays."""
        if skts is None or cyls is None:
            raise NotImplementedError("anerf_b200 needs skts and cyls (the reference's skts=None path is unused)")
        if ray_noise_std > 0.:
            raise NotImplementedError("ray_noise_std > 0 is not supported")
        if subject_idxs is not None:
            raise NotImplementedError("subject_idxs is not supported by the default encoders")
        dev = ray_batch.device
        if dev.type != 'cuda':
            raise RuntimeError("anerf_b200.RayCaster: inputs must be CUDA tensors (no CPU path)")
        N = ray_batch.shape[0]
        J = self._n_joints()
        rays = ray_batch[:, :8].float().contiguous()
        pidx = None
        if pose_idx is not None:
            pidx = pose_idx.to(device=dev, dtype=torch.int32).contiguous()
            if pidx.shape[0] != N:
                raise ValueError(f"pose_idx has {pidx.shape[0]} entries for {N} rays")
            skts_c = skts.float().reshape(-1, J, 4, 4).contiguous()     # one transform set per pose, read through the index
        else:
            skts_c = skts.float().expand(N, J, 4, 4).contiguous()      # differentiable when the pose is being refined
        cyls_c = cyls.float().expand(N, cyls.shape[-1]).contiguous()
        density_scale = preproc_kwargs.get('density_scale', 1.0)
        density_fn = preproc_kwargs.get('density_fn', None)
        use_fc = self.network.use_framecode
        cams_c, eval_mean = None, False
        if use_fc:
            if cams is None:
                raise RuntimeError("opt_framecode networks need `cams`")
            cams_c = cams.float().reshape(-1).expand(N).contiguous()
            eval_mean = (not self.training) and bool(cams_c.max() < 0)     # embedding.py:21
        Sc, Si = int(N_samples), int(N_importance)
        t_rand = u_rand = noise0 = noise1 = None
        if perturb > 0.:
            if pytest:
                np.random.seed(0)
                t_rand = torch.as_tensor(np.random.rand(N, Sc), dtype=torch.float32, device=dev)
                np.random.seed(0)
                u_rand = torch.as_tensor(np.random.rand(N, Si), dtype=torch.float32, device=dev) if Si > 0 else None
            else:
                t_rand = torch.rand(N, Sc, device=dev)
                u_rand = torch.rand(N, Si, device=dev) if Si > 0 else None
        if raw_noise_std > 0.:
            if pytest:
                np.random.seed(0)
                noise0 = torch.as_tensor(np.random.rand(N, Sc) * raw_noise_std, dtype=torch.float32, device=dev)
                np.random.seed(0)
                noise1 = torch.as_tensor(np.random.rand(N, Sc + Si) * raw_noise_std, dtype=torch.float32, device=dev)
            else:
                noise0 = torch.randn(N, Sc, device=dev) * (raw_noise_std * density_scale)
                noise1 = torch.randn(N, Sc + Si, device=dev) * (raw_noise_std * density_scale) if Si > 0 else None
        opts = self._opts(N, Sc, Si, lindisp, density_scale, density_fn, eval_mean)
        names = _lib.param_names(self.network.D, use_fc)
        nets = [self.network] + ([self.network_fine] if Si > 0 else [])
        params = [dict(n.named_parameters())[k] for n in nets for k in names]
        if torch.is_grad_enabled() and (skts_c.requires_grad or any(p.requires_grad for p in params)):
            if eval_mean:
                raise NotImplementedError("gradients through the eval-time mean framecode are not supported")
            aux = dict(rays=rays, cyls=cyls_c, cams=cams_c, t_rand=t_rand, u_rand=u_rand, noise0=noise0, noise1=noise1,
                       n_params0=len(names), pose_idx=pidx)
            with torch.cuda.device(dev):
                outs = _RenderRaysFn.apply(self, opts, aux, skts_c, *params)
            keys = [k for k in OUT_KEYS if Si > 0 or not k.endswith('0')]
            return dict(zip(keys, outs))
        with torch.cuda.device(dev):        # plan buffers, packing kernels and the render launch all on the tensors' device
            p0 = self._packed_image('network')
            p1 = self._packed_image('network_fine') if Si > 0 else None
            out = _lib.render_fwd(self._get_plan(), p0, p1, opts, rays, skts_c.detach(), cyls_c, cams_c, t_rand, u_rand,
                                  noise0, noise1, want_taps=bool(retraw), pose_idx=pidx)
        return out

    @torch.no_grad()
    def render_frame(self, H, W, focal, c2w, skts, cyls, cams=None, pixel_idx=None, chunk=4096, N_samples=64, N_importance=0,
                     lindisp=False, preproc_kwargs={}, center=None, near=0., far=1., out=None, **unused):
        """One frame (or the pixels `pixel_idx` of it) of a posed camera, rays generated inside the kernels and the
        pose read once per frame: the same pixels as the reference's
            render(H, W, focal, rays=get_rays(H, W, focal, c2w)[pixel_idx], skts=skts.expand(n, ...), cyls=..., chunk=chunk)
        (core/utils/ray_utils.py:6-28, run_nerf.py:77-98, core/trainer.py:64-143), chunk by chunk -- the near/far repair of
        rays that miss the cylinder stays a chunk-wide mean -- without materialising rays or per-ray poses.
        c2w [3,4] or [4,4]; skts [J,4,4] or [1,J,4,4]; cyls [5] or [1,5]; pixel_idx: optional int tensor of flat pixel
        indices (e.g. from kp_to_valid_rays).  Returns the output dict for those pixels ([n, ...] tensors)."""
        dev = skts.device
        if dev.type != 'cuda':
            raise RuntimeError("anerf_b200.RayCaster: inputs must be CUDA tensors (no CPU path)")
        J = self._n_joints()
        skts_c = skts.float().reshape(-1, J, 4, 4)[0].contiguous()
        cyl_c = cyls.float().reshape(-1, cyls.shape[-1])[0, :5].contiguous()
        c2w_l = [float(x) for x in torch.as_tensor(c2w, dtype=torch.float32).cpu()[:3, :4].reshape(-1)]
        pix = None if pixel_idx is None else torch.as_tensor(pixel_idx, device=dev).to(torch.int32).contiguous()
        n = H * W if pix is None else int(pix.shape[0])
        Sc, Si = int(N_samples), int(N_importance)
        use_fc = self.network.use_framecode
        cam, eval_mean = 0., False
        if use_fc:
            if cams is None:
                raise RuntimeError("opt_framecode networks need `cams`")
            cam = float(torch.as_tensor(cams).reshape(-1)[0])
            eval_mean = (not self.training) and cam < 0
        density_scale = preproc_kwargs.get('density_scale', 1.0)
        density_fn = preproc_kwargs.get('density_fn', None)
        with torch.cuda.device(dev):
            p0 = self._packed_image('network')
            p1 = self._packed_image('network_fine') if Si > 0 else None
        if out is None:
            f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
            out = dict(rgb_map=f(n, 3), disp_map=f(n), acc_map=f(n), alpha=f(n, Sc + Si if Si > 0 else Sc))
            if Si > 0:
                out.update(rgb0=f(n, 3), disp0=f(n), acc0=f(n), alpha0=f(n, Sc))
        with torch.cuda.device(dev):
            for i in range(0, n, chunk):
                m = min(chunk, n - i)
                opts = self._opts(m, Sc, Si, lindisp, density_scale, density_fn, eval_mean)
                _lib.render_frame(self._get_plan(), p0, p1, opts, c2w_l, focal, center, H, W, skts_c, cyl_c,
                                  pixel0=i if pix is None else 0, pixels=None if pix is None else pix[i:i + m], cam=cam,
                                  near=near, far=far, out=out, ray0=i)
        return out

    @torch.no_grad()
    def render_mesh_density(self, kps, skts, bones, subject_idxs=None, radius=1.0, res=64, render_kwargs=None,
                            netchunk=1024 * 64, v=None, first=0, count=None):
        """[res+1]^3 raw densities around kps[0,0] (reference: core/raycasters.py:579-595: np.meshgrid(t, t, t) with
        t = np.linspace(-radius, radius, res+1), + kps[0,0], density of `network_fine`, reshaped and transposed (1,0)).
        The grid points are generated inside the kernel from (root joint, radius, res): only the densities touch HBM.
        `first`/`count` (ours): a slab of the flattened grid instead of the whole volume -> flat [count] tensor."""
        if v is not None:
            raise NotImplementedError("precomputed v is not supported")
        if skts.shape[0] != 1:
            raise NotImplementedError("density queries take one pose (skts [1,J,4,4])")
        dev = skts.device
        if dev.type != 'cuda':
            raise RuntimeError("anerf_b200.RayCaster: inputs must be CUDA tensors (no CPU path)")
        which = 'network_fine' if self.network_fine is not None else 'network'
        opts = self._opts(0, 64, 0, False, 1.0, None)
        with torch.cuda.device(dev):
            sig = _lib.density_grid(self._get_plan(), self._packed_image(which), opts, kps[0, 0].float().contiguous(), radius, res,
                                    skts[0].float().contiguous(), first, count)
        if first != 0 or count is not None:
            return sig
        n1 = res + 1
        return sig.reshape(n1, n1, n1).transpose(1, 0)

    @torch.no_grad()
    def render_pts_density(self, pts, kps, skts, bones, render_kwargs=None, subject_idxs=None, netchunk=1024 * 64,
                           network=None, color=False, v=None):
        """Raw (pre-activation) density of world points under one pose (core/raycasters.py:597-648).
        pts [P,1,3] or [P,3]; returns [P,1,1] like the reference's batchified forward."""
        if color or v is not None:
            raise NotImplementedError("density_color / precomputed v are not supported")
        if network is None:
            which = 'network_fine' if self.network_fine is not None else 'network'
        else:
            which = 'network_fine' if network is self.network_fine else 'network'
        dev = pts.device
        if dev.type != 'cuda':
            raise RuntimeError("anerf_b200.RayCaster: inputs must be CUDA tensors (no CPU path)")
        if skts.shape[0] != 1:
            raise NotImplementedError("density queries take one pose (skts [1,J,4,4])")
        opts = self._opts(0, 64, 0, False, 1.0, None)
        with torch.cuda.device(dev):
            sig = _lib.density_points(self._get_plan(), self._packed_image(which), opts,
                                      pts.reshape(-1, 3).float().contiguous(), skts[0].float().contiguous())
        return sig.reshape(-1, 1, 1)

    def get_subject_joint_coords(self, subject_idxs=None, device=None):
        return self.joint_coords.to(device)[subject_idxs]

    def update_embed_fns(self, global_step, args):
        freq_target = args.multires - 1
        for fn in (self.embed_fn, self.embeddirs_fn, self.embedbones_fn):
            if fn is not None:
                fn.update_threshold(global_step, args.cutoff_step, args.cutoff_rate, args.freq_schedule_step, freq_target)

    # the reference's nested checkpoint layout (core/raycasters.py:752-788)
    @staticmethod
    def _ckpt_key(k):
        if k.endswith("_fine"):
            return f"{k}_state_dict"
        if k.endswith("_fn"):
            return f"{k.split('_fn')[0]}_state_dict"
        if k == "network":
            return "network_fn_state_dict"
        return f"{k}_state_dict"

    def state_dict(self, *args, **kwargs):
        return {self._ckpt_key(k): m.state_dict() for k, m in self._modules.items() if m is not None}

    def load_state_dict(self, ckpt, strict=True):
        for k, m in self._modules.items():
            if m is None:
                continue
            key = self._ckpt_key(k)
            try:
                m.load_state_dict(ckpt[key], strict=strict)
            except (KeyError, RuntimeError):
                if k.startswith('network'):
                    print('Error occur when loading state dict for network. Try loading with strict=False now')
                    own = m.state_dict()
                    filtered = {n: t for n, t in ckpt[key].items() if n in own and own[n].shape == t.shape}
                    m.load_state_dict(filtered, strict=False)
                else:
                    print(f'Error occurr when loading state dict for {key}. The entity is not in the state dict?')

    def get_embed_fns(self):
        return self.embed_fn, self.embedbones_fn, self.embeddirs_fn

    def get_networks(self):
        return self.network, self.network_fine


class _ModuleHolder(nn.Module):
    """What `render_kwargs_train['ray_caster']` is in place of the reference's nn.DataParallel
    (core/raycasters.py:157): callers reach the caster through `.module` (core/trainer.py:265,270,504).
    Multi-GPU runs are one process per GPU (anerf_b200.parallel), so there is nothing to scatter here."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def _softplus_fn(shift):
    fn = lambda x: torch.nn.functional.softplus(x - shift, beta=1)
    fn.anerf_softplus, fn.anerf_shift = True, shift
    return fn


def _load_ckpt_from_path(ray_caster, optimizer, ckpt_path, finetune=False):
    """core/utils/run_nerf_helpers.py:6-17"""
    ckpt = torch.load(ckpt_path, map_location='cpu', weights_only=False)
    global_step = ckpt["global_step"]
    ray_caster.load_state_dict(ckpt)
    if optimizer is not None and not finetune and "optimizer_state_dict" in ckpt:
        print("load optimizer from ckpt")
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return global_step, ray_caster, optimizer, ckpt


def create_raycaster(args, data_attrs, device=None):
    """Same contract as the reference's factory (core/raycasters.py:17-184): returns
    (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer, loaded_ckpt).
    Flag combinations outside the compiled path raise NotImplementedError here."""
    skel_type = data_attrs["skel_type"]
    n_joints = len(skel_type.joint_names)
    n_framecodes = data_attrs["n_views"] if args.n_framecodes is None else args.n_framecodes
    g = lambda name, default: getattr(args, name, default)

    def unsupported(cond, what):
        if cond:
            raise NotImplementedError(f"anerf_b200: {what} is not implemented (only the default A-NeRF encoders are)")
    unsupported(g('pts_tr_type', 'local') != 'local', f"pts_tr_type={g('pts_tr_type', None)}")
    unsupported(g('kp_dist_type', 'reldist') != 'reldist', f"kp_dist_type={g('kp_dist_type', None)}")
    unsupported(g('view_type', 'relray') != 'relray', f"view_type={g('view_type', None)}")
    unsupported(g('bone_type', 'reldir') != 'reldir', f"bone_type={g('bone_type', None)}")
    unsupported(not g('use_viewdirs', True), "use_viewdirs=False")
    unsupported(not g('use_cutoff', True) or not g('cutoff_viewdir', True) or not g('cutoff_inputs', True),
                "disabling the cutoff (use_cutoff / cutoff_viewdir / cutoff_inputs)")
    unsupported(g('normalize_cutoff', False) or g('cut_to_dist', False) or g('cutoff_shift', False)
                or g('cutoff_bones', False) or g('freq_schedule', False) or g('opt_cutoff', False),
                "normalize_cutoff / cut_to_dist / cutoff_shift / cutoff_bones / freq_schedule / opt_cutoff")
    unsupported(g('multires', 7) != MULTIRES or g('multires_views', 4) not in MULTIRES_VIEWS or g('multires_bones', 0) != 0,
                "multires != 7, multires_views not in {4, 0} or multires_bones != 0")
    unsupported(g('i_embed', 0) != 0, "i_embed != 0")
    unsupported(g('nerf_type', 'nerf') != 'nerf', f"nerf_type={g('nerf_type', None)}")
    if args.density_type not in ('relu', 'softplus'):
        raise NotImplementedError(f'density activation {args.density_type} is undefined')
    if args.netwidth not in (64, 128, 256) or not (2 <= args.netdepth <= 8) or not (1 <= n_joints <= 24):
        raise NotImplementedError("anerf_b200 supports netwidth in {64,128,256}, netdepth 2..8, up to 24 joints")

    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
    cutoff_dist = args.cutoff_mm * args.ext_scale
    embed_fn = CutoffEmbedder(n_joints, MULTIRES, cutoff_dist, n_joints, dist_inputs=False)
    embedbones_fn = Embedder(n_joints * 3, 0)
    embeddirs_fn = CutoffEmbedder(n_joints * 3, g('multires_views', 4), cutoff_dist, n_joints, dist_inputs=True)
    print(f'KPE: RelDist, BPE: VecNorm, VPE: VecNorm')

    output_ch = 5 if args.N_importance > 0 else 4
    nerf_kwargs = dict(D=args.netdepth, W=args.netwidth, input_ch=embed_fn.out_dim, input_ch_bones=embedbones_fn.out_dim,
                       input_ch_views=embeddirs_fn.out_dim, output_ch=output_ch, skips=[4], use_viewdirs=True,
                       use_framecode=bool(args.opt_framecode), framecode_ch=args.framecode_size,
                       n_framecodes=n_framecodes, skel_type=skel_type, density_scale=args.density_scale)
    model = NeRF(**nerf_kwargs)
    single_net = bool(g('single_net', False))
    model_fine = None
    if args.N_importance > 0:          # --single_net: the fine pass re-uses the coarse network (core/raycasters.py:100-104)
        model_fine = model if single_net else NeRF(**nerf_kwargs)
    ray_caster = RayCaster(model, embed_fn, embedbones_fn, embeddirs_fn, network_fine=model_fine,
                           joint_coords=torch.tensor(data_attrs['joint_coords']), single_net=single_net).to(device)

    # trainable variables, with the reference's --fix_layer freezing (core/raycasters.py:186-228)
    if g('finetune', False) and g('fix_layer', 0) > 0:
        for net in (model, model_fine):
            if net is not None:
                for i, l in enumerate(net.pts_linears):
                    if i < args.fix_layer:
                        for p in l.parameters():
                            p.requires_grad = False
    grad_vars = []
    if g('weight_decay', None) is None:
        for m in (model, None if single_net else model_fine, embed_fn, embedbones_fn, embeddirs_fn):
            if m is not None:
                grad_vars += [p for p in m.parameters() if p.requires_grad]
    # torch.optim.Adam's hyper-parameters, state layout and state_dict (core/raycasters.py:116), one launch per step
    # (anerf_b200.optim.FusedAdam); plain torch.optim.Adam when the parameters are not on a GPU
    if torch.device(device).type == 'cuda':
        from .optim import FusedAdam
        optimizer = FusedAdam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))
    else:
        optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start = 0
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        d = os.path.join(args.basedir, args.expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if 'tar' in f and 'pose' not in f] if os.path.isdir(d) else []
    print('Found ckpts', ckpts)
    loaded_ckpt = None
    if len(ckpts) > 0 and not args.no_reload:
        print('Reloading from', ckpts[-1])
        start, ray_caster, optimizer, loaded_ckpt = _load_ckpt_from_path(ray_caster, optimizer, ckpts[-1], args.finetune)
        if args.finetune:
            start = 0
            print(f"set global step to {start}")

    density_fn = torch.nn.functional.relu if args.density_type == 'relu' else _softplus_fn(args.softplus_shift)
    preproc_kwargs = {
        'pts_tr_fn': _EncoderTag('W2LEncoder', n_joints),
        'kp_input_fn': _EncoderTag('RelDist', n_joints),
        'view_input_fn': _EncoderTag('VecNorm', n_joints * 3),
        'bone_input_fn': _EncoderTag('VecNorm', n_joints * 3),
        'density_scale': args.density_scale,
        'density_fn': density_fn,
    }
    render_kwargs_train = {
        'ray_caster': _ModuleHolder(ray_caster),
        'perturb': args.perturb, 'N_importance': args.N_importance, 'N_samples': args.N_samples,
        'use_viewdirs': args.use_viewdirs, 'raw_noise_std': args.raw_noise_std, 'ray_noise_std': args.ray_noise_std,
        'ext_scale': args.ext_scale, 'preproc_kwargs': preproc_kwargs, 'lindisp': args.lindisp,
        'nerf_type': args.nerf_type,
    }
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test['ray_caster'] = ray_caster
    render_kwargs_test['preproc_kwargs'] = dict(preproc_kwargs)
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    render_kwargs_test['ray_noise_std'] = 0.
    print(f"#parameters: {sum(p.numel() for p in model.parameters() if p.requires_grad)}")
    optimizer.zero_grad()
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer, loaded_ckpt


def batchify_rays(rays_flat, chunk=1024 * 32, ray_caster=None, **kwargs):
    """The caller of the boundary, restated (core/trainer.py:64-79): slices every tensor kwarg per chunk.
    Provided so that bench / tests drive the caster exactly as the reference's `render` does."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        kw = {k: (v[i:i + chunk] if torch.is_tensor(v) else v) for k, v in kwargs.items()}
        ret = ray_caster(rays_flat[i:i + chunk], **kw)
        for k, v in ret.items():
            all_ret.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in all_ret.items()}
