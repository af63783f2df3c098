"""ctypes binding of libanerf_b200.so (the C ABI in include/anerf_b200.h).

The library is the only compute path: if it is missing or a call fails, this module raises --
there is no PyTorch or CPU fallback behind these functions."""
import ctypes as C
import os

import torch

# ANERF_B200_LIB points at an alternative build of the same library (tools/ab_variants.py: kernel A/B runs)
_LIB_PATH = os.environ.get("ANERF_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc",
                                                             "libanerf_b200.so")
_lib = None

SYMBOLS = ["anerf_plan_create", "anerf_plan_destroy", "anerf_packed_bytes", "anerf_pack_net",
           "anerf_render_workspace_bytes", "anerf_render_fwd", "anerf_render_fwd_host", "anerf_density_points",
           "anerf_selftest_gemm", "anerf_last_error", "anerf_version", "anerf_debug_set_trace",
           "anerf_render_bwd", "anerf_render_bwd_workspace_bytes", "anerf_selftest_tc_gemm", "anerf_render_frame",
           "anerf_check_status", "anerf_density_grid", "anerf_render_fwd_host_chunked", "anerf_pose_chain_fwd",
           "anerf_pose_chain_bwd", "anerf_pose_chain_bwd_scratch_bytes", "anerf_adam_step", "anerf_mc_count", "anerf_mc_emit", "anerf_sample_rays", "anerf_render_bwd_pass", "anerf_loss_seed",
           "anerf_train_state_bytes", "anerf_render_fwd_train", "anerf_render_bwd_saved"]


class NetConfig(C.Structure):
    _fields_ = [("n_joints", C.c_int32), ("depth", C.c_int32), ("width", C.c_int32), ("skip", C.c_int32),
                ("framecode_ch", C.c_int32), ("n_framecodes", C.c_int32), ("operand_format", C.c_int32),
                ("view_freqs", C.c_int32)]


class NetParams(C.Structure):
    _fields_ = [("pts_w", C.c_void_p * 8), ("pts_b", C.c_void_p * 8), ("alpha_w", C.c_void_p), ("alpha_b", C.c_void_p),
                ("feature_w", C.c_void_p), ("feature_b", C.c_void_p), ("views_w", C.c_void_p), ("views_b", C.c_void_p),
                ("rgb_w", C.c_void_p), ("rgb_b", C.c_void_p), ("framecodes", C.c_void_p)]


class NetGrads(C.Structure):      # same layout as NetParams (gradient buffers, accumulated into)
    _fields_ = NetParams._fields_


class RenderGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0", "alpha0")]


class RenderOpts(C.Structure):
    _fields_ = [("n_rays", C.c_int32), ("n_samples", C.c_int32), ("n_importance", C.c_int32), ("lindisp", C.c_int32),
                ("softplus", C.c_int32), ("eval_mean_framecode", C.c_int32), ("density_scale", C.c_float),
                ("softplus_shift", C.c_float), ("tau_pts", C.c_float), ("tau_views", C.c_float),
                ("cutoff_pts", C.c_float * 24), ("cutoff_views", C.c_float * 24), ("single_net", C.c_int32),
                ("reserved", C.c_int32)]


class RenderInputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rays", "skts", "cyls", "cams", "t_rand", "u_rand", "noise0", "noise1", "pose_idx")] + \
               [("n_poses", C.c_int32), ("reserved", C.c_int32)]


class FrameInputs(C.Structure):
    _fields_ = [("c2w", C.c_float * 12), ("focal_x", C.c_float), ("focal_y", C.c_float), ("center_x", C.c_float),
                ("center_y", C.c_float), ("near", C.c_float), ("far", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("pixel0", C.c_int32), ("pixels", C.c_void_p), ("skts", C.c_void_p), ("cyl", C.c_void_p), ("cam", C.c_float),
                ("reserved", C.c_int32)]


class RenderOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0", "alpha0",
                                          "z_all", "raw")]


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(f"{_LIB_PATH} is missing: build it with `python -m anerf_b200.build` "
                           "(anerf_b200 has no fallback path)")
    lib = C.CDLL(_LIB_PATH)
    lib.anerf_last_error.restype = C.c_char_p
    lib.anerf_packed_bytes.restype = C.c_size_t
    lib.anerf_packed_bytes.argtypes = [C.c_void_p]
    lib.anerf_render_workspace_bytes.restype = C.c_size_t
    lib.anerf_render_workspace_bytes.argtypes = [C.c_int32]
    lib.anerf_plan_create.argtypes = [C.POINTER(NetConfig), C.POINTER(C.c_void_p)]
    lib.anerf_plan_destroy.argtypes = [C.c_void_p]
    lib.anerf_plan_destroy.restype = None
    lib.anerf_pack_net.argtypes = [C.c_void_p, C.POINTER(NetParams), C.c_void_p, C.c_void_p]
    lib.anerf_render_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RenderOpts), C.POINTER(RenderInputs),
                                     C.POINTER(RenderOutputs), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.anerf_render_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RenderOpts), C.POINTER(FrameInputs),
                                       C.POINTER(RenderOutputs), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.anerf_render_fwd_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RenderOpts),
                                          C.POINTER(RenderInputs), C.POINTER(RenderOutputs), C.c_void_p]
    lib.anerf_density_points.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RenderOpts), C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p]
    lib.anerf_density_grid.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(RenderOpts), C.c_void_p, C.c_double, C.c_int32, C.c_int64,
                                       C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.anerf_render_fwd_host_chunked.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RenderOpts), C.c_int32,
                                                  C.POINTER(RenderInputs), C.POINTER(RenderOutputs)]
    lib.anerf_pose_chain_fwd.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.anerf_pose_chain_bwd_scratch_bytes.restype = C.c_size_t
    lib.anerf_pose_chain_bwd_scratch_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.anerf_pose_chain_bwd.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.anerf_adam_step.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                    C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.anerf_mc_count.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]
    lib.anerf_mc_emit.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
    lib.anerf_selftest_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.anerf_render_bwd_workspace_bytes.restype = C.c_size_t
    lib.anerf_render_bwd_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.anerf_render_bwd.argtypes = [C.c_void_p, C.POINTER(NetParams), C.POINTER(NetParams), C.POINTER(RenderOpts),
                                     C.POINTER(RenderInputs), C.c_void_p, C.c_void_p, C.POINTER(RenderGrads),
                                     C.POINTER(NetGrads), C.POINTER(NetGrads), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.anerf_render_bwd_pass.argtypes = [C.c_void_p, C.POINTER(NetParams), C.POINTER(NetParams), C.POINTER(RenderOpts),
                                          C.POINTER(RenderInputs), C.c_void_p, C.c_void_p, C.POINTER(RenderGrads),
                                          C.POINTER(NetGrads), C.POINTER(NetGrads), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]
    lib.anerf_train_state_bytes.restype = C.c_size_t
    lib.anerf_train_state_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.anerf_render_fwd_train.argtypes = [C.c_void_p, C.POINTER(NetParams), C.POINTER(NetParams), C.POINTER(RenderOpts),
                                           C.POINTER(RenderInputs), C.POINTER(RenderOutputs), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.anerf_render_bwd_saved.argtypes = [C.c_void_p, C.POINTER(NetParams), C.POINTER(NetParams), C.POINTER(RenderOpts),
                                           C.POINTER(RenderInputs), C.POINTER(RenderGrads), C.POINTER(NetGrads), C.POINTER(NetGrads),
                                           C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]
    lib.anerf_loss_seed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.anerf_selftest_tc_gemm.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                           C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.anerf_debug_set_trace.argtypes = [C.c_void_p]
    lib.anerf_debug_set_trace.restype = None
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"anerf_b200: {load().anerf_last_error().decode()} (status {rc})")


def check_status():
    """After a stream synchronisation: raises if a kernel recorded a protocol error (anerf_check_status)."""
    check(load().anerf_check_status())


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Plan:
    """Owns an anerf_plan (layer program + K maps of one network configuration)."""

    def __init__(self, n_joints, depth, width, skips=(4,), framecode_ch=0, n_framecodes=0, operand_format=1, view_freqs=4):
        skip = -1
        for s in skips:
            if s < depth - 1:
                if skip >= 0:
                    raise NotImplementedError("only one skip connection is supported")
                skip = int(s)
        self.cfg = NetConfig(n_joints, depth, width, skip, framecode_ch, n_framecodes, operand_format, view_freqs)
        self.handle = C.c_void_p()
        check(load().anerf_plan_create(C.byref(self.cfg), C.byref(self.handle)))
        self.packed_bytes = load().anerf_packed_bytes(self.handle)

    def __del__(self):
        try:
            if self.handle:
                load().anerf_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def pack(self, sd, out=None):
        """sd: mapping of reference state_dict names -> fp32 CUDA tensors.  Returns the packed image (uint8)."""
        dev = sd['pts_linears.0.weight'].device
        if out is None:
            out = torch.empty(self.packed_bytes, dtype=torch.uint8, device=dev)
        keep = []

        def p(name):
            t = sd[name].detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            return t.data_ptr()
        prm = NetParams()
        for i in range(self.cfg.depth):
            prm.pts_w[i] = p(f'pts_linears.{i}.weight')
            prm.pts_b[i] = p(f'pts_linears.{i}.bias')
        prm.alpha_w, prm.alpha_b = p('alpha_linear.weight'), p('alpha_linear.bias')
        prm.feature_w, prm.feature_b = p('feature_linear.weight'), p('feature_linear.bias')
        prm.views_w, prm.views_b = p('views_linears.0.weight'), p('views_linears.0.bias')
        prm.rgb_w, prm.rgb_b = p('rgb_linear.weight'), p('rgb_linear.bias')
        if self.cfg.framecode_ch > 0:
            prm.framecodes = p('framecodes.codes.weight')
        check(load().anerf_pack_net(self.handle, C.byref(prm), _ptr(out), _stream()))
        return out


def make_opts(n_rays, n_samples, n_importance, tau_pts=20., tau_views=20., cutoff_pts=0.5, cutoff_views=0.5,
              n_joints=24, lindisp=False, softplus=False, softplus_shift=0., density_scale=1.,
              eval_mean_framecode=False, single_net=False):
    o = RenderOpts()
    o.single_net = int(single_net)
    o.n_rays, o.n_samples, o.n_importance = n_rays, n_samples, n_importance
    o.lindisp, o.softplus, o.eval_mean_framecode = int(lindisp), int(softplus), int(eval_mean_framecode)
    o.density_scale, o.softplus_shift, o.tau_pts, o.tau_views = density_scale, softplus_shift, tau_pts, tau_views
    for name, v in (("cutoff_pts", cutoff_pts), ("cutoff_views", cutoff_views)):
        arr = getattr(o, name)
        vals = [float(v)] * 24 if not hasattr(v, '__len__') else [float(x) for x in v] + [0.] * (24 - len(v))
        for j in range(24):
            arr[j] = vals[j]
    return o


def render_fwd(plan, packed_coarse, packed_fine, opts, rays, skts, cyls, cams=None, t_rand=None, u_rand=None,
               noise0=None, noise1=None, want_taps=False, keep_nearfar=False, want_z_all=False, pose_idx=None):
    """One chunk on the device.  All tensors fp32 contiguous CUDA.  Returns the reference's output dict
    (core/raycasters.py:711-724) plus 'z_all'/'raw' taps when asked.  pose_idx (int32 [N]): `skts` is then [P,J,4,4],
    one set of bone transforms per pose, read through the index."""
    assert pose_idx is None or (pose_idx.is_cuda and pose_idx.dtype == torch.int32 and pose_idx.is_contiguous())
    N, Sc, Si = opts.n_rays, opts.n_samples, opts.n_importance
    dev = rays.device
    Sf = Sc + Si
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    out = dict(rgb_map=f(N, 3), disp_map=f(N), acc_map=f(N), alpha=f(N, Sf if Si > 0 else Sc))
    if Si > 0:
        out.update(rgb0=f(N, 3), disp0=f(N), acc0=f(N), alpha0=f(N, Sc))
    if want_taps:
        out['raw'] = f(N, Sf if Si > 0 else Sc, 4)
    if (want_taps or want_z_all) and Si > 0:
        out['z_all'] = f(N, Sf)
    ws_bytes = load().anerf_render_workspace_bytes(N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    for t in (rays, skts, cyls, cams, t_rand, u_rand, noise0, noise1):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    rin = RenderInputs(_ptr(rays), _ptr(skts), _ptr(cyls), _ptr(cams), _ptr(t_rand), _ptr(u_rand), _ptr(noise0), _ptr(noise1),
                       _ptr(pose_idx), 0 if pose_idx is None else int(skts.shape[0]), 0)
    rout = RenderOutputs(*[_ptr(out.get(k)) for k in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0",
                                                      "alpha0", "z_all", "raw")])
    check(load().anerf_render_fwd(plan.handle, _ptr(packed_coarse), _ptr(packed_fine), C.byref(opts), C.byref(rin),
                                  C.byref(rout), _ptr(ws), ws_bytes, _stream()))
    if keep_nearfar:        # the repaired near/far of every ray [N,2]: what the backward pass resamples from
        out['nearfar'] = ws.view(torch.float32)[:2 * N].view(N, 2)
    return out


def render_frame(plan, packed_coarse, packed_fine, opts, c2w, focal, center, H, W, skts, cyl, pixel0=0, pixels=None, cam=0.,
                 near=0., far=1., out=None, ray0=0):
    """A chunk of one frame with rays generated in the kernels (C ABI anerf_render_frame).  c2w: 12 floats (rows 0..2 of
    the camera-to-world matrix); skts [J,4,4], cyl [5]: fp32 CUDA tensors; pixels: optional int32 CUDA tensor [n_rays].
    `out`: dict of preallocated full-frame output tensors to write rows [ray0, ray0 + n_rays) of; else fresh tensors."""
    N, Sc, Si = opts.n_rays, opts.n_samples, opts.n_importance
    dev = skts.device
    Sf = Sc + Si
    if out is None:
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        out = dict(rgb_map=f(N, 3), disp_map=f(N), acc_map=f(N), alpha=f(N, Sf if Si > 0 else Sc))
        if Si > 0:
            out.update(rgb0=f(N, 3), disp0=f(N), acc0=f(N), alpha0=f(N, Sc))
        views = out
    else:
        views = {k: v[ray0:ray0 + N] for k, v in out.items()}
    assert skts.is_cuda and skts.dtype == torch.float32 and skts.is_contiguous() and cyl.is_cuda and cyl.is_contiguous()
    assert pixels is None or (pixels.is_cuda and pixels.dtype == torch.int32 and pixels.is_contiguous())
    fr = FrameInputs()
    for i, v in enumerate(c2w):
        fr.c2w[i] = float(v)
    fr.focal_x, fr.focal_y = (float(focal), float(focal)) if not hasattr(focal, '__len__') else (float(focal[0]), float(focal[1]))
    fr.center_x, fr.center_y = (W * 0.5, H * 0.5) if center is None else (float(center[0]), float(center[1]))
    fr.near, fr.far, fr.width, fr.height, fr.pixel0 = near, far, W, H, pixel0
    fr.pixels, fr.skts, fr.cyl, fr.cam = _ptr(pixels), _ptr(skts), _ptr(cyl), float(cam)
    ws_bytes = load().anerf_render_workspace_bytes(N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rout = RenderOutputs(*[_ptr(views.get(k)) for k in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0",
                                                        "alpha0", "z_all", "raw")])
    check(load().anerf_render_frame(plan.handle, _ptr(packed_coarse), _ptr(packed_fine), C.byref(opts), C.byref(fr),
                                    C.byref(rout), _ptr(ws), ws_bytes, _stream()))
    return views


def render_fwd_host(plan, packed_coarse, packed_fine, opts, rays, skts, cyls, cams=None, out=None):
    """Same through host (CPU, ideally pinned) tensors: H2D + kernels + D2H inside the call."""
    N, Sc, Si = opts.n_rays, opts.n_samples, opts.n_importance
    Sf = Sc + Si
    if out is None:
        f = lambda *s: torch.empty(*s, dtype=torch.float32).pin_memory()
        out = dict(rgb_map=f(N, 3), disp_map=f(N), acc_map=f(N), alpha=f(N, Sf if Si > 0 else Sc))
        if Si > 0:
            out.update(rgb0=f(N, 3), disp0=f(N), acc0=f(N), alpha0=f(N, Sc))
    for t in (rays, skts, cyls, cams):
        assert t is None or (not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    rin = RenderInputs(_ptr(rays), _ptr(skts), _ptr(cyls), _ptr(cams), None, None, None, None)
    rout = RenderOutputs(*[_ptr(out.get(k)) for k in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0",
                                                      "alpha0", "z_all", "raw")])
    check(load().anerf_render_fwd_host(plan.handle, _ptr(packed_coarse), _ptr(packed_fine), C.byref(opts), C.byref(rin),
                                       C.byref(rout), _stream()))
    return out


def render_fwd_host_chunked(plan, packed_coarse, packed_fine, opts, chunk, rays, skts, cyls, cams=None, out=None,
                            keys=("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0", "alpha0")):
    """A whole frame (opts.n_rays rays, any number) of host tensors in chunks of `chunk` rays, host<->device copies
    overlapped with the kernels (C ABI anerf_render_fwd_host_chunked).  `keys`: which outputs to produce on the host."""
    N, Sc, Si = opts.n_rays, opts.n_samples, opts.n_importance
    Sf = Sc + Si
    shapes = dict(rgb_map=(N, 3), disp_map=(N,), acc_map=(N,), alpha=(N, Sf if Si > 0 else Sc), rgb0=(N, 3), disp0=(N,),
                  acc0=(N,), alpha0=(N, Sc))
    if out is None:
        out = {k: torch.empty(*shapes[k], dtype=torch.float32).pin_memory() for k in keys if Si > 0 or not k.endswith("0")}
    for t in (rays, skts, cyls, cams):
        assert t is None or (not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    for k, t in out.items():
        assert tuple(t.shape) == shapes[k] and not t.is_cuda and t.is_contiguous()
    rin = RenderInputs(_ptr(rays), _ptr(skts), _ptr(cyls), _ptr(cams), None, None, None, None)
    rout = RenderOutputs(*[_ptr(out.get(k)) for k in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0",
                                                      "alpha0", "z_all", "raw")])
    check(load().anerf_render_fwd_host_chunked(plan.handle, _ptr(packed_coarse), _ptr(packed_fine), C.byref(opts), int(chunk),
                                               C.byref(rin), C.byref(rout)))
    return out


def density_grid(plan, packed, opts, origin, radius, res, skts, first=0, count=None):
    """Raw densities of voxels [first, first+count) of the reference's flattened (res+1)^3 grid around `origin`
    (CUDA tensor, 3 floats), points generated in the kernel (C ABI anerf_density_grid).  Returns sigma [count]."""
    total = (res + 1) ** 3
    count = total - first if count is None else count
    assert origin.is_cuda and origin.dtype == torch.float32 and origin.is_contiguous() and origin.numel() == 3
    assert skts.is_cuda and skts.dtype == torch.float32 and skts.is_contiguous()
    sigma = torch.empty(count, dtype=torch.float32, device=skts.device)
    check(load().anerf_density_grid(plan.handle, _ptr(packed), C.byref(opts), _ptr(origin), float(radius), int(res), int(first),
                                    int(count), _ptr(skts), _ptr(sigma), _stream()))
    return sigma


PARAM_ORDER = (["pts_w", "pts_b"], ["alpha_w", "alpha_b", "feature_w", "feature_b", "views_w", "views_b", "rgb_w", "rgb_b"])


def param_names(depth, framecodes):
    """state_dict names of one network in the order the training entry points take them."""
    names = []
    for i in range(depth):
        names += [f'pts_linears.{i}.weight', f'pts_linears.{i}.bias']
    names += ['alpha_linear.weight', 'alpha_linear.bias', 'feature_linear.weight', 'feature_linear.bias',
              'views_linears.0.weight', 'views_linears.0.bias', 'rgb_linear.weight', 'rgb_linear.bias']
    if framecodes:
        names.append('framecodes.codes.weight')
    return names


def _fill_net_struct(st, depth, tensors, framecodes):
    """tensors: list in param_names() order (entries may be None) -> NetParams / NetGrads fields."""
    it = iter(tensors)
    for i in range(depth):
        w, b = next(it), next(it)
        st.pts_w[i] = None if w is None else w.data_ptr()
        st.pts_b[i] = None if b is None else b.data_ptr()
    for f in PARAM_ORDER[1]:
        t = next(it)
        setattr(st, f, None if t is None else t.data_ptr())
    if framecodes:
        t = next(it)
        st.framecodes = None if t is None else t.data_ptr()
    return st


def render_bwd(plan, opts, params0, params1, rays, skts, cams, t_rand, noise0, noise1, nearfar, z_all, grad_out,
               want0, want1, want_skts, pose_idx=None, into0=None, into1=None, pass_mask=3, g_skts=None, workspace=None,
               state=None):
    """Backward of render_fwd (C ABI anerf_render_bwd).  params0/params1: fp32 CUDA tensors of the coarse / fine
    network in param_names() order; want0/want1: per-parameter flags; grad_out: dict of dL/d(output) tensors (or
    None).  Returns (grads0, grads1, g_skts): gradients (None where not wanted).  into0 / into1: optional lists of
    existing fp32 buffers the kernels ADD the parameter gradients into (entries may be None); everything else comes
    zero-filled out of ONE flat allocation.  state: the buffer render_fwd_train filled for these very inputs -- the
    backward then starts from the kept activations (anerf_render_bwd_saved; nearfar / z_all / workspace are not used)."""
    N, Sc, Si = opts.n_rays, opts.n_samples, opts.n_importance
    dev = rays.device
    depth, fc = plan.cfg.depth, plan.cfg.framecode_ch > 0
    for t in [rays, skts, cams, t_rand, noise0, noise1, nearfar, z_all] + list(params0) + list(params1 or []) + \
            [g for g in grad_out.values() if g is not None]:
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    into0 = into0 or [None] * len(params0)
    into1 = into1 or [None] * len(params1 or [])
    need = [(p, w, t) for p, w, t in list(zip(params0, want0, into0)) + list(zip(params1 or [], want1 or [], into1))]
    n_new = sum(-(-p.numel() // 4) * 4 for p, w, t in need if w and t is None)
    flat = torch.zeros(n_new, dtype=torch.float32, device=dev) if n_new else None
    off = [0]

    def target(p, want, into):
        if not want:
            return None
        if into is not None:
            assert into.is_cuda and into.dtype == torch.float32 and into.is_contiguous() and into.shape == p.shape
            return into
        k = p.numel()
        g = flat[off[0]:off[0] + k].view_as(p)
        off[0] += -(-k // 4) * 4                      # 16-byte aligned pieces
        return g
    g0 = [target(p, w, t) for p, w, t in zip(params0, want0, into0)]
    g1 = [target(p, w, t) for p, w, t in zip(params1, want1, into1)] if params1 is not None else None
    if want_skts and g_skts is None:
        g_skts = torch.zeros_like(skts)
    p0s, g0s = _fill_net_struct(NetParams(), depth, params0, fc), _fill_net_struct(NetGrads(), depth, g0, fc)
    p1s = _fill_net_struct(NetParams(), depth, params1, fc) if params1 is not None else None
    g1s = _fill_net_struct(NetGrads(), depth, g1, fc) if params1 is not None else None
    rin = RenderInputs(_ptr(rays), _ptr(skts), None, _ptr(cams), _ptr(t_rand), None, _ptr(noise0), _ptr(noise1), _ptr(pose_idx),
                       0 if pose_idx is None else int(skts.shape[0]), 0)
    rg = RenderGrads(*[_ptr(grad_out.get(k)) for k in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0", "alpha0")])
    if state is not None:
        assert state.is_cuda and state.dtype == torch.uint8 and state.is_contiguous()
        check(load().anerf_render_bwd_saved(plan.handle, C.byref(p0s), None if p1s is None else C.byref(p1s), C.byref(opts),
                                            C.byref(rin), C.byref(rg), C.byref(g0s), None if g1s is None else C.byref(g1s),
                                            _ptr(g_skts) if want_skts else None, _ptr(state), state.numel(), int(pass_mask), _stream()))
        return g0, g1, g_skts
    ws_bytes = load().anerf_render_bwd_workspace_bytes(plan.handle, N, Sc, Si)
    ws = workspace if workspace is not None and workspace.numel() >= ws_bytes else torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(load().anerf_render_bwd_pass(plan.handle, C.byref(p0s), None if p1s is None else C.byref(p1s), C.byref(opts),
                                       C.byref(rin), _ptr(nearfar), _ptr(z_all), C.byref(rg), C.byref(g0s),
                                       None if g1s is None else C.byref(g1s), _ptr(g_skts) if want_skts else None, _ptr(ws), ws_bytes,
                                       int(pass_mask), _stream()))
    return g0, g1, g_skts


def train_state_bytes(plan, opts):
    """Bytes of the saved-activation state for this batch shape; 0 = too large to keep resident (use render_fwd + render_bwd)."""
    return int(load().anerf_train_state_bytes(plan.handle, opts.n_rays, opts.n_samples, opts.n_importance))


def render_fwd_train(plan, opts, params0, params1, rays, skts, cyls, state, cams=None, t_rand=None, u_rand=None, noise0=None,
                     noise1=None, pose_idx=None):
    """Forward of a training step that keeps its activations in `state` (uint8 CUDA tensor of train_state_bytes()):
    C ABI anerf_render_fwd_train.  params0 / params1: fp32 CUDA parameter tensors in param_names() order (no packed
    image involved).  Returns render_fwd's dict + 'nearfar' [N,2] + 'z_all' [N,Sc+Si]."""
    assert pose_idx is None or (pose_idx.is_cuda and pose_idx.dtype == torch.int32 and pose_idx.is_contiguous())
    assert state.is_cuda and state.dtype == torch.uint8 and state.is_contiguous()
    N, Sc, Si = opts.n_rays, opts.n_samples, opts.n_importance
    dev = rays.device
    Sf = Sc + Si
    depth, fc = plan.cfg.depth, plan.cfg.framecode_ch > 0
    for t in [rays, skts, cyls, cams, t_rand, u_rand, noise0, noise1] + list(params0) + list(params1 or []):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    out = dict(rgb_map=f(N, 3), disp_map=f(N), acc_map=f(N), alpha=f(N, Sf if Si > 0 else Sc), nearfar=f(N, 2))
    if Si > 0:
        out.update(rgb0=f(N, 3), disp0=f(N), acc0=f(N), alpha0=f(N, Sc), z_all=f(N, Sf))
    p0s = _fill_net_struct(NetParams(), depth, params0, fc)
    p1s = _fill_net_struct(NetParams(), depth, params1, fc) if params1 is not None else None
    rin = RenderInputs(_ptr(rays), _ptr(skts), _ptr(cyls), _ptr(cams), _ptr(t_rand), _ptr(u_rand), _ptr(noise0), _ptr(noise1),
                       _ptr(pose_idx), 0 if pose_idx is None else int(skts.shape[0]), 0)
    rout = RenderOutputs(*[_ptr(out.get(k)) for k in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0",
                                                      "alpha0", "z_all", "raw")])
    check(load().anerf_render_fwd_train(plan.handle, C.byref(p0s), None if p1s is None else C.byref(p1s), C.byref(opts),
                                        C.byref(rin), C.byref(rout), _ptr(out['nearfar']), _ptr(state), state.numel(), _stream()))
    return out


def bwd_workspace(plan, opts, device):
    """A reusable workspace tensor for render_bwd (the same size every step of a training loop)."""
    nb = load().anerf_render_bwd_workspace_bytes(plan.handle, opts.n_rays, opts.n_samples, opts.n_importance)
    return torch.empty(nb, dtype=torch.uint8, device=device)


def loss_seed(rgb, acc, target, bg, use_background, mse, weight, sums):
    """Trainer._compute_nerf_loss + its gradient in one launch (C ABI anerf_loss_seed).  bg: [N,3] tensor or a float.
    Returns (d loss / d rgb [N,3], d loss / d acc [N]); adds (sum of loss terms, sum of squared errors) to sums[0:2]."""
    N = rgb.shape[0]
    g_rgb, g_acc = torch.empty_like(rgb), torch.empty_like(acc)
    bg_t = bg if torch.is_tensor(bg) else None
    for t in (rgb, acc, target, bg_t, sums):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    check(load().anerf_loss_seed(_ptr(rgb), _ptr(acc), _ptr(target), _ptr(bg_t), 0.0 if bg_t is not None else float(bg), int(use_background),
                                 int(mse), N, float(weight), _ptr(g_rgb), _ptr(g_acc), _ptr(sums), _stream()))
    return g_rgb, g_acc


def _parents_array(parents):
    arr = (C.c_int32 * len(parents))(*[int(x) for x in parents])
    return arr


def pose_chain_fwd(rots, rest_pose, pelvis, parents, root_id=0):
    """PoseOptLayer.calculate_kinematic's chain (C ABI anerf_pose_chain_fwd): rots [P,J,3,3], rest_pose [1|P,J,3],
    pelvis [P,3] (fp32 CUDA, contiguous) -> (l2ws [P,J,4,4], skts [P,J,4,4], kps [P,J,3])."""
    P, J = rots.shape[:2]
    for t in (rots, rest_pose, pelvis):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=rots.device)
    l2ws, skts, kps = f(P, J, 4, 4), f(P, J, 4, 4), f(P, J, 3)
    check(load().anerf_pose_chain_fwd(P, J, _parents_array(parents), int(root_id), _ptr(rots), _ptr(rest_pose), rest_pose.shape[0],
                                      _ptr(pelvis), _ptr(l2ws), _ptr(skts), _ptr(kps), _stream()))
    return l2ws, skts, kps


def pose_chain_bwd(rots, rest_pose, pelvis, parents, root_id, l2ws, skts, g_skts=None, g_l2ws=None, g_kps=None):
    """-> (g_rots [P,J,3,3], g_pelvis [P,3])."""
    P, J = rots.shape[:2]
    for t in (rots, rest_pose, pelvis, l2ws, skts, g_skts, g_l2ws, g_kps):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    g_rots = torch.empty(P, J, 3, 3, dtype=torch.float32, device=rots.device)
    g_pelvis = torch.empty(P, 3, dtype=torch.float32, device=rots.device)
    nb = load().anerf_pose_chain_bwd_scratch_bytes(P, J)
    scratch = torch.empty(max(nb, 4), dtype=torch.uint8, device=rots.device)
    check(load().anerf_pose_chain_bwd(P, J, _parents_array(parents), int(root_id), _ptr(rots), _ptr(rest_pose), rest_pose.shape[0],
                                      _ptr(pelvis), _ptr(l2ws), _ptr(skts), _ptr(g_skts), _ptr(g_l2ws), _ptr(g_kps), _ptr(g_rots),
                                      _ptr(g_pelvis), _ptr(scratch), nb, _stream()))
    return g_rots, g_pelvis


def density_points(plan, packed, opts, pts, skts):
    sigma = torch.empty(pts.shape[0], dtype=torch.float32, device=pts.device)
    check(load().anerf_density_points(plan.handle, _ptr(packed), C.byref(opts), _ptr(pts), _ptr(skts), pts.shape[0],
                                      _ptr(sigma), _stream()))
    return sigma


def selftest_tc_gemm(A, a_strides, M, K, B, b_strides, N, C, c_strides, bias=None, mask=None, mask_ms=0, relu=False, mode=0,
                     slice_chunks=0):
    """C(m,n) (op)= sum_k A(m,k) B(n,k) through the training path's tensor-core GEMM; strides in elements."""
    check(load().anerf_selftest_tc_gemm(_ptr(A), a_strides[0], a_strides[1], M, K, _ptr(B), b_strides[0], b_strides[1], N,
                                        _ptr(C), c_strides[0], c_strides[1], _ptr(bias), _ptr(mask), mask_ms, int(relu), mode,
                                        slice_chunks, _stream()))
    return C


def selftest_gemm(A, B, fmt):
    N, K = B.shape
    D = torch.empty(2, 256, N, dtype=torch.float32, device=A.device)
    check(load().anerf_selftest_gemm(_ptr(A), _ptr(B), _ptr(D), N, K, fmt, _stream()))
    return D
