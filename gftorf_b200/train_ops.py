"""Fused operators either side of the rasterizer in one training iteration:

  assemble_gaussians   raw parameters (+ deformation outputs) -> rasterizer inputs, differentiable
                       (gaussian_renderer/__init__.py:81-105, scene/gaussian_model.py:123-157)
  fused_loss           lambda * ((1-l) * L(img, gt) + l * (1 - SSIM)) with its image gradient
                       (utils/loss_utils.py:17-33,84-114, train.py:204-223)
  FlatAdam             all parameter groups of scene/gaussian_model.py:247-272 in one launch over a
                       flat buffer (torch.optim.Adam(lr=0, eps=1e-15) semantics)

Host side: Python/PyTorch for memory, streams and autograd bookkeeping; all arithmetic is in
libgftorf_b200.so (include/gftorf_train.h).  No CPU or PyTorch fallback: CPU tensors raise.
"""
import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _capi

ADAM_MAX_SEGMENTS = 24


class GftAssembleArgs(C.Structure):
    _fields_ = [("P", C.c_int), ("M", C.c_int), ("isotropic", C.c_int),
                ("include_static", C.c_int), ("include_dynamic", C.c_int)] + \
               [(n, C.c_void_p) for n in (
                   "xyz", "opacity_raw", "scaling_raw", "rotation_raw", "f_dc_color", "f_rest_color",
                   "f_dc_phase", "f_rest_phase", "f_dc_amp", "f_rest_amp", "dyn_index", "d_xyz",
                   "d_rot", "d_sh", "d_sh_p", "means3D", "opacities", "scales", "rotations", "shs",
                   "shs_p")]


class GftAssembleGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "g_means3D", "g_opacities", "g_scales", "g_rotations", "g_shs", "g_shs_p",
        "g_xyz", "g_opacity_raw", "g_scaling_raw", "g_rotation_raw", "g_f_dc_color",
        "g_f_rest_color", "g_f_dc_phase", "g_f_rest_phase", "g_f_dc_amp", "g_f_rest_amp",
        "g_d_xyz", "g_d_rot", "g_d_sh", "g_d_sh_p")]


class GftLossArgs(C.Structure):
    _fields_ = [("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("kind", C.c_int), ("nch", C.c_int),
                ("w", C.c_float), ("lambda_", C.c_float), ("lambda_dssim", C.c_float),
                ("img", C.c_void_p), ("gt", C.c_void_p), ("grad", C.c_void_p), ("loss", C.c_void_p),
                ("scratch", C.c_void_p)]


class GftAdamSegment(C.Structure):
    _fields_ = [("begin", C.c_longlong), ("end", C.c_longlong), ("lr", C.c_double)]


class GftAdamArgs(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p),
                ("exp_avg_sq", C.c_void_p), ("n_segments", C.c_int), ("step", C.c_int),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("zero_grad", C.c_int), ("seg", GftAdamSegment * ADAM_MAX_SEGMENTS)]


class GftDensifyPlanArgs(C.Structure):
    _fields_ = [("P", C.c_int), ("isotropic", C.c_int), ("size_prune", C.c_int),
                ("max_grad", C.c_float), ("min_opacity", C.c_float), ("extent", C.c_float),
                ("percent_dense", C.c_float)] + \
               [(n, C.c_void_p) for n in ("grad_accum", "denom", "opacity_raw", "scaling_raw", "plan_src",
                                          "plan_kind", "plan_noise", "split_src", "counts", "workspace")]


class GftDensifyApplyArgs(C.Structure):
    _fields_ = [("P_new", C.c_int), ("width", C.c_int), ("mode", C.c_int)] + \
               [(n, C.c_void_p) for n in ("plan_src", "plan_kind", "plan_noise", "param_in", "exp_avg_in",
                                          "exp_avg_sq_in", "param_out", "exp_avg_out", "exp_avg_sq_out",
                                          "noise", "rotation_raw")]


_declared = False


def _lib():
    global _declared
    lib = _capi.lib()
    if not _declared:
        lib.gft_assemble_forward.argtypes = [C.POINTER(GftAssembleArgs), C.c_void_p]
        lib.gft_assemble_forward.restype = C.c_int
        lib.gft_assemble_backward.argtypes = [C.POINTER(GftAssembleArgs), C.POINTER(GftAssembleGrads), C.c_void_p]
        lib.gft_assemble_backward.restype = C.c_int
        lib.gft_fused_loss_scratch_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.gft_fused_loss_scratch_bytes.restype = C.c_size_t
        lib.gft_fused_loss.argtypes = [C.POINTER(GftLossArgs), C.c_void_p]
        lib.gft_fused_loss.restype = C.c_int
        lib.gft_adam_step.argtypes = [C.POINTER(GftAdamArgs), C.c_void_p]
        lib.gft_adam_step.restype = C.c_int
        lib.gft_densify_workspace_bytes.argtypes = [C.c_int]
        lib.gft_densify_workspace_bytes.restype = C.c_size_t
        lib.gft_densify_plan.argtypes = [C.POINTER(GftDensifyPlanArgs), C.c_void_p]
        lib.gft_densify_plan.restype = C.c_int
        lib.gft_densify_apply.argtypes = [C.POINTER(GftDensifyApplyArgs), C.c_void_p]
        lib.gft_densify_apply.restype = C.c_int
        lib.gft_nvls_allreduce_sum.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
        lib.gft_nvls_allreduce_sum.restype = C.c_int
        lib.gft_nvls_allreduce_fused.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                                 C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.gft_nvls_allreduce_fused.restype = C.c_int
        lib.gft_p2p_allreduce_fused.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                                C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.gft_p2p_allreduce_fused.restype = C.c_int
        lib.gft_push_allreduce_fused.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.gft_push_allreduce_fused.restype = C.c_int
        _declared = True
    return lib


def _check(rc, what):
    if rc < 0:
        raise RuntimeError(f"{what} failed ({rc}): " + _capi.lib().gft_last_error().decode("utf-8", "replace"))


def _cuda_f32(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"gftorf_b200 has no CPU path: {name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


# ------------------------------------------------------------------------------------------------
# f1 — assembly
# ------------------------------------------------------------------------------------------------
RAW_NAMES = ("xyz", "opacity_raw", "scaling_raw", "rotation_raw", "f_dc_color", "f_rest_color",
             "f_dc_phase", "f_rest_phase", "f_dc_amp", "f_rest_amp")
DELTA_NAMES = ("d_xyz", "d_rot", "d_sh", "d_sh_p")
OUT_NAMES = ("means3D", "opacities", "scales", "rotations", "shs", "shs_p")


def dyn_index_from_mask(motion_mask: torch.Tensor) -> torch.Tensor:
    """int32 [P]: position of each masked Gaussian among the masked ones, -1 elsewhere.  Changes
    only when the set of Gaussians changes (densification), so it is built once and reused."""
    m = motion_mask.bool()
    idx = torch.cumsum(m.to(torch.int32), 0, dtype=torch.int32) - 1
    return torch.where(m, idx, torch.full_like(idx, -1)).contiguous()


def _fill_fwd_args(a, raw, deltas, dyn_index, outs, isotropic, regions):
    P = int(raw["xyz"].shape[0])
    a.P, a.M = P, int(raw["f_rest_color"].shape[1]) + 1
    a.isotropic = int(bool(isotropic))
    a.include_static, a.include_dynamic = int("static" in regions), int("dynamic" in regions)
    for n in RAW_NAMES:
        setattr(a, n, _ptr(raw[n]))
    a.dyn_index = _ptr(dyn_index)
    for n in DELTA_NAMES:
        setattr(a, n, _ptr(deltas.get(n)))
    for n in OUT_NAMES:
        setattr(a, n, _ptr(outs[n]))
    return a


class _AssembleGaussians(torch.autograd.Function):
    """(10 raw parameters, 4 deformation outputs) -> the six rasterizer inputs."""

    @staticmethod
    def forward(ctx, dyn_index, isotropic, regions, *tensors):
        raw = {n: _cuda_f32(t, n) for n, t in zip(RAW_NAMES, tensors[:10])}
        deltas = {n: (_cuda_f32(t, n) if t is not None else None) for n, t in zip(DELTA_NAMES, tensors[10:14])}
        dev = raw["xyz"].device
        P, M = int(raw["xyz"].shape[0]), int(raw["f_rest_color"].shape[1]) + 1
        f32 = dict(dtype=torch.float32, device=dev)
        outs = dict(means3D=torch.empty((P, 3), **f32), opacities=torch.empty((P, 1), **f32),
                    scales=torch.empty((P, 3), **f32), rotations=torch.empty((P, 4), **f32),
                    shs=torch.empty((P, M, 3), **f32), shs_p=torch.empty((P, M, 2), **f32))
        if dyn_index is not None and (dyn_index.dtype != torch.int32 or not dyn_index.is_cuda):
            raise RuntimeError("dyn_index must be an int32 CUDA tensor (see dyn_index_from_mask)")
        a = _fill_fwd_args(GftAssembleArgs(), raw, deltas, dyn_index, outs, isotropic, regions)
        with torch.cuda.device(dev):
            _check(_lib().gft_assemble_forward(C.byref(a), _stream(dev)), "gft_assemble_forward")
        ctx.isotropic, ctx.regions = isotropic, regions
        ctx.have_delta = [t is not None for t in tensors[10:14]]
        ctx.dyn_index = dyn_index
        saved = [raw[n] for n in RAW_NAMES] + [deltas[n] if deltas[n] is not None else torch.empty(0, **f32)
                                               for n in DELTA_NAMES] + [outs[n] for n in OUT_NAMES]
        ctx.save_for_backward(*saved)
        return tuple(outs[n] for n in OUT_NAMES)

    @staticmethod
    def backward(ctx, *gouts):
        saved = ctx.saved_tensors
        raw = dict(zip(RAW_NAMES, saved[:10]))
        deltas = {n: (t if h else None) for n, t, h in zip(DELTA_NAMES, saved[10:14], ctx.have_delta)}
        outs = dict(zip(OUT_NAMES, saved[14:20]))
        dev = raw["xyz"].device
        a = _fill_fwd_args(GftAssembleArgs(), raw, deltas, ctx.dyn_index, outs, ctx.isotropic, ctx.regions)
        g = GftAssembleGrads()
        gin = []
        for n, t, o in zip(OUT_NAMES, gouts, [outs[n] for n in OUT_NAMES]):
            t = torch.zeros_like(o) if t is None else _cuda_f32(t, "grad_" + n)
            gin.append(t)
            setattr(g, "g_" + n, t.data_ptr())
        graw = [torch.empty_like(raw[n]) for n in RAW_NAMES]
        for n, t in zip(RAW_NAMES, graw):
            setattr(g, "g_" + n, _ptr(t))
        gdel = [torch.empty_like(deltas[n]) if deltas[n] is not None else None for n in DELTA_NAMES]
        for n, t in zip(DELTA_NAMES, gdel):
            setattr(g, "g_" + n, _ptr(t))
        with torch.cuda.device(dev):
            _check(_lib().gft_assemble_backward(C.byref(a), C.byref(g), _stream(dev)), "gft_assemble_backward")
        return (None, None, None) + tuple(graw) + tuple(gdel)


def assemble_gaussians(raw: Dict[str, torch.Tensor], dyn_index: Optional[torch.Tensor] = None,
                       deltas: Optional[Dict[str, torch.Tensor]] = None, isotropic: bool = False,
                       render_regions: Sequence[str] = ("static", "dynamic")) -> Dict[str, torch.Tensor]:
    """`raw`: the GaussianModel's parameter tensors under the names RAW_NAMES (`_xyz`, `_opacity`,
    `_scaling`, `_rotation`, `_features_dc_color`, ... of scene/gaussian_model.py:49-62);
    `deltas`: d_xyz / d_rot / d_sh / d_sh_p from the deformation model, in masked order."""
    deltas = deltas or {}
    out = _AssembleGaussians.apply(dyn_index, bool(isotropic), tuple(render_regions),
                                   *[raw[n] for n in RAW_NAMES], *[deltas.get(n) for n in DELTA_NAMES])
    return dict(zip(OUT_NAMES, out))


# ------------------------------------------------------------------------------------------------
# f3 — loss
# ------------------------------------------------------------------------------------------------
LOSS_KINDS = {"l1": 0, "l2": 1, "weighted_l1": 2, "weighted_l1_quad": 3, "weighted_l2_quad": 4}


def fused_loss(img: torch.Tensor, gt: torch.Tensor, kind: str = "l1", lam: float = 1.0,
               lambda_dssim: float = 0.2, w: float = 0.0, nch: Optional[int] = None,
               loss_out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (loss [1], d loss / d img [C,H,W]).  With `loss_out` the value is added to that
    1-element tensor (the terms of train.py:204-223 accumulate into one scalar)."""
    img, gt = _cuda_f32(img, "img"), _cuda_f32(gt, "gt")
    if img.dim() != 3 or img.shape != gt.shape:
        raise RuntimeError("fused_loss: img and gt must both be [C,H,W]")
    dev = img.device
    Cn, H, W = (int(s) for s in img.shape)
    lib = _lib()
    grad = torch.empty_like(img)
    loss = loss_out if loss_out is not None else torch.zeros(1, dtype=torch.float32, device=dev)
    a = GftLossArgs()
    a.C, a.H, a.W = Cn, H, W
    a.kind, a.nch = LOSS_KINDS[kind], int(nch if nch is not None else Cn)
    a.w, a.lambda_, a.lambda_dssim = float(w), float(lam), float(lambda_dssim)
    a.img, a.gt, a.grad, a.loss = img.data_ptr(), gt.data_ptr(), grad.data_ptr(), loss.data_ptr()
    scratch = None
    if lambda_dssim != 0.0:
        scratch = torch.empty(lib.gft_fused_loss_scratch_bytes(Cn, H, W) // 4, dtype=torch.float32, device=dev)
        a.scratch = scratch.data_ptr()
    with torch.cuda.device(dev):
        _check(lib.gft_fused_loss(C.byref(a), _stream(dev)), "gft_fused_loss")
    return loss, grad


# ------------------------------------------------------------------------------------------------
# f4 — Adam
# ------------------------------------------------------------------------------------------------
class FlatAdam:
    """Adam over one flat fp32 buffer holding every parameter group back to back (16-byte aligned
    slices).  `groups`: [(name, tensor, lr)] in the order of scene/gaussian_model.py:247-272; the
    tensors are COPIED into the flat buffer and `self.params[name]` are the live views the model
    should use from then on (their `.grad` are views of the flat gradient buffer, so backward
    accumulates in place and the NCCL allreduce runs on `self.grad` as it stands)."""

    def __init__(self, groups, betas=(0.9, 0.999), eps=1e-15):
        if len(groups) > ADAM_MAX_SEGMENTS:
            raise RuntimeError(f"at most {ADAM_MAX_SEGMENTS} parameter groups")
        dev = groups[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("gftorf_b200 has no CPU path: parameters must be CUDA tensors")
        self.names, self.lr, self.bounds = [], {}, {}
        cur = 0
        for name, t, lr in groups:
            n = t.numel()
            self.names.append(name)
            self.lr[name] = float(lr)
            self.bounds[name] = (cur, cur + n, tuple(t.shape))
            cur += (n + 3) // 4 * 4
        self.flat = torch.zeros(cur, dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.betas, self.eps, self.step_count = betas, eps, 0
        self.params = {}
        for name, t, _ in groups:
            b, e, shape = self.bounds[name]
            self.flat[b:e].copy_(t.detach().reshape(-1).float())
            p = self.flat[b:e].view(shape).requires_grad_(True)
            p.grad = self.grad[b:e].view(shape)
            self.params[name] = p

    def set_lr(self, name, lr):
        self.lr[name] = float(lr)

    def step(self, zero_grad=True):
        self.step_count += 1
        a = GftAdamArgs()
        a.param, a.grad = self.flat.data_ptr(), self.grad.data_ptr()
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.n_segments, a.step = len(self.names), self.step_count
        a.beta1, a.beta2, a.eps = float(self.betas[0]), float(self.betas[1]), float(self.eps)
        a.zero_grad = int(bool(zero_grad))
        for i, name in enumerate(self.names):
            b, e, _ = self.bounds[name]
            a.seg[i].begin, a.seg[i].end, a.seg[i].lr = b, (e + 3) // 4 * 4, self.lr[name]
        dev = self.flat.device
        with torch.cuda.device(dev):
            _check(_lib().gft_adam_step(C.byref(a), _stream(dev)), "gft_adam_step")


# ------------------------------------------------------------------------------------------------
# f2 — densification
# ------------------------------------------------------------------------------------------------
DENSIFY_GROUPS = ("xyz", "f_dc_color", "f_rest_color", "phase_f_dc", "phase_f_rest", "amp_f_dc",
                  "amp_f_rest", "opacity", "scaling", "rotation", "f_seg_color")


def densify_and_prune(params: Dict[str, torch.Tensor], exp_avg: Optional[Dict[str, torch.Tensor]],
                      exp_avg_sq: Optional[Dict[str, torch.Tensor]], grad_accum: torch.Tensor,
                      denom: torch.Tensor, max_grad: float, min_opacity: float, extent: float,
                      percent_dense: float, size_prune: bool = True, isotropic: bool = False,
                      generator: Optional[torch.Generator] = None, samples: Optional[torch.Tensor] = None):
    """GaussianModel.densify_and_prune (scene/gaussian_model.py:631-646) on the parameter groups of
    the optimizer (names of :247-266) and their Adam moments.  Returns (params, exp_avg,
    exp_avg_sq, info) for the new Gaussian set, in the reference's order; the caller restarts the
    densification statistics from zero, as the reference does (:564-566).  The normal samples of
    the split (:579-581) are drawn here with torch.normal on `generator`, exactly the call the
    reference makes, so seeded runs reproduce it; `samples` ([2 * split, 3], already scaled by the
    standard deviations) replaces the draw (used to replay recorded noise)."""
    lib = _lib()
    xyz = _cuda_f32(params["xyz"], "xyz")
    dev = xyz.device
    P = int(xyz.shape[0])
    if P == 0:
        raise RuntimeError("densify_and_prune: empty Gaussian set")
    if not (max_grad > 0.0):
        raise RuntimeError("densify_and_prune: max_grad must be positive")
    src = {g: _cuda_f32(params[g], g) for g in params}
    i32 = dict(dtype=torch.int32, device=dev)
    plan_src, plan_kind = torch.empty(2 * P, **i32), torch.empty(2 * P, **i32)
    plan_noise, split_src = torch.empty(2 * P, **i32), torch.empty(P, **i32)
    counts = torch.zeros(5, **i32)
    ws = torch.empty(lib.gft_densify_workspace_bytes(P), dtype=torch.uint8, device=dev)
    ga, dn = _cuda_f32(grad_accum, "grad_accum"), _cuda_f32(denom, "denom")
    a = GftDensifyPlanArgs()
    a.P, a.isotropic, a.size_prune = P, int(bool(isotropic)), int(bool(size_prune))
    a.max_grad, a.min_opacity, a.extent, a.percent_dense = float(max_grad), float(min_opacity), float(extent), float(percent_dense)
    a.grad_accum, a.denom = ga.data_ptr(), dn.data_ptr()
    a.opacity_raw, a.scaling_raw = src["opacity"].data_ptr(), src["scaling"].data_ptr()
    a.plan_src, a.plan_kind, a.plan_noise = plan_src.data_ptr(), plan_kind.data_ptr(), plan_noise.data_ptr()
    a.split_src, a.counts, a.workspace = split_src.data_ptr(), counts.data_ptr(), ws.data_ptr()
    with torch.cuda.device(dev):
        _check(lib.gft_densify_plan(C.byref(a), _stream(dev)), "gft_densify_plan")
    n_keep, n_clone, n_split, n_child, P_new = (int(x) for x in counts.tolist())   # the one host sync
    # samples of :579-581 — torch.normal(mean=0, std=get_scaling[selected].repeat(2, 1))
    sel = split_src[:n_split].long()
    sc = src["scaling"][sel]
    stds = torch.exp(sc.repeat(1, 3) if isotropic else sc).repeat(2, 1)
    if samples is not None:
        noise = _cuda_f32(samples, "samples")
        if tuple(noise.shape) != (2 * n_split, 3):
            raise RuntimeError(f"samples must be [{2 * n_split}, 3] for this plan, got {tuple(noise.shape)}")
    else:
        noise = torch.normal(mean=torch.zeros_like(stds), std=stds, generator=generator) if n_split else \
            torch.zeros((0, 3), dtype=torch.float32, device=dev)
    out_p, out_m, out_v = {}, {}, {}
    for g in params:
        t = src[g]
        width = int(t.numel() // P)
        shape = (P_new,) + tuple(t.shape[1:])
        b = GftDensifyApplyArgs()
        b.P_new, b.width = P_new, width
        b.mode = 1 if g == "xyz" else (2 if g == "scaling" else 0)
        b.plan_src, b.plan_kind, b.plan_noise = plan_src.data_ptr(), plan_kind.data_ptr(), plan_noise.data_ptr()
        b.param_in = t.data_ptr()
        out_p[g] = torch.empty(shape, dtype=torch.float32, device=dev)
        b.param_out = _ptr(out_p[g])
        if exp_avg is not None and g in exp_avg:
            mi, vi = _cuda_f32(exp_avg[g], g), _cuda_f32(exp_avg_sq[g], g)
            out_m[g], out_v[g] = torch.empty_like(out_p[g]), torch.empty_like(out_p[g])
            b.exp_avg_in, b.exp_avg_sq_in = mi.data_ptr(), vi.data_ptr()
            b.exp_avg_out, b.exp_avg_sq_out = _ptr(out_m[g]), _ptr(out_v[g])
        if b.mode == 1:
            b.noise, b.rotation_raw = _ptr(noise), src["rotation"].data_ptr()
        with torch.cuda.device(dev):
            _check(lib.gft_densify_apply(C.byref(b), _stream(dev)), "gft_densify_apply")
    info = dict(kept=n_keep, clones=n_clone, split=n_split, children_per_copy=n_child, P_new=P_new)
    return out_p, out_m, out_v, info
