"""Batched views: ONE rasterizer call for several cameras of the same Gaussians.

The reference renders the colour and the ToF camera of an iteration with two rasterizer calls back
to back (gaussian_renderer/__init__.py:107-128), one view per call (train.py:158), one frame per
call in the render-only loop (render.py:95-209).  All of these are independent given the Gaussian
parameters, so the library offers them as one call (include/gftorf.h: gft_forward_views /
gft_backward_views): every kernel covers all views in one launch, the 364 B/Gaussian of parameters
are read once per pass, and the backward writes each parameter-gradient row once — the sum over
the views, which is what autograd's AccumulateGrad hands the optimiser after the reference's
separate calls (train.py:279).  Per view the outputs are those of `GaussianRasterizer` (same 11
tensors, same order); a batch of one IS the reference's call.

    views = [ViewSpec.from_settings(colour_settings), ViewSpec.from_settings(tof_settings)]
    outs = rasterize_views(means3D, means2D, opacities, shs, shs_p, scales, rotations, views, sh_degree)
    color, phasor, depth, normal, acc, entropy, dd, amp_dd, pixels, distribution, radii = outs[0]

`forward_views` / `backward_views` are the same without autograd (bench, render-only sweeps,
gradient buckets).  PyTorch is plumbing only: memory, streams, autograd bookkeeping.
"""
from typing import NamedTuple, Optional, Sequence
import ctypes as C

import torch

from . import _capi
from .rasterizer import (_f32c, _ptr, _prepare_bg, _Workspaces, _ws_tls, _check_rc, _sh_count,
                         _as_float, _r_hint, _R_HISTORY)

MAX_VIEWS = _capi.GFT_MAX_VIEWS


class ViewSpec(NamedTuple):
    """The per-camera fields of GaussianRasterizationSettings (__init__.py:22-40) plus the two
    scalar ToF offsets that the reference passes per call.  `phase_offset` / `dc_offset` = None
    means "the phase_offset / dc_offset argument of rasterize_views" (the optimised Parameters of
    gaussian_renderer/__init__.py:126-127, which the reference hands to the ToF call only)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    campos: torch.Tensor
    near_n: float = 0.01
    far_n: float = 100.0
    depth_range: float = 100.0
    use_view_dependent_phase: bool = False
    phase_offset: Optional[float] = 0.0
    dc_offset: Optional[float] = 0.0

    @classmethod
    def from_settings(cls, s, phase_offset=0.0, dc_offset=0.0):
        return cls(int(s.image_height), int(s.image_width), float(s.tanfovx), float(s.tanfovy), s.bg,
                   s.viewmatrix, s.projmatrix, s.campos, float(s.near_n), float(s.far_n),
                   float(s.depth_range), bool(s.use_view_dependent_phase),
                   None if phase_offset is None else _as_float(phase_offset),
                   None if dc_offset is None else _as_float(dc_offset))


class ViewsForward:
    """Result of forward_views: per-view outputs + the three opaque workspaces for the backward."""
    __slots__ = ("R", "outs", "planes", "pixels", "radii", "geom", "binning", "img", "views",
                 "inputs", "sh_degree", "scale_modifier", "debug", "_keep")

    def view(self, i):
        return self.outs[i]


def _fill_view(va, v, keep):
    H, W = int(v.image_height), int(v.image_width)
    bgc, bg_mode = _prepare_bg(v.bg, H, W)
    vm, pm, cp = _f32c(v.viewmatrix), _f32c(v.projmatrix), _f32c(v.campos)
    keep.extend((bgc, vm, pm, cp))
    va.width, va.height = W, H
    va.background, va.bg_mode = _ptr(bgc), bg_mode
    va.viewmatrix, va.projmatrix, va.campos = _ptr(vm), _ptr(pm), _ptr(cp)
    va.tan_fovx, va.tan_fovy = float(v.tanfovx), float(v.tanfovy)
    va.near_n, va.far_n, va.depth_range = float(v.near_n), float(v.far_n), float(v.depth_range)
    va.use_view_dependent_phase = int(bool(v.use_view_dependent_phase))
    va.phase_offset = 0.0 if v.phase_offset is None else _as_float(v.phase_offset)
    va.dc_offset = 0.0 if v.dc_offset is None else _as_float(v.dc_offset)


def forward_views(means3D, opacities, scales, rotations, shs, shs_p, views: Sequence[ViewSpec],
                  sh_degree, scale_modifier=1.0, colors_precomp=None, phasors_precomp=None,
                  cov3Ds_precomp=None, prefiltered=False, debug=False, R_hint=0,
                  separate_outputs=False) -> ViewsForward:
    """All views of `views` for one set of Gaussians, in one library call (gft_forward_views).
    `separate_outputs`: give every image output its own allocation (the autograd surface: callers
    may then modify any output in place, and keeping one output alive does not pin the others);
    default is one allocation for the 21 planes of all views."""
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("gftorf_b200 has no CPU path: means3D must be a CUDA tensor")
    V = len(views)
    if not 1 <= V <= MAX_VIEWS:
        raise RuntimeError(f"a batch holds 1..{MAX_VIEWS} views (got {V})")
    lib = _capi.lib()
    dev = means3D.device
    P = int(means3D.shape[0])
    f32 = dict(dtype=torch.float32, device=dev)

    # one allocation for the 21 image planes of every view; every element is written by the kernels
    sizes = [21 * int(v.image_height) * int(v.image_width) for v in views]
    flat = torch.empty(1 if separate_outputs else sum(sizes), **f32)
    planes, off = [], 0
    for v, n in zip(views, sizes):
        if separate_outputs:
            planes.append(None)
            continue
        planes.append(flat[off:off + n].view(21, int(v.image_height), int(v.image_width)))
        off += n
    if P == 0:
        pixels = torch.zeros((V, P, 1), **f32)
        radii = torch.zeros((V, P), dtype=torch.int32, device=dev)
    else:
        pixels = torch.empty((V, P, 1), **f32)
        radii = torch.empty((V, P), dtype=torch.int32, device=dev)

    means3D = _f32c(means3D)
    shs, shs_p = _f32c(shs), _f32c(shs_p)
    colors_precomp, phasors_precomp = _f32c(colors_precomp), _f32c(phasors_precomp)
    opacities, scales, rotations = _f32c(opacities), _f32c(scales), _f32c(rotations)
    cov3Ds_precomp = _f32c(cov3Ds_precomp)

    keep = []
    va = (_capi.GftViewArgs * V)()
    outs = []
    for i, v in enumerate(views):
        _fill_view(va[i], v, keep)
        if separate_outputs:
            hw = (int(v.image_height), int(v.image_width))
            color, phasor, depth, normal, acc, entropy, dd, amp_dd, distribution = (
                torch.empty((c,) + hw, **f32) for c in (3, 7, 1, 3, 1, 1, 1, 1, 3))
        else:
            pl = planes[i]
            color, phasor, depth, normal, acc = pl[0:3], pl[3:10], pl[10:11], pl[11:14], pl[14:15]
            entropy, dd, amp_dd, distribution = pl[15:16], pl[16:17], pl[17:18], pl[18:21]
        va[i].out_color, va[i].out_phasor, va[i].out_depth = color.data_ptr(), phasor.data_ptr(), depth.data_ptr()
        va[i].out_normal, va[i].out_acc, va[i].out_entropy = normal.data_ptr(), acc.data_ptr(), entropy.data_ptr()
        va[i].out_depth_distortion, va[i].out_amp_distortion = dd.data_ptr(), amp_dd.data_ptr()
        va[i].out_distribution = distribution.data_ptr()
        # the library needs valid pointers to validate even when P == 0; nothing is written then
        va[i].pixels = pixels[i].data_ptr() if P else flat.data_ptr()
        va[i].radii = radii[i].data_ptr() if P else flat.data_ptr()
        outs.append((color, phasor, depth, normal, acc, entropy, dd, amp_dd, pixels[i], distribution,
                     radii[i]))

    a = _capi.GftForwardViewsArgs()
    a.P, a.sh_degree, a.M, a.M_p, a.n_views = P, int(sh_degree), _sh_count(shs), _sh_count(shs_p), V
    a.means3D, a.shs, a.shs_p = _ptr(means3D), _ptr(shs), _ptr(shs_p)
    a.colors_precomp, a.phasors_precomp = _ptr(colors_precomp), _ptr(phasors_precomp)
    a.opacities, a.scales, a.scale_modifier = _ptr(opacities), _ptr(scales), float(scale_modifier)
    a.rotations, a.cov3D_precomp = _ptr(rotations), _ptr(cov3Ds_precomp)
    a.prefiltered, a.debug = int(bool(prefiltered)), int(bool(debug))
    a.views = va
    a.R_hint = int(R_hint) if R_hint and R_hint > 0 else 0

    ws = _Workspaces(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.gft_forward_views(C.byref(a), ws.cb("geom"), ws.cb("binning"), ws.cb("img"), None,
                                   C.c_void_p(stream))
    _ws_tls.cur = None
    _check_rc(rc, "gft_forward_views")

    f = ViewsForward()
    f.R, f.outs, f.planes, f.pixels, f.radii = rc, outs, planes, pixels, radii
    f.geom, f.binning, f.img = ws.get("geom"), ws.get("binning"), ws.get("img")
    f.views = list(views)
    f.inputs = dict(means3D=means3D, shs=shs, shs_p=shs_p, colors_precomp=colors_precomp,
                    phasors_precomp=phasors_precomp, scales=scales, rotations=rotations,
                    cov3Ds_precomp=cov3Ds_precomp)
    f.sh_degree, f.scale_modifier, f.debug = int(sh_degree), float(scale_modifier), bool(debug)
    f._keep = keep
    return f


GRAD_KEYS = ("color", "phasor", "depth", "acc", "depth_distortion")


def backward_views(fwd: ViewsForward, grads: Sequence[dict], grad_out: Optional[dict] = None,
                   accumulate=False, want_intermediates=False) -> dict:
    """Backward of forward_views.  `grads[i]` holds the pixel gradients of view i under the keys
    color [3,H,W], phasor [7,H,W], depth, acc, depth_distortion [1,H,W] (missing / None = zero).

    Returns a dict with the parameter gradients SUMMED over the views (`means3D`, `opacities`,
    `shs`, `shs_p`, `scales`, `rotations`, `phase_offset`, `dc_offset`, plus `colors_precomp` /
    `cov3Ds_precomp` when those inputs were given) and `means2D` [V,P,3], the per-view screen-space
    gradients whose norms feed densification (scene/gaussian_model.py:650).

    `grad_out` / `accumulate`: as in rasterizer._native_backward — write into (accumulate=False),
    add into (True) or atomically add into ("atomic") the slices of a parallel.GradBucket."""
    lib = _capi.lib()
    inp = fwd.inputs
    means3D = inp["means3D"]
    dev = means3D.device
    P, V = int(means3D.shape[0]), len(fwd.views)
    shs, shs_p = inp["shs"], inp["shs_p"]
    M, M_p = _sh_count(shs), _sh_count(shs_p)
    f32 = dict(dtype=torch.float32, device=dev)
    have_scales = inp["scales"] is not None and inp["scales"].numel() != 0
    have_cp = inp["colors_precomp"] is not None and inp["colors_precomp"].numel() != 0
    have_cov = inp["cov3Ds_precomp"] is not None and inp["cov3Ds_precomp"].numel() != 0

    # one flat allocation: means2D per view, the parameter gradients (unless they go to grad_out),
    # the optional precomp gradients, and the blend-gradient records (cleared by the library)
    sizes = [("means2D", 3 * P * V)]
    if grad_out is None:
        sizes += [("means3D", 3 * P), ("opacities", P), ("shs", 3 * M * P), ("shs_p", 2 * M_p * P),
                  ("scales", 3 * P), ("rotations", 4 * P), ("phase_offset", 1), ("dc_offset", 1)]
    if have_cp or want_intermediates:
        sizes.append(("colors_precomp", 3 * P))
    if have_cov or want_intermediates:
        sizes.append(("cov3Ds_precomp", 6 * P))
    offs, cur = {}, 0
    for name, n in sizes:
        offs[name] = (cur, n)
        cur += (n + 3) // 4 * 4          # keep every slice 16-byte aligned
    scratch_off = cur
    scratch_floats = lib.gft_backward_scratch_bytes_views(P, V) // 4
    flat = torch.empty(cur + scratch_floats, **f32)

    def view(name, *shape):
        o, n = offs[name]
        return flat[o:o + n].view(*shape)

    out = {"means2D": view("means2D", V, P, 3)}
    if grad_out is None:
        out.update(means3D=view("means3D", P, 3), opacities=view("opacities", P, 1),
                   shs=view("shs", P, M, 3), shs_p=view("shs_p", P, M_p, 2),
                   scales=view("scales", P, 3), rotations=view("rotations", P, 4),
                   phase_offset=view("phase_offset", 1), dc_offset=view("dc_offset", 1))
        if not have_scales and P > 0:
            out["scales"].zero_()
            out["rotations"].zero_()
    else:
        for k in ("means3D", "opacities", "shs", "shs_p", "scales", "rotations", "phase_offset", "dc_offset"):
            t = grad_out[k]
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                raise RuntimeError("grad_out tensors must be contiguous fp32 tensors on the input's device")
            out[k] = t
        if not accumulate and not have_scales and P > 0:   # rows the library will not touch
            out["scales"].zero_()
            out["rotations"].zero_()
    if "colors_precomp" in offs:
        out["colors_precomp"] = view("colors_precomp", P, 3)
    if "cov3Ds_precomp" in offs:
        out["cov3Ds_precomp"] = view("cov3Ds_precomp", P, 6)

    keep = []
    va = (_capi.GftViewArgs * V)()
    zeros = {}
    for i, v in enumerate(fwd.views):
        _fill_view(va[i], v, keep)
        H, W = int(v.image_height), int(v.image_width)
        g = grads[i] or {}
        ptrs = []
        for k, ch in zip(GRAD_KEYS, (3, 7, 1, 1, 1)):
            t = g.get(k)
            if t is None:
                t = zeros.get((ch, H, W))
                if t is None:
                    t = zeros[(ch, H, W)] = torch.zeros((ch, H, W), **f32)
            t = _f32c(t)
            keep.append(t)
            ptrs.append(t.data_ptr())
        (va[i].dL_dout_color, va[i].dL_dout_phasor, va[i].dL_dout_depth, va[i].dL_dout_acc,
         va[i].dL_dout_depth_distortion) = ptrs
        va[i].radii = fwd.radii[i].data_ptr() if P else flat.data_ptr()
        va[i].dL_dmeans2D = out["means2D"][i].data_ptr() if P else flat.data_ptr()

    a = _capi.GftBackwardViewsArgs()
    a.P, a.sh_degree, a.M, a.M_p, a.R, a.n_views = P, fwd.sh_degree, M, M_p, int(fwd.R), V
    a.means3D, a.shs, a.shs_p = _ptr(means3D), _ptr(shs), _ptr(shs_p)
    a.colors_precomp, a.phasors_precomp = _ptr(inp["colors_precomp"]), _ptr(inp["phasors_precomp"])
    a.scales, a.scale_modifier, a.rotations = _ptr(inp["scales"]), fwd.scale_modifier, _ptr(inp["rotations"])
    a.cov3D_precomp = _ptr(inp["cov3Ds_precomp"])
    a.geom_buffer, a.binning_buffer, a.img_buffer = _ptr(fwd.geom), _ptr(fwd.binning), _ptr(fwd.img)
    a.views = va
    a.dL_dopacity, a.dL_dmeans3D = _ptr(out["opacities"]), _ptr(out["means3D"])
    a.dL_dsh, a.dL_dsh_p = _ptr(out["shs"]), _ptr(out["shs_p"])
    a.dL_dscales = _ptr(out["scales"]) if have_scales else None
    a.dL_drotations = _ptr(out["rotations"]) if have_scales else None
    a.dL_dphase_offset, a.dL_ddc_offset = out["phase_offset"].data_ptr(), out["dc_offset"].data_ptr()
    a.dL_dcolors = _ptr(out.get("colors_precomp"))
    a.dL_dcov3D = _ptr(out.get("cov3Ds_precomp"))
    a.dL_dconic = a.dL_ddist = a.dL_dndc = None
    a.scratch = flat[scratch_off:].data_ptr()
    a.debug = int(fwd.debug)
    a.accumulate = 0 if (grad_out is None or not accumulate) else (2 if accumulate == "atomic" else 1)

    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.gft_backward_views(C.byref(a), C.c_void_p(stream))
    _check_rc(rc, "gft_backward_views")
    return out


# ------------------------------------------------------------------------------------------------
# autograd surface
# ------------------------------------------------------------------------------------------------
class _RasterizeViews(torch.autograd.Function):
    """Same contract per view as _RasterizeGaussians (__init__.py:69-206); the parameter gradients
    are the sums over the views, `means2D` is [V,P,3] and receives the per-view screen-space
    gradients."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, sh_p, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, phase_offset, dc_offset, views, sh_degree, scale_modifier,
                optimize_phase_offset, optimize_dc_offset, debug):
        phase_f, dc_f = _as_float(phase_offset), _as_float(dc_offset)
        specs = [v._replace(phase_offset=phase_f if v.phase_offset is None else v.phase_offset,
                            dc_offset=dc_f if v.dc_offset is None else v.dc_offset) for v in views]
        key = (means3D.device.index, int(means3D.shape[0]),
               tuple((int(v.image_height), int(v.image_width)) for v in specs))
        fwd = forward_views(means3D, opacities, scales, rotations, sh, sh_p, specs, sh_degree,
                            scale_modifier=scale_modifier, colors_precomp=colors_precomp,
                            cov3Ds_precomp=cov3Ds_precomp, debug=debug, R_hint=_r_hint(key),
                            separate_outputs=True)
        _R_HISTORY[key] = int(fwd.R)
        ctx.fwd = fwd
        ctx.flags = (bool(optimize_phase_offset), bool(optimize_dc_offset))
        flat_out = []
        for i, o in enumerate(fwd.outs):
            # pixels / radii are rows of one [V,P] allocation: returned as independent tensors too
            flat_out.extend(o[:8] + (fwd.pixels[i].clone(), o[9], fwd.radii[i].clone()))
        fwd.outs = None      # the backward needs the workspaces and fwd.radii only
        ctx.mark_non_differentiable(*[t for j, t in enumerate(flat_out) if j % 11 in (8, 10)])
        return tuple(flat_out)

    @staticmethod
    def backward(ctx, *gouts):
        fwd = ctx.fwd
        grads = []
        for i in range(len(fwd.views)):
            g = gouts[11 * i: 11 * i + 11]
            grads.append(dict(color=g[0], phasor=g[1], depth=g[2], acc=g[4], depth_distortion=g[6]))
        out = backward_views(fwd, grads)
        inp = fwd.inputs

        def present(t):
            return t is not None and t.numel() != 0
        return (out["means3D"], out["means2D"],
                out["shs"] if present(inp["shs"]) else None,
                out["shs_p"] if present(inp["shs_p"]) else None,
                out.get("colors_precomp") if present(inp["colors_precomp"]) else None,
                out["opacities"],
                out["scales"] if present(inp["scales"]) else None,
                out["rotations"] if present(inp["rotations"]) else None,
                out.get("cov3Ds_precomp") if present(inp["cov3Ds_precomp"]) else None,
                out["phase_offset"] if ctx.flags[0] else None,
                out["dc_offset"] if ctx.flags[1] else None,
                None, None, None, None, None, None)


def rasterize_views(means3D, means2D, opacities, shs, shs_p, scales, rotations,
                    views: Sequence[ViewSpec], sh_degree, scale_modifier=1.0, colors_precomp=None,
                    cov3D_precomp=None, phase_offset=0.0, dc_offset=0.0,
                    optimize_phase_offset=False, optimize_dc_offset=False, debug=False):
    """Differentiable batched call.  `means2D`: [V,P,3] dummy leaf (zeros) whose .grad receives the
    per-view screen-space gradients, the batched form of the reference's `screenspace_points`
    (gaussian_renderer/__init__.py:27-31).  `phase_offset` / `dc_offset` (floats, or 1-element
    Parameters with optimize_*_offset) apply to the views whose ViewSpec leaves them None; their
    gradient is the sum over all views.  Returns a list of V tuples of the reference's 11
    outputs."""
    if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
    e = torch.Tensor([])
    flat = _RasterizeViews.apply(
        means3D, means2D, e if shs is None else shs, e if shs_p is None else shs_p,
        e if colors_precomp is None else colors_precomp, opacities,
        e if scales is None else scales, e if rotations is None else rotations,
        e if cov3D_precomp is None else cov3D_precomp, phase_offset, dc_offset, list(views),
        int(sh_degree), float(scale_modifier), optimize_phase_offset, optimize_dc_offset, debug)
    return [tuple(flat[11 * i: 11 * i + 11]) for i in range(len(views))]
