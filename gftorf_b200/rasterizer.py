"""Python operator surface of the B200-native RGB+ToF Gaussian rasterizer.

Mirrors submodules/diff-gaussian-rasterization-w-tof/diff_gaussian_rasterization_w_tof/__init__.py
of the reference — same names, argument order, output order and validation errors:

  GaussianRasterizationSettings   (__init__.py:22-40)
  rasterize_gaussians             (__init__.py:42-67)
  _RasterizeGaussians             (__init__.py:69-206)   torch.autograd.Function
  GaussianRasterizer              (__init__.py:208-269)  nn.Module with forward / markVisible

Below the surface the reference calls a pybind/torch extension (`_C.rasterize_gaussians`,
rasterize_points.cu:35-165 / 167-281 / 283-304).  Here the host side allocates with torch and calls
the C-ABI library (include/gftorf.h) through ctypes with raw device pointers and the current CUDA
stream.  PyTorch is plumbing only: memory, streams, autograd bookkeeping.
"""
from typing import NamedTuple, Optional
import ctypes as C
import os
import threading

import torch
import torch.nn as nn

from . import _capi


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item
                      for item in input_tuple]
    return tuple(copied_tensors)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    near_n: Optional[float] = 0.01
    far_n: Optional[float] = 100.0
    depth_range: Optional[float] = 100.0
    use_view_dependent_phase: Optional[bool] = False
    optimize_phase_offset: Optional[bool] = False
    optimize_dc_offset: Optional[bool] = False


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def _ptr(t):
    """Device pointer, or NULL for the reference's "absent" signal (an empty tensor,
    __init__.py:239-254)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t):
    if t is None or t.numel() == 0:
        return t
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _as_float(v):
    if isinstance(v, torch.Tensor):
        return float(v.detach().cpu().numpy().item())
    return float(v)


def _prepare_bg(bg, H, W):
    """Returns (tensor, bg_mode).  The kernels index the map as bg[ch*H*W + pix] with the RENDER
    H, W (forward.cu:644,649), whatever the tensor's own spatial size is.  A
    per-channel constant map (stride-0 expand of 7 values, train.py:127) whose plane size equals
    the render's is passed as 7 constants instead of being materialised (the reference's
    `.contiguous()` writes 28*H*W bytes per call)."""
    if bg.dim() == 3 and bg.shape[0] >= 7 and bg.stride(1) == 0 and bg.stride(2) == 0 \
            and bg.shape[1] * bg.shape[2] == H * W:
        return _f32c(bg[:, 0, 0]), 1
    bgc = _f32c(bg)
    if bgc.numel() < 7 * H * W:
        raise RuntimeError(
            f"bg must hold at least 7*H*W = {7 * H * W} floats (got {bgc.numel()}); the rasterizer "
            "reads 7 background planes with the render resolution as plane stride")
    return bgc, 0


class _Workspaces:
    """Receives the three workspace requests of gft_forward (the reference's resizeFunctional,
    rasterize_points.cu:27-33).

    The three ctypes callbacks are created once per process and find the active request through a
    thread-local; a `_Workspaces` holds only its tensors.  (Per-call closures over `self` form a
    reference cycle that keeps ~150 MB of workspaces alive per forward until the cyclic collector
    runs, which makes the caching allocator fall back to cudaMalloc mid-step.)"""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        _ws_tls.cur = self

    def cb(self, name):
        return _WS_CALLBACKS[name]

    def get(self, name):
        t = self.bufs.get(name)
        if t is None:
            t = torch.empty(0, dtype=torch.uint8, device=self.device)
        return t


_ws_tls = threading.local()


def _make_ws_callback(name):
    def alloc(_ctx, nbytes):
        ws = _ws_tls.cur
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=ws.device)
        ws.bufs[name] = t
        return t.data_ptr()
    return _capi.ALLOC_FN(alloc)


_WS_CALLBACKS = {name: _make_ws_callback(name) for name in ("geom", "binning", "img")}


def _check_rc(rc, what):
    if rc < 0:
        msg = _capi.lib().gft_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


def _sh_count(t):
    return int(t.shape[1]) if (t is not None and t.numel() != 0 and t.dim() >= 2) else 0


def _native_forward(bg, means3D, colors_precomp, phasors_precomp, opacities, scales, rotations,
                    scale_modifier, cov3Ds_precomp, viewmatrix, projmatrix, tanfovx, tanfovy,
                    image_height, image_width, sh, sh_p, degree, campos, prefiltered, debug,
                    near_n, far_n, depth_range, use_view_dependent_phase, phase_offset, dc_offset,
                    R_hint=0, separate_outputs=False):
    """Same argument order and 15-tuple result as `_C.rasterize_gaussians`
    (rasterize_points.cu:35-165).

    `R_hint` (not in the reference): an estimate of num_rendered; > 0 selects the library's hinted
    mode (GftForwardArgs.R_hint) in which the host does not wait for the instance count in the
    middle of the forward.  Results are identical.
    `separate_outputs`: every image output in its own allocation (the public autograd surface —
    outputs of one autograd node that are views of one tensor cannot be modified in place, and
    holding one would pin all 21 planes); default: one allocation for all 21 planes."""
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("gftorf_b200 has no CPU path: means3D must be a CUDA tensor")
    lib = _capi.lib()
    dev = means3D.device
    P, H, W = int(means3D.shape[0]), int(image_height), int(image_width)
    N = H * W
    f32 = dict(dtype=torch.float32, device=dev)

    # one allocation for all 21 image planes; every element is written by the kernels
    if separate_outputs:
        planes = torch.empty((1,), **f32)
        (color, phasor, depth, normal, acc, entropy, depth_distortion, amp_distortion,
         distribution) = (torch.empty((c, H, W), **f32) for c in (3, 7, 1, 3, 1, 1, 1, 1, 3))
    else:
        planes = torch.empty((21, H, W), **f32)
        color, phasor = planes[0:3], planes[3:10]
        depth, normal, acc = planes[10:11], planes[11:14], planes[14:15]
        entropy, depth_distortion, amp_distortion = planes[15:16], planes[16:17], planes[17:18]
        distribution = planes[18:21]
    if P == 0:
        pixels = torch.zeros((P, 1), **f32)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    else:
        pixels = torch.empty((P, 1), **f32)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)

    means3D = _f32c(means3D)
    sh, sh_p = _f32c(sh), _f32c(sh_p)
    colors_precomp, phasors_precomp = _f32c(colors_precomp), _f32c(phasors_precomp)
    opacities, scales, rotations = _f32c(opacities), _f32c(scales), _f32c(rotations)
    cov3Ds_precomp = _f32c(cov3Ds_precomp)
    viewmatrix, projmatrix, campos = _f32c(viewmatrix), _f32c(projmatrix), _f32c(campos)
    bgc, bg_mode = _prepare_bg(bg, H, W)

    a = _capi.GftForwardArgs()
    a.P, a.sh_degree, a.M, a.M_p = P, int(degree), _sh_count(sh), _sh_count(sh_p)
    a.width, a.height = W, H
    a.background, a.bg_mode = _ptr(bgc), bg_mode
    a.means3D, a.shs, a.shs_p = _ptr(means3D), _ptr(sh), _ptr(sh_p)
    a.colors_precomp, a.phasors_precomp = _ptr(colors_precomp), _ptr(phasors_precomp)
    a.opacities, a.scales, a.scale_modifier = _ptr(opacities), _ptr(scales), float(scale_modifier)
    a.rotations, a.cov3D_precomp = _ptr(rotations), _ptr(cov3Ds_precomp)
    a.viewmatrix, a.projmatrix, a.campos = _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos)
    a.tan_fovx, a.tan_fovy = float(tanfovx), float(tanfovy)
    a.prefiltered, a.debug = int(bool(prefiltered)), int(bool(debug))
    a.near_n, a.far_n, a.depth_range = float(near_n), float(far_n), float(depth_range)
    a.use_view_dependent_phase = int(bool(use_view_dependent_phase))
    a.phase_offset, a.dc_offset = float(phase_offset), float(dc_offset)
    a.out_color, a.out_phasor, a.out_depth = color.data_ptr(), phasor.data_ptr(), depth.data_ptr()
    a.out_normal, a.out_acc, a.out_entropy = normal.data_ptr(), acc.data_ptr(), entropy.data_ptr()
    a.out_depth_distortion = depth_distortion.data_ptr()
    a.out_amp_distortion = amp_distortion.data_ptr()
    a.pixels, a.out_distribution, a.radii = _ptr(pixels), distribution.data_ptr(), _ptr(radii)
    if P == 0:  # the library still needs valid pointers to validate; nothing is written to them
        a.pixels = planes.data_ptr()
        a.radii = planes.data_ptr()
    a.R_hint = int(R_hint) if R_hint and R_hint > 0 else 0

    ws = _Workspaces(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.gft_forward(C.byref(a), ws.cb("geom"), ws.cb("binning"), ws.cb("img"), None,
                             C.c_void_p(stream))
    _ws_tls.cur = None
    _check_rc(rc, "gft_forward")
    return (rc, color, phasor, depth, normal, acc, entropy, depth_distortion, amp_distortion,
            pixels, distribution, radii, ws.get("geom"), ws.get("binning"), ws.get("img"))


def _native_backward(bg, means3D, radii, colors_precomp, phasors_precomp, scales, rotations,
                     scale_modifier, cov3Ds_precomp, viewmatrix, projmatrix, tanfovx, tanfovy,
                     grad_out_color, grad_out_phasor, grad_out_depth, grad_out_normal,
                     grad_out_acc, grad_entropy, grad_depth_distortion, grad_amp_distortion,
                     sh, sh_p, degree, campos, geomBuffer, R, binningBuffer, imgBuffer, debug,
                     near_n, far_n, depth_range, use_view_dependent_phase, phase_offset, dc_offset,
                     grad_out=None, accumulate=True, separate_means2D=False):
    """Same argument order and 12-tuple result as `_C.rasterize_gaussians_backward`
    (rasterize_points.cu:167-281).  grad_out_normal / grad_entropy / grad_amp_distortion are
    accepted and ignored, as in the reference (backward.cu never reads them).

    `grad_out` (not in the reference): dict with contiguous fp32 tensors `means3D [P,3]`,
    `shs [P,M,3]`, `shs_p [P,M_p,2]`, `opacities [P,1]`, `scales [P,3]`, `rotations [P,4]`,
    `phase_offset [1]`, `dc_offset [1]` that the parameter gradients are ADDED INTO
    (GftBackwardArgs.accumulate) — the slices of a `parallel.GradBucket`, so the views of a batch
    accumulate in the buffer the collective runs on.  `accumulate=False` makes this call OVERWRITE
    the `grad_out` tensors instead (every row is written, zeros for culled Gaussians): the first
    view of an iteration then needs no zero fill of the bucket.  `accumulate="atomic"` adds with
    atomics, so the backward calls of several views may run concurrently on different streams
    into one (zero-filled) bucket (`parallel.ViewRunner`)."""
    lib = _capi.lib()
    dev = means3D.device
    P = int(means3D.shape[0])
    H, W = int(grad_out_color.shape[1]), int(grad_out_color.shape[2])
    M, M_p = _sh_count(sh), _sh_count(sh_p)
    f32 = dict(dtype=torch.float32, device=dev)

    # One flat allocation for every gradient + the scratch records.  Nothing needs a zero fill
    # except the scratch, which the library clears itself.
    sizes = [("means3D", 3 * P), ("means2D", 3 * P), ("colors", 3 * P),
             ("opacity", P), ("cov3D", 6 * P), ("sh", 3 * M * P), ("sh_p", 2 * M_p * P),
             ("scales", 3 * P), ("rot", 4 * P), ("phase_offset", 1), ("dc_offset", 1)]
    scratch_floats = lib.gft_backward_scratch_bytes(P) // 4
    # keep every slice 16-byte aligned
    offs, cur = {}, 0
    for name, n in sizes:
        offs[name] = (cur, n)
        cur += (n + 3) // 4 * 4
    scratch_off = cur
    flat = torch.empty(cur + scratch_floats, **f32)

    def view(name, *shape):
        o, n = offs[name]
        return flat[o:o + n].view(*shape)

    if grad_out is not None:
        sizes = [(n, (0 if n in ("means3D", "opacity", "sh", "sh_p", "scales", "rot", "phase_offset",
                                 "dc_offset") else k)) for n, k in sizes]
        offs, cur = {}, 0
        for name, n in sizes:
            offs[name] = (cur, n)
            cur += (n + 3) // 4 * 4
        scratch_off = cur
        flat = torch.empty(cur + scratch_floats, **f32)
    dL_dmeans3D, dL_dmeans2D = view("means3D", P, 3) if grad_out is None else None, view("means2D", P, 3)
    if separate_means2D:
        # viewspace_points.grad outlives the call (densification statistics): its own allocation, so
        # it does not pin the other gradients and the scratch records
        dL_dmeans2D = torch.empty((P, 3), **f32)
    dL_dcolors = view("colors", P, 3)
    # grad_phasors_precomp ([P,7] for a [P,2] input in the reference) is not produced
    dL_dphasors = None
    dL_dcov3D = view("cov3D", P, 6)
    have_scales = scales is not None and scales.numel() != 0
    if grad_out is None:
        dL_dopacity = view("opacity", P, 1)
        dL_dsh, dL_dsh_p = view("sh", P, M, 3), view("sh_p", P, M_p, 2)
        dL_dscales, dL_drotations = view("scales", P, 3), view("rot", P, 4)
        dL_dphase_offset, dL_ddc_offset = view("phase_offset", 1), view("dc_offset", 1)
        if not have_scales and P > 0:
            dL_dscales.zero_()
            dL_drotations.zero_()
    else:
        dL_dmeans3D, dL_dopacity = grad_out["means3D"], grad_out["opacities"]
        dL_dsh, dL_dsh_p = grad_out["shs"], grad_out["shs_p"]
        dL_dscales, dL_drotations = grad_out["scales"], grad_out["rotations"]
        dL_dphase_offset, dL_ddc_offset = grad_out["phase_offset"], grad_out["dc_offset"]
        for t in (dL_dmeans3D, dL_dopacity, dL_dsh, dL_dsh_p, dL_dscales, dL_drotations,
                  dL_dphase_offset, dL_ddc_offset):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                raise RuntimeError("grad_out tensors must be contiguous fp32 tensors on the input's device")
        if not accumulate and not have_scales and P > 0:   # rows the library will not touch
            dL_dscales.zero_()
            dL_drotations.zero_()

    means3D = _f32c(means3D)
    sh, sh_p = _f32c(sh), _f32c(sh_p)
    colors_precomp, phasors_precomp = _f32c(colors_precomp), _f32c(phasors_precomp)
    scales, rotations, cov3Ds_precomp = _f32c(scales), _f32c(rotations), _f32c(cov3Ds_precomp)
    viewmatrix, projmatrix, campos = _f32c(viewmatrix), _f32c(projmatrix), _f32c(campos)
    bgc, bg_mode = _prepare_bg(bg, H, W)
    g_color, g_phasor = _f32c(grad_out_color), _f32c(grad_out_phasor)
    g_depth, g_acc, g_dd = _f32c(grad_out_depth), _f32c(grad_out_acc), _f32c(grad_depth_distortion)

    a = _capi.GftBackwardArgs()
    a.P, a.sh_degree, a.M, a.M_p, a.R = P, int(degree), M, M_p, int(R)
    a.width, a.height = W, H
    a.background, a.bg_mode = _ptr(bgc), bg_mode
    a.means3D, a.shs, a.shs_p = _ptr(means3D), _ptr(sh), _ptr(sh_p)
    a.colors_precomp, a.phasors_precomp = _ptr(colors_precomp), _ptr(phasors_precomp)
    a.scales, a.scale_modifier, a.rotations = _ptr(scales), float(scale_modifier), _ptr(rotations)
    a.cov3D_precomp = _ptr(cov3Ds_precomp)
    a.viewmatrix, a.projmatrix, a.campos = _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos)
    a.tan_fovx, a.tan_fovy = float(tanfovx), float(tanfovy)
    a.radii = _ptr(radii)
    a.geom_buffer, a.binning_buffer, a.img_buffer = _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer)
    a.dL_dout_color, a.dL_dout_phasor = _ptr(g_color), _ptr(g_phasor)
    a.dL_dout_depth, a.dL_dout_acc = _ptr(g_depth), _ptr(g_acc)
    a.dL_dout_depth_distortion = _ptr(g_dd)
    a.dL_dmeans2D, a.dL_dopacity, a.dL_dmeans3D = _ptr(dL_dmeans2D), _ptr(dL_dopacity), _ptr(dL_dmeans3D)
    a.dL_dsh, a.dL_dsh_p = _ptr(dL_dsh), _ptr(dL_dsh_p)
    a.dL_dscales = _ptr(dL_dscales) if have_scales else None
    a.dL_drotations = _ptr(dL_drotations) if have_scales else None
    a.dL_dphase_offset, a.dL_ddc_offset = dL_dphase_offset.data_ptr(), dL_ddc_offset.data_ptr()
    a.dL_dcolors, a.dL_dphasors, a.dL_dcov3D = _ptr(dL_dcolors), None, _ptr(dL_dcov3D)
    a.dL_dconic = a.dL_ddist = a.dL_dndc = None
    a.scratch = flat[scratch_off:].data_ptr()
    a.debug = int(bool(debug))
    a.near_n, a.far_n, a.depth_range = float(near_n), float(far_n), float(depth_range)
    a.use_view_dependent_phase = int(bool(use_view_dependent_phase))
    a.phase_offset, a.dc_offset = _as_float(phase_offset), _as_float(dc_offset)
    a.accumulate = 0 if (grad_out is None or not accumulate) else (2 if accumulate == "atomic" else 1)

    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.gft_backward(C.byref(a), C.c_void_p(stream))
    _check_rc(rc, "gft_backward")
    return (dL_dmeans2D, dL_dcolors, dL_dphasors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh,
            dL_dsh_p, dL_dscales, dL_drotations, dL_dphase_offset, dL_ddc_offset)


def _native_mark_visible(means3D, viewmatrix, projmatrix, near_n, far_n):
    """`_C.mark_visible` (rasterize_points.cu:283-304)."""
    lib = _capi.lib()
    dev = means3D.device
    P = int(means3D.shape[0])
    present = torch.zeros((P,), dtype=torch.bool, device=dev)
    if P != 0:
        means3D, viewmatrix, projmatrix = _f32c(means3D), _f32c(viewmatrix), _f32c(projmatrix)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.gft_mark_visible(P, means3D.data_ptr(), viewmatrix.data_ptr(),
                                      projmatrix.data_ptr(), present.data_ptr(), float(near_n),
                                      float(far_n), C.c_void_p(stream))
        _check_rc(rc, "gft_mark_visible")
    return present


class _NativeModule:
    """Stands in for the reference's `_C` extension module (ext.cpp:15-19)."""
    rasterize_gaussians = staticmethod(_native_forward)
    rasterize_gaussians_backward = staticmethod(_native_backward)
    mark_visible = staticmethod(_native_mark_visible)


_C = _NativeModule

# Last instance count per (device, P, H, W): the next forward of the same shape is enqueued in one go
# with a 25 % margin instead of stalling on the count (see GftForwardArgs.R_hint).  GFT_NO_HINT=1
# restores the reference's exact-size behaviour.
_R_HISTORY = {}


_NO_HINT = os.environ.get("GFT_NO_HINT") == "1"      # read once: the forward is launch-bound on small scenes


def _r_hint(key):
    if _NO_HINT:
        return 0
    last = _R_HISTORY.get(key)
    return 0 if last is None else int(last * 1.25) + 4096


# --------------------------------------------------------------------------------------------
# the reference's public surface
# --------------------------------------------------------------------------------------------
def rasterize_gaussians(means3D, means2D, sh, sh_p, colors_precomp, phasors_precomp, opacities,
                        scales, rotations, cov3Ds_precomp, phase_offset, dc_offset,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, sh_p, colors_precomp, phasors_precomp,
                                     opacities, scales, rotations, cov3Ds_precomp, phase_offset,
                                     dc_offset, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, sh_p, colors_precomp, phasors_precomp, opacities,
                scales, rotations, cov3Ds_precomp, phase_offset, dc_offset,
                raster_settings: GaussianRasterizationSettings):
        # phase/dc offsets are floats, or 1-element Parameters when optimised (__init__.py:111-112);
        # converted once here and cached for the backward (the reference syncs again in backward).
        phase_f = _as_float(phase_offset) if raster_settings.optimize_phase_offset else phase_offset
        dc_f = _as_float(dc_offset) if raster_settings.optimize_dc_offset else dc_offset
        args = (
            raster_settings.bg, means3D, colors_precomp, phasors_precomp, opacities, scales,
            rotations, raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
            raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
            raster_settings.image_height, raster_settings.image_width, sh, sh_p,
            raster_settings.sh_degree, raster_settings.campos, raster_settings.prefiltered,
            raster_settings.debug, raster_settings.near_n, raster_settings.far_n,
            raster_settings.depth_range, raster_settings.use_view_dependent_phase,
            _as_float(phase_f), _as_float(dc_f),
        )
        hint_key = (means3D.device.index, int(means3D.shape[0]), int(raster_settings.image_height),
                    int(raster_settings.image_width))
        kw = {"R_hint": _r_hint(hint_key), "separate_outputs": True} if _C is _NativeModule else {}
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy them before they can be corrupted
            try:
                out = _C.rasterize_gaussians(*args, **kw)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            out = _C.rasterize_gaussians(*args, **kw)
        _R_HISTORY[hint_key] = int(out[0])
        (num_rendered, color, phasor, depth, normal, acc, entropy, depth_distortion,
         amp_distortion, pixels, distribution, radii, geomBuffer, binningBuffer, imgBuffer) = out

        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, phasors_precomp, means3D, scales, rotations,
                              cov3Ds_precomp, radii, sh, sh_p, geomBuffer, binningBuffer, imgBuffer)
        ctx.phase_offset = _as_float(phase_f)
        ctx.dc_offset = _as_float(dc_f)
        return (color, phasor, depth, normal, acc, entropy, depth_distortion, amp_distortion,
                pixels, distribution, radii)

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_phasor, grad_out_depth, grad_out_normal,
                 grad_out_acc, grad_entropy, grad_depth_distortion, grad_amp_distortion,
                 grad_pixels, grad_distribution, _):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        (colors_precomp, phasors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
         sh_p, geomBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors

        args = (
            raster_settings.bg, means3D, radii, colors_precomp, phasors_precomp, scales, rotations,
            raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
            raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
            grad_out_color, grad_out_phasor, grad_out_depth, grad_out_normal, grad_out_acc,
            grad_entropy, grad_depth_distortion, grad_amp_distortion, sh, sh_p,
            raster_settings.sh_degree, raster_settings.campos, geomBuffer, num_rendered,
            binningBuffer, imgBuffer, raster_settings.debug, raster_settings.near_n,
            raster_settings.far_n, raster_settings.depth_range,
            raster_settings.use_view_dependent_phase, ctx.phase_offset, ctx.dc_offset,
        )
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                out = _C.rasterize_gaussians_backward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            out = _C.rasterize_gaussians_backward(*args, **({"separate_means2D": True} if _C is _NativeModule else {}))
        (grad_means2D, grad_colors_precomp, grad_phasors_precomp, grad_opacities, grad_means3D,
         grad_cov3Ds_precomp, grad_sh, grad_sh_p, grad_scales, grad_rotations, grad_phase_offset,
         grad_dc_offset) = out

        def present(t):
            return t is not None and t.numel() != 0

        # The reference returns every tensor unconditionally (__init__.py:192-204); autograd
        # drops those whose input does not require grad.  Gradients of absent (empty) inputs are
        # returned as None here, which autograd treats identically.  grad_phasors_precomp is
        # [P,7] for a [P,2] input in the reference — an unusable path there; it
        # is not offered here.
        grads = (
            grad_means3D,
            grad_means2D,
            grad_sh if present(sh) else None,
            grad_sh_p if present(sh_p) else None,
            grad_colors_precomp if present(colors_precomp) else None,
            None,
            grad_opacities,
            grad_scales if present(scales) else None,
            grad_rotations if present(rotations) else None,
            grad_cov3Ds_precomp if present(cov3Ds_precomp) else None,
            grad_phase_offset if raster_settings.optimize_phase_offset else None,
            grad_dc_offset if raster_settings.optimize_dc_offset else None,
            None,
        )
        return grads


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix,
                                      raster_settings.projmatrix, raster_settings.near_n,
                                      raster_settings.far_n)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, shs_p=None, colors_precomp=None,
                phasors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                phase_offset=0.0, dc_offset=0.0):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if shs_p is None:
            shs_p = torch.Tensor([])
        if phasors_precomp is None:
            phasors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])

        return rasterize_gaussians(means3D, means2D, shs, shs_p, colors_precomp, phasors_precomp,
                                   opacities, scales, rotations, cov3D_precomp, phase_offset,
                                   dc_offset, raster_settings)
