"""TEST INFRASTRUCTURE ONLY — drives oracle/_ref/libgftorf_ref.so (the UNMODIFIED reference
kernels behind oracle/ref_shim.cu) from PyTorch the way the reference's own torch binding does.

What the binding does and this file restates (no arithmetic here):
  - every output is a fresh zero-filled tensor: 11 `torch::full/zeros` per forward
    (rasterize_points.cu:80-92), 15 `torch::zeros` per backward (:222-236);
  - every argument goes through `.contiguous()` (:129-150) — a stride-0 expanded background is
    materialised on every call;
  - the three opaque workspaces are torch byte tensors resized by callback (:27-33,94-101);
  - kernels run on the legacy default stream (no stream argument anywhere in the reference).

Only tests/, __graft_entry__.smoke() and bench.py's reference arm import this module.
"""
import ctypes as C
import os

import torch

from gftorf_b200 import _capi
from gftorf_b200.rasterizer import _ptr, _sh_count, _Workspaces, _as_float

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libgftorf_ref.so")


class RefLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "geom_depths", "geom_ndc", "geom_clamped", "geom_clamped_p", "geom_radii", "geom_means2D",
        "geom_cov3D", "geom_conic_opacity", "geom_rgb", "geom_real_img_amp", "geom_dists",
        "geom_pa", "geom_tiles_touched", "geom_point_offsets", "geom_total",
        "img_accum_alpha", "img_w_z_total", "img_w_z2_total", "img_n_contrib", "img_ranges",
        "img_total",
        "bin_point_list", "bin_point_list_unsorted", "bin_keys", "bin_keys_unsorted", "bin_total")]


_lib = None


def available():
    return os.path.exists(REF_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{REF_LIB_PATH} not built (make -C oracle ref)")
        _lib = _capi.declare(C.CDLL(REF_LIB_PATH), prefix="ref_")
        _lib.ref_layout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.POINTER(RefLayout)]
        _lib.ref_layout.restype = None
    return _lib


def _c(t):
    if t is None or t.numel() == 0:
        return t
    return t.contiguous()


def _check(rc, what):
    if rc < 0:
        raise RuntimeError(f"{what} failed: " + lib().ref_last_error().decode("utf-8", "replace"))


def rasterize_gaussians(bg, means3D, colors_precomp, phasors_precomp, opacities, scales, rotations,
                        scale_modifier, cov3Ds_precomp, viewmatrix, projmatrix, tanfovx, tanfovy,
                        image_height, image_width, sh, sh_p, degree, campos, prefiltered, debug,
                        near_n, far_n, depth_range, use_view_dependent_phase, phase_offset,
                        dc_offset):
    """RasterizeGaussiansCUDA, rasterize_points.cu:35-165 (same arguments, same 15-tuple)."""
    L = lib()
    dev = means3D.device
    P, H, W = int(means3D.shape[0]), int(image_height), int(image_width)
    f32 = dict(dtype=torch.float32, device=dev)
    out_color = torch.full((3, H, W), 0.0, **f32)
    out_phasor = torch.full((7, H, W), 0.0, **f32)
    radii = torch.full((P,), 0, dtype=torch.int32, device=dev)
    pixels = torch.zeros((P, 1), **f32)
    out_depth = torch.full((1, H, W), 0.0, **f32)
    out_normal = torch.full((3, H, W), 0.0, **f32)
    out_acc = torch.full((1, H, W), 0.0, **f32)
    out_entropy = torch.full((1, H, W), 0.0, **f32)
    out_dd = torch.full((1, H, W), 0.0, **f32)
    out_ad = torch.full((1, H, W), 0.0, **f32)
    out_distribution = torch.full((3, H, W), 0.0, **f32)
    ws = _Workspaces(dev)
    rendered = 0
    if P != 0:
        bg, means3D = _c(bg), _c(means3D)
        colors_precomp, phasors_precomp = _c(colors_precomp), _c(phasors_precomp)
        opacities, scales, rotations = _c(opacities), _c(scales), _c(rotations)
        cov3Ds_precomp, sh, sh_p = _c(cov3Ds_precomp), _c(sh), _c(sh_p)
        viewmatrix, projmatrix, campos = _c(viewmatrix), _c(projmatrix), _c(campos)
        a = _capi.GftForwardArgs()
        a.P, a.sh_degree, a.M, a.M_p = P, int(degree), _sh_count(sh), _sh_count(sh_p)
        a.width, a.height = W, H
        a.background, a.bg_mode = _ptr(bg), 0
        a.means3D, a.shs, a.shs_p = _ptr(means3D), _ptr(sh), _ptr(sh_p)
        a.colors_precomp, a.phasors_precomp = _ptr(colors_precomp), _ptr(phasors_precomp)
        a.opacities, a.scales, a.scale_modifier = _ptr(opacities), _ptr(scales), float(scale_modifier)
        a.rotations, a.cov3D_precomp = _ptr(rotations), _ptr(cov3Ds_precomp)
        a.viewmatrix, a.projmatrix, a.campos = _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos)
        a.tan_fovx, a.tan_fovy = float(tanfovx), float(tanfovy)
        a.prefiltered, a.debug = int(bool(prefiltered)), int(bool(debug))
        a.near_n, a.far_n, a.depth_range = float(near_n), float(far_n), float(depth_range)
        a.use_view_dependent_phase = int(bool(use_view_dependent_phase))
        a.phase_offset, a.dc_offset = _as_float(phase_offset), _as_float(dc_offset)
        a.out_color, a.out_phasor, a.out_depth = out_color.data_ptr(), out_phasor.data_ptr(), out_depth.data_ptr()
        a.out_normal, a.out_acc, a.out_entropy = out_normal.data_ptr(), out_acc.data_ptr(), out_entropy.data_ptr()
        a.out_depth_distortion, a.out_amp_distortion = out_dd.data_ptr(), out_ad.data_ptr()
        a.pixels, a.out_distribution, a.radii = pixels.data_ptr(), out_distribution.data_ptr(), radii.data_ptr()
        with torch.cuda.device(dev):
            rendered = L.ref_forward(C.byref(a), ws.cb("geom"), ws.cb("binning"), ws.cb("img"),
                                     None, None)
        _check(rendered, "ref_forward")
    return (rendered, out_color, out_phasor, out_depth, out_normal, out_acc, out_entropy, out_dd,
            out_ad, pixels, out_distribution, radii, ws.get("geom"), ws.get("binning"),
            ws.get("img"))


def rasterize_gaussians_backward(bg, means3D, radii, colors_precomp, phasors_precomp, scales,
                                 rotations, scale_modifier, cov3Ds_precomp, viewmatrix, projmatrix,
                                 tanfovx, tanfovy, grad_out_color, grad_out_phasor, grad_out_depth,
                                 grad_out_normal, grad_out_acc, grad_entropy,
                                 grad_depth_distortion, grad_amp_distortion, sh, sh_p, degree,
                                 campos, geomBuffer, R, binningBuffer, imgBuffer, debug, near_n,
                                 far_n, depth_range, use_view_dependent_phase, phase_offset,
                                 dc_offset, return_internal=False):
    """RasterizeGaussiansBackwardCUDA, rasterize_points.cu:167-281 (same 12-tuple)."""
    L = lib()
    dev = means3D.device
    P = int(means3D.shape[0])
    H, W = int(grad_out_color.shape[1]), int(grad_out_color.shape[2])
    M, M_p = _sh_count(sh), _sh_count(sh_p)
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    dL_dmeans3D, dL_dmeans2D = z(P, 3), z(P, 3)
    dL_dcolors, dL_dphasors = z(P, 3), z(P, 7)
    dL_ddist, dL_dndc = z(P, 1), z(P, 1)
    dL_dconic, dL_dopacity, dL_dcov3D = z(P, 2, 2), z(P, 1), z(P, 6)
    dL_dsh, dL_dsh_p = z(P, M, 3), z(P, M_p, 2)
    dL_dscales, dL_drotations = z(P, 3), z(P, 4)
    dL_dphase_offset, dL_ddc_offset = z(1), z(1)
    if P != 0:
        bg, means3D = _c(bg), _c(means3D)
        colors_precomp, phasors_precomp = _c(colors_precomp), _c(phasors_precomp)
        scales, rotations, cov3Ds_precomp = _c(scales), _c(rotations), _c(cov3Ds_precomp)
        sh, sh_p = _c(sh), _c(sh_p)
        viewmatrix, projmatrix, campos = _c(viewmatrix), _c(projmatrix), _c(campos)
        g_color, g_phasor = _c(grad_out_color), _c(grad_out_phasor)
        g_depth, g_acc, g_dd = _c(grad_out_depth), _c(grad_out_acc), _c(grad_depth_distortion)
        a = _capi.GftBackwardArgs()
        a.P, a.sh_degree, a.M, a.M_p, a.R = P, int(degree), M, M_p, int(R)
        a.width, a.height = W, H
        a.background, a.bg_mode = _ptr(bg), 0
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
        a.dL_dmeans2D, a.dL_dopacity, a.dL_dmeans3D = dL_dmeans2D.data_ptr(), dL_dopacity.data_ptr(), dL_dmeans3D.data_ptr()
        a.dL_dsh, a.dL_dsh_p = _ptr(dL_dsh), _ptr(dL_dsh_p)
        a.dL_dscales, a.dL_drotations = dL_dscales.data_ptr(), dL_drotations.data_ptr()
        a.dL_dphase_offset, a.dL_ddc_offset = dL_dphase_offset.data_ptr(), dL_ddc_offset.data_ptr()
        a.dL_dcolors, a.dL_dphasors, a.dL_dcov3D = dL_dcolors.data_ptr(), dL_dphasors.data_ptr(), dL_dcov3D.data_ptr()
        a.dL_dconic, a.dL_ddist, a.dL_dndc = dL_dconic.data_ptr(), dL_ddist.data_ptr(), dL_dndc.data_ptr()
        a.scratch = None
        a.debug = int(bool(debug))
        a.near_n, a.far_n, a.depth_range = float(near_n), float(far_n), float(depth_range)
        a.use_view_dependent_phase = int(bool(use_view_dependent_phase))
        a.phase_offset, a.dc_offset = _as_float(phase_offset), _as_float(dc_offset)
        with torch.cuda.device(dev):
            rc = L.ref_backward(C.byref(a), None)
        _check(rc, "ref_backward")
    out = (dL_dmeans2D, dL_dcolors, dL_dphasors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh,
           dL_dsh_p, dL_dscales, dL_drotations, dL_dphase_offset, dL_ddc_offset)
    if return_internal:
        return out, dict(dL_dconic=dL_dconic, dL_ddist=dL_ddist, dL_dndc=dL_dndc)
    return out


def mark_visible(means3D, viewmatrix, projmatrix, near_n, far_n):
    L = lib()
    P = int(means3D.shape[0])
    present = torch.full((P,), False, dtype=torch.bool, device=means3D.device)
    if P != 0:
        m, v, p = means3D.contiguous(), viewmatrix.contiguous(), projmatrix.contiguous()
        with torch.cuda.device(means3D.device):
            rc = L.ref_mark_visible(P, m.data_ptr(), v.data_ptr(), p.data_ptr(), present.data_ptr(),
                                    float(near_n), float(far_n), None)
        _check(rc, "ref_mark_visible")
    return present


def distCUDA2(points):
    """spatial.cu:15-26"""
    L = lib()
    P = int(points.shape[0])
    means = torch.full((P,), 0.0, dtype=torch.float32, device=points.device)
    pts = points.contiguous()
    with torch.cuda.device(points.device):
        rc = L.ref_dist2(pts.data_ptr(), P, means.data_ptr(), None, None)
    _check(rc, "ref_dist2")
    return means


class RefModule:
    """Same three callables as the reference's `_C` module (ext.cpp:15-19)."""
    rasterize_gaussians = staticmethod(rasterize_gaussians)
    rasterize_gaussians_backward = staticmethod(rasterize_gaussians_backward)
    mark_visible = staticmethod(mark_visible)


_AUTOGRAD = None


def rasterize_autograd(asm, means2D, view):
    """The reference's autograd Function over the reference kernels (oracle/autograd_wrap.py)."""
    from . import autograd_wrap
    global _AUTOGRAD
    if _AUTOGRAD is None:
        _AUTOGRAD = autograd_wrap.make(RefModule)
    return _AUTOGRAD(asm, means2D, view)


def decode_buffers(geom, binning, img, P, R, W, H):
    """Typed views into the reference's opaque buffers, offsets from its own fromChunk walk
    (ref_layout in ref_shim.cu)."""
    L = lib()
    lay = RefLayout()
    L.ref_layout(_ptr(geom), _ptr(img), _ptr(binning), P, R, W, H, C.byref(lay))
    N = W * H
    T = ((W + 15) // 16) * ((H + 15) // 16)

    def v(buf, off, nbytes, dtype):
        return buf[off:off + nbytes].view(dtype)

    out = {}
    if P > 0:
        out.update(
            depths=v(geom, lay.geom_depths, 4 * P, torch.float32),
            ndc=v(geom, lay.geom_ndc, 4 * P, torch.float32),
            clamped=v(geom, lay.geom_clamped, 3 * P, torch.uint8).view(P, 3),
            clamped_p=v(geom, lay.geom_clamped_p, P, torch.uint8),
            means2D=v(geom, lay.geom_means2D, 8 * P, torch.float32).view(P, 2),
            cov3D=v(geom, lay.geom_cov3D, 24 * P, torch.float32).view(P, 6),
            conic_opacity=v(geom, lay.geom_conic_opacity, 16 * P, torch.float32).view(P, 4),
            rgb=v(geom, lay.geom_rgb, 12 * P, torch.float32).view(P, 3),
            real_img_amp=v(geom, lay.geom_real_img_amp, 28 * P, torch.float32).view(P, 7),
            dists=v(geom, lay.geom_dists, 4 * P, torch.float32),
            pa=v(geom, lay.geom_pa, 8 * P, torch.float32).view(P, 2),
            tiles_touched=v(geom, lay.geom_tiles_touched, 4 * P, torch.int32),
            point_offsets=v(geom, lay.geom_point_offsets, 4 * P, torch.int32),
        )
    out.update(
        final_T=v(img, lay.img_accum_alpha, 4 * N, torch.float32),
        w_z_total=v(img, lay.img_w_z_total, 4 * N, torch.float32),
        w_z2_total=v(img, lay.img_w_z2_total, 4 * N, torch.float32),
        n_contrib=v(img, lay.img_n_contrib, 4 * N, torch.int32),
        ranges=v(img, lay.img_ranges, 8 * T, torch.int32).view(T, 2),
    )
    if R > 0:
        out.update(
            point_list=v(binning, lay.bin_point_list, 4 * R, torch.int32),
            keys=v(binning, lay.bin_keys, 8 * R, torch.int64),
        )
    return out
