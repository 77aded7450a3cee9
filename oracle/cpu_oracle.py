"""TEST INFRASTRUCTURE ONLY — Python door onto oracle/libgft_oracle.so (gft_oracle.cpp), the CPU
restatement of the reference rasterizer and distCUDA2.  Same call shapes as the reference's `_C`
module (ext.cpp:15-19) so the parity tests drive product, reference kernels and this oracle with
one harness.  Works on CPU torch tensors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs import this.
"""
import ctypes as C
import os

import numpy as np
import torch

from gftorf_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgft_oracle.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{LIB_PATH} not built (make -C oracle cpu)")
        L = C.CDLL(LIB_PATH)
        L.orc_forward.argtypes = [C.POINTER(_capi.GftForwardArgs), C.POINTER(C.c_void_p)]
        L.orc_forward.restype = C.c_int
        L.orc_backward.argtypes = [C.POINTER(_capi.GftBackwardArgs), C.c_void_p]
        L.orc_backward.restype = C.c_int
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_free.restype = None
        L.orc_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.orc_get.restype = C.c_int
        L.orc_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.orc_mark_visible.restype = C.c_int
        L.orc_dist2.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_dist2.restype = C.c_int
        L.orc_last_error.restype = C.c_char_p
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


class Handle:
    """Owns the oracle's saved forward state (stands where geomBuffer/binningBuffer/imgBuffer
    stand in the reference)."""

    def __init__(self, ptr):
        self.ptr = ptr

    def numel(self):
        return 1

    def __del__(self):
        try:
            if self.ptr and _lib is not None:
                _lib.orc_free(self.ptr)
        except Exception:
            pass
        self.ptr = None


_DTYPES = {
    "depths": np.float32, "ndc": np.float32, "dists": np.float32, "means2D": np.float32,
    "cov3D": np.float32, "conic_opacity": np.float32, "rgb": np.float32, "real_img_amp": np.float32,
    "pa": np.float32, "clamped": np.uint8, "clamped_p": np.uint8, "tiles_touched": np.int32,
    "point_offsets": np.int32, "keys": np.int64, "point_list": np.int32, "ranges": np.int32,
    "final_T": np.float32, "w_z_total": np.float32, "w_z2_total": np.float32, "n_contrib": np.int32,
}
_SHAPES = {"means2D": 2, "cov3D": 6, "conic_opacity": 4, "rgb": 3, "real_img_amp": 7, "pa": 2,
           "clamped": 3, "ranges": 2}


def decode(handle):
    """Saved state as torch tensors, same names as gftorf_b200.debug.decode_buffers."""
    L = lib()
    out = {}
    for name, dt in _DTYPES.items():
        p, n = C.c_void_p(), C.c_size_t()
        if L.orc_get(handle.ptr, name.encode(), C.byref(p), C.byref(n)) != 0:
            continue
        if n.value == 0:
            continue
        nbytes = n.value * np.dtype(dt).itemsize
        buf = (C.c_char * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt).copy()
        if name in _SHAPES:
            arr = arr.reshape(-1, _SHAPES[name])
        out[name] = torch.from_numpy(arr)
    return out


def _p(t):
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _c(t):
    if t is None or t.numel() == 0:
        return t
    return t.detach().float().contiguous()


def _sh_count(t):
    return int(t.shape[1]) if (t is not None and t.numel() != 0 and t.dim() >= 2) else 0


def rasterize_gaussians(bg, means3D, colors_precomp, phasors_precomp, opacities, scales, rotations,
                        scale_modifier, cov3Ds_precomp, viewmatrix, projmatrix, tanfovx, tanfovy,
                        image_height, image_width, sh, sh_p, degree, campos, prefiltered, debug,
                        near_n, far_n, depth_range, use_view_dependent_phase, phase_offset,
                        dc_offset):
    L = lib()
    P, H, W = int(means3D.shape[0]), int(image_height), int(image_width)
    z = lambda *s: torch.zeros(s, dtype=torch.float32)
    out_color, out_phasor, out_depth, out_normal = z(3, H, W), z(7, H, W), z(1, H, W), z(3, H, W)
    out_acc, out_entropy, out_dd, out_ad = z(1, H, W), z(1, H, W), z(1, H, W), z(1, H, W)
    out_distribution, pixels = z(3, H, W), z(P, 1)
    radii = torch.zeros((P,), dtype=torch.int32)
    bg, means3D = _c(bg), _c(means3D)
    colors_precomp, phasors_precomp = _c(colors_precomp), _c(phasors_precomp)
    opacities, scales, rotations = _c(opacities), _c(scales), _c(rotations)
    cov3Ds_precomp, sh, sh_p = _c(cov3Ds_precomp), _c(sh), _c(sh_p)
    viewmatrix, projmatrix, campos = _c(viewmatrix), _c(projmatrix), _c(campos)
    a = _capi.GftForwardArgs()
    a.P, a.sh_degree, a.M, a.M_p = P, int(degree), _sh_count(sh), _sh_count(sh_p)
    a.width, a.height = W, H
    a.background, a.bg_mode = _p(bg), 0
    a.means3D, a.shs, a.shs_p = _p(means3D), _p(sh), _p(sh_p)
    a.colors_precomp, a.phasors_precomp = _p(colors_precomp), _p(phasors_precomp)
    a.opacities, a.scales, a.scale_modifier = _p(opacities), _p(scales), float(scale_modifier)
    a.rotations, a.cov3D_precomp = _p(rotations), _p(cov3Ds_precomp)
    a.viewmatrix, a.projmatrix, a.campos = _p(viewmatrix), _p(projmatrix), _p(campos)
    a.tan_fovx, a.tan_fovy = float(tanfovx), float(tanfovy)
    a.prefiltered, a.debug = int(bool(prefiltered)), int(bool(debug))
    a.near_n, a.far_n, a.depth_range = float(near_n), float(far_n), float(depth_range)
    a.use_view_dependent_phase = int(bool(use_view_dependent_phase))
    a.phase_offset, a.dc_offset = float(phase_offset), float(dc_offset)
    a.out_color, a.out_phasor, a.out_depth = out_color.data_ptr(), out_phasor.data_ptr(), out_depth.data_ptr()
    a.out_normal, a.out_acc, a.out_entropy = out_normal.data_ptr(), out_acc.data_ptr(), out_entropy.data_ptr()
    a.out_depth_distortion, a.out_amp_distortion = out_dd.data_ptr(), out_ad.data_ptr()
    a.pixels, a.out_distribution, a.radii = _p(pixels), out_distribution.data_ptr(), _p(radii)
    h = C.c_void_p()
    rc = L.orc_forward(C.byref(a), C.byref(h))
    handle = Handle(h.value)
    if rc < 0:
        raise RuntimeError("orc_forward failed: " + L.orc_last_error().decode())
    return (rc, out_color, out_phasor, out_depth, out_normal, out_acc, out_entropy, out_dd, out_ad,
            pixels, out_distribution, radii, handle, None, None)


def rasterize_gaussians_backward(bg, means3D, radii, colors_precomp, phasors_precomp, scales,
                                 rotations, scale_modifier, cov3Ds_precomp, viewmatrix, projmatrix,
                                 tanfovx, tanfovy, grad_out_color, grad_out_phasor, grad_out_depth,
                                 grad_out_normal, grad_out_acc, grad_entropy,
                                 grad_depth_distortion, grad_amp_distortion, sh, sh_p, degree,
                                 campos, geomBuffer, R, binningBuffer, imgBuffer, debug, near_n,
                                 far_n, depth_range, use_view_dependent_phase, phase_offset,
                                 dc_offset, return_internal=False):
    L = lib()
    P = int(means3D.shape[0])
    H, W = int(grad_out_color.shape[1]), int(grad_out_color.shape[2])
    M, M_p = _sh_count(sh), _sh_count(sh_p)
    z = lambda *s: torch.zeros(s, dtype=torch.float32)
    dL_dmeans3D, dL_dmeans2D = z(P, 3), z(P, 3)
    dL_dcolors, dL_dphasors, dL_ddist, dL_dndc = z(P, 3), z(P, 7), z(P, 1), z(P, 1)
    dL_dconic, dL_dopacity, dL_dcov3D = z(P, 2, 2), z(P, 1), z(P, 6)
    dL_dsh, dL_dsh_p = z(P, M, 3), z(P, M_p, 2)
    dL_dscales, dL_drotations = z(P, 3), z(P, 4)
    dL_dphase_offset, dL_ddc_offset = z(1), z(1)
    bg, means3D = _c(bg), _c(means3D)
    colors_precomp, phasors_precomp = _c(colors_precomp), _c(phasors_precomp)
    scales, rotations, cov3Ds_precomp = _c(scales), _c(rotations), _c(cov3Ds_precomp)
    sh, sh_p = _c(sh), _c(sh_p)
    viewmatrix, projmatrix, campos = _c(viewmatrix), _c(projmatrix), _c(campos)
    g_color, g_phasor = _c(grad_out_color), _c(grad_out_phasor)
    g_depth, g_acc, g_dd = _c(grad_out_depth), _c(grad_out_acc), _c(grad_depth_distortion)
    radii = radii.contiguous()
    a = _capi.GftBackwardArgs()
    a.P, a.sh_degree, a.M, a.M_p, a.R = P, int(degree), M, M_p, int(R)
    a.width, a.height = W, H
    a.background, a.bg_mode = _p(bg), 0
    a.means3D, a.shs, a.shs_p = _p(means3D), _p(sh), _p(sh_p)
    a.colors_precomp, a.phasors_precomp = _p(colors_precomp), _p(phasors_precomp)
    a.scales, a.scale_modifier, a.rotations = _p(scales), float(scale_modifier), _p(rotations)
    a.cov3D_precomp = _p(cov3Ds_precomp)
    a.viewmatrix, a.projmatrix, a.campos = _p(viewmatrix), _p(projmatrix), _p(campos)
    a.tan_fovx, a.tan_fovy = float(tanfovx), float(tanfovy)
    a.radii = _p(radii)
    a.dL_dout_color, a.dL_dout_phasor = _p(g_color), _p(g_phasor)
    a.dL_dout_depth, a.dL_dout_acc, a.dL_dout_depth_distortion = _p(g_depth), _p(g_acc), _p(g_dd)
    a.dL_dmeans2D, a.dL_dopacity, a.dL_dmeans3D = _p(dL_dmeans2D), _p(dL_dopacity), _p(dL_dmeans3D)
    a.dL_dsh, a.dL_dsh_p = _p(dL_dsh), _p(dL_dsh_p)
    have_scales = scales is not None and scales.numel() != 0
    a.dL_dscales = _p(dL_dscales) if have_scales else None
    a.dL_drotations = _p(dL_drotations) if have_scales else None
    a.dL_dphase_offset, a.dL_ddc_offset = dL_dphase_offset.data_ptr(), dL_ddc_offset.data_ptr()
    a.dL_dcolors, a.dL_dphasors, a.dL_dcov3D = _p(dL_dcolors), _p(dL_dphasors), _p(dL_dcov3D)
    a.dL_dconic, a.dL_ddist, a.dL_dndc = _p(dL_dconic), _p(dL_ddist), _p(dL_dndc)
    a.debug = int(bool(debug))
    a.near_n, a.far_n, a.depth_range = float(near_n), float(far_n), float(depth_range)
    a.use_view_dependent_phase = int(bool(use_view_dependent_phase))
    a.phase_offset, a.dc_offset = float(phase_offset), float(dc_offset)
    rc = L.orc_backward(C.byref(a), geomBuffer.ptr)
    if rc < 0:
        raise RuntimeError("orc_backward failed: " + L.orc_last_error().decode())
    out = (dL_dmeans2D, dL_dcolors, dL_dphasors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh,
           dL_dsh_p, dL_dscales, dL_drotations, dL_dphase_offset, dL_ddc_offset)
    if return_internal:
        return out, dict(dL_dconic=dL_dconic, dL_ddist=dL_ddist, dL_dndc=dL_dndc)
    return out


def mark_visible(means3D, viewmatrix, projmatrix, near_n, far_n):
    P = int(means3D.shape[0])
    present = torch.zeros((P,), dtype=torch.bool)
    if P:
        m, v = _c(means3D), _c(viewmatrix)
        lib().orc_mark_visible(P, m.data_ptr(), v.data_ptr(), present.data_ptr(), float(near_n), float(far_n))
    return present


def distCUDA2(points):
    P = int(points.shape[0])
    out = torch.zeros((P,), dtype=torch.float32)
    if P:
        pts = _c(points)
        lib().orc_dist2(pts.data_ptr(), P, out.data_ptr())
    return out


class OracleModule:
    rasterize_gaussians = staticmethod(rasterize_gaussians)
    rasterize_gaussians_backward = staticmethod(rasterize_gaussians_backward)
    mark_visible = staticmethod(mark_visible)


def rasterize_autograd(asm, view, means2D=None):
    """The reference's autograd Function over the CPU port (oracle/autograd_wrap.py)."""
    from . import autograd_wrap
    global _AUTOGRAD
    try:
        fn = _AUTOGRAD
    except NameError:
        fn = _AUTOGRAD = autograd_wrap.make(OracleModule)
    return fn(asm, means2D, view)
