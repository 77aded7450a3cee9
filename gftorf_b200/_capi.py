"""ctypes binding of the C ABI declared in include/gftorf.h.

This is the only place the shared library is loaded.  There is NO fallback: if
``libgftorf_b200.so`` is missing the import of the rasterizer fails loudly (the reference fails
the same way when its ``_C`` extension is not built,
diff_gaussian_rasterization_w_tof/__init__.py:15).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgftorf_b200.so")

c_float_p = C.c_void_p  # raw device pointers travel as integers
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class GftForwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("sh_degree", C.c_int), ("M", C.c_int), ("M_p", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("background", C.c_void_p), ("bg_mode", C.c_int),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("shs_p", C.c_void_p),
        ("colors_precomp", C.c_void_p), ("phasors_precomp", C.c_void_p),
        ("opacities", C.c_void_p), ("scales", C.c_void_p), ("scale_modifier", C.c_float),
        ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("prefiltered", C.c_int), ("debug", C.c_int),
        ("near_n", C.c_float), ("far_n", C.c_float), ("depth_range", C.c_float),
        ("use_view_dependent_phase", C.c_int),
        ("phase_offset", C.c_float), ("dc_offset", C.c_float),
        ("out_color", C.c_void_p), ("out_phasor", C.c_void_p), ("out_depth", C.c_void_p),
        ("out_normal", C.c_void_p), ("out_acc", C.c_void_p), ("out_entropy", C.c_void_p),
        ("out_depth_distortion", C.c_void_p), ("out_amp_distortion", C.c_void_p),
        ("pixels", C.c_void_p), ("out_distribution", C.c_void_p), ("radii", C.c_void_p),
        ("R_hint", C.c_int),
    ]


class GftBackwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("sh_degree", C.c_int), ("M", C.c_int), ("M_p", C.c_int), ("R", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("background", C.c_void_p), ("bg_mode", C.c_int),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("shs_p", C.c_void_p),
        ("colors_precomp", C.c_void_p), ("phasors_precomp", C.c_void_p),
        ("scales", C.c_void_p), ("scale_modifier", C.c_float),
        ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("radii", C.c_void_p),
        ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p), ("img_buffer", C.c_void_p),
        ("dL_dout_color", C.c_void_p), ("dL_dout_phasor", C.c_void_p),
        ("dL_dout_depth", C.c_void_p), ("dL_dout_acc", C.c_void_p),
        ("dL_dout_depth_distortion", C.c_void_p),
        ("dL_dmeans2D", C.c_void_p), ("dL_dopacity", C.c_void_p), ("dL_dmeans3D", C.c_void_p),
        ("dL_dsh", C.c_void_p), ("dL_dsh_p", C.c_void_p),
        ("dL_dscales", C.c_void_p), ("dL_drotations", C.c_void_p),
        ("dL_dphase_offset", C.c_void_p), ("dL_ddc_offset", C.c_void_p),
        ("dL_dcolors", C.c_void_p), ("dL_dphasors", C.c_void_p), ("dL_dcov3D", C.c_void_p),
        ("dL_dconic", C.c_void_p), ("dL_ddist", C.c_void_p), ("dL_dndc", C.c_void_p),
        ("scratch", C.c_void_p),
        ("debug", C.c_int),
        ("near_n", C.c_float), ("far_n", C.c_float), ("depth_range", C.c_float),
        ("use_view_dependent_phase", C.c_int),
        ("phase_offset", C.c_float), ("dc_offset", C.c_float),
        ("accumulate", C.c_int),
    ]


class GftViewArgs(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("background", C.c_void_p), ("bg_mode", C.c_int),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("near_n", C.c_float), ("far_n", C.c_float), ("depth_range", C.c_float),
        ("use_view_dependent_phase", C.c_int),
        ("phase_offset", C.c_float), ("dc_offset", C.c_float),
        ("out_color", C.c_void_p), ("out_phasor", C.c_void_p), ("out_depth", C.c_void_p),
        ("out_normal", C.c_void_p), ("out_acc", C.c_void_p), ("out_entropy", C.c_void_p),
        ("out_depth_distortion", C.c_void_p), ("out_amp_distortion", C.c_void_p),
        ("pixels", C.c_void_p), ("out_distribution", C.c_void_p), ("radii", C.c_void_p),
        ("dL_dout_color", C.c_void_p), ("dL_dout_phasor", C.c_void_p),
        ("dL_dout_depth", C.c_void_p), ("dL_dout_acc", C.c_void_p),
        ("dL_dout_depth_distortion", C.c_void_p),
        ("dL_dmeans2D", C.c_void_p),
    ]


class GftForwardViewsArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("sh_degree", C.c_int), ("M", C.c_int), ("M_p", C.c_int), ("n_views", C.c_int),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("shs_p", C.c_void_p),
        ("colors_precomp", C.c_void_p), ("phasors_precomp", C.c_void_p),
        ("opacities", C.c_void_p), ("scales", C.c_void_p), ("scale_modifier", C.c_float),
        ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("prefiltered", C.c_int), ("debug", C.c_int),
        ("views", C.POINTER(GftViewArgs)),
        ("R_hint", C.c_int),
    ]


class GftBackwardViewsArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("sh_degree", C.c_int), ("M", C.c_int), ("M_p", C.c_int), ("R", C.c_int),
        ("n_views", C.c_int),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("shs_p", C.c_void_p),
        ("colors_precomp", C.c_void_p), ("phasors_precomp", C.c_void_p),
        ("scales", C.c_void_p), ("scale_modifier", C.c_float),
        ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p), ("img_buffer", C.c_void_p),
        ("views", C.POINTER(GftViewArgs)),
        ("dL_dopacity", C.c_void_p), ("dL_dmeans3D", C.c_void_p),
        ("dL_dsh", C.c_void_p), ("dL_dsh_p", C.c_void_p),
        ("dL_dscales", C.c_void_p), ("dL_drotations", C.c_void_p),
        ("dL_dphase_offset", C.c_void_p), ("dL_ddc_offset", C.c_void_p),
        ("dL_dcolors", C.c_void_p), ("dL_dcov3D", C.c_void_p),
        ("dL_dconic", C.c_void_p), ("dL_ddist", C.c_void_p), ("dL_dndc", C.c_void_p),
        ("scratch", C.c_void_p),
        ("debug", C.c_int), ("accumulate", C.c_int),
    ]


GFT_MAX_VIEWS = 16


class GftWorkspaceLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "geom_cov3D", "geom_rec", "geom_depths", "geom_tiles_touched", "geom_rect",
        "geom_clamped", "geom_pa", "geom_total",
        "bin_point_list", "bin_entries", "bin_total",
        "img_hdr", "img_sub_bins", "img_tile_counts", "img_ranges", "img_state", "img_total")]


# Every symbol include/gftorf.h declares; tests/test_abi.py checks the list against the header.
EXPORTS = (
    "gft_forward", "gft_geom_bytes", "gft_img_bytes", "gft_binning_bytes",
    "gft_backward_scratch_bytes", "gft_backward", "gft_mark_visible",
    "gft_dist2_workspace_bytes", "gft_dist2", "gft_workspace_layout", "gft_last_error",
    "gft_abi_version", "gft_profile_enable", "gft_profile_read", "gft_launch_count",
    "gft_forward_views", "gft_backward_views", "gft_geom_bytes_views", "gft_img_bytes_views",
    "gft_backward_scratch_bytes_views", "gft_workspace_layout_views", "gft_set_option",
)


def declare(lib, prefix="gft_"):
    """Attach argtypes/restypes for the entry points (also used for the reference shim, which
    exports the same signatures under the ``ref_`` prefix — tests and bench only)."""
    f = getattr(lib, prefix + "forward")
    f.argtypes = [C.POINTER(GftForwardArgs), ALLOC_FN, ALLOC_FN, ALLOC_FN, C.c_void_p, C.c_void_p]
    f.restype = C.c_int
    b = getattr(lib, prefix + "backward")
    b.argtypes = [C.POINTER(GftBackwardArgs), C.c_void_p]
    b.restype = C.c_int
    m = getattr(lib, prefix + "mark_visible")
    m.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                  C.c_void_p]
    m.restype = C.c_int
    d = getattr(lib, prefix + "dist2")
    d.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    d.restype = C.c_int
    e = getattr(lib, prefix + "last_error")
    e.argtypes = []
    e.restype = C.c_char_p
    if prefix == "gft_":
        for name in ("gft_geom_bytes", "gft_binning_bytes", "gft_backward_scratch_bytes",
                     "gft_dist2_workspace_bytes"):
            fn = getattr(lib, name)
            fn.argtypes = [C.c_int]
            fn.restype = C.c_size_t
        lib.gft_img_bytes.argtypes = [C.c_int, C.c_int]
        lib.gft_img_bytes.restype = C.c_size_t
        lib.gft_workspace_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.POINTER(GftWorkspaceLayout)]
        lib.gft_workspace_layout.restype = None
        lib.gft_forward_views.argtypes = [C.POINTER(GftForwardViewsArgs), ALLOC_FN, ALLOC_FN, ALLOC_FN,
                                          C.c_void_p, C.c_void_p]
        lib.gft_forward_views.restype = C.c_int
        lib.gft_backward_views.argtypes = [C.POINTER(GftBackwardViewsArgs), C.c_void_p]
        lib.gft_backward_views.restype = C.c_int
        lib.gft_geom_bytes_views.argtypes = [C.c_int, C.c_int]
        lib.gft_geom_bytes_views.restype = C.c_size_t
        lib.gft_img_bytes_views.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.gft_img_bytes_views.restype = C.c_size_t
        lib.gft_backward_scratch_bytes_views.argtypes = [C.c_int, C.c_int]
        lib.gft_backward_scratch_bytes_views.restype = C.c_size_t
        lib.gft_workspace_layout_views.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                                   C.POINTER(C.c_int), C.POINTER(GftWorkspaceLayout)]
        lib.gft_workspace_layout_views.restype = None
        lib.gft_set_option.argtypes = [C.c_char_p, C.c_int]
        lib.gft_set_option.restype = C.c_int
        lib.gft_abi_version.argtypes = []
        lib.gft_abi_version.restype = C.c_int
        lib.gft_profile_enable.argtypes = [C.c_int]
        lib.gft_profile_enable.restype = None
        lib.gft_profile_read.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.c_int]
        lib.gft_profile_read.restype = C.c_int
        lib.gft_launch_count.argtypes = []
        lib.gft_launch_count.restype = C.c_ulonglong
    return lib


ABI_VERSION = 2
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C gftorf_b200/csrc`). There is no CPU or PyTorch fallback.")
        _lib = declare(C.CDLL(LIB_PATH))
        if _lib.gft_abi_version() != ABI_VERSION:
            raise ImportError("libgftorf_b200.so ABI version mismatch")
    return _lib


def profile_enable(on=True):
    lib().gft_profile_enable(1 if on else 0)


def profile_read(cap=64):
    """[(stage name, milliseconds)] of the calls made since the last read, in launch order."""
    ms = (C.c_float * cap)()
    names = (C.c_char_p * cap)()
    n = lib().gft_profile_read(ms, names, cap)
    return [(names[i].decode(), float(ms[i])) for i in range(n)]


def launch_count():
    return int(lib().gft_launch_count())


def set_option(name, value):
    """Flip a kernel tunable / A-B switch (include/gftorf.h: gft_set_option); returns the old value."""
    old = lib().gft_set_option(name.encode(), int(value))
    if old < 0 and name not in ("sort_cap", "sort_radix", "sub_bins", "sort_match", "tile_order", "bwd_ring", "pfwd_minb", "blend_half", "sort_adapt", "bwd_pred", "pbwd_minb", "no_cull"):
        raise KeyError(name)
    return old
