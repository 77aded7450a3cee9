"""Host-side mirror of the reference's Python surface: names, validation errors, background
handling, drop-in import paths, and the no-fallback rule.  Runs without a GPU."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import gftorf_b200  # noqa: E402
from gftorf_b200 import rasterizer  # noqa: E402


def _settings(**kw):
    base = dict(image_height=32, image_width=32, tanfovx=0.5, tanfovy=0.5, bg=torch.zeros(7, 32, 32),
                scale_modifier=1.0, viewmatrix=torch.eye(4), projmatrix=torch.eye(4), sh_degree=0,
                campos=torch.zeros(3), prefiltered=False, debug=False)
    base.update(kw)
    return rasterizer.GaussianRasterizationSettings(**base)


def test_settings_fields_and_defaults_match_reference():
    # diff_gaussian_rasterization_w_tof/__init__.py:22-40
    s = _settings()
    assert s._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                         "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug",
                         "near_n", "far_n", "depth_range", "use_view_dependent_phase",
                         "optimize_phase_offset", "optimize_dc_offset")
    assert (s.near_n, s.far_n, s.depth_range) == (0.01, 100.0, 100.0)
    assert (s.use_view_dependent_phase, s.optimize_phase_offset, s.optimize_dc_offset) == (False,) * 3


def test_validation_errors_match_reference_messages():
    # __init__.py:230-237
    r = rasterizer.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(m, m, torch.zeros(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=torch.zeros(4, 3),
          scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3),
          rotations=torch.ones(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_no_cpu_fallback():
    r = rasterizer.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3),
          rotations=torch.ones(4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        gftorf_b200.distCUDA2(torch.zeros(4, 3))


def test_means3d_shape_error_matches_reference():
    # rasterize_points.cu:69-71
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        rasterizer._native_forward(torch.zeros(7, 8, 8), torch.zeros(4, 2), *([torch.Tensor([])] * 2),
                                   torch.zeros(4, 1), torch.ones(4, 3), torch.ones(4, 4), 1.0,
                                   torch.Tensor([]), torch.eye(4), torch.eye(4), 0.5, 0.5, 8, 8,
                                   torch.zeros(4, 1, 3), torch.Tensor([]), 0, torch.zeros(3), False,
                                   False, 0.01, 100.0, 100.0, False, 0.0, 0.0)


def test_background_modes():
    H, W = 6, 5
    const = torch.arange(7.0).view(7, 1, 1).expand(7, H, W)          # train.py:127
    t, mode = rasterizer._prepare_bg(const, H, W)
    assert mode == 1 and t.shape == (7,) and torch.equal(t, torch.arange(7.0))
    full = torch.rand(7, H, W)
    t, mode = rasterizer._prepare_bg(full, H, W)
    assert mode == 0 and t.data_ptr() == full.data_ptr()
    # colour-sized constant map reused for a smaller ToF view: the kernel's plane stride is the
    # render's H*W, so the map has to be materialised to keep the wrapped indexing
    big_const = torch.arange(7.0).view(7, 1, 1).expand(7, 2 * H, 2 * W)
    t, mode = rasterizer._prepare_bg(big_const, H, W)
    assert mode == 0 and t.is_contiguous() and t.numel() == 7 * 4 * H * W
    with pytest.raises(RuntimeError, match="at least 7\\*H\\*W"):
        rasterizer._prepare_bg(torch.zeros(3, H, W), H, W)


def test_dropin_import_paths():
    # gaussian_renderer/__init__.py:14 and scene/gaussian_model.py:20
    sys.path.insert(0, os.path.join(ROOT, "gftorf_b200", "dropin"))
    try:
        from diff_gaussian_rasterization_w_tof import GaussianRasterizationSettings, GaussianRasterizer
        from simple_knn._C import distCUDA2
        assert GaussianRasterizer is rasterizer.GaussianRasterizer
        assert GaussianRasterizationSettings is rasterizer.GaussianRasterizationSettings
        assert distCUDA2 is gftorf_b200.distCUDA2
    finally:
        sys.path.pop(0)


def test_autograd_function_signature():
    import inspect
    fwd = inspect.signature(rasterizer._RasterizeGaussians.forward)
    assert list(fwd.parameters)[1:] == ["means3D", "means2D", "sh", "sh_p", "colors_precomp",
                                        "phasors_precomp", "opacities", "scales", "rotations",
                                        "cov3Ds_precomp", "phase_offset", "dc_offset", "raster_settings"]
    gr = inspect.signature(rasterizer.GaussianRasterizer.forward)
    assert list(gr.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "shs_p",
                                       "colors_precomp", "phasors_precomp", "scales", "rotations",
                                       "cov3D_precomp", "phase_offset", "dc_offset"]


# ------------------------------------------------------------------------------------------------
# batched views (gftorf_b200.views): host-side contract, no GPU
# ------------------------------------------------------------------------------------------------
def test_viewspec_from_settings_and_batch_limits():
    from gftorf_b200 import views as V
    s = rasterizer.GaussianRasterizationSettings(
        image_height=48, image_width=64, tanfovx=0.5, tanfovy=0.4, bg=torch.zeros(7, 48, 64), scale_modifier=1.0,
        viewmatrix=torch.eye(4), projmatrix=torch.eye(4), sh_degree=3, campos=torch.zeros(3), prefiltered=False,
        debug=False, near_n=0.2, far_n=9.0, depth_range=15.0, use_view_dependent_phase=True)
    v = V.ViewSpec.from_settings(s, phase_offset=torch.tensor([0.25]), dc_offset=0.5)
    assert (v.image_height, v.image_width, v.near_n, v.far_n, v.depth_range) == (48, 64, 0.2, 9.0, 15.0)
    assert v.use_view_dependent_phase is True and v.phase_offset == 0.25 and v.dc_offset == 0.5
    assert V.MAX_VIEWS == 16
    m = torch.zeros(5, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        V.forward_views(m, m[:, :1], m, torch.zeros(5, 4), torch.zeros(5, 16, 3), torch.zeros(5, 16, 2), [v], 3)
    with pytest.raises(RuntimeError, match="num_points, 3"):
        V.forward_views(torch.zeros(5, 2), m[:, :1], m, m, m, m, [v], 3)


def test_rasterize_views_validation_matches_the_single_view_surface():
    from gftorf_b200 import views as V
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        V.rasterize_views(m, m, m[:, :1], None, None, m, m, [], 3)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        V.rasterize_views(m, m, m[:, :1], torch.zeros(4, 16, 3), None, None, None, [], 3)


def test_c_abi_rejects_bad_view_batches_without_a_gpu():
    import ctypes as C
    from gftorf_b200 import _capi
    lib = _capi.lib()
    a = _capi.GftForwardViewsArgs()
    a.P, a.n_views = 10, 0
    cb = _capi.ALLOC_FN(lambda ctx, n: 0)
    assert lib.gft_forward_views(C.byref(a), cb, cb, cb, None, None) == -1
    assert b"n_views" in lib.gft_last_error()
    a.n_views = _capi.GFT_MAX_VIEWS + 1
    assert lib.gft_forward_views(C.byref(a), cb, cb, cb, None, None) == -1
    b = _capi.GftBackwardViewsArgs()
    b.P, b.n_views = 10, 0
    assert lib.gft_backward_views(C.byref(b), None) == -1
    assert lib.gft_set_option(b"no_such_option", 1) < 0
    old = lib.gft_set_option(b"sort_cap", 2048)
    assert lib.gft_set_option(b"sort_cap", old) == 2048


def test_sh_row_staging_piece_map_covers_every_piece_once_without_bank_conflicts():
    """Mirror of `stage_piece<ROW>` (gftorf_b200/csrc/common.cuh): in iteration `it` lane l moves the
    16-byte piece at (row r, first column c).  Over all iterations every piece of the 32 x ROW chunk is
    moved exactly once; within one warp instruction the global addresses fill whole 32-byte sectors
    and each of the four scalar shared-memory accesses (row stride ROW + 1) hits 32 distinct banks."""
    def piece(row, it, lane):
        if row == 32 or it < 8:
            return 4 * it + (lane & 3), 4 * (lane >> 2)
        return (lane & 16) + 4 * (it - 8) + (lane & 3), 32 + 4 * ((lane >> 2) & 3)

    for row in (48, 32):
        seen = set()
        for it in range(row // 4):
            pieces = [piece(row, it, lane) for lane in range(32)]
            for r, c in pieces:
                assert 0 <= r < 32 and 0 <= c <= row - 4 and c % 4 == 0
                assert (r, c) not in seen
                seen.add((r, c))
            for j in range(4):                                     # the four scalar accesses of a piece
                banks = {(r * (row + 1) + c + j) % 32 for r, c in pieces}
                assert len(banks) == 32, (row, it, j)
            sectors = {}
            for r, c in pieces:                                    # 32-byte sectors of global memory
                sectors.setdefault((r * row + c) * 4 // 32, 0)
                sectors[(r * row + c) * 4 // 32] += 16
            assert all(v == 32 for v in sectors.values()), (row, it)
        assert len(seen) == 32 * row // 4
