"""Shared parity harness: runs the product library and a checker (the reference kernels in
oracle/_ref, or the CPU restatement) on the same seeded synthetic inputs and compares them.

Tolerances are the ones BASELINE.json's north_star states:
  integer work (radii, tiles_touched, sort keys, point_list, tile ranges, n_contrib, pixels): bit-exact
  forward images: <= 1e-5 absolute
  gradients: <= 1e-4 relative L2 (atomic ordering differs)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gftorf_b200 import scenes  # noqa: E402

IMG_ATOL = 1e-5
GRAD_REL_L2 = 1e-4


def build_inputs(P, W, H, kind="trained", seed=0, pose="identity", device="cuda", sh_degree=3,
                 depth_range=15.0, bg_hw=None, view_dependent_phase=False, phase_offset=0.0,
                 dc_offset=0.0, sigma_px=1.5, zero_shp_rest=False):
    cam = scenes.make_camera(W, H, depth_range=depth_range, pose=pose, seed=seed)
    cloud = scenes.make_cloud(P, cam, kind=kind, seed=seed, sigma_px=sigma_px)
    if zero_shp_rest:
        # keeps the reference's undefined dL_dPA (defect D1) out of dL_dmeans3D
        cloud["shs_p"][:, 1:, :] = 0.0
    bh, bw = bg_hw if bg_hw else (H, W)
    bg = scenes.make_background(bh, bw, seed=seed)
    grads = scenes.make_pixel_grads(H, W, seed=seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return dict(
        P=P, W=W, H=H, sh_degree=sh_degree,
        means3D=t(cloud["means3D"]), scales=t(cloud["scales"]), rotations=t(cloud["rotations"]),
        opacities=t(cloud["opacities"]), shs=t(cloud["shs"]), shs_p=t(cloud["shs_p"]),
        viewmatrix=t(cam["viewmatrix"]), projmatrix=t(cam["projmatrix"]), campos=t(cam["campos"]),
        tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], near_n=cam["znear"], far_n=cam["zfar"],
        depth_range=cam["depth_range"], bg=t(bg),
        grads={k: t(v) for k, v in grads.items()},
        use_view_dependent_phase=view_dependent_phase, phase_offset=phase_offset,
        dc_offset=dc_offset, empty=torch.Tensor([]),
    )


def call_forward(mod, inp, colors_precomp=None, use_shs_p=True):
    e = inp["empty"]
    sh = e if colors_precomp is not None else inp["shs"]
    cp = colors_precomp if colors_precomp is not None else e
    return mod.rasterize_gaussians(
        inp["bg"], inp["means3D"], cp, e, inp["opacities"], inp["scales"], inp["rotations"], 1.0, e,
        inp["viewmatrix"], inp["projmatrix"], inp["tanfovx"], inp["tanfovy"], inp["H"], inp["W"],
        sh, inp["shs_p"] if use_shs_p else e, inp["sh_degree"], inp["campos"], False, False,
        inp["near_n"], inp["far_n"], inp["depth_range"], inp["use_view_dependent_phase"],
        inp["phase_offset"], inp["dc_offset"])


def call_backward(mod, inp, fwd, colors_precomp=None, use_shs_p=True, **kw):
    e = inp["empty"]
    g = inp["grads"]
    sh = e if colors_precomp is not None else inp["shs"]
    cp = colors_precomp if colors_precomp is not None else e
    R, radii, geom, binning, img = fwd[0], fwd[11], fwd[12], fwd[13], fwd[14]
    zero1 = torch.zeros_like(g["depth"])
    return mod.rasterize_gaussians_backward(
        inp["bg"], inp["means3D"], radii, cp, e, inp["scales"], inp["rotations"], 1.0, e,
        inp["viewmatrix"], inp["projmatrix"], inp["tanfovx"], inp["tanfovy"],
        g["color"], g["phasor"], g["depth"], torch.zeros_like(g["color"]), g["acc"], zero1,
        g["depth_distortion"], zero1, sh, inp["shs_p"] if use_shs_p else e, inp["sh_degree"],
        inp["campos"], geom, R, binning, img, False, inp["near_n"], inp["far_n"],
        inp["depth_range"], inp["use_view_dependent_phase"], inp["phase_offset"], inp["dc_offset"],
        **kw)


FWD_NAMES = ("num_rendered", "color", "phasor", "depth", "normal", "acc", "entropy",
             "depth_distortion", "amp_distortion", "pixels", "distribution", "radii")
BWD_NAMES = ("means2D", "colors_precomp", "phasors_precomp", "opacities", "means3D",
             "cov3Ds_precomp", "sh", "sh_p", "scales", "rotations", "phase_offset", "dc_offset")


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    den = float(torch.linalg.norm(b))
    num = float(torch.linalg.norm(a - b))
    return num / den if den > 0 else num


def nmismatch(a, b):
    return int((a != b).sum().item())


def compare_forward(ours, ref, ours_dec=None, ref_dec=None):
    """Returns {name: value}; integer entries are mismatch counts, float entries max abs diff."""
    rep = {"R_ours": int(ours[0]), "R_ref": int(ref[0])}
    rep["radii"] = nmismatch(ours[11], ref[11])
    rep["pixels"] = nmismatch(ours[9], ref[9])
    for i in (1, 2, 3, 4, 5, 6, 7, 8, 10):
        rep[FWD_NAMES[i]] = float((ours[i] - ref[i]).abs().max().item()) if ours[i].numel() else 0.0
    if ours_dec is not None and ref_dec is not None:
        vis = ref[11] > 0
        rep["V"] = int(vis.sum().item())
        if "tiles_touched" in ours_dec:
            rep["tiles_touched"] = nmismatch(ours_dec["tiles_touched"], ref_dec["tiles_touched"])
            rep["point_offsets"] = nmismatch(ours_dec["point_offsets"], ref_dec["point_offsets"])
            rep["depth_bits"] = nmismatch(ours_dec["depths"].view(torch.int32)[vis],
                                          ref_dec["depths"].view(torch.int32)[vis])
            for k in ("means2D", "conic_opacity", "cov3D"):
                rep[k + "_bits"] = nmismatch(ours_dec[k].contiguous().view(torch.int32)[vis],
                                             ref_dec[k].contiguous().view(torch.int32)[vis])
            for k in ("rgb", "real_img_amp", "dists", "ndc", "pa"):
                a, b = ours_dec[k][vis], ref_dec[k][vis]
                rep[k] = float((a - b).abs().max().item()) if a.numel() else 0.0
            rep["clamped"] = nmismatch(ours_dec["clamped"][vis] != 0, ref_dec["clamped"][vis] != 0)
            rep["clamped_p"] = nmismatch(ours_dec["clamped_p"][vis] != 0, ref_dec["clamped_p"][vis] != 0)
        if "keys" in ours_dec and "keys" in ref_dec and ours[0] == ref[0]:
            rep["keys"] = nmismatch(ours_dec["keys"], ref_dec["keys"])
            rep["point_list"] = nmismatch(ours_dec["point_list"], ref_dec["point_list"])
        rep["ranges"] = nmismatch(ours_dec["ranges"], ref_dec["ranges"])
        rep["n_contrib"] = nmismatch(ours_dec["n_contrib"], ref_dec["n_contrib"])
        for k in ("final_T", "w_z_total", "w_z2_total"):
            rep[k] = float((ours_dec[k] - ref_dec[k]).abs().max().item())
    return rep


def compare_backward(ours, ref, skip=("colors_precomp", "phasors_precomp", "cov3Ds_precomp")):
    rep = {}
    for i, name in enumerate(BWD_NAMES):
        if name in skip or ours[i] is None or ref[i] is None:
            continue
        if ours[i].numel() == 0 and ref[i].numel() == 0:
            continue
        rep[name] = rel_l2(ours[i], ref[i])
    return rep


INT_KEYS = ("radii", "pixels", "tiles_touched", "point_offsets", "depth_bits", "means2D_bits",
            "conic_opacity_bits", "cov3D_bits", "clamped", "clamped_p", "keys", "point_list",
            "ranges", "n_contrib")
IMG_KEYS = ("color", "phasor", "depth", "normal", "acc", "entropy", "depth_distortion",
            "amp_distortion", "distribution")


def assert_forward_parity(rep):
    assert rep["R_ours"] == rep["R_ref"], rep
    for k in INT_KEYS:
        if k in rep:
            assert rep[k] == 0, (k, rep)
    for k in IMG_KEYS:
        assert rep[k] <= IMG_ATOL, (k, rep)


def assert_backward_parity(rep, tol=GRAD_REL_L2):
    for k, v in rep.items():
        assert v <= tol, (k, rep)
