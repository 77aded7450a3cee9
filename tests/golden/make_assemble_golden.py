"""Generates tests/golden/train_assemble.npz by running the reference's OWN render()
(gaussian_renderer/__init__.py:19-139) on CPU in this container with the model getters of
scene/gaussian_model.py — everything up to the two rasterizer calls is the reference's code; the
rasterizer package is replaced by a recorder whose outputs are linear in its inputs, so that
autograd through the reference's assembly yields the raw-parameter gradients for a known
gradient of the rasterizer inputs.
    python tests/golden/make_assemble_golden.py
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
IN_NAMES = ("means3D", "opacities", "scales", "rotations", "shs", "shs_p")
RAW = dict(xyz="_xyz", opacity_raw="_opacity", scaling_raw="_scaling", rotation_raw="_rotation",
           f_dc_color="_features_dc_color", f_rest_color="_features_rest_color", f_dc_phase="_features_dc_phase",
           f_rest_phase="_features_rest_phase", f_dc_amp="_features_dc_amp", f_rest_amp="_features_rest_amp")


class Recorder:
    """Stands in for GaussianRasterizer: records its inputs, returns outputs linear in them."""
    calls = []
    weights = None

    def __init__(self, raster_settings=None):
        pass

    def __call__(self, means3D, means2D, shs=None, shs_p=None, colors_precomp=None, phasors_precomp=None,
                 opacities=None, scales=None, rotations=None, cov3D_precomp=None, phase_offset=0.0, dc_offset=0.0):
        ins = dict(means3D=means3D, opacities=opacities, scales=scales, rotations=rotations, shs=shs, shs_p=shs_p)
        Recorder.calls.append({k: v.detach().clone() for k, v in ins.items()})
        val = sum((ins[k] * Recorder.weights[k]).sum() for k in IN_NAMES)
        img = val * torch.ones(1, 2, 2)
        radii = torch.ones(means3D.shape[0], dtype=torch.int32)
        return (img, img, img, img, img, img, img, img, img, img, radii)


def import_reference():
    sys.path.insert(0, REF)
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))
    ply = types.ModuleType("plyfile"); ply.PlyData = ply.PlyElement = object
    sys.modules.setdefault("plyfile", ply)
    knn = types.ModuleType("simple_knn"); knn_c = types.ModuleType("simple_knn._C"); knn_c.distCUDA2 = lambda x: None
    knn._C = knn_c
    sys.modules.setdefault("simple_knn", knn); sys.modules.setdefault("simple_knn._C", knn_c)
    scene = types.ModuleType("scene"); scene.__path__ = [os.path.join(REF, "scene")]
    sys.modules.setdefault("scene", scene)
    ras = types.ModuleType("diff_gaussian_rasterization_w_tof")
    ras.GaussianRasterizationSettings = lambda **kw: types.SimpleNamespace(**kw)
    ras.GaussianRasterizer = Recorder
    sys.modules["diff_gaussian_rasterization_w_tof"] = ras
    gr = importlib.import_module("gaussian_renderer")
    gm = importlib.import_module("scene.gaussian_model")
    return gr.render, gm.GaussianModel


def make_case(render, GaussianModel, P, isotropic, frac_dynamic, regions, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    M = 16
    raw = dict(xyz=r(P, 3), opacity_raw=r(P, 1) * 2, scaling_raw=r(P, 1 if isotropic else 3) * 0.5 - 2, rotation_raw=r(P, 4),
               f_dc_color=r(P, 1, 3), f_rest_color=r(P, M - 1, 3) * 0.1, f_dc_phase=r(P, 1, 1),
               f_rest_phase=r(P, M - 1, 1) * 0.1, f_dc_amp=r(P, 1, 1), f_rest_amp=r(P, M - 1, 1) * 0.1)
    seg = (torch.rand(P, 1, generator=g) < frac_dynamic).float()
    pc = object.__new__(GaussianModel)
    pc.isotropic = isotropic
    pc.use_view_dependent_phase = False
    pc.active_sh_degree = 3
    pc.setup_functions()
    leaves = {k: torch.nn.Parameter(v.clone()) for k, v in raw.items()}
    for k, a in RAW.items():
        setattr(pc, a, leaves[k])
    pc._features_seg_color = seg
    mask = seg[:, 0] > 0.5
    Nd = int(mask.sum())
    deltas = dict(d_xyz=(r(Nd, 3) * 0.05).requires_grad_(True), d_rot=(r(Nd, 4) * 0.05).requires_grad_(True),
                  d_sh=(r(Nd, M, 3) * 0.05).requires_grad_(True), d_sh_p=(r(Nd, M, 2) * 0.05).requires_grad_(True))
    eye = torch.eye(4)
    cam = types.SimpleNamespace(FoVx=1.0, FoVy=0.8, image_height=8, image_width=8, world_view_transform=eye,
                                full_proj_transform=eye, camera_center=torch.zeros(3), znear=0.1, zfar=10.0,
                                FoVx_tof=1.0, FoVy_tof=0.8, tof_image_height=8, tof_image_width=8,
                                world_view_transform_tof=eye, full_proj_transform_tof=eye, camera_center_tof=torch.zeros(3),
                                depth_range=torch.tensor(15.0), phase_offset=torch.tensor(0.0), dc_offset=torch.tensor(0.0))
    opt = types.SimpleNamespace(optimize_phase_offset=False, optimize_dc_offset=False)
    shapes = dict(means3D=(P, 3), opacities=(P, 1), scales=(P, 3), rotations=(P, 4), shs=(P, M, 3), shs_p=(P, M, 2))
    Recorder.weights = {k: r(*shapes[k]) for k in IN_NAMES}
    Recorder.calls = []
    real_zeros, real_zeros_like = torch.zeros, torch.zeros_like
    strip = lambda k: {kk: vv for kk, vv in k.items() if kk != "device"}
    torch.zeros = lambda *a, **k: real_zeros(*a, **strip(k))
    torch.zeros_like = lambda *a, **k: real_zeros_like(*a, **strip(k))
    try:
        pkg = render(cam, pc, deltas["d_xyz"], deltas["d_rot"], deltas["d_sh"], deltas["d_sh_p"], None, opt,
                     torch.zeros(7, 8, 8), render_regions=list(regions))
    finally:
        torch.zeros, torch.zeros_like = real_zeros, real_zeros_like
    # only the colour rasterizer's image enters the loss: d loss / d (rasterizer input k) = weights[k]
    pkg["render"].mean().backward()
    assert len(Recorder.calls) == 2
    out = {"mask": mask.numpy(), "meta": np.array([1.0 if isotropic else 0.0, float("static" in regions), float("dynamic" in regions)])}
    for k, v in raw.items():
        out["raw/" + k] = v.numpy()
        gr = leaves[k].grad
        out["graw/" + k] = (gr if gr is not None else torch.zeros_like(v)).numpy()
    for k, v in deltas.items():
        out["delta/" + k] = v.detach().numpy()
        out["gdelta/" + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).detach().numpy()
    for k in IN_NAMES:
        out["out/" + k] = Recorder.calls[0][k].numpy()
        assert torch.equal(Recorder.calls[0][k], Recorder.calls[1][k])
        out["gout/" + k] = Recorder.weights[k].numpy()
    return out


def main():
    render, GM = import_reference()
    allc = {}
    cases = dict(static_only=(120, False, 0.0, ("static", "dynamic"), 21),
                 mixed=(120, False, 0.3, ("static", "dynamic"), 22),
                 mixed_static_region=(90, False, 0.3, ("static",), 23),
                 mixed_isotropic=(90, True, 0.5, ("static", "dynamic"), 24),
                 dynamic_region=(80, False, 0.4, ("dynamic",), 25))
    for name, c in cases.items():
        for k, v in make_case(render, GM, *c).items():
            allc[name + "/" + k] = v
        print(name, "ok")
    np.savez_compressed(os.path.join(HERE, "train_assemble.npz"), **allc)


if __name__ == "__main__":
    main()
