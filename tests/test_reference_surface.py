"""The Python surface against the reference's OWN surface, call for call.

The reference's diff_gaussian_rasterization_w_tof/__init__.py is pure Python above its `_C`
extension.  Here it is imported from /root/reference (this container only — the test is skipped
where the reference is absent) with `_C` replaced by a recorder, and our surface is driven with the
same recorder: both must hand the native layer the same positional arguments, in the same order,
forward (27) and backward (36), return the same 11 outputs, and route the 12 native gradients to
the same inputs.  No GPU, no arithmetic: this pins rows a1-a4 of SURVEY.md §8."""
import importlib.util
import os
import sys
import types

import pytest
import torch

from gftorf_b200 import rasterizer

REF_INIT = "/root/reference/submodules/diff-gaussian-rasterization-w-tof/diff_gaussian_rasterization_w_tof/__init__.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF_INIT), reason="reference checkout not present")


class Recorder:
    """Stands in for the `_C` extension module (ext.cpp:15-19)."""

    def __init__(self):
        self.fwd_args, self.bwd_args = [], []

    def rasterize_gaussians(self, *args):
        self.fwd_args.append(args)
        means3D, H, W = args[1], args[13], args[14]
        P = means3D.shape[0]
        img = lambda c: torch.full((c, H, W), float(c))
        buf = lambda n: torch.arange(n, dtype=torch.uint8)
        return (7, img(3), img(7), img(1), img(3), img(1), img(1), img(1), img(1), torch.ones(P, 1), img(3),
                torch.ones(P, dtype=torch.int32), buf(5), buf(6), buf(7))

    def rasterize_gaussians_backward(self, *args):
        self.bwd_args.append(args)
        means3D, sh, sh_p = args[1], args[21], args[22]
        P = means3D.shape[0]
        z = lambda *s: torch.full(s, 0.5)
        # grad_means2D, colors_precomp, phasors_precomp, opacities, means3D, cov3Ds, sh, sh_p, scales, rotations, phase, dc
        return (z(P, 3), z(P, 3), z(P, 2), z(P, 1), z(P, 3), z(P, 6), torch.full_like(sh, 0.5) if sh.numel() else z(P, 0, 3),
                torch.full_like(sh_p, 0.5) if sh_p.numel() else z(P, 0, 2), z(P, 3), z(P, 4), z(1), z(1))

    def mark_visible(self, *args):
        self.fwd_args.append(args)
        return torch.ones(args[0].shape[0], dtype=torch.bool)


def load_reference_surface(rec):
    name = "_ref_surface_pkg"
    for k in [k for k in sys.modules if k.startswith(name)]:
        del sys.modules[k]
    pkg_spec = importlib.util.spec_from_file_location(name, REF_INIT, submodule_search_locations=[])
    mod = importlib.util.module_from_spec(pkg_spec)
    sys.modules[name] = mod
    cmod = types.ModuleType(name + "._C")
    cmod.rasterize_gaussians = rec.rasterize_gaussians
    cmod.rasterize_gaussians_backward = rec.rasterize_gaussians_backward
    cmod.mark_visible = rec.mark_visible
    sys.modules[name + "._C"] = cmod
    mod._C = cmod
    pkg_spec.loader.exec_module(mod)
    return mod


def same(a, b):
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.shape == b.shape and \
            a.dtype == b.dtype and torch.equal(a.detach(), b.detach())
    if isinstance(a, float) or isinstance(b, float):
        return float(a) == pytest.approx(float(b), rel=0, abs=0)
    return a == b


CASES = [
    dict(name="sh", optimize=False, colors=False, cov=False),
    dict(name="sh_optimized_offsets", optimize=True, colors=False, cov=False),
    dict(name="precomputed_colours", optimize=False, colors=True, cov=False),
    dict(name="precomputed_cov", optimize=False, colors=False, cov=True),
]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_same_native_calls_as_the_reference_surface(case, monkeypatch):
    P, H, W = 6, 8, 10
    g = torch.Generator().manual_seed(1)
    base = dict(means3D=torch.randn(P, 3, generator=g), opacities=torch.rand(P, 1, generator=g),
                shs=torch.randn(P, 16, 3, generator=g), shs_p=torch.randn(P, 16, 2, generator=g),
                scales=torch.rand(P, 3, generator=g), rotations=torch.randn(P, 4, generator=g),
                colors=torch.rand(P, 3, generator=g), cov=torch.rand(P, 6, generator=g))

    def run(surface, rec):
        leaves = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        m2d = torch.zeros(P, 3, requires_grad=True)
        phase = torch.nn.Parameter(torch.tensor([0.25])) if case["optimize"] else 0.25
        dc = torch.nn.Parameter(torch.tensor([0.05])) if case["optimize"] else 0.05
        s = surface.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=0.6, tanfovy=0.5, bg=torch.rand(7, H, W, generator=torch.Generator().manual_seed(2)),
            scale_modifier=1.25, viewmatrix=torch.eye(4) * 2, projmatrix=torch.eye(4) * 3, sh_degree=2,
            campos=torch.tensor([1.0, 2.0, 3.0]), prefiltered=False, debug=False, near_n=0.2, far_n=9.0,
            depth_range=12.0, use_view_dependent_phase=True, optimize_phase_offset=case["optimize"],
            optimize_dc_offset=case["optimize"])
        kw = dict(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs_p=leaves["shs_p"],
                  phase_offset=phase, dc_offset=dc)
        if case["colors"]:
            kw["colors_precomp"] = leaves["colors"]
        else:
            kw["shs"] = leaves["shs"]
        if case["cov"]:
            kw["cov3D_precomp"] = leaves["cov"]
        else:
            kw["scales"], kw["rotations"] = leaves["scales"], leaves["rotations"]
        out = surface.GaussianRasterizer(s)(**kw)
        diff = [o for o in out if o.requires_grad]
        torch.autograd.backward(diff, [torch.full_like(o, 2.0) for o in diff])
        grads = {k: (None if v.grad is None else v.grad.clone()) for k, v in leaves.items()}
        grads["means2D"] = None if m2d.grad is None else m2d.grad.clone()
        if case["optimize"]:
            grads["phase"], grads["dc"] = phase.grad, dc.grad
        vis = surface.GaussianRasterizer(s).markVisible(leaves["means3D"].detach())
        return out, grads, vis

    rec_ref, rec_ours = Recorder(), Recorder()
    ref = load_reference_surface(rec_ref)
    out_r, grads_r, vis_r = run(ref, rec_ref)
    monkeypatch.setattr(rasterizer, "_C", rec_ours)
    out_o, grads_o, vis_o = run(rasterizer, rec_ours)

    # forward: 27 positional arguments, same order and values (+ the mark_visible call)
    assert len(rec_ref.fwd_args) == len(rec_ours.fwd_args) == 2
    for a_r, a_o in zip(rec_ref.fwd_args, rec_ours.fwd_args):
        assert len(a_r) == len(a_o)
        for i, (x, y) in enumerate(zip(a_r, a_o)):
            assert same(x, y), ("forward arg", i, x, y)
    assert len(rec_ref.fwd_args[0]) == 27
    # backward: 36 positional arguments
    assert len(rec_ref.bwd_args) == len(rec_ours.bwd_args) == 1
    assert len(rec_ref.bwd_args[0]) == len(rec_ours.bwd_args[0]) == 36
    for i, (x, y) in enumerate(zip(rec_ref.bwd_args[0], rec_ours.bwd_args[0])):
        if i in (34, 35) and case["optimize"]:
            # the reference hands the Parameter itself to a C++ float argument (pybind converts it);
            # we convert once in the forward — same value
            assert float(x) == float(y), ("backward arg", i)
            continue
        assert same(x, y), ("backward arg", i, x, y)
    # outputs: 11, same order / shapes / values
    assert len(out_r) == len(out_o) == 11
    for x, y in zip(out_r, out_o):
        assert same(x, y)
    assert torch.equal(vis_r, vis_o)
    # gradients reach the same inputs with the same values
    for k in grads_r:
        gr, go = grads_r[k], grads_o[k]
        if k == "colors" and not case["colors"] or k == "shs" and case["colors"] or \
                k == "cov" and not case["cov"] or k in ("scales", "rotations") and case["cov"]:
            assert go is None                      # input not part of this call at all
            continue
        assert (gr is None) == (go is None), k
        if gr is not None:
            assert torch.equal(gr, go), k
