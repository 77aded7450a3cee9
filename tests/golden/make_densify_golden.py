"""Generates tests/golden/train_densify.npz by running the reference's OWN
GaussianModel.densify_and_prune (scene/gaussian_model.py:631-646, with the clone / split / prune /
optimizer-surgery methods it calls) on CPU in this container.

The class hard-codes device="cuda" and imports CUDA extensions, so the script (a) stubs the
modules that are not installed here (plyfile, cv2, simple_knn._C, the `scene` package __init__),
(b) strips the device argument from torch.zeros while the method runs, (c) records the samples
torch.normal returns, so that the split can be replayed with the same noise.  The model object is
created without __init__ and given exactly the attributes the methods read.
    python tests/golden/make_densify_golden.py
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GROUPS = ("xyz", "f_dc_color", "f_rest_color", "phase_f_dc", "phase_f_rest", "amp_f_dc", "amp_f_rest",
          "opacity", "scaling", "rotation", "f_seg_color")
ATTR = dict(xyz="_xyz", f_dc_color="_features_dc_color", f_rest_color="_features_rest_color",
            phase_f_dc="_features_dc_phase", phase_f_rest="_features_rest_phase", amp_f_dc="_features_dc_amp",
            amp_f_rest="_features_rest_amp", opacity="_opacity", scaling="_scaling", rotation="_rotation",
            f_seg_color="_features_seg_color")


def import_reference_model():
    sys.path.insert(0, REF)
    for name in ("cv2",):
        sys.modules.setdefault(name, types.ModuleType(name))
    ply = types.ModuleType("plyfile")
    ply.PlyData = ply.PlyElement = object
    sys.modules.setdefault("plyfile", ply)
    knn = types.ModuleType("simple_knn"); knn_c = types.ModuleType("simple_knn._C")
    knn_c.distCUDA2 = lambda x: None
    knn._C = knn_c
    sys.modules.setdefault("simple_knn", knn); sys.modules.setdefault("simple_knn._C", knn_c)
    scene = types.ModuleType("scene"); scene.__path__ = [os.path.join(REF, "scene")]   # skip scene/__init__.py
    sys.modules.setdefault("scene", scene)
    return importlib.import_module("scene.gaussian_model").GaussianModel


def make_case(GaussianModel, P, isotropic, seed, size_threshold):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    params = dict(xyz=r(P, 3), f_dc_color=r(P, 1, 3), f_rest_color=r(P, 15, 3) * 0.1, phase_f_dc=r(P, 1, 1),
                  phase_f_rest=r(P, 15, 1) * 0.1, amp_f_dc=r(P, 1, 1), amp_f_rest=r(P, 15, 1) * 0.1,
                  opacity=r(P, 1) * 2, scaling=r(P, 1 if isotropic else 3) * 1.2 - 3.0, rotation=r(P, 4),
                  f_seg_color=torch.rand(P, 1, generator=g))
    m = object.__new__(GaussianModel)
    m.isotropic = isotropic
    m.percent_dense = 0.01
    m.setup_functions()
    leaves = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
    for k, a in ATTR.items():
        setattr(m, a, leaves[k])
    m._phase_offset = torch.nn.Parameter(torch.zeros(1))
    m._dc_offset = torch.nn.Parameter(torch.zeros(1))
    groups = [{"params": [leaves[k]], "lr": 1e-3, "name": k} for k in GROUPS]
    groups += [{"params": [m._phase_offset], "lr": 0.0, "name": "phase_offset"},
               {"params": [m._dc_offset], "lr": 0.0, "name": "dc_offset"}]
    m.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    # one optimizer step with random gradients fills exp_avg / exp_avg_sq; parameters move by lr
    for k in GROUPS:
        leaves[k].grad = torch.randn(leaves[k].shape, generator=g) * 0.01
    m.optimizer.step()
    before = {k: getattr(m, ATTR[k]).detach().clone() for k in GROUPS}
    ea = {k: m.optimizer.state[getattr(m, ATTR[k])]["exp_avg"].clone() for k in GROUPS}
    es = {k: m.optimizer.state[getattr(m, ATTR[k])]["exp_avg_sq"].clone() for k in GROUPS}
    m.xyz_gradient_accum = torch.rand(P, 1, generator=g)
    m.denom = torch.randint(0, 4, (P, 1), generator=g).float()
    m.max_radii2D = torch.rand(P, generator=g) * 40
    acc, den = m.xyz_gradient_accum.clone(), m.denom.clone()

    samples = []
    real_zeros, real_normal = torch.zeros, torch.normal
    torch.zeros = lambda *a, **k: real_zeros(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})

    def rec_normal(*a, **k):
        out = real_normal(*a, **k, generator=g)
        samples.append(out.detach().clone())
        return out
    torch.normal = rec_normal
    try:
        kw = dict(max_grad=0.3, min_opacity=0.2, extent=4.0)
        m.densify_and_prune(kw["max_grad"], kw["min_opacity"], kw["extent"], size_threshold)
    finally:
        torch.zeros, torch.normal = real_zeros, real_normal
    out = {}
    for k in GROUPS:
        t = getattr(m, ATTR[k])
        out["in/" + k], out["in_m/" + k], out["in_v/" + k] = before[k].numpy(), ea[k].numpy(), es[k].numpy()
        out["out/" + k] = t.detach().numpy()
        st = m.optimizer.state[t]
        out["out_m/" + k], out["out_v/" + k] = st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()
    out["acc"], out["den"] = acc.numpy(), den.numpy()
    out["samples"] = samples[0].numpy() if samples else np.zeros((0, 3), np.float32)
    out["meta"] = np.array([kw["max_grad"], kw["min_opacity"], kw["extent"], 0.01, 1.0 if size_threshold else 0.0,
                            1.0 if isotropic else 0.0])
    return out


def main():
    GM = import_reference_model()
    allc = {}
    for name, (P, iso, seed, thr) in dict(aniso=(300, False, 11, 20), iso=(250, True, 12, 20),
                                          nosize=(200, False, 13, None)).items():
        for k, v in make_case(GM, P, iso, seed, thr).items():
            allc[name + "/" + k] = v
        print(name, P, "->", allc[name + "/out/xyz"].shape[0], "split samples", allc[name + "/samples"].shape[0])
    np.savez_compressed(os.path.join(HERE, "train_densify.npz"), **allc)


if __name__ == "__main__":
    main()
