"""Where does a densification iteration go?  Host and device time of densify_and_prune, the FlatAdam
rebuild, and the first rasterizer iteration after the Gaussian count changed (C3 shape)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import test_train_loop as tl  # noqa: E402
from gftorf_b200 import train_ops as T, views as V  # noqa: E402

dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["c3"]
params, views = bench.build_scene(wl, seed=0, device=dev)
inp = {k: params[k] for k in ("means3D", "opacities", "scales", "rotations", "shs", "shs_p")}
model0 = tl.initial_model(inp)
names = list(tl.LRS)
specs = [bench.view_spec(v) for v in views]
fa = T.FlatAdam([(n, model0[n], tl.LRS[n]) for n in names])
kw = dict(max_grad=2e-4, min_opacity=0.005, extent=5.0, percent_dense=0.01)


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"  {label:34s} {(t1 - t0) * 1e3:8.2f} ms", flush=True)
    return t1


def iteration(fa):
    raw = {tl.RAW_OF[n]: fa.params[n] for n in names}
    asm = T.assemble_gaussians(raw)
    Pn = asm["means3D"].shape[0]
    m2d = torch.zeros((2, Pn, 3), device=dev, requires_grad=True)
    oc, ot = V.rasterize_views(asm["means3D"], m2d, asm["opacities"], asm["shs"], asm["shs_p"], asm["scales"],
                               asm["rotations"], specs, 3)
    torch.autograd.backward([oc[0], ot[1][5:6]], [torch.ones_like(oc[0]), torch.ones_like(ot[1][5:6])])
    fa.step(zero_grad=True)
    return m2d, oc


for rnd in range(3):
    for _ in range(3):
        m2d, oc = iteration(fa)
    torch.cuda.synchronize()
    print("round", rnd, "P =", fa.params["xyz"].shape[0])
    t = time.perf_counter()
    m2d, oc = iteration(fa)
    t = tick("plain iteration", t)
    acc = torch.norm(m2d.grad[0][:, :2], dim=-1, keepdim=True)
    den = (oc[10] > 0).float().unsqueeze(1)
    P0 = fa.params["xyz"].shape[0]
    prm = {n: fa.params[n].detach() for n in names}
    prm["f_seg_color"] = torch.zeros(P0, 1, device=dev)
    ea = {n: fa.exp_avg[fa.bounds[n][0]:fa.bounds[n][1]].view(fa.bounds[n][2]) for n in names}
    es = {n: fa.exp_avg_sq[fa.bounds[n][0]:fa.bounds[n][1]].view(fa.bounds[n][2]) for n in names}
    ea["f_seg_color"] = es["f_seg_color"] = torch.zeros(P0, 1, device=dev)
    t = tick("statistics + views of the moments", t)
    p_new, m_new, v_new, info = T.densify_and_prune(prm, ea, es, acc, den, generator=torch.Generator(device=dev).manual_seed(3), **kw)
    t = tick("densify_and_prune", t)
    fa2 = T.FlatAdam([(n, p_new[n], tl.LRS[n]) for n in names])
    t = tick("FlatAdam construction", t)
    for n in names:
        lo, hi, shp = fa2.bounds[n]
        fa2.exp_avg[lo:hi].view(shp).copy_(m_new[n])
        fa2.exp_avg_sq[lo:hi].view(shp).copy_(v_new[n])
    t = tick("moment copies", t)
    fa = fa2
    del prm, ea, es, p_new, m_new, v_new, fa2
    m2d, oc = iteration(fa)
    t = tick("first iteration at the new size", t)
    m2d, oc = iteration(fa)
    t = tick("second iteration at the new size", t)
