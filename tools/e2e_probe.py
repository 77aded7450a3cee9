"""Times the pieces of bench.py's e2e step (public autograd surface, host buffers)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from gftorf_b200 import rasterizer

dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["c2"]
params, views = bench.build_scene(wl, 0, dev)
names = ("means3D", "opacities", "shs", "shs_p", "scales", "rotations")
host_params = {k: params[k].cpu().pin_memory() for k in names}
host_out = {k: torch.empty_like(host_params[k]).pin_memory() for k in names}
v = views[0]
S, R = rasterizer.GaussianRasterizationSettings, rasterizer.GaussianRasterizer

def T():
    torch.cuda.synchronize(); return time.perf_counter()

for it in range(6):
    t0 = T()
    dp = {k: host_params[k].to(dev, non_blocking=True) for k in names}
    t1 = T()
    leaves = {k: dp[k].requires_grad_(True) for k in names}
    m2d = torch.zeros_like(dp["means3D"], requires_grad=True)
    s = S(image_height=v["H"], image_width=v["W"], tanfovx=v["tanfovx"], tanfovy=v["tanfovy"], bg=v["bg"],
          scale_modifier=1.0, viewmatrix=v["viewmatrix"], projmatrix=v["projmatrix"], sh_degree=3,
          campos=v["campos"], prefiltered=False, debug=False, near_n=v["near_n"], far_n=v["far_n"],
          depth_range=v["depth_range"])
    out = R(s)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
               shs_p=leaves["shs_p"], scales=leaves["scales"], rotations=leaves["rotations"])
    t2 = T()
    g = v["grads"]
    torch.autograd.backward([out[0], out[1], out[2], out[4], out[6]],
                            [g["color"], g["phasor"], g["depth"], g["acc"], g["depth_distortion"]])
    t3 = T()
    for k in names:
        host_out[k].copy_(dp[k].grad, non_blocking=True)
    t4 = T()
    print(f"it{it}: h2d {1e3*(t1-t0):.2f} ms  fwd {1e3*(t2-t1):.2f}  bwd(autograd) {1e3*(t3-t2):.2f}  d2h {1e3*(t4-t3):.2f}")
