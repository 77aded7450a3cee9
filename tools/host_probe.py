"""Where does the host time of one step go?  Wall-clock per phase with and without device syncs."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from gftorf_b200 import rasterizer

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
params, views = bench.build_scene(wl, 0, "cuda")
empty = torch.Tensor([])
z3 = torch.zeros_like(views[0]["grads"]["color"]); z1 = torch.zeros_like(views[0]["grads"]["depth"])
mod = rasterizer._C
v = views[0]
for _ in range(5):
    f = mod.rasterize_gaussians(*bench.fwd_args(params, v, empty))
    b = mod.rasterize_gaussians_backward(*bench.bwd_args(params, v, f, empty, z3, z1))
torch.cuda.synchronize()

def wall(fn, n=50, sync=True):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        if sync: torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    ts.sort()
    return ts[len(ts) // 2], r

t_fwd_sync, f = wall(lambda: mod.rasterize_gaussians(*bench.fwd_args(params, v, empty)))
t_fwd_nosync, f = wall(lambda: mod.rasterize_gaussians(*bench.fwd_args(params, v, empty)), sync=False)
t_bwd_sync, b = wall(lambda: mod.rasterize_gaussians_backward(*bench.bwd_args(params, v, f, empty, z3, z1)))
t_bwd_nosync, b = wall(lambda: mod.rasterize_gaussians_backward(*bench.bwd_args(params, v, f, empty, z3, z1)), sync=False)
print(f"forward : wall+sync {t_fwd_sync:.0f} us, host-only return {t_fwd_nosync:.0f} us")
print(f"backward: wall+sync {t_bwd_sync:.0f} us, host-only return {t_bwd_nosync:.0f} us")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    f = mod.rasterize_gaussians(*bench.fwd_args(params, v, empty))
    b = mod.rasterize_gaussians_backward(*bench.bwd_args(params, v, f, empty, z3, z1))
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
