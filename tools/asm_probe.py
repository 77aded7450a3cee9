"""Where does the time of assemble_gaussians go: kernels or host bookkeeping?"""
import os, sys, time, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_train_ops as tt
from gftorf_b200 import train_ops as T
P = 300000
raw = {k: v.cuda().requires_grad_(True) for k, v in tt.make_raw(P, seed=5).items()}
mask = torch.zeros(P, dtype=torch.bool, device="cuda"); mask[::4] = True
deltas = {k: v.cuda().requires_grad_(True) for k, v in tt.make_deltas(int(mask.sum()), seed=6).items()}
dyn = T.dyn_index_from_mask(mask)
def ev(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1e3 / n
o = T.assemble_gaussians(raw, dyn, deltas)
g = [torch.randn_like(o[k]) for k in T.OUT_NAMES]
print("forward only           gpu %.4f ms  wall %.4f ms" % ev(lambda: T.assemble_gaussians(raw, dyn, deltas)))
with torch.no_grad():
    print("forward only, no_grad  gpu %.4f ms  wall %.4f ms" % ev(lambda: T.assemble_gaussians(raw, dyn, deltas)))
def fb():
    for t in list(raw.values()) + list(deltas.values()): t.grad = None
    o = T.assemble_gaussians(raw, dyn, deltas)
    torch.autograd.backward([o[k] for k in T.OUT_NAMES], g)
print("forward + backward     gpu %.4f ms  wall %.4f ms" % ev(fb))
