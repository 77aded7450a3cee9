"""Upper bound for running the two views of an iteration concurrently (two host threads, two
streams) instead of back to back.  Separate gradient outputs per view (no shared bucket)."""
import gc, os, sys, time, threading
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from gftorf_b200 import rasterizer
wl = bench.WORKLOADS[os.environ.get("WL", "c2")]
params, views = bench.build_scene(wl, 0, "cuda")
empty = torch.Tensor([])
z3 = torch.zeros_like(views[0]["grads"]["color"]); z1 = torch.zeros_like(views[0]["grads"]["depth"])
z3b = torch.zeros_like(views[1]["grads"]["color"]); z1b = torch.zeros_like(views[1]["grads"]["depth"])
mod = rasterizer._C
hint = {}
def one(v, a3, a1):
    f = mod.rasterize_gaussians(*bench.fwd_args(params, v, empty), R_hint=hint.get(id(v), 0))
    hint[id(v)] = int(f[0] * 1.25) + 4096
    mod.rasterize_gaussians_backward(*bench.bwd_args(params, v, f, empty, a3, a1))
def seq():
    one(views[0], z3, z1); one(views[1], z3b, z1b)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
class Worker(threading.Thread):
    def __init__(self, i):
        super().__init__(daemon=True); self.i = i; self.go = threading.Event(); self.done = threading.Event(); self.start()
    def run(self):
        torch.cuda.set_device(0)
        while True:
            self.go.wait(); self.go.clear()
            with torch.cuda.stream(streams[self.i]):
                one(views[self.i], (z3, z3b)[self.i], (z1, z1b)[self.i])
            self.done.set()
workers = [Worker(0), Worker(1)]
def par():
    main = torch.cuda.current_stream()
    for s in streams: s.wait_stream(main)
    for w in workers: w.go.set()
    for w in workers: w.done.wait(); w.done.clear()
    for s in streams: main.wait_stream(s)
def timeit(fn, K=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); gc.collect(); gc.disable()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K): fn()
    b.record(); torch.cuda.synchronize(); gc.enable()
    return a.elapsed_time(b) / K
print("sequential %.4f ms/step" % timeit(seq))
print("two streams %.4f ms/step" % timeit(par))
print("sequential %.4f ms/step" % timeit(seq))
