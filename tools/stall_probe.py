"""Which side stalls in an outlier step: the host (enqueue wall time) or the GPU (event time)?"""
import gc, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from gftorf_b200 import rasterizer
wl = bench.WORKLOADS["c2"]
params, views = bench.build_scene(wl, 0, "cuda")
empty = torch.Tensor([])
z3 = torch.zeros_like(views[0]["grads"]["color"]); z1 = torch.zeros_like(views[0]["grads"]["depth"])
mod = rasterizer._C
hint = {}
def step():
    for v in views:
        f = mod.rasterize_gaussians(*bench.fwd_args(params, v, empty), R_hint=hint.get(id(v), 0))
        hint[id(v)] = int(f[0] * 1.25) + 4096
        mod.rasterize_gaussians_backward(*bench.bwd_args(params, v, f, empty, z3, z1))
for _ in range(10): step()
torch.cuda.synchronize(); gc.collect(); gc.disable()
K = 300
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
host = []
for i in range(K):
    evs[i][0].record(); t0 = time.perf_counter(); step(); host.append((time.perf_counter() - t0) * 1e3); evs[i][1].record()
torch.cuda.synchronize()
gpu = [a.elapsed_time(b) for a, b in evs]
order = sorted(range(K), key=lambda i: -gpu[i])[:8]
print("median gpu %.3f host %.3f" % (sorted(gpu)[K // 2], sorted(host)[K // 2]))
for i in order:
    print(f"step {i}: gpu {gpu[i]:.2f} ms, host enqueue {host[i]:.2f} ms")
try:
    print(open("/sys/fs/cgroup/cpu.max").read().strip(), "| cpu.stat:", open("/sys/fs/cgroup/cpu.stat").read().replace("\n", " "))
except Exception as e:
    print("cgroup:", e)
print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)), "load:", os.getloadavg())
