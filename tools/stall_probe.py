"""Outlier steps: who stalls?  Per step: GPU event time, host time per call, the scheduler's
run-queue delay for this thread (/proc/thread-self/schedstat), context switches."""
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
def sched():
    a = open("/proc/thread-self/schedstat").read().split()
    return int(a[0]), int(a[1])
def ctx():
    d = {}
    for l in open("/proc/thread-self/status"):
        if "ctxt_switches" in l:
            k, v = l.split(":"); d[k.strip()] = int(v)
    return d["voluntary_ctxt_switches"], d["nonvoluntary_ctxt_switches"]
def step(rec):
    fs = []
    for v in views:
        t0 = time.perf_counter()
        f = mod.rasterize_gaussians(*bench.fwd_args(params, v, empty), R_hint=hint.get(id(v), 0))
        rec.append(time.perf_counter() - t0)
        hint[id(v)] = int(f[0] * 1.25) + 4096
        fs.append(f)
    for v, f in zip(views, fs):
        t0 = time.perf_counter()
        mod.rasterize_gaussians_backward(*bench.bwd_args(params, v, f, empty, z3, z1))
        rec.append(time.perf_counter() - t0)
for _ in range(10): step([])
torch.cuda.synchronize(); gc.collect(); gc.disable()
K = int(os.environ.get("K", "400"))
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
rows = []
tstart = time.perf_counter()
for i in range(K):
    rec = []
    s0, c0 = sched(), ctx()
    evs[i][0].record(); t0 = time.perf_counter(); step(rec); host = time.perf_counter() - t0; evs[i][1].record()
    s1, c1 = sched(), ctx()
    ms_ = torch.cuda.memory_stats()
    rows.append((t0 - tstart, host, rec, (s1[0] - s0[0]) / 1e6, (s1[1] - s0[1]) / 1e6, c1[0] - c0[0], c1[1] - c0[1],
                 ms_["num_device_alloc"], ms_["num_device_free"], ms_["reserved_bytes.all.current"] >> 20, ms_["allocated_bytes.all.current"] >> 20))
    if os.environ.get("SYNC_EACH") == "1":
        torch.cuda.synchronize()
torch.cuda.synchronize()
gpu = [a.elapsed_time(b) for a, b in evs]
order = sorted(range(K), key=lambda i: -gpu[i])[:12]
print("median gpu %.3f host %.3f ms; total gpu %.1f ms over %d steps" % (sorted(gpu)[K // 2], sorted(r[1] for r in rows)[K // 2] * 1e3, sum(gpu), K))
for i in sorted(order):
    t, host, rec, oncpu, rundelay, vol, nonvol, nda, ndf, rsv, alc = rows[i]
    pda = rows[i - 1][7] if i else 0
    print(f"step {i} @{t*1e3:8.1f} ms: gpu {gpu[i]:7.2f} host {host*1e3:7.2f} | calls(ms) " + " ".join(f"{x*1e3:.2f}" for x in rec) +
          f" | on-cpu {oncpu:.2f} runq {rundelay:.2f} v{vol}/nv{nonvol} | cudaMallocs {nda} (+{nda - pda}) frees {ndf} reserved {rsv} MiB allocated {alc} MiB")
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.stat", "/proc/pressure/cpu", "/proc/loadavg"):
    try: print(p, "->", open(p).read().strip().replace("\n", " "))
    except Exception as e: print(p, "->", type(e).__name__)
os.system("ps -eo pid,pcpu,etimes,comm --sort=-pcpu | head -8")
os.system("nvidia-smi --query-gpu=persistence_mode,clocks.sm,power.draw,temperature.gpu --format=csv,noheader")
