#!/usr/bin/env python
"""bench.py — measures the hot path BASELINE.json names: the RGB+ToF Gaussian rasterizer's
forward+backward inside a training iteration, on synthetic F-TöRF-shaped scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|cpu] [--workload c2]

A STEP is one training iteration's worth of rasterizer work for one camera pair, the call pattern
of gaussian_renderer/__init__.py:107-128 + train.py:279: forward of the colour view, forward of
the ToF view, backward of both (every output computed, all five consumed output gradients dense).
At N > 1 every rank does one such step on its own views (weak scaling) and the step ends with the
NCCL sum-allreduce of the flat per-Gaussian gradient bucket.

  value  : Mpix/s of pixels taken through forward+backward, whole job, inputs resident in HBM,
           called through the C-ABI entry points (gft_forward_views / gft_backward_views: the
           two views of the step in ONE batched call; the parameter gradients are written
           straight into the gradient bucket, each row once).  ms_per_step is BASELINE's
           "fwd+bwd ms/iter".  Per-kernel times (stage_ms_per_step, roofline) come from a second
           pass of the same steps with the library's per-stage CUDA events switched on.
  e2e    : the same metric through the public autograd surface (gftorf_b200.rasterize_views +
           backward()) with HOST buffers: pinned-host -> device copies of all Gaussian parameters,
           cameras, background and pixel gradients, and device -> host reads of the parameter
           gradients and the rendered images, all inside the timed region; copies of neighbouring
           steps overlap the kernels (three streams, double-buffered staging).
  blocks : the other BASELINE configs, each with its own reference-arm number in the reference
           line: c1 (small scene), c4_batch8 (8 cameras x (1080p + 640x480), 2 M Gaussians,
           cameras sharded over the ranks, strong scaling, allreduce), c5_sweep (120-frame
           trajectory at 1080p, 1 M and 8 M Gaussians, frames sharded over the ranks), and at
           N = 1 c3_iter (a full fused training iteration incl. densification), next_rows (the
           SURVEY §8f operators) and cpu_pipeline (the §8d CPU baseline).
  --impl reference : the UNMODIFIED reference kernels (oracle/_ref/libgftorf_ref.so, built from
           /root/reference by oracle/Makefile) driven the way the reference's torch binding drives
           them (oracle/ref_driver.py) — same workloads, same metric, on the GPU.  The reference has
           no CPU implementation, so its "own implementation of the path" is this.
           Falls back to the CPU port (oracle/gft_oracle.cpp) only if that library is absent.
  --impl cpu : the CPU port on the host cores (the cpu_baseline, as a full line).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gftorf_b200 import scenes  # noqa: E402

WORKLOADS = {
    # configs[1] of BASELINE.json: F-TöRF synthetic-shaped scene, 300k Gaussians, 640x480 RGB + ToF
    "c2": dict(P=300000, color=(640, 480), tof=(640, 480), depth_range=15.0, kind="trained",
               desc="BASELINE configs[1]: 300k Gaussians, 640x480 RGB view + 640x480 ToF view, SH degree 3"),
    # the same shape at INITIALISATION (dataset_readers.py:891-903): uniform cloud, isotropic scale from
    # distCUDA2, opacity 0.1 — huge splats (median radius ~50 px), R/V ~ 25
    "c2_init": dict(P=300000, color=(640, 480), tof=(640, 480), depth_range=15.0, kind="init",
                    desc="configs[1] shape at initialisation: 300k init-like Gaussians, 640x480 x2"),
    # configs[0]: the CPU-runnable case
    "c1": dict(P=20000, color=(320, 240), tof=(320, 240), depth_range=15.0, kind="trained",
               desc="BASELINE configs[0]: 20k Gaussians, 320x240"),
    # configs[2] shape (TöRF real-shaped): 500k Gaussians, ToF 320x240 + colour 640x480
    "c3": dict(P=500000, color=(640, 480), tof=(320, 240), depth_range=10.0, kind="trained",
               desc="BASELINE configs[2] shape: 500k Gaussians, 640x480 colour + 320x240 ToF"),
    # configs[3] per-camera shape: 2M Gaussians, 1080p RGB + 640x480 ToF
    "c4": dict(P=2000000, color=(1920, 1080), tof=(640, 480), depth_range=15.0, kind="trained",
               desc="BASELINE configs[3] per-camera shape: 2M Gaussians, 1080p colour + 640x480 ToF"),
    # configs[4] upper end: 8M Gaussians (screen-space sigma 1 px), two 1080p frames per step
    "c5": dict(P=8000000, color=(1920, 1080), tof=(1920, 1080), depth_range=15.0, kind="trained", sigma_px=1.0,
               desc="BASELINE configs[4] upper end: 8M Gaussians, two 1920x1080 frames"),
}
PARAM_NAMES = ("means3D", "opacities", "shs", "shs_p", "scales", "rotations")


# --------------------------------------------------------------------------------------------
def _t(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def make_view(P, wh, depth_range, kind, seed, device, cloud=None, bg_hw=None, sigma_px=1.5, pose="identity",
              bg=None, grads=None):
    W, H = wh
    cam = scenes.make_camera(W, H, depth_range=depth_range, pose=pose, seed=seed)
    if cloud is None:
        cloud = scenes.make_cloud(P, cam, kind=kind, seed=seed, sigma_px=sigma_px)
    if bg is None:
        bh, bw = bg_hw if bg_hw else (H, W)
        bg = _t(scenes.make_background(bh, bw, seed=seed), device)
    if grads is None:
        grads = {k: _t(v, device) for k, v in scenes.make_pixel_grads(H, W, seed=seed).items()}
    view = dict(W=W, H=H, viewmatrix=_t(cam["viewmatrix"], device), projmatrix=_t(cam["projmatrix"], device),
                campos=_t(cam["campos"], device), tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
                near_n=cam["znear"], far_n=cam["zfar"], depth_range=cam["depth_range"], bg=bg, grads=grads)
    return view, cloud


def build_scene(wl, seed, device):
    color, cloud = make_view(wl["P"], wl["color"], wl["depth_range"], wl["kind"], seed, device,
                             sigma_px=wl.get("sigma_px", 1.5))
    # the bg map is sized from the colour camera and reused for the ToF view (train.py:121-128)
    tof, _ = make_view(wl["P"], wl["tof"], wl["depth_range"], wl["kind"], seed, device, cloud=cloud,
                       bg=color["bg"])
    params = {k: _t(cloud[k], device) for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p")}
    return params, [color, tof]


def build_cameras(wl, n_cams, seed, device, with_tof=True):
    """One Gaussian cloud seen by `n_cams` cameras on a trajectory around the first (identity) pose:
    per camera a colour view (+ a ToF view).  Backgrounds and pixel gradients are shared by the
    views of one resolution (their content does not matter for the timing)."""
    first, cloud = make_view(wl["P"], wl["color"], wl["depth_range"], wl["kind"], seed, device,
                             sigma_px=wl.get("sigma_px", 1.5))
    tgrads = {k: _t(v, device) for k, v in scenes.make_pixel_grads(wl["tof"][1], wl["tof"][0], seed=seed).items()} \
        if with_tof else None
    cams = []
    for c in range(n_cams):
        pose = "identity" if c == 0 else "orbit"
        col = first if c == 0 else make_view(wl["P"], wl["color"], wl["depth_range"], wl["kind"], seed + 100 + c,
                                             device, cloud=cloud, pose=pose, bg=first["bg"], grads=first["grads"])[0]
        views = [col]
        if with_tof:
            views.append(make_view(wl["P"], wl["tof"], wl["depth_range"], wl["kind"], seed + 100 + c, device,
                                   cloud=cloud, pose=pose, bg=first["bg"], grads=tgrads)[0])
        cams.append(views)
    params = {k: _t(cloud[k], device) for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p")}
    return params, cams


def view_spec(v):
    """A bench view as the batched API's ViewSpec."""
    from gftorf_b200 import views as V
    return V.ViewSpec(v["H"], v["W"], v["tanfovx"], v["tanfovy"], v["bg"], v["viewmatrix"], v["projmatrix"],
                      v["campos"], v["near_n"], v["far_n"], v["depth_range"])


def fwd_args(params, v, empty):
    return (v["bg"], params["means3D"], empty, empty, params["opacities"], params["scales"],
            params["rotations"], 1.0, empty, v["viewmatrix"], v["projmatrix"], v["tanfovx"],
            v["tanfovy"], v["H"], v["W"], params["shs"], params["shs_p"], 3, v["campos"], False,
            False, v["near_n"], v["far_n"], v["depth_range"], False, 0.0, 0.0)


def bwd_args(params, v, f, empty, zero3, zero1):
    g = v["grads"]
    return (v["bg"], params["means3D"], f[11], empty, empty, params["scales"], params["rotations"],
            1.0, empty, v["viewmatrix"], v["projmatrix"], v["tanfovx"], v["tanfovy"], g["color"],
            g["phasor"], g["depth"], zero3, g["acc"], zero1, g["depth_distortion"], zero1,
            params["shs"], params["shs_p"], 3, v["campos"], f[12], f[0], f[13], f[14], False,
            v["near_n"], v["far_n"], v["depth_range"], False, 0.0, 0.0)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), through NVML, from
    the timing thread itself: a sample right after a timed step has been enqueued (its kernels and
    its exchange still running), outside the step's own CUDA events, on every rank at the same
    steps (a delay on one rank only is measured by the others inside their allreduce) and on a
    handful of steps only (the two NVML calls take ~0.1-1 ms of host time).  (A background
    poller — round 1 ran `nvidia-smi -lms`, an NVML thread was tried first this round — takes the
    driver's lock under a kernel launch every now and then: one 2.7-8 ms straggler step per run in
    SCALE_r01 and again with the NVML thread.)  Falls back to one nvidia-smi query after the region
    when NVML is not importable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index, every=1):
        self.index, self.every, self.n = index, max(1, every), 0
        self.sm, self.mx, self.reasons, self.nv, self.h, self.cost = [], [], set(), None, None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
            for _ in range(3):      # the FIRST call of each query takes 2-5 ms (seen as a straggler at the
                pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)          # first sampled step
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)                # on every rank)
        except Exception:
            self.nv = None

    def sample(self):
        self.n += 1
        if self.nv is None or (self.n % self.every):
            return
        try:
            t0 = time.perf_counter()
            self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            self.cost.append((time.perf_counter() - t0) * 1e3)
            for name, bit in self.REASONS:
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def result(self):
        if self.nv is None:
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                p = [x.strip() for x in subprocess.check_output(
                    ["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                    text=True, timeout=20).split(",")]
                reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6])
                           if v.lower().startswith("active")]
                return {"sm_mhz": float(p[0]), "sm_max_mhz": float(p[1]), "reasons": reasons, "samples": 1,
                        "source": "nvidia-smi, one query right after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"], "source": None}
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"], "source": "nvml"}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)) if self.mx else None,
                "sm_mhz_min": float(min(self.sm)), "reasons": sorted(self.reasons), "samples": len(self.sm),
                "host_ms_per_sample": round(sorted(self.cost)[len(self.cost) // 2], 3) if self.cost else None,
                "source": "nvml, sampled by the timing thread right after enqueuing a timed step (its kernels still "
                          "running), on every rank, every K//6-th step"}


def stage_bytes(stage, P, V, R, N, T, nviews=1):
    """Algorithmic bytes per launch of each stage (SURVEY §8d split per kernel).  P is
    counted once per launch (the batched kernels read the Gaussian parameters once for all views);
    V, R, N, T are sums over the views of the launch."""
    return {
        "preprocess_fwd": 12 * P + 12 * nviews * P + 352 * min(V, P) + 112 * V,
        "tile_scan": 8 * T,
        "scatter_entries": 12 * V + 8 * R,
        "tile_sort": 8 * R + 12 * R,
        "blend_fwd": 76 * R + 128 * N,
        "zero_grad_records": 64 * nviews * P,
        "blend_bwd": 76 * R + 96 * N + 80 * V,
        "preprocess_bwd": 4 * nviews * P + 364 * min(V, P) + 120 * V + 364 * P + 12 * nviews * P,
    }.get(stage, 0)


def step_bytes(P, Vs, Rs, Ns, Ts):
    """fwd+bwd compulsory traffic of one rasterizer call, summed over the views."""
    return sum(504 * P + 936 * Vs[i] + 188 * Rs[i] + 224 * Ns[i] + 8 * Ts[i] for i in range(len(Vs)))


def numa_pin(local):
    """Bind this rank's host threads (and with them the first-touch placement of its pinned
    buffers) to the NUMA node its GPU hangs off.  No-op on single-node hosts."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        if node < 0 or len(nodes) < 2:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
class Timer:
    """K timed steps after W warm-up steps: L2 flushed before every step (outside its events),
    per-step CUDA events, barrier + synchronize on both sides, sum of the per-step durations,
    MAX over ranks."""

    def __init__(self, dev, world, flush):
        self.dev, self.world, self.flush = dev, world, flush

    def barrier(self):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(self, step_fn, K, W, flush=True, sampler=None, after_warmup=None):
        # collect BEFORE the warm-up and keep the collector off from there on: a collection between
        # the warm-up and the timed steps returns cached blocks to the allocator in a different
        # order, and the first timed step then pays for fresh cudaMallocs (seen as one 1.5-2x
        # straggler at the start of every timed region, in both arms)
        gc.collect()
        gc.disable()      # a cyclic-GC pause inside a 2 ms step would be charged to the step
        out = None
        for _ in range(W):
            # keep the previous step's result alive while the next one runs, exactly like the timed
            # loop below: the allocator then owns two sets of workspaces before the timing starts
            # (otherwise the SECOND timed step pays the cudaMalloc of the second set: 5-240 ms,
            # the "one straggler per run" of round 1)
            out = step_fn()
        self.barrier()
        if after_warmup is not None:
            after_warmup()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        host = []
        t0 = time.perf_counter()
        for i in range(K):
            if flush:
                self.flush.zero_()             # 256 MiB > 126 MB L2, untimed
            evs[i][0].record()
            h0 = time.perf_counter()
            out = step_fn()
            host.append((time.perf_counter() - h0) * 1e3)
            evs[i][1].record()
            if sampler is not None:
                # clocks under load: the step's last kernels (and its exchange) are still running
                # when the host gets here.  Sampling BEFORE the step delayed rank 0's start by the
                # NVML call, and the other ranks measured that delay inside their allreduce.
                sampler.sample()
        self.barrier()
        self.wall_ms = (time.perf_counter() - t0) * 1e3
        gc.enable()
        ms = [a.elapsed_time(b) for a, b in evs]
        self.steps_ms = ms
        self.host_ms = host                    # host time spent inside step_fn (enqueue + its own waits)
        total = torch.tensor([sum(ms)], dtype=torch.float64, device=self.dev)
        self.rank_totals_ms = [float(total.item())]
        self.rank_steps_ms = [ms]
        if self.world > 1:
            mine = torch.tensor(ms, dtype=torch.float64, device=self.dev)
            every = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(every, mine)
            self.rank_steps_ms = [[float(x) for x in t.tolist()] for t in every]
            self.rank_totals_ms = [sum(r) for r in self.rank_steps_ms]
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item()), out


def slowest_step(timer):
    """Which timed step was the slowest, its device time and the host time spent inside step_fn."""
    ms, host = timer.steps_ms, timer.host_ms
    w = max(range(len(ms)), key=lambda i: ms[i])
    return {"index": w, "gpu_ms": round(ms[w], 4), "host_ms_in_step_fn": round(host[w], 4),
            "median_host_ms_in_step_fn": round(sorted(host)[len(host) // 2], 4)}


def min_med_max(ms):
    s = sorted(ms)
    return [round(s[0], 4), round(s[len(s) // 2], 4), round(s[-1], 4)]


class Ours:
    """The product: one batched call per step (gftorf_b200.views)."""
    name = "ours"

    def __init__(self, dev):
        from gftorf_b200 import views as V, parallel, _capi
        self.V, self.parallel, self.capi, self.dev = V, parallel, _capi, dev
        self.hints = {}

    def make_bucket(self, params):
        scal = [torch.zeros(1, device=self.dev), torch.zeros(1, device=self.dev)]
        # N > 1: the bucket lives in symmetric memory and the exchange is whichever of NCCL and the
        # library's NVLS kernels is fastest on this very buffer (timed once, at construction)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        b = self.parallel.GradBucket(params, scal, symmetric="auto" if multi else False)
        forced = os.environ.get("GFT_BENCH_EXCHANGE")          # A/B runs: nccl | nvls | nvls_fused
        if multi and forced in ("nccl", "nvls", "nvls_fused") and getattr(b, "_symm", None) is not None:
            b.mode = forced
        return b, b.grad_out()

    def forward(self, key, params, views, specs=None):
        specs = specs or [view_spec(v) for v in views]
        f = self.V.forward_views(params["means3D"], params["opacities"], params["scales"], params["rotations"],
                                 params["shs"], params["shs_p"], specs, 3, R_hint=self.hints.get(key, 0))
        # the largest instance count seen under this key (+25 %) sizes the binning workspace, so the
        # forward is enqueued without a host round trip in the middle
        self.hints[key] = max(self.hints.get(key, 0), int(f.R * 1.25) + 4096)
        return f

    def fwd_bwd(self, key, params, views, bucket, go, accumulate=False, specs=None):
        """Forward + backward of `views` (<= 16 per call); gradients written into the bucket (the
        first call of an iteration overwrites it: no zero fill)."""
        stats = []
        for c0 in range(0, len(views), self.V.MAX_VIEWS):
            chunk = views[c0:c0 + self.V.MAX_VIEWS]
            f = self.forward((key, c0), params, chunk, None if specs is None else specs[c0:c0 + self.V.MAX_VIEWS])
            self.V.backward_views(f, [v["grads"] for v in chunk], grad_out=go, accumulate=accumulate or c0 > 0)
            stats.append(f)
        return stats

    def render(self, key, params, views, specs=None):
        return self.forward(key, params, views, specs)


class Reference:
    """The unmodified reference kernels, one view per call, on the default stream, gradients of the
    views summed the way autograd's AccumulateGrad does (the first is kept, the others are added)."""
    name = "reference"

    def __init__(self, dev):
        from oracle import ref_driver
        self.mod, self.dev = ref_driver.RefModule, dev
        self.empty = torch.Tensor([])
        self.zeros = {}
        self.acc = None

    def make_bucket(self, params):
        return None, None

    def _z(self, v):
        k = (v["H"], v["W"])
        if k not in self.zeros:
            self.zeros[k] = (torch.zeros((3, v["H"], v["W"]), device=self.dev), torch.zeros((1, v["H"], v["W"]), device=self.dev))
        return self.zeros[k]

    def fwd_bwd(self, key, params, views, bucket, go, accumulate=False, specs=None):
        acc, stats = None, []
        for v in views:
            z3, z1 = self._z(v)
            f = self.mod.rasterize_gaussians(*fwd_args(params, v, self.empty))
            b = self.mod.rasterize_gaussians_backward(*bwd_args(params, v, f, self.empty, z3, z1))
            g = dict(means3D=b[4], opacities=b[3], shs=b[6], shs_p=b[7], scales=b[8], rotations=b[9])
            if acc is None:
                acc = g
            else:
                for k in PARAM_NAMES:
                    acc[k] += g[k]
            stats.append((f[0], f[11]))
        self.acc = acc
        return stats

    def allreduce(self, params):
        if self.acc is None:       # a rank without views contributes zeros
            self.acc = {k: torch.zeros_like(params[k]) for k in PARAM_NAMES}
        for k in PARAM_NAMES:
            dist.all_reduce(self.acc[k])
        self.acc = None

    def render(self, key, params, views, specs=None):
        return [self.mod.rasterize_gaussians(*fwd_args(params, v, self.empty)) for v in views]


def view_stats(arm, stats, views):
    """(V per view, R per view) of the last step."""
    if arm.name == "ours":
        Vs, Rs = [], []
        from gftorf_b200 import debug
        for f in stats:
            Vs += [int((f.radii[i] > 0).sum().item()) for i in range(len(f.views))]
            dec = debug.decode_views(f.geom, f.binning, f.img, int(f.radii.shape[1]), f.R,
                                     [(int(v.image_width), int(v.image_height)) for v in f.views])
            Rs += [d["num_rendered"] for d in dec]
        return Vs, Rs
    return [int((s[1] > 0).sum().item()) for s in stats], [int(s[0]) for s in stats]


# --------------------------------------------------------------------------------------------
def run_gpu(args, impl):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = numa_pin(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    arm = Ours(dev) if impl == "ours" else Reference(dev)
    wl = WORKLOADS[args.workload]
    # weak scaling: every rank renders a camera pair of the SAME synthetic scene (seed 0), so the
    # per-GPU work is identical at every N and the driver's efficiency measures the exchange, not
    # the luck of the per-rank random clouds (round 1 seeded by rank: the slowest cloud set the pace)
    params, views = build_scene(wl, seed=0, device=dev)
    P = wl["P"]
    npix = sum(v["W"] * v["H"] for v in views)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    timer = Timer(dev, world, flush)
    bucket, go = arm.make_bucket(params)

    def step_resident():
        stats = arm.fwd_bwd("main", params, views, bucket, go)
        if world > 1:
            bucket.allreduce() if impl == "ours" else arm.allreduce(params)
        return stats

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local, every=max(1, K // 6)) if not args.no_clocks else None   # every rank, same steps
    n0 = [0]

    def mark():               # launch counting starts after the warm-up (smem attribute calls etc.)
        n0[0] = arm.capi.launch_count() if impl == "ours" else 0
    total_ms, stats = timer.run(step_resident, K, W, sampler=sampler, after_warmup=mark)
    launches = (arm.capi.launch_count() - n0[0]) if impl == "ours" else None
    step_ms = list(timer.steps_ms)
    slowest = slowest_step(timer)
    per_rank = [round(t / K, 4) for t in timer.rank_totals_ms]
    per_rank_med = [min_med_max(r)[1] for r in timer.rank_steps_ms]
    if rank == 0 and world > 1 and os.environ.get("GFT_BENCH_TRACE_RANKS"):
        for i in range(K):
            print("step", i, " ".join(f"{r[i]:.3f}" for r in timer.rank_steps_ms), file=sys.stderr)
    wall_ms = timer.wall_ms
    clocks = sampler.result() if sampler is not None else None

    # per-kernel durations: a second pass of the same steps with the library's stage events on
    stages, stage_steps = [], max(10, K // 4)
    if impl == "ours":
        arm.capi.profile_read(8192)
        arm.capi.profile_enable(True)
        for _ in range(stage_steps):
            flush.zero_()
            step_resident()
        torch.cuda.synchronize()
        stages = arm.capi.profile_read(8192)
        arm.capi.profile_enable(False)

    if args.resident_only:
        if rank == 0:
            share = {}
            for name, ms in stages:
                share[name] = share.get(name, 0.0) + ms / stage_steps
            print(json.dumps({"resident_only": True, "impl": impl, "ms_per_step": round(total_ms / K, 4),
                              "step_ms_min_med_max": min_med_max(step_ms),
                              "ms_per_step_by_rank": per_rank, "median_step_ms_by_rank": per_rank_med,
                              "exchange": getattr(bucket, "mode", None) if world > 1 else None,
                              "stage_ms_per_step": {k: round(v, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])},
                              "gpu_launches": launches}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    e2e = e2e_bench(arm, impl, params, views, dev, timer, K, W, world)

    # render-only throughput (forward only, all outputs) — BASELINE's second metric
    render_ms, _ = timer.run(lambda: arm.render("render", params, views), K, W)

    Vs, Rs = view_stats(arm, stats, views)
    del stats
    blocks = {}
    want = set(args.blocks.split(",")) if args.blocks not in ("all", "none") else None

    def on(name):
        return args.blocks == "all" or (want is not None and name in want)
    if on("c1"):
        blocks["c1"] = small_scene_block(arm, impl, dev, timer, rank, world)
    if on("c4"):
        blocks["c4_batch8"] = c4_block(arm, impl, dev, timer, rank, world)
    if on("c5"):
        blocks["c5_sweep"] = c5_block(arm, impl, dev, timer, rank, world, args.c5_gaussians)
    if world == 1 and on("c3"):
        blocks["c3_iter"] = c3_iter_block(impl, dev, flush)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    mpix = npix * world / 1e6
    value = mpix * K / (total_ms / 1e3)
    line = {
        "metric": "rasterizer fwd+bwd throughput (RGB+ToF+depth)", "value": round(value, 3),
        "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(total_ms / K, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}; trained-like cloud (SURVEY §8d), "
                               "one camera pair per rank per step (the same scene on every rank), fwd colour + fwd ToF + bwd both"
                               + ("; + NCCL allreduce of the per-Gaussian gradients" if world > 1 else ""),
                   "P": P, "views_per_step_per_rank": len(views),
                   "l2": "flushed between timed steps (256 MiB memset, outside the per-step events)",
                   "timing": "sum of per-step CUDA-event durations, max over ranks"},
        "call_shape": ("one batched call for the views of a step (gft_forward_views / gft_backward_views)"
                       if impl == "ours" else "one call per view, default stream (the reference's call shape)"),
        "fwd_bwd_ms_per_iter": round(total_ms / K, 4),
        "step_ms_min_med_max": min_med_max(step_ms),
        "slowest_step": slowest,
        "ms_per_step_by_rank": per_rank,
        "median_step_ms_by_rank": per_rank_med,
        "wall_ms_total_incl_flush": round(wall_ms, 2),
        "render_mpix_s": round(mpix * K / (render_ms / 1e3), 3),
        "render_ms_per_step": round(render_ms / K, 4),
        "e2e": e2e,
        "clocks": clocks,
        "numa_node": numa_node,
        "views": {"V": Vs, "R": Rs},
        "blocks": blocks,
    }
    if impl == "ours" and world > 1:
        line["exchange"] = {"mode": bucket.mode, "autotune_ms": bucket.tuning, "bytes": bucket.nbytes()}
    Ns = [v["W"] * v["H"] for v in views]
    Ts = [((v["W"] + 15) // 16) * ((v["H"] + 15) // 16) for v in views]
    if impl == "ours":
        line["gpu_launches"] = int(launches)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        per = {}
        for name, ms in stages:
            per.setdefault(name, []).append(ms)
        mean_ms = {k: float(np.mean(v)) for k, v in per.items()}
        calls_per_step = {k: len(v) / stage_steps for k, v in per.items()}
        step_share = {k: mean_ms[k] * calls_per_step[k] for k in mean_ms}
        dom = max(step_share, key=step_share.get) if step_share else None
        if dom:
            byts = float(stage_bytes(dom, P, sum(Vs), sum(Rs), sum(Ns), sum(Ts), nviews=len(views)))
            ach = byts / (mean_ms[dom] * 1e-3) / 1e9
            traffic = issue_busy = src = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = tj.get(dom)
                issue_busy = tj.get("_issue_slots_busy_per_active_cycle", {}).get(dom)
                src = "profiles/traffic.json <- " + str(tj.get("_source"))
            except Exception:
                pass
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 2),
                                "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 5),
                                "traffic": traffic, "issue_slots_busy": issue_busy,
                                "traffic_source": src,     # ncu capture committed under profiles/, not measured in this run
                                "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s",
                                "algorithmic_bytes_per_launch": int(byts),
                                "kernel_ms_per_launch": round(mean_ms[dom], 4),
                                "views_per_launch": len(views),
                                "note": "the blend kernels are issue/latency bound by construction "
                                        "(per-pair ALU + MUFU + shuffles), so their HBM fraction is small; "
                                        "see profiles/ for issue-slot utilisation"}
            total_bytes = step_bytes(P, Vs, Rs, Ns, Ts)
            line["step_roofline"] = {"algorithmic_bytes_per_step": int(total_bytes),
                                     "achieved_gbs": round(total_bytes / (total_ms / K * 1e-3) / 1e9, 2),
                                     "frac": round(total_bytes / (total_ms / K * 1e-3) / 1e9 / peak, 5)}
            line["stage_ms_per_step"] = {k: round(v, 4) for k, v in sorted(step_share.items(), key=lambda kv: -kv[1])}
            line["stage_hbm_frac"] = {
                k: round(stage_bytes(k, P, sum(Vs), sum(Rs), sum(Ns), sum(Ts), nviews=len(views)) / (mean_ms[k] * 1e-3) / 1e9 / peak, 4)
                for k in mean_ms if stage_bytes(k, P, 1, 1, 1, 1)}
        if world == 1:
            if on("next"):
                line["next_rows"] = next_rows_bench(P, views[0]["H"], views[0]["W"], dev, flush, peak)
            line["cpu_baseline"] = cpu_baseline(args, wl, bound_s=12.0)
            if on("cpu"):
                line["blocks"]["cpu_pipeline"] = cpu_pipeline_block(bound_s=15.0)
    else:
        line["impl"] = "reference"
        line["cpu_baseline"] = {"value": line["value"], "unit": "Mpix/s", "cores": 0, "kind": "reference",
                                "sample": "full workload; the reference's own implementation of this path is "
                                          "CUDA-only (no CPU rasterizer exists, SURVEY §8c), so this arm runs its "
                                          "unmodified kernels (oracle/_ref) on the same GPU"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
def e2e_bench(arm, impl, params, views, dev, timer, K, W, world):
    """Host buffers in, host buffers out.  Ours: the public batched autograd surface, the copies of
    neighbouring steps overlapping the kernels (upload stream / compute stream / download stream,
    triple-buffered device and host staging).  Reference: its stock path — sequential, default
    stream (its kernels take no stream argument).

    At N > 1 ours runs the data-parallel step the way a host-resident (offloaded) optimizer would
    drive it: the replicated parameters reach the GPUs ONCE — rank r uploads slice r of the flat
    parameter buffer and an NVLink all-gather completes it on every GPU — and the summed gradients
    leave ONCE — a reduce-scatter over NVLink, rank r downloads slice r.  Every rank still uploads
    its own views' data and downloads its own images.  (The eight ranks of a box share its PCIe
    switches: eight full copies of the same 109 MB each way took 16 ms per step, copies alone.)"""
    names = PARAM_NAMES
    rank = dist.get_rank() if world > 1 else 0
    sharded = impl == "ours" and world > 1
    host_params = {k: params[k].cpu().pin_memory() for k in names}
    host_views = []
    for v in views:
        hv = {k: v[k].cpu().pin_memory() for k in ("viewmatrix", "projmatrix", "campos")}
        hv["grads"] = {k: g.cpu().pin_memory() for k, g in v["grads"].items()}
        host_views.append(hv)
    host_bg = views[0]["bg"].cpu().pin_memory()
    h2d_bytes = sum(t.numel() * 4 for t in host_params.values()) + host_bg.numel() * 4 + sum(
        sum(t.numel() * 4 for t in (hv["viewmatrix"], hv["projmatrix"], hv["campos"])) +
        sum(g.numel() * 4 for g in hv["grads"].values()) for hv in host_views)
    NB = 3 if impl == "ours" else 1      # staging sets in flight: the host enqueues up to two steps ahead of the link
    host_imgs = [[torch.empty((11, v["H"], v["W"]), dtype=torch.float32).pin_memory() for v in views] for _ in range(NB)]
    npix_all = sum(v["W"] * v["H"] for v in views) * world
    n_par = sum(host_params[k].numel() for k in names)
    if sharded:
        # flat parameter / gradient layout in PARAM_NAMES order, padded so that it splits evenly
        n_pad = (n_par + 4 * world - 1) // (4 * world) * (4 * world)
        shard = n_pad // world
        host_flat = torch.zeros(n_pad, dtype=torch.float32).pin_memory()
        o = 0
        for k in names:
            host_flat[o:o + host_params[k].numel()].copy_(host_params[k].reshape(-1))
            o += host_params[k].numel()
        host_in_shard = host_flat[rank * shard:(rank + 1) * shard]
        host_out_shard = [torch.empty(shard, dtype=torch.float32).pin_memory() for _ in range(NB)]
        g_up, g_down = dist.new_group(), dist.new_group()      # own communicators: no queueing behind each other
        h2d_bytes += (shard - n_par) * 4
        d2h_bytes = shard * 4 + sum(t.numel() * 4 for t in host_imgs[0])
    else:
        host_out_grads = [{k: torch.empty_like(host_params[k]).pin_memory() for k in names} for _ in range(NB)]
        d2h_bytes = sum(t.numel() * 4 for t in host_out_grads[0].values()) + sum(t.numel() * 4 for t in host_imgs[0])

    # persistent device-side staging (what an application keeps): the H2D copies land in these
    def staging():
        if sharded:
            flat = torch.empty(n_pad, dtype=torch.float32, device=dev)
            prm, o = {}, 0
            for k in names:
                prm[k] = flat[o:o + params[k].numel()].view(params[k].shape)
                o += params[k].numel()
            s = {"flat": flat, "params": prm}
        else:
            s = {"params": {k: torch.empty_like(params[k]) for k in names}}
        s.update({"bg": torch.empty_like(views[0]["bg"]), "views": []})
        for v in views:
            dv = {k: torch.empty_like(v[k]) for k in ("viewmatrix", "projmatrix", "campos")}
            dv["grads"] = {k: torch.empty_like(g) for k, g in v["grads"].items()}
            s["views"].append(dv)
        return s
    stage = [staging() for _ in range(NB)]

    def upload(s):
        with torch.no_grad():
            if sharded:
                mine = s["flat"][rank * shard:(rank + 1) * shard]
                mine.copy_(host_in_shard, non_blocking=True)
                dist.all_gather_into_tensor(s["flat"], mine, group=g_up)      # in place, over NVLink
            else:
                for k in names:
                    s["params"][k].copy_(host_params[k], non_blocking=True)
            s["bg"].copy_(host_bg, non_blocking=True)
            for hv, dv in zip(host_views, s["views"]):
                for k in ("viewmatrix", "projmatrix", "campos"):
                    dv[k].copy_(hv[k], non_blocking=True)
                for k, t in hv["grads"].items():
                    dv["grads"][k].copy_(t, non_blocking=True)

    if impl != "ours":
        empty = torch.Tensor([])
        s = stage[0]

        def step_ref():
            upload(s)
            acc = None
            for v, dv, himg in zip(views, s["views"], host_imgs[0]):
                vv = dict(v, viewmatrix=dv["viewmatrix"], projmatrix=dv["projmatrix"], campos=dv["campos"],
                          bg=s["bg"], grads=dv["grads"])
                z3, z1 = arm._z(v)
                f = arm.mod.rasterize_gaussians(*fwd_args(s["params"], vv, empty))
                b = arm.mod.rasterize_gaussians_backward(*bwd_args(s["params"], vv, f, empty, z3, z1))
                g = dict(means3D=b[4], opacities=b[3], shs=b[6], shs_p=b[7], scales=b[8], rotations=b[9])
                if acc is None:
                    acc = g
                else:
                    for k in names:
                        acc[k] += g[k]
                himg[0:3].copy_(f[1], non_blocking=True)
                himg[3:10].copy_(f[2], non_blocking=True)
                himg[10:11].copy_(f[3], non_blocking=True)
            for k in names:
                host_out_grads[0][k].copy_(acc[k], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        ms, _ = timer.run(step_ref, K, W)
        return {"value": round(npix_all / 1e6 * K / (ms / 1e3), 3), "unit": "Mpix/s", "ms_per_step": round(ms / K, 4),
                "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "api": "reference kernels via oracle/ref_driver.py (binding restated), sequential, default stream"}

    from gftorf_b200 import views as V
    if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    comp = torch.cuda.current_stream(dev)
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    P = params["means3D"].shape[0]
    for s in stage:
        for k in names:
            s["params"][k].requires_grad_(True)
        s["m2d"] = torch.zeros((len(views), P, 3), device=dev, requires_grad=True)
    ev_up = [torch.cuda.Event() for _ in range(NB)]
    ev_comp = [torch.cuda.Event() for _ in range(NB)]
    ev_down = [torch.cuda.Event() for _ in range(NB)]

    trace = [] if os.environ.get("GFT_E2E_TRACE") else None

    def mark(stream, tag, i):
        if trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            trace.append((tag, i, e, time.perf_counter()))

    def prefetch(i):
        """Enqueue the upload of step i's inputs (pipelined mode)."""
        b = i % NB
        up.wait_event(ev_comp[b])          # the staging set is free once its previous compute is done
        with torch.cuda.stream(up):
            mark(up, "U0", i)
            upload(stage[b])
            ev_up[b].record(up)
            mark(up, "U1", i)

    def one(i, pipelined, more=True):
        b = i % NB if pipelined else 0
        s = stage[b]
        if pipelined:
            # the forward call blocks the host until the GPU has counted the tile instances, so the
            # NEXT step's upload is enqueued first: the link never waits for the host
            if more:
                prefetch(i + 1)
            comp.wait_event(ev_up[b])
            mark(comp, "C0", i)
        else:
            upload(s)
        for k in names:
            s["params"][k].grad = None
        s["m2d"].grad = None
        specs = [V.ViewSpec(v["H"], v["W"], v["tanfovx"], v["tanfovy"], s["bg"], dv["viewmatrix"], dv["projmatrix"],
                            dv["campos"], v["near_n"], v["far_n"], v["depth_range"]) for v, dv in zip(views, s["views"])]
        outs = V.rasterize_views(s["params"]["means3D"], s["m2d"], s["params"]["opacities"], s["params"]["shs"],
                                 s["params"]["shs_p"], s["params"]["scales"], s["params"]["rotations"], specs, 3)
        tens, gr = [], []
        for o, dv in zip(outs, s["views"]):
            g = dv["grads"]
            tens += [o[0], o[1], o[2], o[4], o[6]]
            gr += [g["color"], g["phasor"], g["depth"], g["acc"], g["depth_distortion"]]
        if pipelined:
            mark(comp, "Cf", i)
        torch.autograd.backward(tens, gr)
        if sharded:
            with torch.no_grad():
                pad = [torch.zeros(n_pad - n_par, dtype=torch.float32, device=dev)] if n_pad > n_par else []
                s["gflat"] = torch.cat([s["params"][k].grad.reshape(-1) for k in names] + pad)
        ev_comp[b].record(comp)
        tgt = down if pipelined else comp
        if pipelined:
            mark(comp, "C1", i)
            ev_down[b].synchronize()           # host: the previous results in this host set have landed
            down.wait_event(ev_comp[b])
            mark(down, "D0", i)
        with torch.cuda.stream(tgt), torch.no_grad():
            for o, himg in zip(outs, host_imgs[b]):
                himg[0:3].copy_(o[0], non_blocking=True)
                himg[3:10].copy_(o[1], non_blocking=True)
                himg[10:11].copy_(o[2], non_blocking=True)
                for t in (o[0], o[1], o[2]):
                    t.record_stream(tgt)
            if sharded:
                gflat, gsh = s["gflat"], torch.empty(shard, dtype=torch.float32, device=dev)
                dist.reduce_scatter_tensor(gsh, gflat, group=g_down)          # summed over the ranks, slice r stays here
                host_out_shard[b].copy_(gsh, non_blocking=True)
                gflat.record_stream(tgt)
            else:
                for k in names:
                    gk = s["params"][k].grad
                    host_out_grads[b][k].copy_(gk, non_blocking=True)
                    gk.record_stream(tgt)
            ev_down[b].record(tgt)
            if pipelined:
                mark(tgt, "D1", i)
        if not pipelined:
            comp.synchronize()

    # what the host link allows: the step's copies alone, each direction by itself and both at once
    dimgs = [torch.empty((11, v["H"], v["W"]), dtype=torch.float32, device=dev) for v in views]

    def copies(do_up, do_down, n=8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record(comp)
        up.wait_event(e0)
        down.wait_event(e0)
        for _ in range(n):
            if do_up:
                with torch.cuda.stream(up):
                    upload(stage[0])
            if do_down:
                with torch.cuda.stream(down), torch.no_grad():
                    for d, himg in zip(dimgs, host_imgs[0]):
                        himg.copy_(d, non_blocking=True)
                    if sharded:
                        host_out_shard[0].copy_(stage[0]["flat"][rank * shard:(rank + 1) * shard], non_blocking=True)
                    else:
                        for k in names:
                            host_out_grads[0][k].copy_(stage[0]["params"][k].detach(), non_blocking=True)
        comp.wait_stream(up)
        comp.wait_stream(down)
        e1.record(comp)
        torch.cuda.synchronize(dev)
        return round(e0.elapsed_time(e1) / n, 4)
    copies(True, True, 2)
    link = {"h2d_ms": copies(True, False), "d2h_ms": copies(False, True), "both_ms": copies(True, True)}

    # serial: upload -> compute -> download, one step at a time (round 1's e2e shape)
    it = [0]

    def step_serial():
        one(it[0], False)
        it[0] += 1
    Ks = max(5, K // 4)
    serial_ms, _ = timer.run(step_serial, Ks, 3)
    serial_ms /= Ks
    shard_check = None
    if sharded:
        # the sharded transfers carry what the full copies carried: every GPU holds the whole
        # parameter set after the all-gather, and slice r of the reduce-scatter is the sum over the
        # ranks (all ranks render the same scene here: N x the local gradient)
        torch.cuda.synchronize(dev)
        s0 = stage[0]
        ok_p = bool(torch.equal(s0["flat"][:n_par], torch.cat([params[k].reshape(-1) for k in names])))
        want = world * s0["gflat"][rank * shard:(rank + 1) * shard].double().cpu()
        got = host_out_shard[0].double()
        err = float((got - want).norm() / max(float(want.norm()), 1e-30))
        shard_check = {"params_identical_after_all_gather": ok_p, "grad_slice_rel_err_vs_N_x_local": round(err, 8)}
        if not ok_p or err > 1e-3:
            raise RuntimeError(f"sharded host I/O check failed: {shard_check}")

    # pipelined: whole region timed with one event pair on the compute stream, every copy inside
    for e in ev_comp + ev_down:
        e.record(comp)
    gc.collect()
    gc.disable()
    prefetch(0)
    for i in range(W):
        one(i, True, more=i + 1 < W)
    comp.wait_stream(down)
    comp.wait_stream(up)
    timer.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(comp)
    up.wait_event(e0)                      # every copy of the K timed steps lies between the two events
    prefetch(W)
    for i in range(K):
        one(W + i, True, more=i + 1 < K)
    comp.wait_stream(down)
    comp.wait_stream(up)
    e1.record(comp)
    timer.barrier()
    gc.enable()
    tot = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms = float(tot.item())
    if trace is not None:      # GPU and host timeline of the timed steps (ms since the first event), for diagnosis
        t_host0 = None
        for tag, i, e, th in trace:
            if i < W or i > W + 6:
                continue
            t_host0 = th if t_host0 is None else t_host0
            print(f"e2e-trace step {i - W} {tag}: gpu {e0.elapsed_time(e):8.3f} ms   host-enqueue {(th - t_host0) * 1e3:8.3f} ms",
                  file=sys.stderr)
    return {"value": round(npix_all / 1e6 * K / (ms / 1e3), 3), "unit": "Mpix/s", "ms_per_step": round(ms / K, 4),
            "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
            "serial_ms_per_step": round(serial_ms, 4),
            "host_link_copies_only": dict(link, note="the same H2D / D2H copies of one step with no kernels between them: "
                                                     "each direction alone, then both concurrently (the floor of the pipelined step)"),
            "l2": "every step's inputs arrive from the host or over NVLink (150 MB) and its results leave (136 MB): "
                  "the working set of a step exceeds the 126 MB L2; no extra flush in the pipelined region",
            "api": "gftorf_b200.rasterize_views(...) + torch.autograd.backward; H2D of step i+1 and D2H of step i-1 "
                   "overlap the kernels of step i (3 streams, triple-buffered pinned and device staging; the upload of step i+1 is "
                   "enqueued before the forward call of step i, which blocks the host while it learns the instance count); "
                   "serial_ms_per_step = the same calls with upload -> compute -> download one after the other"
                   + ("; N > 1: data-parallel step with host I/O sharded over the ranks — rank r uploads slice r of the "
                      "replicated parameters (NVLink all-gather completes them on every GPU) and downloads slice r of the "
                      "summed gradients (NVLink reduce-scatter), plus its own views' inputs and images; the byte counts are "
                      "per rank" if sharded else ""),
            "sharded_host_io": bool(sharded), "sharded_host_io_check": shard_check}


# --------------------------------------------------------------------------------------------
def small_scene_block(arm, impl, dev, timer, rank, world):
    """BASELINE configs[0] shape on the GPU: 20k Gaussians, 320x240 x2 — launch-bound."""
    wl = WORKLOADS["c1"]
    params, views = build_scene(wl, seed=0, device=dev)
    bucket, go = arm.make_bucket(params)
    ms, _ = timer.run(lambda: arm.fwd_bwd("c1", params, views, bucket, go), 30, 5)
    return {"workload": wl["desc"], "ms_per_step": round(ms / 30, 4), "step_ms_min_med_max": min_med_max(timer.steps_ms)}


def c4_block(arm, impl, dev, timer, rank, world):
    """BASELINE configs[3]: a multi-view batch of 8 cameras per iteration (train.py:158 renders one
    camera, gaussian_renderer/__init__.py:107-128 its colour + ToF view), 2 M Gaussians, 1080p RGB +
    640x480 ToF, cameras sharded over the ranks (parallel.shard_views), one allreduce of the
    per-Gaussian gradients per iteration.  STRONG scaling: the 8-camera batch is fixed."""
    from gftorf_b200 import parallel
    wl = WORKLOADS["c4"]
    params, cams = build_cameras(wl, 8, seed=0, device=dev)        # the same batch on every rank
    mine = parallel.shard_views(8, rank, world)
    views = [v for c in mine for v in cams[c]]
    bucket, go = arm.make_bucket(params)
    specs = [view_spec(v) for v in views] if impl == "ours" else None

    def step():
        stats = arm.fwd_bwd("c4", params, views, bucket, go, specs=specs) if views else []
        if not views and impl == "ours":
            bucket.zero()
        if world > 1:
            bucket.allreduce() if impl == "ours" else arm.allreduce(params)
        return stats
    K = 6 if impl == "ours" else 4
    ms, stats = timer.run(step, K, 3)
    out = {"workload": "BASELINE configs[3]: 8 cameras x (1920x1080 colour + 640x480 ToF), 2M Gaussians, "
                       f"{len(mine)} camera(s) on this rank, fwd+bwd of all views + gradient allreduce",
           "scaling": "strong", "cameras_per_iter": 8, "n_gpus": world,
           "ms_per_iter": round(ms / K, 3), "ms_per_iter_median": min_med_max(timer.steps_ms)[1],
           "iter_ms_min_med_max": min_med_max(timer.steps_ms), "slowest_iter": slowest_step(timer),
           "mpix_s": round(8 * (1920 * 1080 + 640 * 480) / 1e6 * K / (ms / 1e3), 2)}
    if rank == 0 and stats:
        Vs, Rs = view_stats(arm, stats, views)
        out["rank0_views"] = {"V": Vs[:2], "R": Rs[:2], "R_total": int(sum(Rs))}
    if impl == "ours" and world > 1:
        out["exchange"] = {"mode": bucket.mode, "autotune_ms": bucket.tuning, "bytes": bucket.nbytes()}
    del params, cams, views, bucket, stats
    torch.cuda.empty_cache()
    return out


def c5_block(arm, impl, dev, timer, rank, world, sizes):
    """BASELINE configs[4]: render-only sweep, 120 distinct poses of a trajectory at 1920x1080,
    frames r, r+n, ... on rank r (parallel.shard_frames, no collective), forward only, all outputs
    (render.py:95-209 renders one frame per call; ours batches 4 frames per call)."""
    from gftorf_b200 import parallel
    out = {}
    for P in sizes:
        wl = dict(WORKLOADS["c5"], P=P, sigma_px=1.0 if P >= 4000000 else 1.5)
        params, cams = build_cameras(wl, 120, seed=1, device=dev, with_tof=False)
        frames = [cams[i][0] for i in parallel.shard_frames(120, rank, world)]
        specs = [view_spec(v) for v in frames] if impl == "ours" else None
        B = 4

        def sweep(n=None):
            fr = frames if n is None else frames[:n]
            for c0 in range(0, len(fr), B):
                # one size hint for the whole sweep: the frames of a trajectory have similar instance counts
                arm.render(("c5", P), params, fr[c0:c0 + B], None if specs is None else specs[c0:c0 + B])
        sweep()                        # warm-up over the whole trajectory: allocator, largest instance count
        timer.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sweep()
        e1.record()
        timer.barrier()
        tot = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        ms = float(tot.item())
        out[f"P{P // 1000000}M"] = {"frames": 120, "frames_this_rank": len(frames), "ms_total": round(ms, 2),
                                    "ms_per_frame_per_gpu": round(ms / max(1, len(frames)), 3),
                                    "mpix_s": round(120 * 1920 * 1080 / 1e6 / (ms / 1e3), 1)}
        del params, cams, frames
        torch.cuda.empty_cache()
    out["workload"] = "BASELINE configs[4]: 120-frame trajectory at 1920x1080, forward only, frames sharded over the ranks"
    out["scaling"] = "strong"
    return out


def c3_iter_block(impl, dev, flush, iters=18, densify_every=6):
    """BASELINE configs[2]: a FULL training iteration incl. densification as one timed unit
    (train.py:165-279,441-474): assemble the rasterizer inputs from the raw parameters -> colour +
    ToF view -> losses (l1 + SSIM on colour, weighted L2 on a quad) -> backward -> Adam; every
    `densify_every` iterations densify_and_prune + optimizer-state surgery.  Ours: the fused
    operators around the batched rasterizer.  Reference arm: the reference's PyTorch operator
    chain (oracle/train_oracle.py restates it) + torch.optim.Adam around the reference kernels."""
    import test_train_loop as tl
    from oracle import train_oracle as orc
    wl = WORKLOADS["c3"]
    params, views = build_scene(wl, seed=0, device=dev)
    inp = dict(means3D=params["means3D"], opacities=params["opacities"], scales=params["scales"],
               rotations=params["rotations"], shs=params["shs"], shs_p=params["shs_p"])
    model0 = tl.initial_model(inp)
    names = list(tl.LRS)
    g = torch.Generator(device=dev).manual_seed(1)
    cv, tv = views
    gt_color = torch.rand(3, cv["H"], cv["W"], device=dev, generator=g)
    gt_quad = torch.rand(1, tv["H"], tv["W"], device=dev, generator=g) * 0.05
    kw = dict(max_grad=2e-4, min_opacity=0.005, extent=5.0, percent_dense=0.01)
    times, sizes = [], []

    if impl == "ours":
        from gftorf_b200 import train_ops as T, views as V
        specs = [view_spec(cv), view_spec(tv)]
        fa = T.FlatAdam([(n, model0[n], tl.LRS[n]) for n in names])
        for it in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            raw = {tl.RAW_OF[n]: fa.params[n] for n in names}
            asm = T.assemble_gaussians(raw)
            Pn = asm["means3D"].shape[0]
            m2d = torch.zeros((2, Pn, 3), device=dev, requires_grad=True)
            oc, ot = V.rasterize_views(asm["means3D"], m2d, asm["opacities"], asm["shs"], asm["shs_p"],
                                       asm["scales"], asm["rotations"], specs, 3)
            total = torch.zeros(1, device=dev)
            _, g_img = T.fused_loss(oc[0].detach(), gt_color, "l1", 1.0, 0.2, loss_out=total)
            quad = ot[1][5:6]
            _, g_quad = T.fused_loss(quad.detach(), gt_quad, "weighted_l2_quad", 0.5, 0.2, w=0.01, loss_out=total)
            torch.autograd.backward([oc[0], quad], [g_img, g_quad])
            fa.step(zero_grad=True)
            if (it + 1) % densify_every == 0:
                # densification statistics (scene/gaussian_model.py:648-654) from the colour view
                acc = torch.norm(m2d.grad[0][:, :2], dim=-1, keepdim=True)
                den = (oc[10] > 0).float().unsqueeze(1)
                P0 = fa.params["xyz"].shape[0]
                prm = {n: fa.params[n].detach() for n in names}
                prm["f_seg_color"] = torch.zeros(P0, 1, device=dev)
                ea = {n: fa.exp_avg[fa.bounds[n][0]:fa.bounds[n][1]].view(fa.bounds[n][2]) for n in names}
                es = {n: fa.exp_avg_sq[fa.bounds[n][0]:fa.bounds[n][1]].view(fa.bounds[n][2]) for n in names}
                ea["f_seg_color"] = es["f_seg_color"] = torch.zeros(P0, 1, device=dev)
                p_new, m_new, v_new, info = T.densify_and_prune(prm, ea, es, acc, den,
                                                                generator=torch.Generator(device=dev).manual_seed(3), **kw)
                steps_done = fa.t if hasattr(fa, "t") else None
                fa = T.FlatAdam([(n, p_new[n], tl.LRS[n]) for n in names])
                for n in names:
                    lo, hi, shp = fa.bounds[n]
                    fa.exp_avg[lo:hi].view(shp).copy_(m_new[n])
                    fa.exp_avg_sq[lo:hi].view(shp).copy_(v_new[n])
                if steps_done is not None:
                    fa.t = steps_done
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
            sizes.append(int(fa.params["xyz"].shape[0]))
    else:
        from oracle import ref_driver
        fn = ref_driver.rasterize_autograd
        pb = {n: model0[n].clone().requires_grad_(True) for n in names}
        opt = torch.optim.Adam([{"params": [pb[n]], "lr": tl.LRS[n], "name": n} for n in names], lr=0.0, eps=1e-15)
        for it in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rawb = {tl.RAW_OF[n]: pb[n] for n in names}
            asm = orc.assemble(rawb)
            m2c = torch.zeros_like(asm["means3D"], requires_grad=True)
            m2t = torch.zeros_like(asm["means3D"], requires_grad=True)
            oc = fn(asm, m2c, cv)
            ot = fn(asm, m2t, tv)
            loss = orc.loss_term(oc[0], gt_color, "l1", 1.0, 0.2) + \
                orc.loss_term(ot[1][5:6], gt_quad, "weighted_l2_quad", 0.5, 0.2, w=0.01)
            opt.zero_grad()
            loss.backward()
            opt.step()
            if (it + 1) % densify_every == 0:
                acc = torch.norm(m2c.grad[:, :2], dim=-1, keepdim=True)
                den = (oc[10] > 0).float().unsqueeze(1)
                P0 = pb["xyz"].shape[0]
                prm = {n: pb[n].detach() for n in names}
                prm["f_seg_color"] = torch.zeros(P0, 1, device=dev)
                st = {n: opt.state[pb[n]] for n in names}
                ea = {n: st[n]["exp_avg"] for n in names}
                es = {n: st[n]["exp_avg_sq"] for n in names}
                ea["f_seg_color"] = es["f_seg_color"] = torch.zeros(P0, 1, device=dev)
                z = lambda s: torch.normal(mean=torch.zeros_like(s), std=s, generator=torch.Generator(device=dev).manual_seed(3))
                p_ref, m_ref, v_ref = orc.densify_and_prune(prm, ea, es, acc, den, normal_fn=z, **kw)
                steps = {n: st[n]["step"] for n in names}
                pb = {n: p_ref[n].clone().requires_grad_(True) for n in names}
                opt = torch.optim.Adam([{"params": [pb[n]], "lr": tl.LRS[n], "name": n} for n in names], lr=0.0, eps=1e-15)
                for n in names:
                    opt.state[pb[n]] = {"step": steps[n], "exp_avg": m_ref[n].clone(), "exp_avg_sq": v_ref[n].clone()}
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
            sizes.append(int(pb["xyz"].shape[0]))
    warm = 2
    plain = [t for i, t in enumerate(times) if i >= warm and (i + 1) % densify_every != 0]
    dens = [t for i, t in enumerate(times) if i >= warm and (i + 1) % densify_every == 0]
    return {"workload": "BASELINE configs[2]: 500k Gaussians, 640x480 colour + 320x240 ToF, full training "
                        f"iteration (assemble, 2 views, losses, backward, Adam), densify_and_prune every {densify_every}",
            "ms_per_iter_mean_incl_densify": round(float(np.mean(times[warm:])), 3),
            "ms_per_plain_iter": round(float(np.median(plain)), 3),
            "ms_per_densify_iter": round(float(np.median(dens)), 3) if dens else None,
            "ms_per_densify_iter_each": [round(float(t), 3) for t in dens],   # the first pays one-time costs (lazy module loads, generator set-up)
            "gaussians_start_end": [sizes[0], sizes[-1]], "iters": iters}


# --------------------------------------------------------------------------------------------
def next_rows_bench(P, H, W, dev, flush, peak_gbs, iters=20):
    """SURVEY §8f rows (f1 assembly, f3 loss, f4 Adam) and distCUDA2 at the bench workload's size:
    ms per call of the fused CUDA operator (CUDA events, L2 flushed between calls), its
    algorithmic bytes against the HBM peak, and the reference's own implementation of the same step
    — its PyTorch operator chain (oracle/train_oracle.py restates it line for line) — on the same
    GPU.  Reported beside the headline metric, not part of it."""
    from gftorf_b200 import train_ops as T
    from oracle import train_oracle as orc
    import test_train_ops as tt

    def time_ms(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    out = {}
    # ---- f1: 25 % of the Gaussians dynamic, deformation outputs present ------------------------
    raw = {k: v.to(dev).requires_grad_(True) for k, v in tt.make_raw(P, seed=5).items()}
    mask = torch.zeros(P, dtype=torch.bool, device=dev)
    mask[::4] = True
    Nd = int(mask.sum())
    deltas = {k: v.to(dev).requires_grad_(True) for k, v in tt.make_deltas(Nd, seed=6).items()}
    dyn = T.dyn_index_from_mask(mask)
    gouts = None

    def asm_ours():
        o = T.assemble_gaussians(raw, dyn, deltas)
        nonlocal gouts
        if gouts is None:
            gouts = [torch.randn_like(o[k]) for k in T.OUT_NAMES]
        torch.autograd.backward([o[k] for k in T.OUT_NAMES], gouts, inputs=list(raw.values()) + list(deltas.values()))

    def asm_ref():
        o = orc.assemble(raw, mask, deltas)
        torch.autograd.backward([o[k] for k in T.OUT_NAMES], gouts, inputs=list(raw.values()) + list(deltas.values()))

    def clear():
        for t in list(raw.values()) + list(deltas.values()):
            t.grad = None

    t_o = time_ms(lambda: (clear(), asm_ours()))
    t_r = time_ms(lambda: (clear(), asm_ref()))
    byts = (91 * 4 * 2) * P + 91 * 4 * Nd + (91 * 4 * 2 + 48) * P + 91 * 4 * Nd   # fwd in+out (+deltas), bwd in+saved+out (+delta grads)
    out["f1_assemble_fwd_bwd"] = {"ms": round(t_o, 4), "reference_torch_ms": round(t_r, 4),
                                  "algorithmic_bytes": int(byts), "achieved_gbs": round(byts / t_o / 1e6, 1),
                                  "hbm_frac": round(byts / t_o / 1e6 / peak_gbs, 4),
                                  "workload": f"P={P}, 25% dynamic, sh_degree 3 (incl. autograd bookkeeping)"}
    del raw, deltas, gouts
    # ---- f3: colour loss (l1 + ssim) on the colour view ----------------------------------------
    img = torch.rand(3, H, W, device=dev)
    gt = torch.rand(3, H, W, device=dev)

    def loss_ref():
        leaf = img.detach().requires_grad_(True)
        orc.loss_term(leaf, gt, "l1", 1.0, 0.2).backward()

    t_o = time_ms(lambda: T.fused_loss(img, gt, "l1", 1.0, 0.2))
    t_r = time_ms(loss_ref)
    byts = 11 * 4 * img.numel()
    out["f3_loss_l1_ssim"] = {"ms": round(t_o, 4), "reference_torch_ms": round(t_r, 4),
                              "algorithmic_bytes": int(byts), "achieved_gbs": round(byts / t_o / 1e6, 1),
                              "hbm_frac": round(byts / t_o / 1e6 / peak_gbs, 4),
                              "workload": f"3x{H}x{W}, value + gradient"}
    # ---- f4: the reference's param groups at this P ---------------------------------------------
    shapes = [("xyz", (P, 3), 1.6e-4), ("f_dc_color", (P, 1, 3), 2.5e-3), ("f_rest_color", (P, 15, 3), 1.25e-4),
              ("phase_f_dc", (P, 1, 1), 1e-3), ("phase_f_rest", (P, 15, 1), 5e-5), ("amp_f_dc", (P, 1, 1), 1e-3),
              ("amp_f_rest", (P, 15, 1), 5e-5), ("opacity", (P, 1), 0.05), ("scaling", (P, 3), 5e-3),
              ("rotation", (P, 4), 1e-3), ("f_seg_color", (P, 1), 0.0), ("phase_offset", (1,), 0.0),
              ("dc_offset", (1,), 0.0)]
    init = [torch.randn(sh, device=dev) for _, sh, _ in shapes]
    fa = T.FlatAdam([(n, t, lr) for (n, _, lr), t in zip(shapes, init)])
    fa.grad.normal_()
    tp = [t.clone().requires_grad_(True) for t in init]
    for t in tp:
        t.grad = torch.randn_like(t)
    opt = torch.optim.Adam([{"params": [t], "lr": lr} for t, (_, _, lr) in zip(tp, shapes)], lr=0.0, eps=1e-15)
    t_o = time_ms(lambda: fa.step(zero_grad=False))
    t_r = time_ms(opt.step)
    byts = 28 * fa.flat.numel()
    out["f4_adam_13_groups"] = {"ms": round(t_o, 4), "reference_torch_ms": round(t_r, 4),
                                "algorithmic_bytes": int(byts), "achieved_gbs": round(byts / t_o / 1e6, 1),
                                "hbm_frac": round(byts / t_o / 1e6 / peak_gbs, 4),
                                "workload": f"P={P}, {fa.flat.numel()} parameters, 13 groups (torch.optim.Adam default = foreach)"}
    # ---- a18: distCUDA2 (initialisation) on a uniform cloud of the workload's size ---------------
    from gftorf_b200 import distCUDA2
    pts = torch.rand(P, 3, device=dev) * 4.0 - 2.0
    t_o = time_ms(lambda: distCUDA2(pts))
    entry = {"ms": round(t_o, 4), "algorithmic_bytes": 32 * P, "achieved_gbs": round(32 * P / t_o / 1e6, 1),
             "workload": f"P={P} uniform points"}
    try:
        from oracle import ref_driver
        if ref_driver.available():
            entry["reference_kernels_ms"] = round(time_ms(lambda: ref_driver.distCUDA2(pts)), 4)
    except Exception:          # the reference library did not travel: report ours alone
        entry["reference_kernels_ms"] = None
    out["a18_distCUDA2"] = entry
    return out


# --------------------------------------------------------------------------------------------
def cpu_baseline(args, wl, bound_s=20.0):
    """The CPU port (oracle/gft_oracle.cpp, OpenMP over all host cores) on a bounded sample of the
    workload: the colour view's forward+backward, repeated until ~bound_s seconds are spent."""
    try:
        from oracle import cpu_oracle
        if not cpu_oracle.available():
            return {"value": None, "unit": "Mpix/s", "cores": 0, "kind": "port",
                    "sample": "oracle/libgft_oracle.so not built"}
        params, views = build_scene(wl, seed=0, device="cpu")
        v = views[0]
        empty = torch.Tensor([])
        z3, z1 = torch.zeros_like(v["grads"]["color"]), torch.zeros_like(v["grads"]["depth"])
        times = []
        t_start = time.perf_counter()
        while True:
            t0 = time.perf_counter()
            f = cpu_oracle.rasterize_gaussians(*fwd_args(params, v, empty))
            cpu_oracle.rasterize_gaussians_backward(*bwd_args(params, v, f, empty, z3, z1))
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > bound_s or len(times) >= 10:
                break
        best = float(np.median(times))
        return {"value": round(v["W"] * v["H"] / 1e6 / best, 4), "unit": "Mpix/s",
                "cores": cpu_oracle.num_threads(), "kind": "port",
                "ms_per_view_fwd_bwd": round(best * 1e3, 1),
                "sample": f"colour view only ({v['W']}x{v['H']}, P={wl['P']}), fwd+bwd, "
                          f"median of {len(times)} runs, OpenMP threads = host cores; the §8(d) pipeline "
                          "baseline (deform MLP + losses + Adam around this rasterizer, config C1) is "
                          "blocks.cpu_pipeline"}
    except Exception as ex:  # never take the GPU line down with the baseline
        return {"value": None, "unit": "Mpix/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}


def cpu_pipeline_block(bound_s=15.0):
    """SURVEY §8(d) CPU baseline, config C1 (20k Gaussians, 320x240 colour + ToF): the reference
    PIPELINE on the host cores — deform MLP (utils/time_utils.py:56-127) on the dynamic quarter of
    the Gaussians, assembly, the faithful CPU transcription of the rasterizer forward/backward
    (oracle/gft_oracle.cpp, OpenMP) for both views, the losses (utils/loss_utils.py) and Adam, in
    PyTorch CPU with all host threads.  A reported baseline, not an optimisation target."""
    try:
        from oracle import cpu_oracle, train_oracle as orc
        import test_train_loop as tl
        if not cpu_oracle.available():
            return {"value": None, "sample": "oracle/libgft_oracle.so not built"}
        n_threads = os.cpu_count() or 1
        torch.set_num_threads(n_threads)
        wl = WORKLOADS["c1"]
        params, views = build_scene(wl, seed=0, device="cpu")
        model0 = tl.initial_model(params)
        names = list(tl.LRS)
        P = wl["P"]
        mask = torch.zeros(P, dtype=torch.bool)
        mask[::4] = True
        deform = orc.DeformNetwork()
        pb = {n: model0[n].clone().requires_grad_(True) for n in names}
        opt = torch.optim.Adam([{"params": [pb[n]], "lr": tl.LRS[n]} for n in names] +
                               [{"params": list(deform.parameters()), "lr": 8e-4}], lr=0.0, eps=1e-15)
        cv, tv = views
        g = torch.Generator().manual_seed(1)
        gt_color = torch.rand(3, cv["H"], cv["W"], generator=g)
        gt_quad = torch.rand(1, tv["H"], tv["W"], generator=g) * 0.05
        times, t_start = [], time.perf_counter()
        while True:
            t0 = time.perf_counter()
            rawb = {tl.RAW_OF[n]: pb[n] for n in names}
            d_xyz, d_rot, d_sh, d_sh_p = deform(pb["xyz"][mask].detach(), torch.full((int(mask.sum()), 1), 0.37))
            asm = orc.assemble(rawb, mask, dict(d_xyz=d_xyz, d_rot=d_rot, d_sh=d_sh, d_sh_p=d_sh_p))
            oc = cpu_oracle.rasterize_autograd(asm, cv)
            ot = cpu_oracle.rasterize_autograd(asm, tv)
            loss = orc.loss_term(oc[0], gt_color, "l1", 1.0, 0.2) + \
                orc.loss_term(ot[1][5:6], gt_quad, "weighted_l2_quad", 0.5, 0.2, w=0.01)
            opt.zero_grad()
            loss.backward()
            opt.step()
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > bound_s or len(times) >= 12:
                break
        best = float(np.median(times[1:] if len(times) > 1 else times))
        npix = cv["W"] * cv["H"] + tv["W"] * tv["H"]
        return {"value": round(npix / 1e6 / best, 4), "unit": "Mpix/s", "ms_per_iter": round(best * 1e3, 1),
                "cores": n_threads, "omp_threads": cpu_oracle.num_threads(), "kind": "port",
                "sample": f"config C1: P={P}, 320x240 colour + ToF, 25% dynamic Gaussians through the deform MLP "
                          f"(D=8, W=256), l1+SSIM and weighted-L2 losses, Adam over 10 groups + MLP; median of {len(times)} iterations"}
    except Exception as ex:
        return {"value": None, "sample": f"failed: {type(ex).__name__}: {ex}"}


def run_cpu(args):
    wl = WORKLOADS[args.workload]
    cb = cpu_baseline(args, wl, bound_s=30.0)
    line = {"metric": "rasterizer fwd+bwd throughput (RGB+ToF+depth)", "value": cb["value"],
            "unit": "Mpix/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb.get("ms_per_view_fwd_bwd"), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.workload}: {wl['desc']} (CPU port, bounded sample)"},
            "cpu_baseline": cb,
            "blocks": {"cpu_pipeline": cpu_pipeline_block(bound_s=30.0)},
            "e2e": {"value": cb["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--no-clocks", action="store_true", help="do not sample the clocks during the run")
    ap.add_argument("--resident-only", action="store_true",
                    help="skip the e2e / render-only / block / cpu_baseline legs (for runs under ncu)")
    ap.add_argument("--blocks", default="all",
                    help="all | none | comma list of c1,c4,c5,c3,next,cpu: the other BASELINE configs beside the headline")
    ap.add_argument("--c5-gaussians", type=lambda s: [int(x) for x in s.split(",")], default=[1000000, 8000000])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "cpu":
        if rank == 0:
            run_cpu(args)
        return
    if args.impl == "reference":
        from oracle import ref_driver
        if not ref_driver.available():
            if rank == 0:
                run_cpu(args)     # the oracle always exists: CPU port as the reference arm
            return
    run_gpu(args, args.impl)


if __name__ == "__main__":
    main()
