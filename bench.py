#!/usr/bin/env python
"""bench.py — measures the hot path BASELINE.json names: the RGB+ToF Gaussian rasterizer's
forward+backward inside a training iteration, on synthetic F-TöRF-shaped scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|cpu] [--workload c2]

A STEP is one training iteration's worth of rasterizer work for one camera pair, exactly the call
pattern of gaussian_renderer/__init__.py:107-128 + train.py:279: forward of the colour view,
forward of the ToF view, backward of both (every output computed, all five consumed output
gradients dense).  At N > 1 every rank does one such step on its own views (weak scaling) and the
step ends with the NCCL sum-allreduce of the flat per-Gaussian gradient bucket.

  value  : Mpix/s of pixels taken through forward+backward, whole job, inputs resident in HBM,
           called through the C-ABI entry points (gftorf_b200.rasterizer._C);
           ms_per_step is BASELINE's "fwd+bwd ms/iter".  The two views of a step are in flight
           concurrently (parallel.ViewRunner: one stream + host thread per view, gradients added
           into the zero-filled bucket with atomics); --sequential-views runs them back to back.
           Per-kernel times (stage_ms_per_step, roofline) come from a second, sequential pass of
           the same steps, because concurrent kernels share the SMs.
  e2e    : the same metric through the public autograd surface (GaussianRasterizer + backward(),
           the two views through ViewRunner) with HOST buffers: pinned-host -> device copies of
           all Gaussian parameters, cameras, background and pixel gradients, and device -> host
           reads of the parameter gradients and the rendered images, all inside the timed region.
  next_rows : the SURVEY §8f operators (assembly, loss, Adam) and distCUDA2 at the workload's
           size, each beside the reference's implementation of the same step on the same GPU.
  --impl reference : the UNMODIFIED reference kernels (oracle/_ref/libgftorf_ref.so, built from
           /root/reference by oracle/Makefile) driven the way the reference's torch binding drives
           them (oracle/ref_driver.py) — same workload, same metric, on the GPU.  The reference has
           no CPU implementation, so its "own implementation of the path" is this.
           Falls back to the CPU port (oracle/gft_oracle.cpp) only if that library is absent.
  --impl cpu : the CPU port on the host cores (the cpu_baseline, as a full line).
"""
import argparse
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
    # configs[3] per-view shape: 2M Gaussians, 1080p RGB + 640x480 ToF
    "c4": dict(P=2000000, color=(1920, 1080), tof=(640, 480), depth_range=15.0, kind="trained",
               desc="BASELINE configs[3] per-view shape: 2M Gaussians, 1080p colour + 640x480 ToF"),
    # configs[4] upper end: 8M Gaussians (screen-space sigma 1 px), two 1080p frames per step
    "c5": dict(P=8000000, color=(1920, 1080), tof=(1920, 1080), depth_range=15.0, kind="trained", sigma_px=1.0,
               desc="BASELINE configs[4] upper end: 8M Gaussians, two 1920x1080 frames"),
}


# --------------------------------------------------------------------------------------------
def make_view(P, wh, depth_range, kind, seed, device, cloud=None, bg_hw=None, sigma_px=1.5):
    W, H = wh
    cam = scenes.make_camera(W, H, depth_range=depth_range, seed=seed)
    if cloud is None:
        cloud = scenes.make_cloud(P, cam, kind=kind, seed=seed, sigma_px=sigma_px)
    bh, bw = bg_hw if bg_hw else (H, W)
    bg = scenes.make_background(bh, bw, seed=seed)
    grads = scenes.make_pixel_grads(H, W, seed=seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    view = dict(W=W, H=H, viewmatrix=t(cam["viewmatrix"]), projmatrix=t(cam["projmatrix"]),
                campos=t(cam["campos"]), tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
                near_n=cam["znear"], far_n=cam["zfar"], depth_range=cam["depth_range"], bg=t(bg),
                grads={k: t(v) for k, v in grads.items()})
    return view, cloud


def build_scene(wl, seed, device):
    color, cloud = make_view(wl["P"], wl["color"], wl["depth_range"], wl["kind"], seed, device,
                             sigma_px=wl.get("sigma_px", 1.5))
    # the bg map is sized from the colour camera and reused for the ToF view (train.py:121-128)
    tof, _ = make_view(wl["P"], wl["tof"], wl["depth_range"], wl["kind"], seed, device, cloud=cloud,
                       bg_hw=(wl["color"][1], wl["color"][0]))
    tof["bg"] = color["bg"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    params = {k: t(cloud[k]) for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p")}
    return params, [color, tof]


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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=100):
        self.index, self.proc, self.lines, self.interval_ms = index, None, [], interval_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", str(self.interval_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def stage_bytes(stage, P, V, R, N, T):
    """Algorithmic bytes per launch of each stage (SURVEY §8d split per kernel)."""
    return {
        "preprocess_fwd": 24 * P + 464 * V,
        "duplicate_keys": 12 * R,
        "radix_sort": 24 * R,
        "identify_ranges": 8 * R + 8 * T,
        "blend_fwd": 76 * R + 128 * N,
        "zero_grad_records": 64 * P,
        "blend_bwd": 76 * R + 96 * N + 80 * V,
        "preprocess_bwd": 388 * P + 472 * V,
    }.get(stage, 0)


# --------------------------------------------------------------------------------------------
def run_gpu(args, impl):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from gftorf_b200 import rasterizer, _capi, parallel
    if impl == "ours":
        mod = rasterizer._C
    else:
        from oracle import ref_driver
        mod = ref_driver.RefModule

    wl = WORKLOADS[args.workload]
    params, views = build_scene(wl, seed=rank, device=dev)   # each rank: its own views of the batch
    empty = torch.Tensor([])
    zero3 = {id(v): torch.zeros_like(v["grads"]["color"]) for v in views}
    zero1 = {id(v): torch.zeros_like(v["grads"]["depth"]) for v in views}
    P = wl["P"]
    npix = sum(v["W"] * v["H"] for v in views)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # the flat gradient bucket [P*91 + 2 scalars]; the backward adds each view's gradients into it
    scal = [torch.zeros(1, device=dev), torch.zeros(1, device=dev)]
    bucket = parallel.GradBucket(params, scal)
    grad_out = bucket.grad_out() if impl == "ours" else None
    bucket_views = dict(zip(bucket.names, bucket.views()[:len(bucket.names)]))

    r_last = {}

    def fwd(v):
        if impl != "ours":
            return mod.rasterize_gaussians(*fwd_args(params, v, empty))
        # hinted mode: the previous step's instance count (+25 %) sizes the binning workspace, so
        # the forward is enqueued without a host round trip in the middle
        hint = int(r_last[id(v)] * 1.25) + 4096 if id(v) in r_last else 0
        f = mod.rasterize_gaussians(*fwd_args(params, v, empty), R_hint=hint)
        r_last[id(v)] = f[0]
        return f

    runner = parallel.ViewRunner(len(views), dev) if (impl == "ours" and not args.sequential_views) else None
    if runner is not None and hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        # the leaves live on the main stream, the views' backward nodes on the runner's streams
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)

    def one_view_atomic(v):
        f = fwd(v)
        mod.rasterize_gaussians_backward(
            *bwd_args(params, v, f, empty, zero3[id(v)], zero1[id(v)]), grad_out=grad_out, accumulate="atomic")
        return (f[0], f[11])

    def step_resident():
        if runner is not None:
            # the two views of the iteration run concurrently (one stream + host thread each) and
            # add their gradients into the zero-filled bucket with atomics
            bucket.zero()
            stats = runner.run([(lambda v=v: one_view_atomic(v)) for v in views])
            if world > 1:
                bucket.allreduce()
            return stats
        return step_sequential()

    def step_sequential():
        if impl != "ours":
            bucket.zero()
        stats = []
        for vi, v in enumerate(views):
            f = fwd(v)
            if impl == "ours":
                # the first view overwrites the bucket (no zero fill), the others add into it
                mod.rasterize_gaussians_backward(
                    *bwd_args(params, v, f, empty, zero3[id(v)], zero1[id(v)]), grad_out=grad_out,
                    accumulate=vi > 0)
            else:
                # the reference's gradients of the two views add in autograd's AccumulateGrad
                b = mod.rasterize_gaussians_backward(*bwd_args(params, v, f, empty, zero3[id(v)], zero1[id(v)]))
                for name, t in (("means3D", b[4]), ("shs", b[6]), ("shs_p", b[7]), ("opacities", b[3]),
                                ("scales", b[8]), ("rotations", b[9])):
                    bucket_views[name] += t
            stats.append((f[0], f[11]))
        if world > 1:
            bucket.allreduce()
        return stats

    # ---- e2e: public autograd surface, host buffers --------------------------------------------
    names = ("means3D", "opacities", "shs", "shs_p", "scales", "rotations")
    host_params = {k: params[k].cpu().pin_memory() for k in names}
    host_views = []
    for v in views:
        hv = {k: v[k].cpu().pin_memory() for k in ("viewmatrix", "projmatrix", "campos")}
        hv["grads"] = {k: g.cpu().pin_memory() for k, g in v["grads"].items()}
        host_views.append(hv)
    host_bg = views[0]["bg"].cpu().pin_memory()
    host_out_grads = {k: torch.empty_like(host_params[k]).pin_memory() for k in names}
    host_imgs = [torch.empty((11, v["H"], v["W"]), dtype=torch.float32).pin_memory() for v in views]
    h2d_bytes = sum(t.numel() * 4 for t in host_params.values()) + host_bg.numel() * 4 + sum(
        sum(t.numel() * 4 for t in (hv["viewmatrix"], hv["projmatrix"], hv["campos"])) +
        sum(g.numel() * 4 for g in hv["grads"].values()) for hv in host_views)
    d2h_bytes = sum(t.numel() * 4 for t in host_out_grads.values()) + sum(t.numel() * 4 for t in host_imgs)

    if impl == "ours":
        Settings, Raster = rasterizer.GaussianRasterizationSettings, rasterizer.GaussianRasterizer
    else:
        # the reference's own autograd Function would sit here; its binding is restated by
        # ref_driver, so the e2e arm drives forward/backward explicitly with the same copies
        Settings = Raster = None

    # persistent device-side staging (what an application keeps): the H2D copies land in these
    dev_params = {k: torch.empty_like(params[k]) for k in names}
    dev_bg = torch.empty_like(views[0]["bg"])
    dev_views = []
    for v in views:
        dv = {k: torch.empty_like(v[k]) for k in ("viewmatrix", "projmatrix", "campos")}
        dv["grads"] = {k: torch.empty_like(g) for k, g in v["grads"].items()}
        dev_views.append(dv)
    if impl == "ours":
        for k in names:
            dev_params[k].requires_grad_(True)
        dev_m2d = torch.zeros_like(params["means3D"], requires_grad=True)

    side = {id(v): (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)) for v in views}

    def e2e_view_job(v, hv, dv, himg):
        # one view through the public surface on the runner's stream: forward, autograd backward
        # (gradients accumulate into the shared leaves).  Its host copies ride on two side streams:
        # the pixel gradients come up while the forward runs, the images go down while the
        # backward runs.
        cur = torch.cuda.current_stream()
        up, down = side[id(v)]
        with torch.no_grad():
            for k in ("viewmatrix", "projmatrix", "campos"):
                dv[k].copy_(hv[k], non_blocking=True)
            up.wait_stream(cur)
            with torch.cuda.stream(up):
                for k, t in hv["grads"].items():
                    dv["grads"][k].copy_(t, non_blocking=True)
        g = dv["grads"]
        s = Settings(image_height=v["H"], image_width=v["W"], tanfovx=v["tanfovx"], tanfovy=v["tanfovy"],
                     bg=dev_bg, scale_modifier=1.0, viewmatrix=dv["viewmatrix"], projmatrix=dv["projmatrix"],
                     sh_degree=3, campos=dv["campos"], prefiltered=False, debug=False, near_n=v["near_n"],
                     far_n=v["far_n"], depth_range=v["depth_range"])
        out = Raster(s)(means3D=dev_params["means3D"], means2D=dev_m2d, opacities=dev_params["opacities"],
                        shs=dev_params["shs"], shs_p=dev_params["shs_p"], scales=dev_params["scales"],
                        rotations=dev_params["rotations"])
        down.wait_stream(cur)
        with torch.cuda.stream(down), torch.no_grad():
            himg[0:3].copy_(out[0], non_blocking=True)
            himg[3:10].copy_(out[1], non_blocking=True)
            himg[10:11].copy_(out[2], non_blocking=True)
        cur.wait_stream(up)
        torch.autograd.backward([out[0], out[1], out[2], out[4], out[6]],
                                [g["color"], g["phasor"], g["depth"], g["acc"], g["depth_distortion"]])
        cur.wait_stream(down)

    def step_e2e_concurrent():
        with torch.no_grad():
            for k in names:
                dev_params[k].copy_(host_params[k], non_blocking=True)
            dev_bg.copy_(host_bg, non_blocking=True)
        for k in names:
            dev_params[k].grad = None
        dev_m2d.grad = None
        runner.run([(lambda a=a: e2e_view_job(*a)) for a in zip(views, host_views, dev_views, host_imgs)])
        for k in names:
            host_out_grads[k].copy_(dev_params[k].grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def step_e2e():
        if runner is not None:
            return step_e2e_concurrent()
        with torch.no_grad():
            for k in names:
                dev_params[k].copy_(host_params[k], non_blocking=True)
            dev_bg.copy_(host_bg, non_blocking=True)
            for hv, dv in zip(host_views, dev_views):
                for k in ("viewmatrix", "projmatrix", "campos"):
                    dv[k].copy_(hv[k], non_blocking=True)
                for k, t in hv["grads"].items():
                    dv["grads"][k].copy_(t, non_blocking=True)
        dp, bg = dev_params, dev_bg
        acc = None
        if impl == "ours":
            for k in names:
                dp[k].grad = None
            dev_m2d.grad = None
        for v, dv, himg in zip(views, dev_views, host_imgs):
            vm, pm, cp, g = dv["viewmatrix"], dv["projmatrix"], dv["campos"], dv["grads"]
            if impl == "ours":
                s = Settings(image_height=v["H"], image_width=v["W"], tanfovx=v["tanfovx"],
                             tanfovy=v["tanfovy"], bg=bg, scale_modifier=1.0, viewmatrix=vm,
                             projmatrix=pm, sh_degree=3, campos=cp, prefiltered=False, debug=False,
                             near_n=v["near_n"], far_n=v["far_n"], depth_range=v["depth_range"])
                out = Raster(s)(means3D=dp["means3D"], means2D=dev_m2d, opacities=dp["opacities"],
                                shs=dp["shs"], shs_p=dp["shs_p"], scales=dp["scales"],
                                rotations=dp["rotations"])
                torch.autograd.backward(
                    [out[0], out[1], out[2], out[4], out[6]],
                    [g["color"], g["phasor"], g["depth"], g["acc"], g["depth_distortion"]])
                himg[0:3].copy_(out[0], non_blocking=True)
                himg[3:10].copy_(out[1], non_blocking=True)
                himg[10:11].copy_(out[2], non_blocking=True)
            else:
                vv = dict(v, viewmatrix=vm, projmatrix=pm, campos=cp, bg=bg, grads=g)
                f = mod.rasterize_gaussians(*fwd_args(dp, vv, empty))
                b = mod.rasterize_gaussians_backward(*bwd_args(dp, vv, f, empty, zero3[id(v)], zero1[id(v)]))
                grads_now = dict(means3D=b[4], opacities=b[3], shs=b[6], shs_p=b[7], scales=b[8], rotations=b[9])
                acc = grads_now if acc is None else {k: acc[k] + grads_now[k] for k in names}
                himg[0:3].copy_(f[1], non_blocking=True)
                himg[3:10].copy_(f[2], non_blocking=True)
                himg[10:11].copy_(f[3], non_blocking=True)
        for k in names:
            gsrc = dp[k].grad if impl == "ours" else acc[k]
            host_out_grads[k].copy_(gsrc, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, K, W, profile=False):
        for _ in range(W):
            step_fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        if profile and impl == "ours":
            _capi.profile_read(4096)
            _capi.profile_enable(True)
        n0 = _capi.launch_count() if impl == "ours" else 0
        import gc
        gc.collect()
        gc.disable()      # a cyclic-GC pause inside a 2 ms step would be charged to the step
        t0 = time.perf_counter()
        stats = None
        for i in range(K):
            flush.zero_()                      # L2 flush (256 MiB > 126 MB L2), untimed
            evs[i][0].record()
            stats = step_fn()
            evs[i][1].record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        gc.enable()
        launches = (_capi.launch_count() - n0) if impl == "ours" else None
        stages = []
        if profile and impl == "ours":
            stages = _capi.profile_read(4096)
            _capi.profile_enable(False)
        ms = [a.elapsed_time(b) for a, b in evs]
        timed.last_steps = ms
        total = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item()), wall, launches, stages, stats

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local, args.clock_interval_ms)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    total_ms, wall_ms, launches, stages, stats = timed(step_resident, K, W, profile=runner is None)
    step_ms = sorted(timed.last_steps)
    if runner is not None:
        # per-kernel durations need the kernels one at a time: a second, sequential pass of the same
        # steps (untimed for `value`) feeds the stage table and the roofline entry
        Kp = max(10, K // 4)
        _, _, _, stages, _ = timed(step_sequential, Kp, 3, profile=True)
    stage_steps = Kp if runner is not None else K
    clocks = sampler.stop() if (rank == 0 and not args.no_clocks) else None
    if args.resident_only:
        if rank == 0:
            share = {}
            for name, ms in stages:
                share[name] = share.get(name, 0.0) + ms / stage_steps
            print(json.dumps({"resident_only": True, "ms_per_step": round(total_ms / K, 4),
                              "step_ms_min_med_max": [round(step_ms[0], 4), round(step_ms[len(step_ms) // 2], 4), round(step_ms[-1], 4)],
                              "stage_ms_per_step": {k: round(v, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])},
                              "gpu_launches": launches}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    e2e_ms, _, _, _, _ = timed(step_e2e, K, W)

    # render-only throughput (forward only, all outputs) — BASELINE's second metric
    def step_render():
        if runner is not None:       # frames are independent: two in flight (render-only)
            runner.run([(lambda v=v: fwd(v)) for v in views])
            return
        for v in views:
            fwd(v)
    render_ms, _, _, _, _ = timed(step_render, K, W)

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
                               "one camera pair per rank per step, fwd colour + fwd ToF + bwd both"
                               + ("; + NCCL allreduce of the [P,91] gradient bucket" if world > 1 else ""),
                   "P": P, "views_per_step_per_rank": len(views),
                   "l2": "flushed between timed steps (256 MiB memset, outside the per-step events)",
                   "timing": "sum of per-step CUDA-event durations, max over ranks",
                   "views": "concurrent (one stream per view)" if runner is not None else "sequential"},
        "fwd_bwd_ms_per_iter": round(total_ms / K, 4),
        "step_ms_min_med_max": [round(step_ms[0], 4), round(step_ms[len(step_ms) // 2], 4), round(step_ms[-1], 4)],
        "wall_ms_total_incl_flush": round(wall_ms, 2),
        "render_mpix_s": round(mpix * K / (render_ms / 1e3), 3),
        "render_ms_per_step": round(render_ms / K, 4),
        "e2e": {"value": round(mpix * K / (e2e_ms / 1e3), 3), "unit": "Mpix/s",
                "ms_per_step": round(e2e_ms / K, 4), "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes),
                "api": ("GaussianRasterizer(...) + torch.autograd.backward"
                        + (", the two views through parallel.ViewRunner" if runner is not None else "")) if impl == "ours"
                       else "reference kernels via oracle/ref_driver.py (binding restated)"},
        "clocks": clocks,
    }
    if impl == "ours":
        line["gpu_launches"] = int(launches)
        # roofline of the dominant kernel, from the per-stage CUDA events of the timed region
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
            # V, R of the view with the larger share are close; use the mean over the step's views
            Vs = [int((s[1] > 0).sum().item()) for s in stats]
            Rs = [int(s[0]) for s in stats]
            Ns = [v["W"] * v["H"] for v in views]
            Ts = [((v["W"] + 15) // 16) * ((v["H"] + 15) // 16) for v in views]
            byts = float(np.mean([stage_bytes(dom, P, Vs[i], Rs[i], Ns[i], Ts[i]) for i in range(len(views))]))
            ach = byts / (mean_ms[dom] * 1e-3) / 1e9
            traffic = issue_busy = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = tj.get(dom)
                issue_busy = tj.get("_issue_slots_busy_per_active_cycle", {}).get(dom)
            except Exception:
                pass
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 2),
                                "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 5),
                                "traffic": traffic,
                                "issue_slots_busy": issue_busy,   # ncu smsp__issue_active per active cycle (profiles/)
                                "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s",
                                "algorithmic_bytes_per_launch": int(byts),
                                "kernel_ms_per_launch": round(mean_ms[dom], 4),
                                "note": "the blend kernels are issue/latency bound by construction "
                                        "(per-pair ALU + MUFU + shuffles), so their HBM fraction is small; "
                                        "see profiles/ for issue-slot utilisation"}
            total_bytes = sum(504 * P + 936 * Vs[i] + 188 * Rs[i] + 224 * Ns[i] + 8 * Ts[i] for i in range(len(views)))
            line["step_roofline"] = {"algorithmic_bytes_per_step": int(total_bytes),
                                     "achieved_gbs": round(total_bytes / (total_ms / K * 1e-3) / 1e9, 2),
                                     "frac": round(total_bytes / (total_ms / K * 1e-3) / 1e9 / peak, 5),
                                     "V": Vs, "R": Rs}
            line["stage_ms_per_step"] = {k: round(v, 4) for k, v in sorted(step_share.items(), key=lambda kv: -kv[1])}
            if runner is not None:
                line["stage_ms_note"] = ("per-kernel times from a sequential pass of the same steps "
                                         f"({stage_steps} steps); the timed steps run the two views concurrently")
        if world == 1:
            line["next_rows"] = next_rows_bench(P, views[0]["H"], views[0]["W"], dev, flush, peak)
            line["cpu_baseline"] = cpu_baseline(args, wl, bound_s=20.0)
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
def next_rows_bench(P, H, W, dev, flush, peak_gbs, iters=20):
    """SURVEY §8f rows built so far (f1 assembly, f3 loss, f4 Adam) at the bench workload's size:
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
    except Exception as e:          # the reference library did not travel: report ours alone
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
                          f"median of {len(times)} runs, OpenMP threads = host cores"}
    except Exception as ex:  # never take the GPU line down with the baseline
        return {"value": None, "unit": "Mpix/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}


def run_cpu(args):
    wl = WORKLOADS[args.workload]
    cb = cpu_baseline(args, wl, bound_s=30.0)
    line = {"metric": "rasterizer fwd+bwd throughput (RGB+ToF+depth)", "value": cb["value"],
            "unit": "Mpix/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb.get("ms_per_view_fwd_bwd"), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.workload}: {wl['desc']} (CPU port, bounded sample)"},
            "cpu_baseline": cb,
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
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi during the run")
    ap.add_argument("--sequential-views", action="store_true",
                    help="run the two views of a step back to back on one stream (default: concurrently)")
    ap.add_argument("--clock-interval-ms", type=int, default=100)
    ap.add_argument("--resident-only", action="store_true",
                    help="skip the e2e / render-only / cpu_baseline legs (for runs under ncu)")
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
