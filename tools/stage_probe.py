"""Per-stage CUDA-event times of one batched step (forward_views + backward_views) on a bench
workload, for A/B runs of the kernel options:
    python tools/stage_probe.py [--workload c2] [--steps 30] [--opt bwd_pred=0 --opt pbwd_minb=3 ...] [--single]
--single runs the views one call each (a batch of one per view) instead of one batch."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--single", action="store_true")
    args = ap.parse_args()
    import bench
    from gftorf_b200 import views as V, _capi, parallel
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    for o in args.opt:
        k, v = o.split("=")
        _capi.set_option(k, int(v))
    wl = bench.WORKLOADS[args.workload]
    params, views = bench.build_scene(wl, seed=0, device=dev)
    specs = [bench.view_spec(v) for v in views]
    grads = [v["grads"] for v in views]
    bucket = parallel.GradBucket(params, [torch.zeros(1, device=dev), torch.zeros(1, device=dev)])
    go = bucket.grad_out()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    groups = [[i] for i in range(len(specs))] if args.single else [list(range(len(specs)))]
    hint = {}

    def step():
        Rs = []
        for gi, grp in enumerate(groups):
            f = V.forward_views(params["means3D"], params["opacities"], params["scales"], params["rotations"],
                                params["shs"], params["shs_p"], [specs[i] for i in grp], 3,
                                R_hint=hint.get(gi, 0))
            hint[gi] = int(f.R * 1.25) + 4096
            V.backward_views(f, [grads[i] for i in grp], grad_out=go, accumulate=gi > 0)
            Rs.append(f.R)
        return Rs

    for _ in range(5):
        Rs = step()
    torch.cuda.synchronize()
    _capi.profile_read(4096)
    _capi.profile_enable(True)
    ev = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    stages = _capi.profile_read(4096)
    _capi.profile_enable(False)
    per = {}
    for n, ms in stages:
        per.setdefault(n, []).append(ms)
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    print(json.dumps({"workload": args.workload, "opts": args.opt, "single": args.single, "R": Rs,
                      "step_ms_min_med": [round(ms[0], 4), round(ms[len(ms) // 2], 4)],
                      "stage_ms_per_step": {k: round(float(np.sum(v)) / args.steps, 4) for k, v in per.items()}}))


if __name__ == "__main__":
    main()
