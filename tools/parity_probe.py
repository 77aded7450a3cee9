"""Prints a detailed parity report (product library vs the reference kernels in oracle/_ref) for a
list of synthetic cases.  Diagnostic companion of tests/test_gpu_parity.py; run on a GPU box:
    python tools/parity_probe.py [--cases small,c1,c2] [--json gpurun_out/parity.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import harness  # noqa: E402
from gftorf_b200 import rasterizer, debug  # noqa: E402
from gftorf_b200.knn import distCUDA2  # noqa: E402
from oracle import ref_driver  # noqa: E402

CASES = {
    "tiny": dict(P=64, W=48, H=40, kind="trained", seed=0, sigma_px=4.0),
    "small": dict(P=2000, W=160, H=120, kind="trained", seed=1, pose="orbit"),
    "init_small": dict(P=2000, W=160, H=120, kind="init", seed=2),
    "c1": dict(P=20000, W=320, H=240, kind="trained", seed=0),
    "c1_r0": dict(P=20000, W=320, H=240, kind="trained", seed=0, zero_shp_rest=True),
    "c2_r0": dict(P=300000, W=640, H=480, kind="trained", seed=0, zero_shp_rest=True),
    "c1_init": dict(P=20000, W=320, H=240, kind="init", seed=0),
    "c2": dict(P=300000, W=640, H=480, kind="trained", seed=0),
    "c2_orbit_vdp": dict(P=300000, W=640, H=480, kind="trained", seed=3, pose="orbit",
                         view_dependent_phase=True, phase_offset=0.3, dc_offset=0.1),
    "big": dict(P=2000000, W=1920, H=1080, kind="trained", seed=0),
}


def timed(fn, n=5):
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        ev0.record()
        fn()
        ev1.record()
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="tiny,small,init_small,c1,c1_init,c2,c2_orbit_vdp")
    ap.add_argument("--json", default="")
    ap.add_argument("--time", action="store_true")
    args = ap.parse_args()
    out = {}
    for name in args.cases.split(","):
        spec = CASES[name]
        inp = harness.build_inputs(device="cuda", **spec)
        P, W, H = inp["P"], inp["W"], inp["H"]
        ours = harness.call_forward(rasterizer._C, inp)
        ref = harness.call_forward(ref_driver.RefModule, inp)
        torch.cuda.synchronize()
        od = debug.decode_buffers(ours[12], ours[13], ours[14], P, ours[0], W, H)
        rd = ref_driver.decode_buffers(ref[12], ref[13], ref[14], P, ref[0], W, H)
        rep = harness.compare_forward(ours, ref, od, rd)
        ob = harness.call_backward(rasterizer._C, inp, ours)
        rb = harness.call_backward(ref_driver.RefModule, inp, ref)
        torch.cuda.synchronize()
        brep = harness.compare_backward(ob, rb)
        if ob[7].numel():
            brep["sh_p_row0"] = harness.rel_l2(ob[7][0], rb[7][0])
        rep["bwd"] = brep
        if args.time:
            rep["ms"] = dict(
                ours_fwd=timed(lambda: harness.call_forward(rasterizer._C, inp)),
                ref_fwd=timed(lambda: harness.call_forward(ref_driver.RefModule, inp)),
                ours_bwd=timed(lambda: harness.call_backward(rasterizer._C, inp, ours)),
                ref_bwd=timed(lambda: harness.call_backward(ref_driver.RefModule, inp, ref)),
            )
        out[name] = rep
        print(name, json.dumps(rep), flush=True)

    # knn
    for P in (1, 3, 4, 100, 5000, 100000):
        g = torch.Generator(device="cpu").manual_seed(P)
        pts = (torch.rand((P, 3), generator=g) * 4 - 2).cuda()
        a = distCUDA2(pts)
        b = ref_driver.distCUDA2(pts)
        torch.cuda.synchronize()
        same = bool(torch.equal(a.view(torch.int32), b.view(torch.int32)))
        print("knn", P, "bit-identical" if same else
              f"MISMATCH n={(a != b).sum().item()} maxabs={(a - b).abs().max().item()}", flush=True)
        out[f"knn_{P}"] = same
    if args.json:
        os.makedirs(os.path.dirname(args.json), exist_ok=True)
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
