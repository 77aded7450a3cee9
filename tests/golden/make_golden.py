"""Generates the golden fixtures of tests/golden/*.npz by running the UNMODIFIED reference kernels
(oracle/_ref/libgftorf_ref.so, built by oracle/Makefile from /root/reference) on a GPU.

The reference ships no golden vectors of its own, so these are outputs of the
reference itself on seeded synthetic inputs; inputs are stored next to the outputs so the fixtures
do not depend on a random generator.  Run on a B200 box:
    python tests/golden/make_golden.py --out gpurun_out/golden
and copy the .npz files into tests/golden/.  (Last generated: see tests/golden/MANIFEST.json.)

Gradient fixtures: the reference's dL_dsh_p is only defined for Gaussian 0 (out-of-bounds local
read, DESIGN.md defect D1), which also leaks into dL_dmeans3D through the SH direction term unless
the higher-order phase/amplitude SH coefficients are zero; cases with `zero_shp_rest` are the ones
whose dL_dmeans3D is meaningful, and only row 0 of dL_dsh_p ever is.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import harness  # noqa: E402
from oracle import ref_driver  # noqa: E402

CASES = {
    # partial tiles in y (40 = 2.5 tiles), large splats
    "tiny": dict(P=64, W=48, H=40, kind="trained", seed=0, sigma_px=4.0, zero_shp_rest=True),
    # rotated camera, view-dependent phase, offsets, partial tiles in both axes
    "orbit": dict(P=400, W=72, H=56, kind="trained", seed=1, pose="orbit", sigma_px=2.5,
                  view_dependent_phase=True, phase_offset=0.3, dc_offset=0.1),
    "orbit_r0": dict(P=400, W=72, H=56, kind="trained", seed=1, pose="orbit", sigma_px=2.5,
                     view_dependent_phase=True, phase_offset=0.3, dc_offset=0.1,
                     zero_shp_rest=True),
    # the reference's initialisation: huge isotropic splats, identity rotations, opacity 0.1
    "init": dict(P=300, W=64, H=48, kind="init", seed=2),
    # saturating pixels: early termination, n_contrib < list length
    "dense": dict(P=1500, W=48, H=32, kind="trained", seed=3, sigma_px=5.0, zero_shp_rest=True),
    # SH degree 1 with 16-coefficient tensors (training starts at degree 0 and grows)
    "deg1": dict(P=200, W=48, H=48, kind="trained", seed=4, sigma_px=3.0, sh_degree=1,
                 zero_shp_rest=True),
}
KNN_SIZES = (1, 3, 4, 100, 2000)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    manifest = {"generator": "tests/golden/make_golden.py", "device": torch.cuda.get_device_name(0),
                "torch": torch.__version__, "cases": {}}
    for name, spec in CASES.items():
        inp = harness.build_inputs(device="cuda", **spec)
        P, W, H = inp["P"], inp["W"], inp["H"]
        fwd = harness.call_forward(ref_driver.RefModule, inp)
        torch.cuda.synchronize()
        dec = ref_driver.decode_buffers(fwd[12], fwd[13], fwd[14], P, fwd[0], W, H)
        bwd, internal = harness.call_backward(ref_driver.RefModule, inp, fwd, return_internal=True)
        torch.cuda.synchronize()
        vis = (fwd[11] > 0)
        out = {"spec": json.dumps(spec)}
        for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p", "viewmatrix",
                  "projmatrix", "campos", "bg"):
            out["in_" + k] = inp[k].cpu().numpy()
        for k, v in inp["grads"].items():
            out["g_" + k] = v.cpu().numpy()
        out["scalars"] = np.array([inp["tanfovx"], inp["tanfovy"], inp["near_n"], inp["far_n"],
                                   inp["depth_range"], inp["phase_offset"], inp["dc_offset"],
                                   float(inp["use_view_dependent_phase"]), float(inp["sh_degree"])],
                                  dtype=np.float64)
        out["R"] = np.int64(fwd[0])
        for i, k in enumerate(harness.FWD_NAMES):
            if i >= 1:
                out["f_" + k] = fwd[i].cpu().numpy()
        # per-Gaussian state is only defined for visible Gaussians (uninitialised otherwise)
        for k in ("tiles_touched", "point_offsets"):
            out["s_" + k] = dec[k].cpu().numpy()
        for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "real_img_amp", "dists",
                  "ndc", "pa", "clamped", "clamped_p"):
            v = dec[k].clone()
            v[~vis] = 0
            out["s_" + k] = v.cpu().numpy()
        for k in ("final_T", "w_z_total", "w_z2_total", "n_contrib", "ranges"):
            out["s_" + k] = dec[k].cpu().numpy()
        if fwd[0] > 0:
            out["s_keys"] = dec["keys"].cpu().numpy()
            out["s_point_list"] = dec["point_list"].cpu().numpy()
        for i, k in enumerate(harness.BWD_NAMES):
            out["b_" + k] = bwd[i].cpu().numpy()
        for k, v in internal.items():
            out["b_" + k] = v.cpu().numpy()
        np.savez_compressed(os.path.join(args.out, f"raster_{name}.npz"), **out)
        manifest["cases"][name] = dict(spec=spec, R=int(fwd[0]), V=int(vis.sum().item()))
        print(name, "R", fwd[0], "V", int(vis.sum().item()), flush=True)

    knn = {}
    for P in KNN_SIZES:
        rng = np.random.default_rng(100 + P)
        pts = (rng.random((P, 3), dtype=np.float32) * 4 - 2).astype(np.float32)
        if P >= 100:
            pts[7] = pts[3]  # coincident points count with distance 0
        d = ref_driver.distCUDA2(torch.from_numpy(pts).cuda())
        torch.cuda.synchronize()
        knn[f"pts_{P}"] = pts
        knn[f"out_{P}"] = d.cpu().numpy()
    np.savez_compressed(os.path.join(args.out, "knn.npz"), **knn)
    manifest["knn_sizes"] = list(KNN_SIZES)
    with open(os.path.join(args.out, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
