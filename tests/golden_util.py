"""Loads tests/golden/raster_*.npz (outputs of the unmodified reference kernels on a B200, made by
tests/golden/make_golden.py) into the input dict shape tests/harness.py uses."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("tiny", "orbit", "orbit_r0", "init", "dense", "deg1")


def load(name, device="cpu"):
    z = np.load(os.path.join(GOLDEN_DIR, f"raster_{name}.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    sc = z["scalars"]
    spec = json.loads(str(z["spec"]))
    inp = dict(
        P=int(z["in_means3D"].shape[0]), W=int(spec["W"]), H=int(spec["H"]), sh_degree=int(sc[8]),
        means3D=t(z["in_means3D"]), scales=t(z["in_scales"]), rotations=t(z["in_rotations"]),
        opacities=t(z["in_opacities"]), shs=t(z["in_shs"]), shs_p=t(z["in_shs_p"]),
        viewmatrix=t(z["in_viewmatrix"]), projmatrix=t(z["in_projmatrix"]), campos=t(z["in_campos"]),
        tanfovx=float(sc[0]), tanfovy=float(sc[1]), near_n=float(sc[2]), far_n=float(sc[3]),
        depth_range=float(sc[4]), phase_offset=float(sc[5]), dc_offset=float(sc[6]),
        use_view_dependent_phase=bool(sc[7]), bg=t(z["in_bg"]),
        grads={k[2:]: t(z[k]) for k in z.files if k.startswith("g_")},
        empty=torch.Tensor([]),
    )
    gold = {k: torch.from_numpy(np.ascontiguousarray(z[k])) for k in z.files
            if k[:2] in ("f_", "s_", "b_")}
    gold["R"] = int(z["R"])
    gold["spec"] = spec
    return inp, gold


def load_knn():
    z = np.load(os.path.join(GOLDEN_DIR, "knn.npz"))
    return {k: torch.from_numpy(np.ascontiguousarray(z[k])) for k in z.files}
