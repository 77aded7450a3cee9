"""Generates tests/golden/train_*.npz by IMPORTING the reference's own Python functions from
/root/reference (this container only; the GPU box has no /root/reference) and running them on CPU:

  utils/loss_utils.py     l1_loss, l2_loss, weighted_l1_loss, weighted_l1_loss_quad,
                          weighted_l2_loss_quad, ssim   (+ autograd gradients), combined as in
                          train.py:204-223
  utils/general_utils.py  build_rotation, inverse_sigmoid

Inputs are stored next to the outputs.  The fixtures pin oracle/train_oracle.py (CPU tests) and the
CUDA operators (GPU tests) to the reference itself.
    python tests/golden/make_train_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))      # imported, never used by the functions below
    L = load(os.path.join(REF, "utils", "loss_utils.py"), "ref_loss_utils")
    cases = [
        dict(name="l1_ssim", C=3, H=37, W=53, kind="l1", lam=1.0, ld=0.2, w=0.0, nch=3),
        dict(name="wl1_ssim", C=3, H=40, W=40, kind="weighted_l1", lam=0.7, ld=0.2, w=0.01, nch=3),
        dict(name="wl1_2of7", C=7, H=33, W=20, kind="weighted_l1", lam=2.0, ld=0.0, w=0.1, nch=2),
        dict(name="wl2quad_ssim", C=1, H=64, W=48, kind="weighted_l2_quad", lam=1.5, ld=0.2, w=0.05, nch=1),
        dict(name="wl1quad_ssim", C=1, H=30, W=31, kind="weighted_l1_quad", lam=1.0, ld=0.1, w=0.05, nch=1),
        dict(name="l2_small", C=2, H=5, W=7, kind="l2", lam=1.0, ld=0.5, w=0.0, nch=2),
    ]
    out = {}
    for c in cases:
        g = torch.Generator().manual_seed(len(c["name"]) * 1000 + c["H"])
        img = (torch.rand(c["C"], c["H"], c["W"], generator=g) + 0.05).requires_grad_(True)
        gt = (img.detach() + 0.2 * torch.randn(img.shape, generator=g)).clamp(0, 1.2)
        k = c["kind"]
        if k == "l1":
            el = L.l1_loss(img, gt)
        elif k == "l2":
            el = L.l2_loss(img, gt)
        elif k == "weighted_l1":
            el = L.weighted_l1_loss(img, gt, c["w"], c["nch"])
        elif k == "weighted_l1_quad":
            el = L.weighted_l1_loss_quad(img, gt, c["w"])
        else:
            el = L.weighted_l2_loss_quad(img, gt, c["w"])
        # train.py:204-223
        loss = c["lam"] * ((1.0 - c["ld"]) * el + c["ld"] * (1.0 - L.ssim(img, gt))) if c["ld"] != 0.0 \
            else c["lam"] * el
        loss.backward()
        n = c["name"]
        out[n + "/img"] = img.detach().numpy()
        out[n + "/gt"] = gt.numpy()
        out[n + "/loss"] = np.float64(loss.item())
        out[n + "/grad"] = img.grad.numpy()
        out[n + "/ssim"] = np.float64(L.ssim(img.detach(), gt).item())
        out[n + "/meta"] = np.array([c["lam"], c["ld"], c["w"], c["nch"]], dtype=np.float64)
        out[n + "/kind"] = np.array(c["kind"])
    np.savez_compressed(os.path.join(HERE, "train_loss.npz"), **out)

    G = load(os.path.join(REF, "utils", "general_utils.py"), "ref_general_utils")
    torch.manual_seed(3)
    q = torch.randn(64, 4)
    real_zeros = torch.zeros
    torch.zeros = lambda *a, **k: real_zeros(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})  # build_rotation asks for device='cuda'
    try:
        R = G.build_rotation(q)
    finally:
        torch.zeros = real_zeros
    x = torch.rand(64) * 0.98 + 0.01
    np.savez_compressed(os.path.join(HERE, "train_misc.npz"), q=q.numpy(), R=R.numpy(), x=x.numpy(),
                        inv_sigmoid=G.inverse_sigmoid(x).numpy())
    print("wrote train_loss.npz, train_misc.npz")


if __name__ == "__main__":
    main()
