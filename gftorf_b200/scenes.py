"""Synthetic F-TöRF / TöRF-shaped inputs for tests and benchmarks.

No datasets ship with the reference (data/ is git-ignored upstream), so every measured workload is
generated here, deterministically from a seed, with the camera model and initialisation
distributions of the reference:
  - projection: utils/graphics_utils.py:55-75 (getProjectionMatrix), matrices stored transposed
    (column-major) as scene/cameras.py:121-129 does, `projmatrix` = view @ proj (cameras.py:129);
  - near/far from depth_range: scene/dataset_readers.py:465-466 with the factors of
    run_optimize.py:30-31,84-85;
  - "init" cloud: uniform in the frustum box, identity rotation, opacity 0.1, isotropic scale
    (scene/dataset_readers.py:891-903, scene/gaussian_model.py:180-236);
  - "trained" cloud: depth sheets, ~1.5 px screen-space sigma, random rotations/opacities,
    small higher-order SH — gives the overdraw of an optimised scene.
All arrays are float32 numpy; `to_torch` moves them to a device.
"""
import math

import numpy as np

SH_C0 = 0.28209479177387814


def make_camera(width, height, fovx=1.2, depth_range=15.0, min_depth_fac=0.01, max_depth_fac=0.45,
                pose="identity", seed=0):
    znear = min_depth_fac * depth_range * 0.9
    zfar = max_depth_fac * depth_range * 1.1
    focal = width / (2.0 * math.tan(fovx / 2.0))
    fovy = 2.0 * math.atan(height / (2.0 * focal))
    tanx, tany = math.tan(fovx / 2.0), math.tan(fovy / 2.0)

    # world -> view (row-major, points as columns); identity for F-TöRF (scene/torf_utils.py:321)
    w2c = np.eye(4, dtype=np.float64)
    if pose != "identity":
        rng = np.random.default_rng(1000 + seed)
        ang = rng.normal(0.0, 0.15, size=3)
        cx, sx = math.cos(ang[0]), math.sin(ang[0])
        cy, sy = math.cos(ang[1]), math.sin(ang[1])
        cz, sz = math.cos(ang[2]), math.sin(ang[2])
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        w2c[:3, :3] = Rz @ Ry @ Rx
        w2c[:3, 3] = rng.normal(0.0, 0.2, size=3)

    top, right = tany * znear, tanx * znear
    Pm = np.zeros((4, 4), dtype=np.float64)
    Pm[0, 0] = 2.0 * znear / (2.0 * right)
    Pm[1, 1] = 2.0 * znear / (2.0 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)

    view_t = np.float32(w2c).T.copy()                      # world_view_transform
    proj_t = np.float32(Pm).T.copy()                       # projection_matrix
    full_t = (view_t @ proj_t).astype(np.float32)          # full_proj_transform
    campos = np.linalg.inv(view_t.astype(np.float64))[3, :3].astype(np.float32)
    return dict(width=int(width), height=int(height), tanfovx=tanx, tanfovy=tany,
                viewmatrix=view_t, projmatrix=full_t, campos=campos, znear=float(znear),
                zfar=float(zfar), depth_range=float(depth_range), focal=float(focal),
                c2w=np.linalg.inv(w2c))


def _to_world(cam, pts_cam):
    c2w = cam["c2w"]
    return (pts_cam @ c2w[:3, :3].T + c2w[:3, 3]).astype(np.float32)


def make_cloud(P, cam, kind="trained", seed=0, sh_coeffs=16, sigma_px=1.5, sheets=4):
    """Gaussian parameters in the reference's activated form (what render() hands the rasterizer:
    scales after exp, opacities after sigmoid, rotations normalised; gaussian_model.py:43,131)."""
    rng = np.random.default_rng(seed)
    znear, zfar = cam["znear"], cam["zfar"]
    tanx, tany = cam["tanfovx"], cam["tanfovy"]
    if kind == "init":
        # uniform in the frustum's bounding box; a share of the points falls outside the view
        z = rng.uniform(znear, zfar, size=P)
        x = rng.uniform(-tanx * zfar, tanx * zfar, size=P)
        y = rng.uniform(-tany * zfar, tany * zfar, size=P)
        pts_cam = np.stack([x, y, z], 1)
        vol = (2 * tanx * zfar) * (2 * tany * zfar) * (zfar - znear)
        # isotropic scale ~ sqrt(mean squared 3-NN distance) of a uniform cloud of this density
        s = 0.55 * (vol / max(P, 1)) ** (1.0 / 3.0) * np.exp(rng.normal(0.0, 0.2, size=P))
        scales = np.repeat(s[:, None], 3, 1)
        rots = np.zeros((P, 4)); rots[:, 0] = 1.0
        opac = np.full((P, 1), 0.1)
        shs = np.zeros((P, sh_coeffs, 3))
        shs_p = np.zeros((P, sh_coeffs, 2))
        shs_p[:, 0, 0] = (rng.uniform(0.0, 2.0 * math.pi, size=P) - 0.5) / SH_C0
        shs_p[:, 0, 1] = (0.1 - 0.5) / SH_C0
    else:
        zs = np.linspace(4.0 * znear, 0.8 * zfar, sheets)
        z = zs[rng.integers(0, sheets, size=P)] + rng.normal(0.0, 0.02 * zfar, size=P)
        z = np.clip(z, 0.5 * znear, 1.05 * zfar)       # a few land outside [near, far]
        x = rng.uniform(-1.1, 1.1, size=P) * tanx * z
        y = rng.uniform(-1.1, 1.1, size=P) * tany * z
        pts_cam = np.stack([x, y, z], 1)
        scales = (np.abs(z)[:, None] / cam["focal"]) * np.exp(
            rng.normal(math.log(sigma_px), 0.7, size=(P, 3)))
        q = rng.normal(size=(P, 4))
        rots = q / np.linalg.norm(q, axis=1, keepdims=True)
        opac = 1.0 / (1.0 + np.exp(-rng.normal(0.0, 2.0, size=(P, 1))))
        shs = rng.normal(0.0, 0.05, size=(P, sh_coeffs, 3))
        shs[:, 0, :] = (rng.uniform(0.0, 1.0, size=(P, 3)) - 0.5) / SH_C0
        shs_p = rng.normal(0.0, 0.05, size=(P, sh_coeffs, 2))
        shs_p[:, 0, 0] = (rng.uniform(0.0, 2.0 * math.pi, size=P) - 0.5) / SH_C0
        shs_p[:, 0, 1] = (rng.uniform(0.02, 0.5, size=P) * 4.0 - 0.5) / SH_C0
    f = np.float32
    return dict(means3D=_to_world(cam, pts_cam), scales=scales.astype(f), rotations=rots.astype(f),
                opacities=opac.astype(f), shs=shs.astype(f), shs_p=shs_p.astype(f))


def make_background(height, width, seed=0):
    """bg = rand(7,H,W)*2-1 seeded per iteration (train.py:123-124)."""
    rng = np.random.default_rng(7000 + seed)
    return (rng.random((7, height, width), dtype=np.float32) * 2.0 - 1.0).astype(np.float32)


def make_pixel_grads(height, width, seed=0):
    """Dense, non-trivial dL/d(out) for the five outputs the backward consumes."""
    rng = np.random.default_rng(9000 + seed)
    n = lambda c: rng.normal(0.0, 1.0, size=(c, height, width)).astype(np.float32)
    return dict(color=n(3), phasor=n(7), depth=n(1) * 0.1, acc=n(1), depth_distortion=n(1))


def to_torch(d, device):
    import torch
    out = {}
    for k, v in d.items():
        out[k] = torch.from_numpy(np.ascontiguousarray(v)).to(device) if isinstance(v, np.ndarray) else v
    return out
