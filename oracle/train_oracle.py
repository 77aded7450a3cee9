"""TEST INFRASTRUCTURE ONLY — CPU restatement of the PyTorch operator chains that sit either side
of the rasterizer in the reference's training iteration.  The reference implements
these steps in Python/PyTorch, so the restatement is PyTorch too (fp32, CPU or any device), written
from the cited lines; autograd through it IS the reference's backward.

  assemble()      gaussian_renderer/__init__.py:81-105 + scene/gaussian_model.py:35-43,123-157
  loss_term()     utils/loss_utils.py:17-33 (l1 / l2 / weighted variants), :70-114 (SSIM),
                  combined as in train.py:204-223
  adam_step()     torch.optim.Adam as configured at scene/gaussian_model.py:272
                  (lr per group, betas (0.9, 0.999), eps 1e-15, no weight decay, no amsgrad) —
                  a third-party algorithm (PyTorch, pinned 1.12.1 in the reference's
                  environment.yml:11); restated in numpy from the published update rule and pinned
                  against torch.optim.Adam of this image in tests/test_train_ops.py.

Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ---- f1 -----------------------------------------------------------------------------------------
def assemble(raw, motion_mask=None, deltas=None, isotropic=False, render_regions=("static", "dynamic")):
    """Returns the six rasterizer inputs.  `raw` holds the GaussianModel's parameters
    (xyz, opacity_raw, scaling_raw, rotation_raw, f_dc_color, f_rest_color, f_dc_phase,
    f_rest_phase, f_dc_amp, f_rest_amp); `deltas` the deformation outputs in masked order."""
    deltas = deltas or {}
    xyz = raw["xyz"]
    P = xyz.shape[0]
    if motion_mask is None:
        motion_mask = torch.zeros(P, dtype=torch.bool, device=xyz.device)
    # the model's getters (gaussian_model.py:123-157)
    scaling = torch.exp(raw["scaling_raw"].repeat(1, 3) if isotropic else raw["scaling_raw"])
    opacity = torch.sigmoid(raw["opacity_raw"])
    feat_color = torch.cat((raw["f_dc_color"], raw["f_rest_color"]), dim=1)
    feat_phasor = torch.cat((torch.cat((raw["f_dc_phase"], raw["f_dc_amp"]), dim=-1),
                             torch.cat((raw["f_rest_phase"], raw["f_rest_amp"]), dim=-1)), dim=1)
    out = dict(means3D=torch.zeros_like(xyz), opacities=torch.zeros_like(opacity),
               scales=torch.zeros_like(scaling), rotations=torch.zeros_like(raw["rotation_raw"]),
               shs=torch.zeros_like(feat_color), shs_p=torch.zeros_like(feat_phasor))
    m = motion_mask
    if "static" in render_regions:       # __init__.py:81-88
        out["means3D"][~m] = xyz[~m]
        out["opacities"][~m] = opacity[~m]
        out["scales"][~m] = scaling[~m]
        out["rotations"][~m] = F.normalize(raw["rotation_raw"])[~m]
        out["shs"][~m] = feat_color[~m]
        out["shs_p"][~m] = feat_phasor[~m]
    if "dynamic" in render_regions:      # __init__.py:89-96
        zero = lambda t: 0.0 if t is None else t
        out["means3D"][m] = xyz[m] + zero(deltas.get("d_xyz"))
        out["opacities"][m] = opacity[m]
        out["scales"][m] = scaling[m]
        out["rotations"][m] = F.normalize(raw["rotation_raw"][m] + zero(deltas.get("d_rot")))
        out["shs"][m] = feat_color[m] + zero(deltas.get("d_sh"))
        out["shs_p"][m] = feat_phasor[m] + zero(deltas.get("d_sh_p"))
    return out


# ---- f3 -----------------------------------------------------------------------------------------
def _window(channels, size=11, sigma=1.5):
    g = torch.tensor([math.exp(-(x - size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(size)])
    g = g / g.sum()
    w2 = g.unsqueeze(1).mm(g.unsqueeze(0)).float()
    return w2.expand(channels, 1, size, size).contiguous()


def ssim(img1, img2, size=11):
    """loss_utils.py:84-114 with size_average=True; inputs [C,H,W]."""
    C = img1.shape[-3]
    win = _window(C, size).to(img1)
    conv = lambda t: F.conv2d(t, win, padding=size // 2, groups=C)
    mu1, mu2 = conv(img1), conv(img2)
    s11 = conv(img1 * img1) - mu1 * mu1
    s22 = conv(img2 * img2) - mu2 * mu2
    s12 = conv(img1 * img2) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))
    return m.mean()


def elementwise_term(img, gt, kind, w=0.0, nch=None):
    if kind == "l1":                      # :17-18
        return (img - gt).abs().mean()
    if kind == "l2":                      # :20-21
        return ((img - gt) ** 2).mean()
    if kind == "weighted_l1":             # :23-25
        n = img.shape[0] if nch is None else nch
        weight = w + torch.sqrt(torch.sum(img ** 2, dim=0)).detach()
        return ((img[:n] - gt[:n]) / weight).abs().mean()
    if kind == "weighted_l1_quad":        # :27-29
        weight = w + img.detach().abs()
        return ((img - gt) / weight).abs().mean()
    if kind == "weighted_l2_quad":        # :31-33
        weight = w + img.detach().abs()
        return torch.square((img - gt) / weight).mean()
    raise ValueError(kind)


def loss_term(img, gt, kind="l1", lam=1.0, lambda_dssim=0.2, w=0.0, nch=None):
    """train.py:204-223: lam * ((1 - lambda_dssim) * L + lambda_dssim * (1 - ssim))."""
    L = elementwise_term(img, gt, kind, w, nch)
    if lambda_dssim == 0.0:
        return lam * L
    return lam * ((1.0 - lambda_dssim) * L + lambda_dssim * (1.0 - ssim(img, gt)))


def ssim_bruteforce(img1, img2, size=11, sigma=1.5):
    """Independent pin for ssim(): explicit loops over pixels and taps, float64, zero padding."""
    a, b = img1.double().numpy(), img2.double().numpy()
    C, H, W = a.shape
    g = np.array([math.exp(-(x - size // 2) ** 2 / (2 * sigma ** 2)) for x in range(size)])
    g = g / g.sum()
    h = size // 2
    tot = 0.0
    for c in range(C):
        for y in range(H):
            for x in range(W):
                mu1 = mu2 = xx = yy = xy = 0.0
                for dy in range(-h, h + 1):
                    for dx in range(-h, h + 1):
                        yy_, xx_ = y + dy, x + dx
                        if 0 <= yy_ < H and 0 <= xx_ < W:
                            wgt = g[dy + h] * g[dx + h]
                            u, v = a[c, yy_, xx_], b[c, yy_, xx_]
                            mu1 += wgt * u; mu2 += wgt * v
                            xx += wgt * u * u; yy += wgt * v * v; xy += wgt * u * v
                s1, s2, s12 = xx - mu1 * mu1, yy - mu2 * mu2, xy - mu1 * mu2
                tot += ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))
    return tot / (C * H * W)


# ---- f4 -----------------------------------------------------------------------------------------
def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-15):
    """One Adam update on float32 numpy arrays, in place; `step` counts from 1.
    m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;
    p <- p - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""
    f = np.float32
    m += (g - m) * f(1.0 - beta1)
    v *= f(beta2)
    v += f(1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
    p += f(-(lr / bc1)) * (m / denom)


# ---- f2 -----------------------------------------------------------------------------------------
GROUPS = ("xyz", "f_dc_color", "f_rest_color", "phase_f_dc", "phase_f_rest", "amp_f_dc", "amp_f_rest",
          "opacity", "scaling", "rotation", "f_seg_color")


def _rotation_matrix(r):
    """utils/general_utils.py:91-112 (normalises the quaternion first)."""
    q = r / torch.sqrt((r * r).sum(dim=1))[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.zeros((q.shape[0], 3, 3), dtype=r.dtype, device=r.device)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (y * z + w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def densify_and_prune(params, exp_avg, exp_avg_sq, grad_accum, denom, max_grad, min_opacity, extent,
                      percent_dense, size_prune=True, isotropic=False, normal_fn=None):
    """scene/gaussian_model.py:631-646 (densify_and_prune) with the clone (:607-629), split
    (:568-605), prune (:493-513) and optimizer-state surgery (:473-536) it calls, on plain tensors.
    `params`, `exp_avg`, `exp_avg_sq`: dicts over GROUPS.  `normal_fn(std)` stands for
    torch.normal(mean=0, std=std) (:581).  Returns the three dicts for the new Gaussian set; the
    densification statistics and max_radii2D restart from zero (:564-566), which also makes the
    screen-size criterion of :639 vacuous."""
    get_scaling = lambda p: torch.exp(p["scaling"].repeat(1, 3) if isotropic else p["scaling"])
    P = params["xyz"].shape[0]
    grads = grad_accum / denom
    grads[grads.isnan()] = 0.0
    p, m, v = dict(params), dict(exp_avg), dict(exp_avg_sq)

    def append(new):
        for g in GROUPS:
            p[g] = torch.cat((p[g], new[g]), dim=0)
            m[g] = torch.cat((m[g], torch.zeros_like(new[g])), dim=0)
            v[g] = torch.cat((v[g], torch.zeros_like(new[g])), dim=0)

    def prune(mask):
        keep = ~mask
        for g in GROUPS:
            p[g], m[g], v[g] = p[g][keep], m[g][keep], v[g][keep]

    # clone (:607-629)
    sel = torch.norm(grads, dim=-1) >= max_grad
    sel = sel & (get_scaling(p).max(dim=1).values <= percent_dense * extent)
    append({g: p[g][sel] for g in GROUPS})
    # split (:568-605), N = 2
    n_now = p["xyz"].shape[0]
    padded = torch.zeros(n_now, dtype=grads.dtype, device=grads.device)
    padded[:P] = grads.squeeze(-1)
    sel = (padded >= max_grad) & (get_scaling(p).max(dim=1).values > percent_dense * extent)
    stds = get_scaling(p)[sel].repeat(2, 1)
    samples = normal_fn(stds)
    rots = _rotation_matrix(p["rotation"][sel]).repeat(2, 1, 1)
    new = {g: p[g][sel].repeat(2, *([1] * (p[g].dim() - 1))) for g in GROUPS}
    new["xyz"] = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + p["xyz"][sel].repeat(2, 1)
    if isotropic:
        new["scaling"] = torch.log(torch.exp(p["scaling"][sel]).repeat(2, 1) / (0.8 * 2))
    else:
        new["scaling"] = torch.log(get_scaling(p)[sel].repeat(2, 1) / (0.8 * 2))
    append(new)
    prune(torch.cat((sel, torch.zeros(2 * int(sel.sum()), dtype=torch.bool, device=sel.device))))
    # prune (:636-644); max_radii2D was reset to zero by the two appends, so the screen-size term is False
    mask = (torch.sigmoid(p["opacity"]) < min_opacity).squeeze(-1)
    if size_prune:
        ws = get_scaling(p).max(dim=1).values
        mask = mask | (ws > 0.05 * extent) | (ws < 0.001 * extent)
    prune(mask)
    return p, m, v


# ---- the deformation MLP (bench.py's CPU pipeline baseline) -----------------------------------------
class DeformNetwork(torch.nn.Module):
    """utils/time_utils.py:56-127: positional encodings of xyz (10 frequencies) and t (6), an
    8-layer 256-wide ReLU MLP with a skip connection into layer 5, linear heads for d_xyz, d_rot
    and the four 16-coefficient SH deltas.  As in the reference's forward (:126-127), d_rot and
    the phase/amplitude SH deltas are returned as zeros; initialisation as initialize_weights
    (:87-104) with xavier_init_dxyz off."""

    def __init__(self, D=8, W=256, xyz_multires=10, t_multires=6, sh_degree=3):
        super().__init__()
        self.skips = [D // 2]
        self.xf = 2.0 ** torch.linspace(0.0, xyz_multires - 1, xyz_multires)
        self.tf = 2.0 ** torch.linspace(0.0, t_multires - 1, t_multires)
        in_ch = 3 + 3 * 2 * xyz_multires + 1 + 2 * t_multires
        self.linear = torch.nn.ModuleList(
            [torch.nn.Linear(in_ch, W)] +
            [torch.nn.Linear(W, W) if i not in self.skips else torch.nn.Linear(W + in_ch, W) for i in range(D - 1)])
        n = (1 + sh_degree) ** 2
        self.xyz_warp, self.rot = torch.nn.Linear(W, 3), torch.nn.Linear(W, 4)
        self.r, self.g, self.b, self.a = (torch.nn.Linear(W, n) for _ in range(4))
        for l in self.linear:
            torch.nn.init.xavier_normal_(l.weight)
            torch.nn.init.constant_(l.bias, 0.0)
        for l in (self.xyz_warp, self.rot, self.r, self.g, self.b, self.a):
            torch.nn.init.normal_(l.weight, mean=0.0, std=1e-5)
            torch.nn.init.constant_(l.bias, 0.0)

    @staticmethod
    def _embed(x, freqs):
        out = [x]
        for f in freqs:
            out += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(out, -1)

    def forward(self, x, t):
        emb = torch.cat([self._embed(x, self.xf), self._embed(t, self.tf)], dim=-1)
        h = emb
        for i, l in enumerate(self.linear):
            h = F.relu(l(h))
            if i in self.skips:
                h = torch.cat([emb, h], -1)
        d_xyz = self.xyz_warp(h)
        d_sh = torch.stack([self.r(h), self.g(h), self.b(h)], dim=-1)
        d_rot = torch.zeros_like(self.rot(h))
        d_sh_p = torch.zeros(d_sh.shape[0], d_sh.shape[1], 2, dtype=d_sh.dtype, device=d_sh.device)
        return d_xyz, d_rot, d_sh, d_sh_p
