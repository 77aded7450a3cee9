"""How good is the blend kernels' per-patch cull?  (analysis helper, CPU only, not a test)

For a bench workload's colour view it runs the CPU oracle's forward, decodes the per-tile lists
and counts, for a sample of tiles and each of the eight 8x4 patches a warp owns:

  candidates   list entries below the patch's furthest last-contributor (what the warp walks)
  box hits     entries whose conservative alpha >= 1/255 box overlaps the patch (what it evaluates)
  ellipse hits entries whose alpha >= 1/255 ELLIPSE overlaps the patch rectangle (an exact cull)
  replays      entries with at least one contributing pixel in the patch (what it has to replay)
  lanes        contributing pixels per replay

    python tests/analysis_cull_stats.py [workload] [n_tiles]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import bench  # noqa: E402
import cpu_oracle  # noqa: E402


def min_quadratic_over_rect(A, B, C, cx, cy, x0, x1, y0, y1):
    """min over the rectangle of q(d) = A dx^2 + 2 B dx dy + C dy^2, d = (x, y) - (cx, cy)."""
    inside = (cx >= x0) & (cx <= x1) & (cy >= y0) & (cy <= y1)
    best = np.full(A.shape, np.inf)
    for yy in (y0, y1):           # horizontal edges: minimise over x
        dy = yy - cy
        x = np.clip(cx - B * dy / A, x0, x1)
        dx = x - cx
        best = np.minimum(best, A * dx * dx + 2 * B * dx * dy + C * dy * dy)
    for xx in (x0, x1):           # vertical edges
        dx = xx - cx
        y = np.clip(cy - B * dx / C, y0, y1)
        dy = y - cy
        best = np.minimum(best, A * dx * dx + 2 * B * dx * dy + C * dy * dy)
    return np.where(inside, 0.0, best)


def main():
    wl_name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n_tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    wl = bench.WORKLOADS[wl_name]
    params, views = bench.build_scene(wl, 0, "cpu")
    v = views[0]
    e = torch.empty(0)
    out = cpu_oracle.rasterize_gaussians(
        v["bg"], params["means3D"], e, e, params["opacities"], params["scales"], params["rotations"], 1.0, e,
        v["viewmatrix"], v["projmatrix"], v["tanfovx"], v["tanfovy"], v["H"], v["W"], params["shs"],
        params["shs_p"], 3, v["campos"], False, False, v["near_n"], v["far_n"], v["depth_range"], True, 0.0, 0.0)
    st = cpu_oracle.decode(out[12])
    W, H = v["W"], v["H"]
    gx = (W + 15) // 16
    m2 = st["means2D"].numpy().astype(np.float64)
    co = st["conic_opacity"].numpy().astype(np.float64)
    pl = st["point_list"].numpy()
    rng = st["ranges"].numpy()
    ncontrib = st["n_contrib"].numpy().reshape(H, W)
    r = np.random.default_rng(0)
    lens = rng[:, 1] - rng[:, 0]
    tiles = r.choice(np.nonzero(lens > 0)[0], size=min(n_tiles, int((lens > 0).sum())), replace=False)
    tot = dict(cand=0, box=0, ell=0, replay=0, lanes=0)
    for t in tiles:
        ids = pl[rng[t, 0]:rng[t, 1]]
        n = len(ids)
        x, y = m2[ids, 0], m2[ids, 1]
        A, B, C, op = co[ids, 0], co[ids, 1], co[ids, 2], co[ids, 3]
        # alpha >= 1/255  <=>  q(d) <= t2 = 2 ln(255 op)   (power = -q/2)
        t2 = 2.0 * np.log(np.maximum(255.0 * op, 1e-30))
        det = A * C - B * B
        ex = np.sqrt(np.maximum(t2, 0) * C / det)
        ey = np.sqrt(np.maximum(t2, 0) * A / det)
        alive = op >= 1.0 / 255.0
        tx, ty = (t % gx) * 16, (t // gx) * 16
        pos = np.arange(n)
        for w in range(8):
            x0, y0 = tx + (w & 1) * 8, ty + (w >> 1) * 4
            xs, ys = np.arange(x0, x0 + 8), np.arange(y0, y0 + 4)
            okx, oky = xs < W, ys < H
            last = np.zeros((4, 8), dtype=np.int64)
            last[np.ix_(oky, okx)] = ncontrib[np.ix_(ys[oky], xs[okx])]
            wmax = last.max()
            cand = pos < wmax
            box = cand & alive & ~((x + ex < x0) | (x - ex > x0 + 7) | (y + ey < y0) | (y - ey > y0 + 3))
            qmin = min_quadratic_over_rect(A, B, C, x, y, x0, x0 + 7, y0, y0 + 3)
            ell = box & (qmin <= t2)
            k = np.nonzero(box)[0]
            dx = x[k, None, None] - xs[None, None, :]
            dy = y[k, None, None] - ys[None, :, None]
            power = -0.5 * (A[k, None, None] * dx * dx + C[k, None, None] * dy * dy) - B[k, None, None] * dx * dy
            alpha = np.minimum(0.99, op[k, None, None] * np.exp(power))
            contrib = (power <= 0) & (alpha >= 1.0 / 255.0) & (k[:, None, None] < last[None])
            lanes = contrib.reshape(len(k), -1).sum(1)
            tot["cand"] += int(cand.sum())
            tot["box"] += int(box.sum())
            tot["ell"] += int(ell.sum())
            tot["replay"] += int((lanes > 0).sum())
            tot["lanes"] += int(lanes.sum())
    print(wl_name, "tiles sampled:", len(tiles), tot)
    print("box hits / replay      %.2f" % (tot["box"] / tot["replay"]))
    print("ellipse hits / replay  %.2f" % (tot["ell"] / tot["replay"]))
    print("lanes / replay         %.1f of 32" % (tot["lanes"] / tot["replay"]))
    print("candidates / box hit   %.2f" % (tot["cand"] / tot["box"]))


if __name__ == "__main__":
    main()
