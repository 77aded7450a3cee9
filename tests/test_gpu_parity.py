"""GPU parity tests proper: the product library (through its C ABI, via the Python surface) against
  (1) the UNMODIFIED reference kernels (oracle/_ref/libgftorf_ref.so) on the same GPU,
  (2) the committed golden fixtures (reference outputs), and
  (3) the CPU restatement (oracle/libgft_oracle.so) where the reference itself is undefined.
Tolerances are north_star's: integers bit-exact, images <= 1e-5 abs, gradients <= 1e-4 rel-L2.
"""
import math
import os

import pytest
import torch

import golden_util
import harness
from gftorf_b200 import rasterizer, debug, distCUDA2
from oracle import ref_driver, cpu_oracle

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref not built")
needs_oracle = pytest.mark.skipif(not cpu_oracle.available(), reason="CPU oracle not built")

CASES = {
    "tiny": dict(P=64, W=48, H=40, kind="trained", seed=0, sigma_px=4.0),
    "ragged": dict(P=777, W=101, H=67, kind="trained", seed=5, sigma_px=3.0, pose="orbit"),
    "small_orbit": dict(P=2000, W=160, H=120, kind="trained", seed=1, pose="orbit"),
    "init_small": dict(P=2000, W=160, H=120, kind="init", seed=2),
    "c1": dict(P=20000, W=320, H=240, kind="trained", seed=0),
    "c1_init": dict(P=20000, W=320, H=240, kind="init", seed=1),
    "c1_dense": dict(P=20000, W=320, H=240, kind="trained", seed=2, sigma_px=6.0),
    # active SH degree below the stored one: training starts at 0 and adds a degree per 1000
    # iterations (train.py:152-153); coefficients above the active degree must get zero gradients
    "deg0": dict(P=3000, W=160, H=120, kind="trained", seed=6, pose="orbit", sh_degree=0),
    "deg1": dict(P=3000, W=160, H=120, kind="trained", seed=7, pose="orbit", sh_degree=1),
    "deg2": dict(P=3000, W=160, H=120, kind="trained", seed=8, pose="orbit", sh_degree=2),
    "c2": dict(P=300000, W=640, H=480, kind="trained", seed=0),
    "c2_vdp": dict(P=300000, W=640, H=480, kind="trained", seed=3, pose="orbit",
                   view_dependent_phase=True, phase_offset=0.3, dc_offset=0.1),
    "c3_tof": dict(P=500000, W=320, H=240, kind="trained", seed=4, depth_range=10.0,
                   bg_hw=(480, 640)),
    # BASELINE configs[2], the colour view (TöRF real-shaped: 500k Gaussians, depth_range 10)
    "c3_color": dict(P=500000, W=640, H=480, kind="trained", seed=4, depth_range=10.0),
    # configs[1] at initialisation: uniform cloud, isotropic splats with a median radius of ~50 px
    # (R/V ~ 25): thousands of instances per tile, the long-segment path of the tile sort
    "c2_init": dict(P=300000, W=640, H=480, kind="init", seed=5),
}


def run_both(spec, **fw):
    inp = harness.build_inputs(device="cuda", **spec)
    ours = harness.call_forward(rasterizer._C, inp, **fw)
    ref = harness.call_forward(ref_driver.RefModule, inp, **fw)
    torch.cuda.synchronize()
    return inp, ours, ref


def decoded(inp, ours, ref):
    P, W, H = inp["P"], inp["W"], inp["H"]
    od = debug.decode_buffers(ours[12], ours[13], ours[14], P, ours[0], W, H)
    rd = ref_driver.decode_buffers(ref[12], ref[13], ref[14], P, ref[0], W, H)
    return od, rd


@needs_ref
@pytest.mark.parametrize("name", list(CASES))
def test_forward_vs_reference_kernels(name):
    inp, ours, ref = run_both(CASES[name])
    od, rd = decoded(inp, ours, ref)
    rep = harness.compare_forward(ours, ref, od, rd)
    harness.assert_forward_parity(rep)
    assert rep["R_ours"] > 0 and rep["V"] > 0


@needs_ref
@pytest.mark.parametrize("name", ["tiny", "ragged", "small_orbit", "init_small", "deg0", "deg1", "deg2",
                                  "c1", "c1_dense", "c2", "c2_vdp", "c3_tof", "c3_color", "c2_init"])
def test_backward_vs_reference_kernels(name):
    # zero higher-order phase/amp SH keeps the reference's undefined dL_dPA out of
    # dL_dmeans3D; dL_dsh_p is compared on the one row the reference defines (Gaussian 0).
    spec = dict(CASES[name], zero_shp_rest=True)
    inp, ours, ref = run_both(spec)
    ob = harness.call_backward(rasterizer._C, inp, ours)
    rb = harness.call_backward(ref_driver.RefModule, inp, ref)
    torch.cuda.synchronize()
    scale = float(rb[8].double().norm())
    for i, k in enumerate(harness.BWD_NAMES):
        if k in ("colors_precomp", "phasors_precomp", "cov3Ds_precomp"):
            continue
        if k == "sh_p":
            assert harness.rel_l2(ob[i][0], rb[i][0]) <= harness.GRAD_REL_L2
            continue
        err = float((ob[i].double() - rb[i].double()).norm())
        refn = float(rb[i].double().norm())
        # the two scalar offsets are single float32 sums over all Gaussians of signed terms, which
        # the reference accumulates with one atomic per Gaussian in arbitrary order
        # (backward.cu:556,567): their own run-to-run spread is a few 1e-4 on large clouds
        tol = 1e-3 if k in ("phase_offset", "dc_offset") else harness.GRAD_REL_L2
        assert err <= tol * max(refn, 1e-3 * scale), (k, err, refn)


@needs_ref
def test_sh_p_gradient_rows_by_rotating_each_gaussian_to_index_zero():
    """The reference defines dL_dsh_p only for Gaussian 0.  Swap Gaussian k with Gaussian 0, read
    the reference's row 0, and compare with OUR row k of the unswapped scene."""
    spec = dict(P=300, W=64, H=48, kind="trained", seed=7, sigma_px=3.0, view_dependent_phase=True,
                phase_offset=0.2, dc_offset=0.05)
    inp = harness.build_inputs(device="cuda", **spec)
    ours = harness.call_forward(rasterizer._C, inp)
    ob = harness.call_backward(rasterizer._C, inp, ours)
    vis = (ours[11] > 0).nonzero().flatten().tolist()
    checked = 0
    for k in vis[:12]:
        perm = torch.arange(inp["P"], device="cuda")
        perm[0], perm[k] = k, 0
        sw = dict(inp)
        for name in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p"):
            sw[name] = inp[name][perm].contiguous()
        ref = harness.call_forward(ref_driver.RefModule, sw)
        rb = harness.call_backward(ref_driver.RefModule, sw, ref)
        torch.cuda.synchronize()
        if float(rb[7][0].norm()) == 0.0:
            continue
        assert harness.rel_l2(ob[7][k], rb[7][0]) <= harness.GRAD_REL_L2, k
        checked += 1
    assert checked >= 6


def _oracle_run(inp):
    cin = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    cin["grads"] = {k: v.cpu() for k, v in inp["grads"].items()}
    cf = harness.call_forward(cpu_oracle.OracleModule, cin)
    cb = harness.call_backward(cpu_oracle.OracleModule, cin, cf)
    return cf, cb


def _assert_gradients_vs_oracle(ob, cb, tol):
    scale = float(cb[8].double().norm())
    for i, k in enumerate(harness.BWD_NAMES):
        if k in ("colors_precomp", "phasors_precomp", "cov3Ds_precomp"):
            continue
        a, b = ob[i].cpu().double(), cb[i].double()
        err = float((a - b).norm())
        refn = float(b.norm())
        assert err <= tol * max(refn, 1e-3 * scale), (k, err, refn)


@needs_oracle
@pytest.mark.parametrize("spec", [
    dict(P=3000, W=128, H=96, kind="trained", pose="orbit", view_dependent_phase=True,
         phase_offset=0.3, dc_offset=0.1),
    dict(P=1500, W=96, H=80, kind="init"),
])
def test_full_gradients_vs_cpu_oracle(spec):
    """Every gradient, including all rows of dL_dsh_p and dL_dmeans3D with non-zero SH_p, against
    the CPU restatement (which is pinned to the reference on everything the reference defines).
    libm's expf is not CUDA's, so a pair sitting exactly on the alpha >= 1/255 threshold can flip
    between the two; such a scene is not skipped: the next seed is taken, and one of the first
    four seeds must be flip-free."""
    checked = 0
    for seed in (11, 12, 13, 14):
        inp = harness.build_inputs(device="cuda", seed=seed, **spec)
        ours = harness.call_forward(rasterizer._C, inp)
        ob = harness.call_backward(rasterizer._C, inp, ours)
        cf, cb = _oracle_run(inp)
        assert ours[0] == cf[0]
        assert torch.equal(ours[11].cpu(), cf[11])
        flips = harness.nmismatch(ours[9].cpu(), cf[9])
        assert flips <= 2, flips
        if flips:
            continue
        _assert_gradients_vs_oracle(ob, cb, harness.GRAD_REL_L2)
        checked += 1
        break
    assert checked == 1


@needs_oracle
def test_c2_size_gradients_with_live_sh_p_vs_cpu_oracle():
    """BASELINE configs[1] size (300k Gaussians, 640x480), ALL SH_p coefficients non-zero: the
    reference's dL_dsh_p / the SH_p part of dL_dmeans3D are undefined there, so the
    full-size check of those rows is against the CPU restatement — every gradient tensor as a
    whole, and separately 4000 sampled rows so a localised error cannot hide in the norm.  A few
    threshold flips (libm vs CUDA expf) are expected at this size; they perturb single pixels by
    alpha ~ 1/255, which the loosened tolerance below covers — the comparison always runs."""
    spec = dict(P=300000, W=640, H=480, kind="trained", seed=0, view_dependent_phase=True,
                phase_offset=0.2, dc_offset=0.05)
    inp = harness.build_inputs(device="cuda", **spec)
    assert float(inp["shs_p"][:, 1:, :].abs().max()) > 0
    ours = harness.call_forward(rasterizer._C, inp)
    ob = harness.call_backward(rasterizer._C, inp, ours)
    cf, cb = _oracle_run(inp)
    assert ours[0] == cf[0]
    assert torch.equal(ours[11].cpu(), cf[11])
    flips = harness.nmismatch(ours[9].cpu(), cf[9])
    assert flips <= 64, flips
    tol = harness.GRAD_REL_L2 if flips == 0 else 5e-4
    _assert_gradients_vs_oracle(ob, cb, tol)
    vis = (cf[11] > 0).nonzero().flatten()
    g = torch.Generator().manual_seed(1)
    rows = vis[torch.randperm(vis.numel(), generator=g)[:4000]]
    for i, k in ((7, "sh_p"), (4, "means3D"), (6, "sh")):
        a, b = ob[i].cpu().double()[rows], cb[i].double()[rows]
        assert float((a - b).norm()) <= tol * float(b.norm()), (k, flips)
        assert float(b.norm()) > 0


@pytest.mark.parametrize("name", golden_util.CASES)
def test_golden_fixtures(name):
    """Product vs committed reference outputs (no reference library needed at run time)."""
    inp, gold = golden_util.load(name, device="cuda")
    f = harness.call_forward(rasterizer._C, inp)
    b = harness.call_backward(rasterizer._C, inp, f)
    torch.cuda.synchronize()
    d = debug.decode_buffers(f[12], f[13], f[14], inp["P"], f[0], inp["W"], inp["H"])
    assert f[0] == gold["R"]
    assert torch.equal(f[11].cpu(), gold["f_radii"])
    assert torch.equal(f[9].cpu(), gold["f_pixels"])
    for k in ("tiles_touched", "point_offsets", "keys", "point_list", "ranges", "n_contrib"):
        if "s_" + k in gold:
            assert torch.equal(d[k].cpu(), gold["s_" + k]), k
    for i, k in enumerate(harness.FWD_NAMES):
        if i >= 1 and k not in ("pixels", "radii"):
            assert (f[i].cpu() - gold["f_" + k]).abs().max() <= harness.IMG_ATOL, k
    zero_rest = bool(gold["spec"].get("zero_shp_rest", False)) or gold["spec"]["kind"] == "init"
    scale = float(gold["b_scales"].double().norm())
    for i, k in enumerate(harness.BWD_NAMES):
        if k in ("colors_precomp", "phasors_precomp", "cov3Ds_precomp"):
            continue
        if k == "sh_p":
            assert harness.rel_l2(b[i][0].cpu(), gold["b_sh_p"][0]) <= harness.GRAD_REL_L2
            continue
        if k == "means3D" and not zero_rest:
            continue
        err = float((b[i].cpu().double() - gold["b_" + k].double()).norm())
        refn = float(gold["b_" + k].double().norm())
        assert err <= harness.GRAD_REL_L2 * max(refn, 1e-3 * scale), (k, err, refn)


# ------------------------------------------------------------------------------------------------
# edge cases the reference's binding handles (rasterize_points.cu:104,238)
# ------------------------------------------------------------------------------------------------
@needs_ref
def test_no_gaussians():
    inp = harness.build_inputs(device="cuda", P=0, W=40, H=24)
    ours = harness.call_forward(rasterizer._C, inp)
    ref = harness.call_forward(ref_driver.RefModule, inp)
    assert ours[0] == 0 == ref[0]
    for i in range(1, 12):
        assert ours[i].shape == ref[i].shape
        assert torch.equal(ours[i], ref[i])  # all zeros, NOT T*bg (rasterize_points.cu:104)
    ob = harness.call_backward(rasterizer._C, inp, ours)
    rb = harness.call_backward(ref_driver.RefModule, inp, ref)
    for i, (a, b) in enumerate(zip(ob, rb)):
        if harness.BWD_NAMES[i] == "phasors_precomp":
            continue  # not produced by the product
        assert a.shape == b.shape and torch.equal(a, b)


@needs_ref
def test_everything_culled_renders_background():
    inp = harness.build_inputs(device="cuda", P=500, W=40, H=24, seed=3)
    inp["means3D"] = inp["means3D"] * torch.tensor([1.0, 1.0, -1.0], device="cuda")  # behind camera
    ours = harness.call_forward(rasterizer._C, inp)
    ref = harness.call_forward(ref_driver.RefModule, inp)
    assert ours[0] == 0 == ref[0]
    assert int((ours[11] != 0).sum()) == 0
    for i in range(1, 12):
        assert torch.equal(ours[i], ref[i])
    assert torch.equal(ours[1], inp["bg"][0:3])           # T = 1: colour = bg planes 0..2
    assert torch.equal(ours[2], inp["bg"][0:7])
    ob = harness.call_backward(rasterizer._C, inp, ours)
    for g in ob:
        assert g is None or float(g.abs().sum()) == 0.0


@needs_ref
def test_non_finite_positions_are_dropped_like_the_reference():
    """NaN / inf positions never reach the sort: their radius converts to 0 (forward.cu:272-278) and
    the Gaussian is dropped; everything else must stay bit-identical and their gradient rows zero."""
    inp = harness.build_inputs(device="cuda", P=3000, W=96, H=64, seed=12)
    bad = torch.arange(0, 3000, 37, device="cuda")
    inp["means3D"] = inp["means3D"].clone()
    inp["means3D"][bad[0::3]] = float("nan")
    inp["means3D"][bad[1::3], 2] = float("inf")
    inp["means3D"][bad[2::3], 0] = float("nan")
    ours = harness.call_forward(rasterizer._C, inp)
    ref = harness.call_forward(ref_driver.RefModule, inp)
    od, rd = decoded(inp, ours, ref)
    harness.assert_forward_parity(harness.compare_forward(ours, ref, od, rd))
    assert int((ours[11][bad] != 0).sum()) == 0 and ours[0] > 0
    for i in range(1, 11):
        assert bool(torch.isfinite(ours[i]).all())
    ob = harness.call_backward(rasterizer._C, inp, ours)
    for g in ob:
        if g is not None and g.dim() >= 1 and g.shape[0] == 3000:
            assert bool(torch.isfinite(g).all()) and float(g[bad].abs().sum()) == 0.0


@needs_ref
def test_single_gaussian():
    inp = harness.build_inputs(device="cuda", P=1, W=33, H=17, seed=9, sigma_px=6.0)
    inp["means3D"] = torch.tensor([[0.02, -0.01, 2.0]], device="cuda")
    ours = harness.call_forward(rasterizer._C, inp)
    ref = harness.call_forward(ref_driver.RefModule, inp)
    od, rd = decoded(inp, ours, ref)
    harness.assert_forward_parity(harness.compare_forward(ours, ref, od, rd))
    ob = harness.call_backward(rasterizer._C, inp, ours)
    rb = harness.call_backward(ref_driver.RefModule, inp, ref)
    # P = 1: index 0 is the only Gaussian, so even dL_dsh_p / dL_dmeans3D are defined in the reference
    harness.assert_backward_parity(harness.compare_backward(ob, rb))


@needs_ref
def test_render_flow_path_precomputed_colours_no_phasor():
    """gaussian_renderer/__init__.py:194-202: colors_precomp, no shs_p.  The reference blends
    uninitialised phasor features there; only colour/depth/acc/dd/radii/pixels and
    dL_dcolors are meaningful.  We define the phasor features as 0."""
    inp = harness.build_inputs(device="cuda", P=4000, W=96, H=64, seed=4, sigma_px=3.0)
    cp = torch.rand((4000, 3), device="cuda")
    inp["bg"] = torch.zeros_like(inp["bg"])  # bg_map_flow is all zeros, train.py:128
    ours = harness.call_forward(rasterizer._C, inp, colors_precomp=cp, use_shs_p=False)
    ref = harness.call_forward(ref_driver.RefModule, inp, colors_precomp=cp, use_shs_p=False)
    assert ours[0] == ref[0]
    assert torch.equal(ours[11], ref[11]) and torch.equal(ours[9], ref[9])
    for i in (1, 3, 5, 7):
        assert torch.equal(ours[i], ref[i])
    assert float(ours[2].abs().max()) == 0.0
    inp["grads"]["phasor"] = torch.zeros_like(inp["grads"]["phasor"])
    ob = harness.call_backward(rasterizer._C, inp, ours, colors_precomp=cp, use_shs_p=False)
    rb = harness.call_backward(ref_driver.RefModule, inp, ref, colors_precomp=cp, use_shs_p=False)
    for t in ob:
        assert t is None or bool(torch.isfinite(t).all())
    if all(bool(torch.isfinite(t).all()) for t in rb):
        for i in (0, 1, 3, 8, 9):  # means2D, colors_precomp, opacities, scales, rotations
            assert harness.rel_l2(ob[i], rb[i]) <= harness.GRAD_REL_L2, harness.BWD_NAMES[i]
    else:
        # the reference multiplied uninitialised (NaN/Inf) phasor features by the zero phasor
        # gradient: its gradients are undefined on this run.  Ours are finite (checked above) and
        # are checked against the CPU restatement instead, which defines those features as 0 too.
        cin = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
        cin["grads"] = {k: v.cpu() for k, v in inp["grads"].items()}
        cf = harness.call_forward(cpu_oracle.OracleModule, cin, colors_precomp=cp.cpu(), use_shs_p=False)
        cb = harness.call_backward(cpu_oracle.OracleModule, cin, cf, colors_precomp=cp.cpu(), use_shs_p=False)
        tol = harness.GRAD_REL_L2 if harness.nmismatch(ours[9].cpu(), cf[9]) == 0 else 1e-3
        for i in (0, 1, 3, 8, 9):
            assert harness.rel_l2(ob[i].cpu(), cb[i]) <= tol, harness.BWD_NAMES[i]


@needs_ref
def test_precomputed_covariance_path():
    inp = harness.build_inputs(device="cuda", P=3000, W=96, H=64, seed=6, sigma_px=3.0,
                               zero_shp_rest=True)
    ours = harness.call_forward(rasterizer._C, inp)
    od = debug.decode_buffers(ours[12], ours[13], ours[14], 3000, ours[0], 96, 64)
    cov = od["cov3D"].clone()
    cov[ours[11] <= 0] = 0
    e = inp["empty"]

    def fwd(mod):
        return mod.rasterize_gaussians(
            inp["bg"], inp["means3D"], e, e, inp["opacities"], e, e, 1.0, cov, inp["viewmatrix"],
            inp["projmatrix"], inp["tanfovx"], inp["tanfovy"], 64, 96, inp["shs"], inp["shs_p"], 3,
            inp["campos"], False, False, inp["near_n"], inp["far_n"], inp["depth_range"], False,
            0.0, 0.0)
    a, b = fwd(rasterizer._C), fwd(ref_driver.RefModule)
    assert a[0] == b[0] and torch.equal(a[11], b[11])
    for i in range(1, 11):
        assert torch.equal(a[i], b[i])
    g = inp["grads"]
    z1 = torch.zeros_like(g["depth"])

    def bwd(mod, f):
        return mod.rasterize_gaussians_backward(
            inp["bg"], inp["means3D"], f[11], e, e, e, e, 1.0, cov, inp["viewmatrix"],
            inp["projmatrix"], inp["tanfovx"], inp["tanfovy"], g["color"], g["phasor"], g["depth"],
            torch.zeros_like(g["color"]), g["acc"], z1, g["depth_distortion"], z1, inp["shs"],
            inp["shs_p"], 3, inp["campos"], f[12], f[0], f[13], f[14], False, inp["near_n"],
            inp["far_n"], inp["depth_range"], False, 0.0, 0.0)
    ga, gb = bwd(rasterizer._C, a), bwd(ref_driver.RefModule, b)
    assert harness.rel_l2(ga[5], gb[5]) <= harness.GRAD_REL_L2      # dL_dcov3D
    assert harness.rel_l2(ga[4], gb[4]) <= harness.GRAD_REL_L2      # dL_dmeans3D
    assert float(ga[8].abs().sum()) == 0.0 and float(gb[8].abs().sum()) == 0.0


@needs_ref
def test_constant_background_is_not_materialised_and_matches():
    inp = harness.build_inputs(device="cuda", P=2000, W=80, H=48, seed=8, sigma_px=3.0)
    const = torch.tensor([0.1, -0.2, 0.3, 0.4, -0.5, 0.6, 0.7], device="cuda")
    inp["bg"] = const.view(7, 1, 1).expand(7, 48, 80)   # train.py:127
    ours = harness.call_forward(rasterizer._C, inp)
    ref = harness.call_forward(ref_driver.RefModule, inp)
    for i in range(1, 12):
        assert torch.equal(ours[i], ref[i])
    ob = harness.call_backward(rasterizer._C, inp, ours)
    rb = harness.call_backward(ref_driver.RefModule, inp, ref)
    assert harness.rel_l2(ob[3], rb[3]) <= harness.GRAD_REL_L2


def test_tile_segments_are_sorted_by_depth_then_index():
    """The tile-segmented binning: every tile's slice of the sorted entries is ordered by
    (float_bits(view_z), Gaussian index) — the order of the reference's stable radix sort of
    (tile << 32 | depth) keys — carries exactly its Gaussians' depth bits, and `ranges` is the
    exclusive scan of the per-tile instance counts with (0,0) for empty tiles."""
    inp = harness.build_inputs(device="cuda", **CASES["c1"])
    f = harness.call_forward(rasterizer._C, inp)
    d = debug.decode_buffers(f[12], f[13], f[14], inp["P"], f[0], inp["W"], inp["H"])
    k = d["keys"]
    assert bool((k[1:] >= k[:-1]).all())
    same = k[1:] == k[:-1]
    assert bool((d["point_list"][1:][same] > d["point_list"][:-1][same]).all())
    g = d["point_list"].long()
    assert torch.equal(d["entries"] & 0xffffffff, g)
    zbits = d["depths"].view(torch.int32).long()[g]
    assert torch.equal(k & 0xffffffff, zbits)
    cnt = d["tile_counts"].long()
    assert int(cnt.sum()) == f[0] == d["num_rendered"]
    x = torch.cumsum(cnt, 0) - cnt
    rng = d["ranges"].long()
    nz = cnt > 0
    assert torch.equal(rng[nz, 0], x[nz]) and torch.equal(rng[nz, 1], (x + cnt)[nz])
    assert int(rng[~nz].abs().sum()) == 0
    # tile membership: Gaussian g is listed in tile t iff t lies in g's tile rectangle
    rect = d["rect"].long()
    T = rng.shape[0]
    gx = (inp["W"] + 15) // 16
    tile_of = torch.repeat_interleave(torch.arange(T, device="cuda"), cnt)
    tx, ty = tile_of % gx, tile_of // gx
    r = rect[g]
    assert bool(((tx >= r[:, 0]) & (tx < r[:, 2]) & (ty >= r[:, 1]) & (ty < r[:, 3])).all())
    assert torch.equal(d["tiles_touched"].long(), torch.bincount(g, minlength=inp["P"]))


@needs_ref
@pytest.mark.parametrize("deg,M", [(0, 1), (1, 4), (2, 9), (1, 16)])
def test_truncated_sh_tensors_vs_reference_kernels(deg, M):
    """max_sh_degree < 3: the SH tensors hold only (deg+1)^2 coefficients (M != 16 takes the
    unstaged path of both preprocess kernels).  Forward bit-exact, gradients within tolerance."""
    inp = harness.build_inputs(device="cuda", P=2500, W=128, H=96, kind="trained", seed=10 + M, pose="orbit",
                               sh_degree=deg, zero_shp_rest=True)
    inp["shs"] = inp["shs"][:, :M, :].contiguous()
    inp["shs_p"] = inp["shs_p"][:, :M, :].contiguous()
    ours = harness.call_forward(rasterizer._C, inp)
    ref = harness.call_forward(ref_driver.RefModule, inp)
    assert ours[0] == ref[0]
    for i in range(1, 12):
        assert torch.equal(ours[i], ref[i]), harness.FWD_NAMES[i]
    ob = harness.call_backward(rasterizer._C, inp, ours)
    rb = harness.call_backward(ref_driver.RefModule, inp, ref)
    scale = float(rb[8].double().norm())
    for i, k in enumerate(harness.BWD_NAMES):
        if k in ("colors_precomp", "phasors_precomp", "cov3Ds_precomp"):
            continue
        a, b = (ob[i][0], rb[i][0]) if k == "sh_p" else (ob[i], rb[i])
        assert a.shape == b.shape, k
        err = float((a.double() - b.double()).norm())
        assert err <= harness.GRAD_REL_L2 * max(float(b.double().norm()), 1e-3 * scale), (k, err)


@pytest.mark.parametrize("name", ["c1_init", "c1_dense"])
def test_tile_sort_paths_agree_bitwise(name):
    """The tile sort keeps segments of up to `sort_cap` entries in shared memory and runs the wide
    stages of longer ones on global memory.  Forcing tiny capacities sends every tile of an
    initialisation-like cloud (thousands of instances per tile) through the long-segment path;
    the sorted lists and the images must not change by a bit."""
    from gftorf_b200 import _capi
    inp = harness.build_inputs(device="cuda", **CASES[name])
    outs = {}
    # (sort_cap, sort_radix, sort_adapt)
    configs = [(0, 1, 1), (256, 1, 1), (1024, 1, 1), (8192, 1, 1), (0, 0, 1), (256, 0, 1), (0, 1, 0), (8192, 1, 0)]
    for cap, radix, adapt in configs:
        old = _capi.set_option("sort_cap", cap)
        old_r = _capi.set_option("sort_radix", radix)
        old_a = _capi.set_option("sort_adapt", adapt)
        try:
            f = harness.call_forward(rasterizer._C, inp)
            d = debug.decode_buffers(f[12], f[13], f[14], inp["P"], f[0], inp["W"], inp["H"])
            outs[(cap, radix, adapt)] = (d["keys"].clone(), d["point_list"].clone(), f[1].clone(), f[2].clone())
        finally:
            _capi.set_option("sort_cap", old)
            _capi.set_option("sort_radix", old_r)
            _capi.set_option("sort_adapt", old_a)
    assert int(d["tile_counts"].max()) > 256
    for cfg in configs[1:]:
        for a, b in zip(outs[configs[0]], outs[cfg]):
            assert torch.equal(a, b), cfg


@needs_ref
@pytest.mark.parametrize("spread", ["wide", "narrow", "clustered", "two_values"])
def test_tile_sort_over_depth_spreads(spread):
    """The per-tile radix sort looks only at the bits in which a tile's depths differ (at most the
    top 24 of them) and leaves the rest to a short odd-even fix-up; depth distributions that
    stress each branch: depths over nine binades (more than 24 varying bits, low bits left to the
    fix-up), depths a few ulps apart (one or two passes), thousands of depths inside one bucket
    next to far outliers (the fix-up gives up and the bitonic network sorts), and two distinct
    values only.  The lists must be the reference's, bit for bit."""
    inp = harness.build_inputs(device="cuda", P=40000, W=160, H=120, kind="trained", seed=82, sigma_px=3.0)
    P = inp["P"]
    g = torch.Generator(device="cpu").manual_seed(5)
    z0 = inp["means3D"][:, 2].clone()
    if spread == "wide":
        zt = torch.exp(torch.empty(P).uniform_(math.log(0.25), math.log(120.0), generator=g)).cuda()
    elif spread == "narrow":
        zt = (2.0 + torch.randint(0, 3000, (P,), generator=g).float() * 2.3841858e-07).cuda()
    elif spread == "clustered":
        zt = (3.0 + torch.randint(0, 6, (P,), generator=g).float() * 2.3841858e-07).cuda()
        far = torch.arange(P, device="cuda") % 50 == 0
        zt = torch.where(far, torch.full_like(zt, 90.0), zt)
        zt = torch.where(torch.arange(P, device="cuda") % 50 == 1, torch.full_like(zt, 0.3), zt)
    else:
        zt = torch.where(torch.arange(P, device="cuda") % 2 == 0, torch.tensor(1.5, device="cuda"),
                         torch.tensor(1.5000001, device="cuda"))
    f = (zt / z0).unsqueeze(1)
    inp["means3D"] = inp["means3D"] * f              # same pixel, new depth (identity camera)
    inp["scales"] = inp["scales"] * f
    ours = harness.call_forward(rasterizer._C, inp)
    ref = harness.call_forward(ref_driver.RefModule, inp)
    od, rd = decoded(inp, ours, ref)
    assert ours[0] == ref[0] and ours[0] > 20000
    assert torch.equal(od["point_list"], rd["point_list"])
    assert torch.equal(od["keys"], rd["keys"])
    for i in range(1, 12):
        assert torch.equal(ours[i], ref[i]), harness.FWD_NAMES[i]


def test_equal_depths_keep_ascending_index_order():
    """Gaussians at EXACTLY the same view depth: the reference's stable sort lists them in
    ascending index order inside every tile; the tile sort must too (its radix passes see only the
    depth bits, the slots come from atomics — the tie check has to catch it)."""
    inp = harness.build_inputs(device="cuda", P=6000, W=160, H=120, kind="trained", seed=81, sigma_px=4.0)
    m = inp["means3D"].clone()
    m[:, 2] = torch.where(torch.arange(6000, device="cuda") % 3 == 0, torch.full_like(m[:, 2], 2.0), m[:, 2])
    inp["means3D"] = m                      # identity camera: view z = world z
    for radix in (1, 0):
        from gftorf_b200 import _capi
        old = _capi.set_option("sort_radix", radix)
        try:
            f = harness.call_forward(rasterizer._C, inp)
        finally:
            _capi.set_option("sort_radix", old)
        d = debug.decode_buffers(f[12], f[13], f[14], inp["P"], f[0], inp["W"], inp["H"])
        k = d["keys"]
        same = k[1:] == k[:-1]
        assert int(same.sum()) > 1000
        assert bool((k[1:] >= k[:-1]).all())
        assert bool((d["point_list"][1:][same] > d["point_list"][:-1][same]).all())
        if ref_driver.available():
            ref = harness.call_forward(ref_driver.RefModule, inp)
            rd = ref_driver.decode_buffers(ref[12], ref[13], ref[14], inp["P"], ref[0], inp["W"], inp["H"])
            assert torch.equal(d["point_list"], rd["point_list"])
            for i in range(1, 12):
                assert torch.equal(f[i], ref[i]), harness.FWD_NAMES[i]


def test_subtile_culling_is_exact():
    """The conservative per-warp culling must not change a single bit of the result."""
    for name in ("c1", "c1_init", "c1_dense"):
        inp = harness.build_inputs(device="cuda", **CASES[name])
        from gftorf_b200 import _capi
        a = harness.call_forward(rasterizer._C, inp)
        _capi.set_option("no_cull", 1)
        try:
            b = harness.call_forward(rasterizer._C, inp)
        finally:
            _capi.set_option("no_cull", 0)
        assert not torch.equal(debug.decode_buffers(a[12], a[13], a[14], inp["P"], a[0], inp["W"], inp["H"])["extents"],
                               debug.decode_buffers(b[12], b[13], b[14], inp["P"], b[0], inp["W"], inp["H"])["extents"])
        for i in range(1, 12):
            assert torch.equal(a[i], b[i]), (name, harness.FWD_NAMES[i])
        ga = harness.call_backward(rasterizer._C, inp, a)
        gb = harness.call_backward(rasterizer._C, inp, b)
        scale = float(ga[8].double().norm())
        for x, y in zip(ga, gb):
            if x is None:
                continue
            err = float((x.double() - y.double()).norm())
            # two runs differ by the order of the floating-point atomics only
            assert err <= harness.GRAD_REL_L2 * max(float(y.double().norm()), 1e-2 * scale)


@pytest.mark.parametrize("name", ["ragged", "c1", "c1_init"])
def test_blend_backward_variants_agree(name):
    """The blend backward exists in four shapes (block double buffer or mbarrier ring of 32-Gaussian
    chunks; branch-free or branchy replay).  They replay the same pairs with the same arithmetic,
    so they differ by the order of the floating-point atomics only."""
    from gftorf_b200 import _capi
    inp = harness.build_inputs(device="cuda", **CASES[name])
    f = harness.call_forward(rasterizer._C, inp)
    res = {}
    for ring in (0, 1):
        for pred in (0, 1):
            o1, o2 = _capi.set_option("bwd_ring", ring), _capi.set_option("bwd_pred", pred)
            try:
                res[(ring, pred)] = harness.call_backward(rasterizer._C, inp, f)
                torch.cuda.synchronize()
            finally:
                _capi.set_option("bwd_ring", o1)
                _capi.set_option("bwd_pred", o2)
    base = res[(0, 0)]
    scale = float(base[8].double().norm())
    for key, g in res.items():
        for i, (x, y) in enumerate(zip(g, base)):
            if x is None:
                continue
            err = float((x.double() - y.double()).norm())
            assert err <= harness.GRAD_REL_L2 * max(float(y.double().norm()), 1e-2 * scale), (key, harness.BWD_NAMES[i])


@pytest.mark.parametrize("name", ["tiny", "ragged", "c1", "c1_init", "c1_dense", "c2"])
def test_half_patch_blend_is_exact(name):
    """Option blend_half: the blend warps walk their 8x4 patch as two independent 4x4 halves.  Per
    pixel the same candidates arrive in the same order with the same arithmetic: every forward
    output bit for bit (images, pixels counts, n_contrib), gradients up to atomic order."""
    from gftorf_b200 import _capi
    inp = harness.build_inputs(device="cuda", **CASES[name])
    f0 = harness.call_forward(rasterizer._C, inp)
    g0 = harness.call_backward(rasterizer._C, inp, f0)
    old = _capi.set_option("blend_half", 1 - _capi.set_option("blend_half", 0))
    try:
        f1 = harness.call_forward(rasterizer._C, inp)
        g1 = harness.call_backward(rasterizer._C, inp, f1)
        torch.cuda.synchronize()
    finally:
        _capi.set_option("blend_half", old)
    for i in range(1, 12):
        assert torch.equal(f0[i], f1[i]), harness.FWD_NAMES[i]
    d0 = debug.decode_buffers(f0[12], f0[13], f0[14], inp["P"], f0[0], inp["W"], inp["H"])
    d1 = debug.decode_buffers(f1[12], f1[13], f1[14], inp["P"], f1[0], inp["W"], inp["H"])
    for k in ("n_contrib", "final_T", "w_z_total", "w_z2_total"):
        assert torch.equal(d0[k], d1[k]), k
    scale = float(g0[8].double().norm())
    for i, (x, y) in enumerate(zip(g1, g0)):
        if x is None:
            continue
        err = float((x.double() - y.double()).norm())
        assert err <= harness.GRAD_REL_L2 * max(float(y.double().norm()), 1e-2 * scale), harness.BWD_NAMES[i]


def test_debug_mode_runs():
    inp = harness.build_inputs(device="cuda", **CASES["tiny"])
    e = inp["empty"]
    out = rasterizer._C.rasterize_gaussians(
        inp["bg"], inp["means3D"], e, e, inp["opacities"], inp["scales"], inp["rotations"], 1.0, e,
        inp["viewmatrix"], inp["projmatrix"], inp["tanfovx"], inp["tanfovy"], inp["H"], inp["W"],
        inp["shs"], inp["shs_p"], 3, inp["campos"], False, True, inp["near_n"], inp["far_n"],
        inp["depth_range"], False, 0.0, 0.0)
    assert out[0] > 0


# ------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full sizes
# ------------------------------------------------------------------------------------------------
@needs_ref
def test_full_size_c4_view_vs_reference():
    """One 1080p view of config C4 (2M Gaussians): direct comparison still fits in seconds."""
    spec = dict(P=2000000, W=1920, H=1080, kind="trained", seed=0, zero_shp_rest=True)
    inp, ours, ref = run_both(spec)
    od, rd = decoded(inp, ours, ref)
    harness.assert_forward_parity(harness.compare_forward(ours, ref, od, rd))
    ob = harness.call_backward(rasterizer._C, inp, ours)
    rb = harness.call_backward(ref_driver.RefModule, inp, ref)
    for i, k in enumerate(harness.BWD_NAMES):
        if k in ("colors_precomp", "phasors_precomp", "cov3Ds_precomp", "sh_p"):
            continue
        assert harness.rel_l2(ob[i], rb[i]) <= harness.GRAD_REL_L2, k


def test_full_size_properties():
    spec = dict(P=2000000, W=1920, H=1080, kind="trained", seed=1)
    inp = harness.build_inputs(device="cuda", **spec)
    f = harness.call_forward(rasterizer._C, inp)
    P, W, H, R = inp["P"], inp["W"], inp["H"], f[0]
    d = debug.decode_buffers(f[12], f[13], f[14], P, R, W, H)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    bits = 32 + T.bit_length()
    keys = d["keys"] & ((1 << bits) - 1)
    assert bool((keys[1:] >= keys[:-1]).all())                      # sortedness
    tiles = (d["keys"] >> 32)
    rng = d["ranges"].long()
    lens = (rng[:, 1] - rng[:, 0])
    assert int(lens.sum()) == R                                      # ranges partition [0, R)
    nz = lens > 0
    assert bool((tiles[rng[nz, 0]] == torch.arange(T, device="cuda")[nz]).all())
    assert bool((tiles[rng[nz, 1] - 1] == torch.arange(T, device="cuda")[nz]).all())
    assert int(d["tiles_touched"].long().sum()) == R
    assert int(d["point_offsets"][-1]) == R
    # stable sort: equal (tile, depth) keys keep ascending Gaussian index
    same = keys[1:] == keys[:-1]
    assert bool((d["point_list"][1:][same] > d["point_list"][:-1][same]).all())
    # n_contrib never exceeds the tile list length of its pixel's tile
    ncon = d["n_contrib"].view(H, W).long()
    ty = torch.arange(H, device="cuda") // 16
    tx = torch.arange(W, device="cuda") // 16
    tl = lens.view((H + 15) // 16, (W + 15) // 16)[ty][:, tx]
    assert bool((ncon <= tl).all())
    # acc + final_T: 0 <= T <= 1, acc = sum of weights <= 1 - T (+ rounding)
    Tf = d["final_T"].view(H, W)
    assert bool(((Tf >= 0) & (Tf <= 1)).all())
    assert float((f[5][0] + Tf - 1).abs().max()) < 1e-4
    # determinism: forward twice, bit-identical everywhere
    f2 = harness.call_forward(rasterizer._C, inp)
    for i in range(1, 12):
        assert torch.equal(f[i], f2[i])
    # linearity of the backward in dL/dout
    g1 = harness.call_backward(rasterizer._C, inp, f)
    inp2 = dict(inp)
    inp2["grads"] = {k: 2.0 * v for k, v in inp["grads"].items()}
    g2 = harness.call_backward(rasterizer._C, inp2, f)
    for a, b, k in zip(g1, g2, harness.BWD_NAMES):
        if k in ("colors_precomp", "phasors_precomp", "cov3Ds_precomp") or a is None:
            continue
        assert harness.rel_l2(2.0 * a, b) <= harness.GRAD_REL_L2, k


# ------------------------------------------------------------------------------------------------
# distCUDA2 and markVisible
# ------------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 33, 1000, 1025, 20000, 300000])
def test_dist2_bit_identical_to_reference(P):
    g = torch.Generator().manual_seed(P)
    pts = (torch.rand((P, 3), generator=g) * 4 - 2).cuda()
    if P >= 33:
        pts[7] = pts[3]
        pts[8] = pts[3]
    a, b = distCUDA2(pts), ref_driver.distCUDA2(pts)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))


@needs_ref
def test_dist2_clustered_and_degenerate_clouds():
    g = torch.Generator().manual_seed(5)
    clustered = torch.cat([torch.randn((5000, 3), generator=g) * 0.01 + c
                           for c in (torch.tensor([0., 0, 0]), torch.tensor([5., 5, 5]),
                                     torch.tensor([-3., 2, 9]))]).cuda()
    planar = torch.rand((8000, 3), generator=g).cuda()
    planar[:, 2] = 1.0
    same = torch.ones((100, 3)).cuda()
    for pts in (clustered, planar, same):
        a, b = distCUDA2(pts), ref_driver.distCUDA2(pts)
        torch.cuda.synchronize()
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))


def test_dist2_golden():
    k = golden_util.load_knn()
    for name in [n for n in k if n.startswith("pts_")]:
        out = distCUDA2(k[name].cuda()).cpu()
        assert torch.equal(out.view(torch.int32), k["out_" + name[4:]].view(torch.int32)), name
    assert distCUDA2(torch.zeros((0, 3), device="cuda")).shape == (0,)


@needs_ref
def test_mark_visible():
    inp = harness.build_inputs(device="cuda", **CASES["c1"])
    s = rasterizer.GaussianRasterizationSettings(
        image_height=240, image_width=320, tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"],
        bg=inp["bg"], scale_modifier=1.0, viewmatrix=inp["viewmatrix"], projmatrix=inp["projmatrix"],
        sh_degree=3, campos=inp["campos"], prefiltered=False, debug=False, near_n=inp["near_n"],
        far_n=inp["far_n"])
    vis = rasterizer.GaussianRasterizer(s).markVisible(inp["means3D"])
    ref = ref_driver.mark_visible(inp["means3D"], inp["viewmatrix"], inp["projmatrix"],
                                  inp["near_n"], inp["far_n"])
    assert vis.dtype == torch.bool and torch.equal(vis, ref)
    assert 0 < int(vis.sum()) < inp["P"]


# ------------------------------------------------------------------------------------------------
# the public autograd surface, end to end
# ------------------------------------------------------------------------------------------------
@needs_ref
def test_autograd_surface_matches_reference_binding():
    spec = dict(P=5000, W=128, H=96, kind="trained", seed=13, zero_shp_rest=True)
    inp = harness.build_inputs(device="cuda", **spec)
    phase = torch.nn.Parameter(torch.tensor([0.25], device="cuda"))
    dc = torch.nn.Parameter(torch.tensor([0.05], device="cuda"))
    names = ("means3D", "opacities", "shs", "shs_p", "scales", "rotations")
    leaves = {k: inp[k].clone().requires_grad_(True) for k in names}
    means2D = torch.zeros_like(inp["means3D"], requires_grad=True)
    s = rasterizer.GaussianRasterizationSettings(
        image_height=96, image_width=128, tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"],
        bg=inp["bg"], scale_modifier=1.0, viewmatrix=inp["viewmatrix"], projmatrix=inp["projmatrix"],
        sh_degree=3, campos=inp["campos"], prefiltered=False, debug=False, near_n=inp["near_n"],
        far_n=inp["far_n"], depth_range=inp["depth_range"], use_view_dependent_phase=False,
        optimize_phase_offset=True, optimize_dc_offset=True)
    out = rasterizer.GaussianRasterizer(s)(
        means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
        shs=leaves["shs"], shs_p=leaves["shs_p"], scales=leaves["scales"],
        rotations=leaves["rotations"], phase_offset=phase, dc_offset=dc)
    assert len(out) == 11 and out[10].dtype == torch.int32
    g = inp["grads"]
    loss = (out[0] * g["color"]).sum() + (out[1] * g["phasor"]).sum() + (out[2] * g["depth"]).sum() \
        + (out[4] * g["acc"]).sum() + (out[6] * g["depth_distortion"]).sum()
    loss.backward()
    inp["phase_offset"], inp["dc_offset"] = 0.25, 0.05
    ref = harness.call_forward(ref_driver.RefModule, inp)
    rb = harness.call_backward(ref_driver.RefModule, inp, ref)
    for i in range(10):
        assert torch.equal(out[i], ref[i + 1]), harness.FWD_NAMES[i + 1]
    pairs = dict(means3D=rb[4], opacities=rb[3], shs=rb[6], scales=rb[8], rotations=rb[9])
    for k, r in pairs.items():
        assert harness.rel_l2(leaves[k].grad, r) <= harness.GRAD_REL_L2, k
    assert harness.rel_l2(means2D.grad, rb[0]) <= harness.GRAD_REL_L2
    assert harness.rel_l2(leaves["shs_p"].grad[0], rb[7][0]) <= harness.GRAD_REL_L2
    assert harness.rel_l2(phase.grad, rb[10]) <= harness.GRAD_REL_L2
    assert harness.rel_l2(dc.grad, rb[11]) <= harness.GRAD_REL_L2
    assert float(means2D.grad[:, 2].abs().max()) == 0.0


def test_accumulate_mode_adds_views_into_the_bucket():
    """GftBackwardArgs.accumulate: two views' gradients added straight into a GradBucket equal the
    sum of the two plain backward results; rows of Gaussians culled in both views stay zero."""
    from gftorf_b200 import parallel
    inp_a = harness.build_inputs(device="cuda", P=6000, W=128, H=96, kind="trained", seed=31)
    inp_b = harness.build_inputs(device="cuda", P=6000, W=96, H=80, kind="trained", seed=31, pose="orbit")
    for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p"):
        inp_b[k] = inp_a[k]                       # same Gaussians, two cameras
    params = {k: inp_a[k] for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p")}
    scal = [torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")]
    bucket = parallel.GradBucket(params, scal)
    go = bucket.grad_out()
    plain, vis = [], None
    for inp in (inp_a, inp_b):
        f = harness.call_forward(rasterizer._C, inp)
        plain.append(harness.call_backward(rasterizer._C, inp, f))
        acc = harness.call_backward(rasterizer._C, inp, f, grad_out=go)
        assert harness.rel_l2(acc[0], plain[-1][0]) <= harness.GRAD_REL_L2   # means2D stays per view (atomic order differs)
        vis = (f[11] > 0) if vis is None else (vis | (f[11] > 0))
    pairs = dict(means3D=4, shs=6, shs_p=7, opacities=3, scales=8, rotations=9)
    for name, i in pairs.items():
        expect = plain[0][i] + plain[1][i]
        assert harness.rel_l2(go[name], expect) <= harness.GRAD_REL_L2, name
        assert float(go[name][~vis].abs().sum()) == 0.0, name
    assert harness.rel_l2(go["phase_offset"], plain[0][10] + plain[1][10]) <= harness.GRAD_REL_L2
    assert harness.rel_l2(go["dc_offset"], plain[0][11] + plain[1][11]) <= harness.GRAD_REL_L2
    assert bucket.flat.data_ptr() == go["means3D"].data_ptr()
    # accumulate=False on the first view overwrites whatever the bucket held: no zero fill needed
    summed = bucket.flat.clone()
    bucket.flat.fill_(float("nan"))
    for vi, inp in enumerate((inp_a, inp_b)):
        f = harness.call_forward(rasterizer._C, inp)
        harness.call_backward(rasterizer._C, inp, f, grad_out=go, accumulate=vi > 0)
    for name, t in go.items():        # (the alignment padding between slices is never touched)
        assert bool(torch.isfinite(t).all()), name
    bucket.flat.nan_to_num_(nan=0.0)
    assert harness.rel_l2(bucket.flat, summed) <= harness.GRAD_REL_L2


def test_concurrent_views_through_the_autograd_surface():
    """Two views of one iteration issued through parallel.ViewRunner with the PUBLIC surface
    (GaussianRasterizer + autograd.backward from the two host threads): the leaves' accumulated
    gradients equal those of the two views run one after the other."""
    from gftorf_b200 import parallel
    inp_a = harness.build_inputs(device="cuda", P=8000, W=160, H=120, kind="trained", seed=51)
    inp_b = harness.build_inputs(device="cuda", P=8000, W=128, H=96, kind="trained", seed=51, pose="orbit")
    names = ("means3D", "opacities", "shs", "shs_p", "scales", "rotations")

    def run(concurrent):
        leaves = {k: inp_a[k].clone().requires_grad_(True) for k in names}
        m2d = torch.zeros_like(inp_a["means3D"], requires_grad=True)

        def view(inp):
            s = rasterizer.GaussianRasterizationSettings(
                image_height=inp["H"], image_width=inp["W"], tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"],
                bg=inp["bg"], scale_modifier=1.0, viewmatrix=inp["viewmatrix"], projmatrix=inp["projmatrix"],
                sh_degree=3, campos=inp["campos"], prefiltered=False, debug=False, near_n=inp["near_n"],
                far_n=inp["far_n"], depth_range=inp["depth_range"])
            out = rasterizer.GaussianRasterizer(s)(
                means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
                shs_p=leaves["shs_p"], scales=leaves["scales"], rotations=leaves["rotations"])
            g = inp["grads"]
            torch.autograd.backward([out[0], out[1], out[2], out[4], out[6]],
                                    [g["color"], g["phasor"], g["depth"], g["acc"], g["depth_distortion"]])
            return out[0].clone()

        if concurrent:
            runner = parallel.ViewRunner(2)
            imgs = runner.run([lambda: view(inp_a), lambda: view(inp_b)])
            runner.close()
        else:
            imgs = [view(inp_a), view(inp_b)]
        torch.cuda.synchronize()
        return imgs, {k: leaves[k].grad.clone() for k in names}

    imgs_s, grads_s = run(False)
    imgs_c, grads_c = run(True)
    for a, b in zip(imgs_s, imgs_c):
        assert torch.equal(a, b)
    for k in names:
        assert harness.rel_l2(grads_c[k], grads_s[k]) <= harness.GRAD_REL_L2, k


def test_concurrent_views_with_atomic_accumulation():
    """parallel.ViewRunner: the two views of an iteration on two streams / host threads, gradients
    added into one zero-filled bucket with atomics.  Same images as the sequential calls (bit for
    bit), same gradient sums within the gradient tolerance; errors in a view reach the caller."""
    from gftorf_b200 import parallel
    inp_a = harness.build_inputs(device="cuda", P=20000, W=320, H=240, kind="trained", seed=41)
    inp_b = harness.build_inputs(device="cuda", P=20000, W=256, H=192, kind="trained", seed=41, pose="orbit")
    for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p"):
        inp_b[k] = inp_a[k]
    params = {k: inp_a[k] for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p")}
    bucket = parallel.GradBucket(params, [torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")])
    go = bucket.grad_out()
    seq = []
    for inp in (inp_a, inp_b):
        f = harness.call_forward(rasterizer._C, inp)
        seq.append((f, harness.call_backward(rasterizer._C, inp, f)))
    runner = parallel.ViewRunner(2)

    def view(inp):
        f = harness.call_forward(rasterizer._C, inp)
        b = harness.call_backward(rasterizer._C, inp, f, grad_out=go, accumulate="atomic")
        return f, b

    for _ in range(3):                       # repeated: a race would not show every time
        bucket.zero()
        res = runner.run([lambda: view(inp_a), lambda: view(inp_b)])
        torch.cuda.synchronize()
        for (f, b), (fs, bs) in zip(res, seq):
            assert f[0] == fs[0]
            for i in range(1, 12):
                assert torch.equal(f[i], fs[i]), harness.FWD_NAMES[i]
            assert harness.rel_l2(b[0], bs[0]) <= harness.GRAD_REL_L2      # means2D, per view
        for name, i in dict(means3D=4, shs=6, shs_p=7, opacities=3, scales=8, rotations=9).items():
            assert harness.rel_l2(go[name], seq[0][1][i] + seq[1][1][i]) <= harness.GRAD_REL_L2, name

    def boom():
        raise ValueError("view failed")
    with pytest.raises(ValueError, match="view failed"):
        runner.run([lambda: view(inp_a), boom])
    torch.cuda.synchronize()
    runner.close()


def test_hinted_forward_is_identical_and_survives_a_bad_hint():
    """GftForwardArgs.R_hint: same bits as the exact mode for a generous hint, an exact hint, and a
    hint that is too small (overflow -> the tail of the pipeline re-runs with the right size)."""
    inp = harness.build_inputs(device="cuda", **CASES["c1"])
    e = inp["empty"]

    def fwd(hint):
        return rasterizer._C.rasterize_gaussians(
            inp["bg"], inp["means3D"], e, e, inp["opacities"], inp["scales"], inp["rotations"], 1.0, e,
            inp["viewmatrix"], inp["projmatrix"], inp["tanfovx"], inp["tanfovy"], inp["H"], inp["W"],
            inp["shs"], inp["shs_p"], 3, inp["campos"], False, False, inp["near_n"], inp["far_n"],
            inp["depth_range"], False, 0.0, 0.0, R_hint=hint)
    exact = fwd(0)
    R = exact[0]
    gb_exact = harness.call_backward(rasterizer._C, inp, exact)
    for hint in (2 * R, R + 1, R, R - 1, R // 3, 1):
        f = fwd(hint)
        assert f[0] == R
        for i in range(1, 12):
            assert torch.equal(f[i], exact[i]), (hint, harness.FWD_NAMES[i])
        g = harness.call_backward(rasterizer._C, inp, f)
        for x, y in zip(g, gb_exact):
            if x is not None:
                assert harness.rel_l2(x, y) <= 1e-5, hint
    # a hint with nothing to render
    inp0 = harness.build_inputs(device="cuda", P=500, W=40, H=24, seed=3)
    inp0["means3D"] = inp0["means3D"] * torch.tensor([1.0, 1.0, -1.0], device="cuda")
    z = rasterizer._C.rasterize_gaussians(
        inp0["bg"], inp0["means3D"], e, e, inp0["opacities"], inp0["scales"], inp0["rotations"], 1.0, e,
        inp0["viewmatrix"], inp0["projmatrix"], inp0["tanfovx"], inp0["tanfovy"], 24, 40, inp0["shs"],
        inp0["shs_p"], 3, inp0["campos"], False, False, inp0["near_n"], inp0["far_n"],
        inp0["depth_range"], False, 0.0, 0.0, R_hint=5000)
    assert z[0] == 0 and torch.equal(z[1], inp0["bg"][0:3])


# ------------------------------------------------------------------------------------------------
# the remaining BASELINE configs: 8 M Gaussians at 1080p (configs[4]), a camera that moves every
# frame (render.py:95-209), and batches of views (configs[1]-[3]: colour + ToF camera per
# iteration; configs[3]: 8 cameras per iteration)
# ------------------------------------------------------------------------------------------------
@needs_ref
def test_full_size_c5_forward_vs_reference():
    """BASELINE configs[4] upper end: 8 M Gaussians (screen-space sigma 1 px) at 1920x1080,
    forward only: every integer stage and every image plane against the reference kernels."""
    spec = dict(P=8000000, W=1920, H=1080, kind="trained", seed=2, sigma_px=1.0)
    inp, ours, ref = run_both(spec)
    od, rd = decoded(inp, ours, ref)
    rep = harness.compare_forward(ours, ref, od, rd)
    harness.assert_forward_parity(rep)
    assert rep["R_ours"] > 15000000


def _orbit_views(n, W, H, device, depth_range=15.0):
    """A trajectory around the scene: every frame has its own pose and the instance count moves by
    far more than the 25 % margin of the R hint between some frames (zoom by field of view)."""
    from gftorf_b200 import scenes
    out = []
    fovs = (1.2, 0.3, 1.2, 0.25, 1.1, 0.2, 1.2, 0.4, 1.2)
    for i in range(n):
        fov = fovs[i % len(fovs)]                    # wide / strong zoom alternate: R jumps both ways
        cam = scenes.make_camera(W, H, fovx=fov, depth_range=depth_range, pose="orbit" if i else "identity", seed=100 + i)
        out.append(cam)
    return out


@needs_ref
def test_moving_camera_sequence_through_the_autograd_surface():
    """render.py:95-209 renders a trajectory frame by frame.  The public surface sizes the binning
    workspace from the previous frame's instance count (+25 %); on a moving camera that estimate
    is sometimes too small and the tail of the forward re-runs.  Every frame must equal the
    reference kernels bit for bit, whichever path it took."""
    from gftorf_b200 import scenes
    W, H, P = 320, 240, 60000
    base = harness.build_inputs(device="cuda", P=P, W=W, H=H, kind="trained", seed=17, sigma_px=2.5)
    cams = _orbit_views(9, W, H, "cuda")
    t = lambda a: torch.from_numpy(a).cuda()
    means2D = torch.zeros_like(base["means3D"])
    Rs, overflowed = [], 0
    rasterizer._R_HISTORY.clear()
    for cam in cams:
        inp = dict(base, viewmatrix=t(cam["viewmatrix"]), projmatrix=t(cam["projmatrix"]),
                   campos=t(cam["campos"]), tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
                   near_n=cam["znear"], far_n=cam["zfar"])
        s = rasterizer.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"], bg=inp["bg"],
            scale_modifier=1.0, viewmatrix=inp["viewmatrix"], projmatrix=inp["projmatrix"], sh_degree=3,
            campos=inp["campos"], prefiltered=False, debug=False, near_n=inp["near_n"], far_n=inp["far_n"],
            depth_range=inp["depth_range"])
        hint = rasterizer._r_hint((0, P, H, W))
        with torch.no_grad():
            out = rasterizer.GaussianRasterizer(s)(
                means3D=inp["means3D"], means2D=means2D, opacities=inp["opacities"], shs=inp["shs"],
                shs_p=inp["shs_p"], scales=inp["scales"], rotations=inp["rotations"])
        ref = harness.call_forward(ref_driver.RefModule, inp)
        torch.cuda.synchronize()
        for i in range(11):
            assert torch.equal(out[i], ref[i + 1]), (len(Rs), harness.FWD_NAMES[i + 1])
        if Rs and ref[0] > hint:
            overflowed += 1
        Rs.append(ref[0])
    assert overflowed >= 1, Rs       # the too-small-hint path really ran
    assert min(Rs) > 0


def _two_camera_inputs(P, wh_a, wh_b, seed, kind="trained", **kw):
    """The same Gaussians seen by two cameras of different resolution; the background map is sized
    from the first camera and reused for the second (train.py:121-128, quirk A.7-3)."""
    (Wa, Ha), (Wb, Hb) = wh_a, wh_b
    a = harness.build_inputs(device="cuda", P=P, W=Wa, H=Ha, kind=kind, seed=seed, **kw)
    b = harness.build_inputs(device="cuda", P=P, W=Wb, H=Hb, kind=kind, seed=seed, pose="orbit",
                             bg_hw=(Ha, Wa) if Wb * Hb <= Wa * Ha else None, **kw)
    for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p"):
        b[k] = a[k]
    if Wb * Hb <= Wa * Ha:
        b["bg"] = a["bg"]
    return a, b


def _spec(inp):
    from gftorf_b200 import views as V
    return V.ViewSpec(inp["H"], inp["W"], inp["tanfovx"], inp["tanfovy"], inp["bg"], inp["viewmatrix"],
                      inp["projmatrix"], inp["campos"], inp["near_n"], inp["far_n"], inp["depth_range"],
                      inp["use_view_dependent_phase"], inp["phase_offset"], inp["dc_offset"])


@needs_ref
@pytest.mark.parametrize("shape", [
    dict(P=6000, wh_a=(128, 96), wh_b=(96, 80), seed=31),
    dict(P=20000, wh_a=(320, 240), wh_b=(320, 240), seed=32, view_dependent_phase=True, phase_offset=0.2,
         dc_offset=0.1),
    dict(P=300000, wh_a=(640, 480), wh_b=(640, 480), seed=0),           # BASELINE configs[1]
    dict(P=500000, wh_a=(640, 480), wh_b=(320, 240), seed=4, depth_range=10.0),   # configs[2]
    dict(P=30000, wh_a=(160, 120), wh_b=(101, 67), seed=33, kind="init"),
])
def test_view_batch_equals_the_reference_calls_view_by_view(shape):
    """gft_forward_views / gft_backward_views: one call for the colour and the ToF camera of an
    iteration.  Per view every integer stage and image is the reference kernels' result bit for
    bit; the parameter gradients are the sum of the reference's two backward calls (what
    AccumulateGrad produces, train.py:279); means2D gradients stay per view."""
    from gftorf_b200 import views as V
    shape = dict(shape)
    a, b = _two_camera_inputs(shape.pop("P"), shape.pop("wh_a"), shape.pop("wh_b"), shape.pop("seed"),
                              zero_shp_rest=True, **shape)
    fwd = V.forward_views(a["means3D"], a["opacities"], a["scales"], a["rotations"], a["shs"], a["shs_p"],
                          [_spec(a), _spec(b)], 3)
    dec = debug.decode_views(fwd.geom, fwd.binning, fwd.img, a["P"], fwd.R,
                             [(a["W"], a["H"]), (b["W"], b["H"])])
    refs = []
    for i, inp in enumerate((a, b)):
        ref = harness.call_forward(ref_driver.RefModule, inp)
        torch.cuda.synchronize()
        rd = ref_driver.decode_buffers(ref[12], ref[13], ref[14], inp["P"], ref[0], inp["W"], inp["H"])
        ours = (dec[i]["num_rendered"],) + tuple(fwd.outs[i])
        rep = harness.compare_forward(ours, ref, dec[i], rd)
        harness.assert_forward_parity(rep)
        refs.append(ref)
    assert fwd.R == refs[0][0] + refs[1][0]
    out = V.backward_views(fwd, [a["grads"], b["grads"]])
    rbs = [harness.call_backward(ref_driver.RefModule, inp, ref) for inp, ref in zip((a, b), refs)]
    torch.cuda.synchronize()
    scale = float((rbs[0][8] + rbs[1][8]).double().norm())
    for name, i in dict(means3D=4, shs=6, opacities=3, scales=8, rotations=9, phase_offset=10, dc_offset=11).items():
        expect = (rbs[0][i] + rbs[1][i]).double()
        err = float((out[name].double() - expect).norm())
        # the quaternion gradient is a difference of terms a hundred times its own size (here
        # |dL/drot| ~ 1e-2 |dL/dscale|), so its noise floor sits at float rounding of THOSE terms;
        # the scalar offsets are single sums over all Gaussians in arbitrary atomic order
        tol = {"rotations": 3e-4, "phase_offset": 1e-3, "dc_offset": 1e-3}.get(name, harness.GRAD_REL_L2)
        bound = tol * max(float(expect.norm()), 1e-3 * scale)
        if name == "rotations":
            # ... so the floor is stated against those terms: 1e-5 of |dL/dscale| — float32 rounding
            # (6e-8) through the ~100 operations of the covariance chain and atomic sums in
            # arbitrary order; observed 2e-6 ... 5e-6 of |dL/dscale| from run to run
            bound = max(bound, 1e-5 * scale)
        assert err <= bound, (name, err, bound)
    assert harness.rel_l2(out["shs_p"][0], rbs[0][7][0] + rbs[1][7][0]) <= harness.GRAD_REL_L2
    for i in range(2):
        assert harness.rel_l2(out["means2D"][i], rbs[i][0]) <= harness.GRAD_REL_L2


def test_view_batch_gradients_equal_single_view_calls_with_live_sh_p():
    """With all SH_p coefficients live (where the reference is undefined, D1) the batch must still
    equal OUR single-view calls: same images bit for bit, gradients = sum of the two calls."""
    from gftorf_b200 import views as V
    a, b = _two_camera_inputs(9000, (160, 120), (128, 96), 35, view_dependent_phase=True,
                              phase_offset=0.3, dc_offset=0.05)
    fwd = V.forward_views(a["means3D"], a["opacities"], a["scales"], a["rotations"], a["shs"], a["shs_p"],
                          [_spec(a), _spec(b)], 3)
    out = V.backward_views(fwd, [a["grads"], b["grads"]])
    singles = []
    for i, inp in enumerate((a, b)):
        f = harness.call_forward(rasterizer._C, inp)
        for j in range(1, 12):
            assert torch.equal(f[j], fwd.outs[i][j - 1]), (i, harness.FWD_NAMES[j])
        singles.append(harness.call_backward(rasterizer._C, inp, f))
    for name, i in dict(means3D=4, shs=6, shs_p=7, opacities=3, scales=8, rotations=9, phase_offset=10,
                        dc_offset=11).items():
        assert harness.rel_l2(out[name], singles[0][i] + singles[1][i]) <= harness.GRAD_REL_L2, name
    for i in range(2):
        assert harness.rel_l2(out["means2D"][i], singles[i][0]) <= harness.GRAD_REL_L2
    # Gaussians culled in both views: zero rows
    vis = (fwd.radii[0] > 0) | (fwd.radii[1] > 0)
    assert int((~vis).sum()) > 0
    for name in ("means3D", "shs", "shs_p", "opacities", "scales", "rotations"):
        assert float(out[name][~vis].abs().sum()) == 0.0, name


def test_view_batch_of_eight_cameras_and_bucket_accumulation():
    """BASELINE configs[3] shape in small: 8 cameras x (colour + ToF resolution) = 16 views in one
    call, gradients written straight into a GradBucket; a second call accumulates into it."""
    from gftorf_b200 import views as V, parallel, scenes
    P = 12000
    base = harness.build_inputs(device="cuda", P=P, W=192, H=108, kind="trained", seed=61, sigma_px=2.0)
    t = lambda x: torch.from_numpy(x).cuda()
    specs, grads = [], []
    for c in range(8):
        for (W, H) in ((192, 108), (64, 48)):
            cam = scenes.make_camera(W, H, pose="orbit" if c else "identity", seed=200 + c)
            specs.append(V.ViewSpec(H, W, cam["tanfovx"], cam["tanfovy"], base["bg"], t(cam["viewmatrix"]),
                                    t(cam["projmatrix"]), t(cam["campos"]), cam["znear"], cam["zfar"],
                                    cam["depth_range"]))
            grads.append({k: t(v) for k, v in scenes.make_pixel_grads(H, W, seed=300 + c).items()})
    args = (base["means3D"], base["opacities"], base["scales"], base["rotations"], base["shs"], base["shs_p"])
    fwd = V.forward_views(*args, specs, 3)
    assert len(fwd.outs) == 16 and fwd.R > 0
    out = V.backward_views(fwd, grads)
    # view by view through the single-view entry point
    acc = None
    for i, sp in enumerate(specs):
        f1 = V.forward_views(*args, [sp], 3)
        for j in range(11):
            assert torch.equal(f1.outs[0][j], fwd.outs[i][j]), (i, j)
        o1 = V.backward_views(f1, [grads[i]])
        assert harness.rel_l2(out["means2D"][i], o1["means2D"][0]) <= harness.GRAD_REL_L2
        acc = {k: v.clone() for k, v in o1.items()} if acc is None else {k: acc[k] + o1[k] for k in acc}
    for k in ("means3D", "shs", "shs_p", "opacities", "scales", "rotations", "phase_offset", "dc_offset"):
        assert harness.rel_l2(out[k], acc[k]) <= harness.GRAD_REL_L2, k
    # bucket: first call overwrites (no zero fill), second call adds
    params = {k: base[k] for k in ("means3D", "scales", "rotations", "opacities", "shs", "shs_p")}
    bucket = parallel.GradBucket(params, [torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")])
    bucket.flat.fill_(float("nan"))
    go = bucket.grad_out()
    fa = V.forward_views(*args, specs[:8], 3)
    V.backward_views(fa, grads[:8], grad_out=go, accumulate=False)
    fb = V.forward_views(*args, specs[8:], 3)
    V.backward_views(fb, grads[8:], grad_out=go, accumulate=True)
    for k in ("means3D", "shs", "shs_p", "opacities", "scales", "rotations", "phase_offset", "dc_offset"):
        assert harness.rel_l2(go[k], out[k]) <= harness.GRAD_REL_L2, k


@needs_ref
def test_rasterize_views_autograd_matches_two_reference_calls():
    """The differentiable batched surface: outputs are independent tensors (in-place ops allowed),
    leaves receive the summed gradients, means2D [V,P,3] the per-view ones."""
    from gftorf_b200 import views as V
    a, b = _two_camera_inputs(7000, (128, 96), (96, 64), 71, zero_shp_rest=True)
    names = ("means3D", "opacities", "shs", "shs_p", "scales", "rotations")
    leaves = {k: a[k].clone().requires_grad_(True) for k in names}
    m2d = torch.zeros((2, 7000, 3), device="cuda", requires_grad=True)
    outs = V.rasterize_views(leaves["means3D"], m2d, leaves["opacities"], leaves["shs"], leaves["shs_p"],
                             leaves["scales"], leaves["rotations"], [_spec(a), _spec(b)], 3)
    loss = 0
    for o, inp in zip(outs, (a, b)):
        g = inp["grads"]
        assert len(o) == 11 and o[10].dtype == torch.int32
        loss = loss + (o[0] * g["color"]).sum() + (o[1] * g["phasor"]).sum() + (o[2] * g["depth"]).sum() \
            + (o[4] * g["acc"]).sum() + (o[6] * g["depth_distortion"]).sum()
    loss.backward()
    outs[0][0].clamp_(0, 1)          # outputs are independent tensors: in-place ops are legal
    refs = [harness.call_forward(ref_driver.RefModule, inp) for inp in (a, b)]
    rbs = [harness.call_backward(ref_driver.RefModule, inp, r) for inp, r in zip((a, b), refs)]
    for o, r in zip(outs, refs):
        for i in range(1, 10):
            assert torch.equal(o[i], r[i + 1]), harness.FWD_NAMES[i + 1]
    for k, i in dict(means3D=4, opacities=3, shs=6, scales=8, rotations=9).items():
        assert harness.rel_l2(leaves[k].grad, rbs[0][i] + rbs[1][i]) <= harness.GRAD_REL_L2, k
    for v in range(2):
        assert harness.rel_l2(m2d.grad[v], rbs[v][0]) <= harness.GRAD_REL_L2


@needs_ref
def test_rasterize_views_optimised_offsets_go_to_the_tof_view_only():
    """gaussian_renderer/__init__.py:126-127 hands the optimised phase / dc offsets to the ToF call
    only.  In the batched surface the ToF view leaves its offsets None (= the call's Parameters),
    the colour view keeps 0: outputs and the two scalar gradients equal the reference's calls."""
    from gftorf_b200 import views as V
    a, b = _two_camera_inputs(5000, (128, 96), (96, 64), 73, zero_shp_rest=True)
    phase = torch.nn.Parameter(torch.tensor([0.25], device="cuda"))
    dc = torch.nn.Parameter(torch.tensor([0.05], device="cuda"))
    names = ("means3D", "opacities", "shs", "shs_p", "scales", "rotations")
    leaves = {k: a[k].clone().requires_grad_(True) for k in names}
    m2d = torch.zeros((2, 5000, 3), device="cuda", requires_grad=True)
    specs = [_spec(a), _spec(b)._replace(phase_offset=None, dc_offset=None)]
    outs = V.rasterize_views(leaves["means3D"], m2d, leaves["opacities"], leaves["shs"], leaves["shs_p"],
                             leaves["scales"], leaves["rotations"], specs, 3, phase_offset=phase, dc_offset=dc,
                             optimize_phase_offset=True, optimize_dc_offset=True)
    # as in render(): the colour call's phasor image is discarded, the ToF call's colour image too
    loss = (outs[0][0] * a["grads"]["color"]).sum() + (outs[0][2] * a["grads"]["depth"]).sum() \
        + (outs[1][1] * b["grads"]["phasor"]).sum() + (outs[1][2] * b["grads"]["depth"]).sum()
    loss.backward()
    b2 = dict(b, phase_offset=0.25, dc_offset=0.05)
    refs = [harness.call_forward(ref_driver.RefModule, inp) for inp in (a, b2)]
    for o, r in zip(outs, refs):
        for i in range(1, 10):
            assert torch.equal(o[i], r[i + 1]), harness.FWD_NAMES[i + 1]
    rbs = []
    for vi, (inp, r) in enumerate(zip((a, b2), refs)):
        g = dict(inp["grads"], acc=torch.zeros_like(inp["grads"]["acc"]),
                 depth_distortion=torch.zeros_like(inp["grads"]["depth_distortion"]))
        g["phasor" if vi == 0 else "color"] = torch.zeros_like(g["phasor" if vi == 0 else "color"])
        rbs.append(harness.call_backward(ref_driver.RefModule, dict(inp, grads=g), r))
    assert float(rbs[0][10].abs().sum()) == 0.0
    assert harness.rel_l2(phase.grad, rbs[0][10] + rbs[1][10]) <= 1e-3
    assert harness.rel_l2(dc.grad, rbs[0][11] + rbs[1][11]) <= 1e-3
    assert harness.rel_l2(leaves["means3D"].grad, rbs[0][4] + rbs[1][4]) <= harness.GRAD_REL_L2


def test_public_outputs_are_independent_tensors():
    """In-place edits of a returned image must be legal under autograd (the reference returns
    separate tensors), and the screen-space gradient must not pin the scratch allocation."""
    inp = harness.build_inputs(device="cuda", **CASES["tiny"])
    names = ("means3D", "opacities", "shs", "shs_p", "scales", "rotations")
    leaves = {k: inp[k].clone().requires_grad_(True) for k in names}
    m2d = torch.zeros_like(inp["means3D"], requires_grad=True)
    s = rasterizer.GaussianRasterizationSettings(
        image_height=inp["H"], image_width=inp["W"], tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"],
        bg=inp["bg"], scale_modifier=1.0, viewmatrix=inp["viewmatrix"], projmatrix=inp["projmatrix"],
        sh_degree=3, campos=inp["campos"], prefiltered=False, debug=False, near_n=inp["near_n"],
        far_n=inp["far_n"], depth_range=inp["depth_range"])
    out = rasterizer.GaussianRasterizer(s)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                          shs=leaves["shs"], shs_p=leaves["shs_p"], scales=leaves["scales"],
                                          rotations=leaves["rotations"])
    ptrs = {t.untyped_storage().data_ptr() for t in out}
    assert len(ptrs) == 11
    depth = out[2].clone()
    out[2].clamp_(min=0.5)           # raised "view ... output of a function that returns multiple views" before
    (out[0].sum() + depth.sum()).backward()
    assert m2d.grad.untyped_storage().nbytes() == m2d.grad.numel() * 4
