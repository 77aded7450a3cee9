"""CPU oracle (oracle/gft_oracle.cpp) vs the golden fixtures — outputs of the UNMODIFIED reference
kernels on a B200 (tests/golden/make_golden.py).  This is what pins the oracle."""
import pytest
import torch

import golden_util
import harness
from oracle import cpu_oracle

pytestmark = pytest.mark.skipif(not cpu_oracle.available(), reason="oracle/libgft_oracle.so not built")

# libm's expf/sinf/cosf are not CUDA's: per-Gaussian phasor features and blended images agree to
# float rounding, not bit for bit.
FEATURE_ATOL = 1e-5
# pixels whose alpha lands within an ulp of 1/255 or of the 1e-4 transmittance cut may flip
MAX_THRESHOLD_FLIPS = 2


@pytest.mark.parametrize("name", golden_util.CASES)
def test_forward_matches_reference_golden(name):
    inp, gold = golden_util.load(name)
    f = harness.call_forward(cpu_oracle.OracleModule, inp)
    d = cpu_oracle.decode(f[12])
    vis = gold["f_radii"] > 0
    assert f[0] == gold["R"]
    assert torch.equal(f[11], gold["f_radii"])
    # integer chain: bit-exact (IEEE + - * / sqrt fma only)
    for k in ("tiles_touched", "point_offsets", "keys", "point_list", "ranges"):
        if "s_" + k in gold:
            assert torch.equal(d[k], gold["s_" + k]), k
    for k in ("depths", "means2D", "cov3D", "conic_opacity"):
        assert torch.equal(d[k].view(torch.int32)[vis], gold["s_" + k].view(torch.int32)[vis]), k
    assert torch.equal(d["clamped"][vis] != 0, gold["s_clamped"][vis] != 0)
    assert torch.equal(d["clamped_p"][vis] != 0, gold["s_clamped_p"][vis] != 0)
    for k in ("rgb", "real_img_amp", "dists", "ndc", "pa"):
        assert (d[k][vis] - gold["s_" + k][vis]).abs().max() <= FEATURE_ATOL, k
    # threshold-dependent integers: bounded number of flips
    assert harness.nmismatch(d["n_contrib"], gold["s_n_contrib"]) <= MAX_THRESHOLD_FLIPS
    assert harness.nmismatch(f[9], gold["f_pixels"]) <= MAX_THRESHOLD_FLIPS
    if harness.nmismatch(d["n_contrib"], gold["s_n_contrib"]) == 0:
        for i, k in enumerate(harness.FWD_NAMES):
            if i >= 1 and k not in ("pixels", "radii"):
                assert (f[i] - gold["f_" + k]).abs().max() <= harness.IMG_ATOL, k


@pytest.mark.parametrize("name", golden_util.CASES)
def test_backward_matches_reference_golden(name):
    inp, gold = golden_util.load(name)
    f = harness.call_forward(cpu_oracle.OracleModule, inp)
    b, internal = harness.call_backward(cpu_oracle.OracleModule, inp, f, return_internal=True)
    zero_rest = bool(gold["spec"].get("zero_shp_rest", False)) or gold["spec"]["kind"] == "init"
    scale = max(float(gold["b_scales"].norm()), 1e-20)
    for i, k in enumerate(harness.BWD_NAMES):
        if k in ("colors_precomp", "phasors_precomp", "cov3Ds_precomp"):
            continue
        if k == "sh_p":
            # reference defect D1: only Gaussian 0's row is defined
            assert harness.rel_l2(b[i][0], gold["b_sh_p"][0]) <= harness.GRAD_REL_L2
            continue
        if k == "means3D" and not zero_rest:
            continue  # polluted by the same undefined read through the SH direction term
        err = float((b[i].double() - gold["b_" + k].double()).norm())
        ref = float(gold["b_" + k].double().norm())
        # rotations of isotropic Gaussians have an analytically zero gradient: compare absolutely
        assert err <= harness.GRAD_REL_L2 * max(ref, 1e-3 * scale), (k, err, ref)
    for k, v in internal.items():
        assert harness.rel_l2(v, gold["b_" + k]) <= harness.GRAD_REL_L2, k


def test_knn_matches_reference_golden_bitwise():
    k = golden_util.load_knn()
    for name in [n for n in k if n.startswith("pts_")]:
        P = name[4:]
        out = cpu_oracle.distCUDA2(k[name])
        assert torch.equal(out.view(torch.int32), k["out_" + P].view(torch.int32)), P
