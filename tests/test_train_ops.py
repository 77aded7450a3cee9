"""The §8f operators (assembly, fused loss, flat Adam): oracle pins on CPU, parity on the GPU.

CPU part: the restated oracle (oracle/train_oracle.py) is pinned against independent computations —
torch.optim.Adam itself, a brute-force SSIM, and hand-derived gradients.  GPU part (-m gpu): the
CUDA kernels behind include/gftorf_train.h against that oracle and its autograd, through the C ABI.
Tolerances are floating point: 1e-6 relative for streaming arithmetic, 2e-5 for SSIM (separable
vs 2-D window summation order)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import train_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------
def make_raw(P, M=16, isotropic=False, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    raw = dict(xyz=r(P, 3), opacity_raw=r(P, 1) * 2, scaling_raw=r(P, 1 if isotropic else 3) * 0.5 - 2,
               rotation_raw=r(P, 4), f_dc_color=r(P, 1, 3), f_rest_color=r(P, M - 1, 3) * 0.1,
               f_dc_phase=r(P, 1, 1), f_rest_phase=r(P, M - 1, 1) * 0.1, f_dc_amp=r(P, 1, 1),
               f_rest_amp=r(P, M - 1, 1) * 0.1)
    return {k: v.to(device) for k, v in raw.items()}


def make_deltas(Nd, M=16, seed=1, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g) * 0.05
    return {k: v.to(device) for k, v in dict(d_xyz=r(Nd, 3), d_rot=r(Nd, 4), d_sh=r(Nd, M, 3), d_sh_p=r(Nd, M, 2)).items()}


def rel(a, b):
    if b is None:               # autograd reports "input unused" as None; the kernels write zeros
        b = torch.zeros_like(a)
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


# ------------------------------------------------------------------------------------------------
# CPU: the oracle is pinned
# ------------------------------------------------------------------------------------------------
def test_adam_restatement_matches_torch_optim_adam():
    torch.manual_seed(0)
    shapes, lrs = [(1000, 3), (1000, 15, 3), (1000, 1), (7,)], [1.6e-4, 1.25e-4, 0.05, 0.0]
    params = [torch.randn(s) for s in shapes]
    tp = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(tp, lrs)], lr=0.0, eps=1e-15)
    npp = [p.numpy().copy() for p in params]
    m = [np.zeros_like(a) for a in npp]
    v = [np.zeros_like(a) for a in npp]
    for step in range(1, 8):
        grads = [torch.randn(s) * (10.0 ** -(step % 4)) for s in shapes]
        for p, g in zip(tp, grads):
            p.grad = g.clone()
        opt.step()
        for a, g, mm, vv, lr in zip(npp, grads, m, v, lrs):
            orc.adam_step(a, g.numpy(), mm, vv, lr, step)
    for a, p, lr in zip(npp, tp, lrs):
        assert rel(torch.from_numpy(a), p.detach()) < 2e-7
        if lr == 0.0:
            assert np.array_equal(a, p.detach().numpy())   # lr 0: the parameter must not move


def test_ssim_restatement_matches_bruteforce():
    torch.manual_seed(1)
    a, b = torch.rand(2, 13, 17), torch.rand(2, 13, 17)
    assert abs(float(orc.ssim(a, b)) - orc.ssim_bruteforce(a, b)) < 2e-6
    assert abs(float(orc.ssim(a, a)) - 1.0) < 1e-6


def test_assemble_restatement_semantics():
    raw = make_raw(50)
    mask = torch.zeros(50, dtype=torch.bool)
    mask[::3] = True
    d = make_deltas(int(mask.sum()))
    o = orc.assemble(raw, mask, d)
    assert torch.allclose(o["rotations"].norm(dim=1), torch.ones(50), atol=1e-6)
    assert torch.equal(o["shs"][~mask][:, 0], raw["f_dc_color"][~mask][:, 0])
    assert torch.allclose(o["means3D"][mask], raw["xyz"][mask] + d["d_xyz"])
    only_static = orc.assemble(raw, mask, d, render_regions=("static",))
    assert float(only_static["means3D"][mask].abs().max()) == 0.0 and float(only_static["opacities"][mask].abs().max()) == 0.0
    iso = orc.assemble(make_raw(8, isotropic=True), None, None, isotropic=True)
    assert torch.equal(iso["scales"][:, 0], iso["scales"][:, 2])


GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_loss_cases():
    z = np.load(os.path.join(GOLDEN, "train_loss.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    out = []
    for n in names:
        lam, ld, w, nch = (float(x) for x in z[n + "/meta"])
        out.append(dict(name=n, img=torch.from_numpy(z[n + "/img"]), gt=torch.from_numpy(z[n + "/gt"]),
                        loss=float(z[n + "/loss"]), grad=torch.from_numpy(z[n + "/grad"]), ssim=float(z[n + "/ssim"]),
                        kw=dict(kind=str(z[n + "/kind"]), lam=lam, lambda_dssim=ld, w=w, nch=int(nch))))
    return out


def test_loss_restatement_matches_the_reference_functions():
    """tests/golden/train_loss.npz was produced by the reference's OWN utils/loss_utils.py
    (tests/golden/make_train_golden.py imports it from /root/reference): values and gradients."""
    for c in golden_loss_cases():
        leaf = c["img"].clone().requires_grad_(True)
        val = orc.loss_term(leaf, c["gt"], **c["kw"])
        val.backward()
        assert abs(float(val) - c["loss"]) <= 1e-6 * max(1.0, abs(c["loss"])), c["name"]
        assert rel(leaf.grad, c["grad"]) < 1e-5, c["name"]
        assert abs(float(orc.ssim(c["img"], c["gt"])) - c["ssim"]) < 1e-6, c["name"]


def test_rotation_restatement_matches_the_reference_function():
    """build_rotation / inverse_sigmoid of the reference's utils/general_utils.py (fixture)."""
    z = np.load(os.path.join(GOLDEN, "train_misc.npz"))
    R = orc._rotation_matrix(torch.from_numpy(z["q"]))
    assert float((R - torch.from_numpy(z["R"])).abs().max()) < 1e-6
    x = torch.from_numpy(z["x"])
    assert float((torch.log(x / (1 - x)) - torch.from_numpy(z["inv_sigmoid"])).abs().max()) < 1e-6


def test_train_header_symbols_are_exported():
    src = open(os.path.join(ROOT, "include", "gftorf_train.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(gft_[a-z0-9_]+)\s*\(", src)))
    assert names == ["gft_adam_step", "gft_assemble_backward", "gft_assemble_forward", "gft_densify_apply",
                     "gft_densify_plan", "gft_densify_workspace_bytes", "gft_fused_loss",
                     "gft_fused_loss_scratch_bytes", "gft_nvls_allreduce_fused", "gft_nvls_allreduce_sum",
                     "gft_p2p_allreduce_fused", "gft_push_allreduce_fused"]
    from gftorf_b200 import _capi
    lib = C.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n


def test_train_ctypes_structs_match_the_c_layout(tmp_path):
    import subprocess
    from gftorf_b200 import train_ops as T
    structs = {"GftAssembleArgs": T.GftAssembleArgs, "GftAssembleGrads": T.GftAssembleGrads,
               "GftLossArgs": T.GftLossArgs, "GftAdamArgs": T.GftAdamArgs, "GftAdamSegment": T.GftAdamSegment,
               "GftDensifyPlanArgs": T.GftDensifyPlanArgs, "GftDensifyApplyArgs": T.GftDensifyApplyArgs}
    cname = lambda f: "lambda" if f == "lambda_" else f
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gftorf_train.h"', "int main(void){"]
    for s, cls in structs.items():
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f, *_ in cls._fields_:
            lines.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {cname(f)}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for s, cls in structs.items():
        assert int(out[s]) == C.sizeof(cls), s
        for f, *_ in cls._fields_:
            assert int(out[f"{s}.{f}"]) == getattr(cls, f).offset, (s, f)


def test_train_ops_refuse_cpu_tensors():
    from gftorf_b200 import train_ops as T
    with pytest.raises(RuntimeError, match="no CPU path"):
        T.fused_loss(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        T.assemble_gaussians(make_raw(4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        T.FlatAdam([("xyz", torch.zeros(4, 3), 1e-3)])


# ------------------------------------------------------------------------------------------------
# GPU: parity through the C ABI
# ------------------------------------------------------------------------------------------------
ASM_CASES = [
    dict(P=1, mask=None, iso=False, regions=("static", "dynamic")),
    dict(P=1000, mask=None, iso=False, regions=("static", "dynamic")),
    dict(P=777, mask=3, iso=False, regions=("static", "dynamic")),
    dict(P=777, mask=3, iso=False, regions=("static",)),
    dict(P=777, mask=2, iso=True, regions=("dynamic",)),
    dict(P=5000, mask=1, iso=False, regions=("static", "dynamic")),   # every Gaussian dynamic
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ASM_CASES)
def test_assemble_forward_backward_vs_oracle(case):
    from gftorf_b200 import train_ops as T
    P = case["P"]
    raw_c = make_raw(P, isotropic=case["iso"], seed=P)
    mask = None
    deltas_c = None
    if case["mask"]:
        mask = torch.zeros(P, dtype=torch.bool)
        mask[::case["mask"]] = True
        deltas_c = make_deltas(int(mask.sum()), seed=P + 1)
    # oracle with autograd
    leaves = {k: v.clone().requires_grad_(True) for k, v in raw_c.items()}
    dl = {k: v.clone().requires_grad_(True) for k, v in deltas_c.items()} if deltas_c else None
    o = orc.assemble(leaves, mask, dl, case["iso"], case["regions"])
    gouts = {k: torch.randn(v.shape, generator=torch.Generator().manual_seed(7 + i)) for i, (k, v) in enumerate(o.items())}
    torch.autograd.backward([o[k] for k in T.OUT_NAMES], [gouts[k] for k in T.OUT_NAMES])
    # product
    raw_g = {k: v.cuda().requires_grad_(True) for k, v in raw_c.items()}
    dg = {k: v.cuda().requires_grad_(True) for k, v in deltas_c.items()} if deltas_c else None
    dyn = T.dyn_index_from_mask(mask.cuda()) if mask is not None else None
    og = T.assemble_gaussians(raw_g, dyn, dg, case["iso"], case["regions"])
    torch.autograd.backward([og[k] for k in T.OUT_NAMES], [gouts[k].cuda() for k in T.OUT_NAMES])
    for k in T.OUT_NAMES:
        assert og[k].shape == o[k].shape
        assert rel(og[k], o[k].detach()) < 1e-6, k
    for k in raw_c:
        assert rel(raw_g[k].grad, leaves[k].grad) < 2e-6, k
    if dl:
        for k in dl:
            assert rel(dg[k].grad, dl[k].grad) < 2e-6, k


LOSS_CASES = [
    dict(C=3, H=48, W=64, kind="l1", lam=1.0, ld=0.2),
    dict(C=3, H=37, W=53, kind="l1", lam=0.7, ld=0.2),            # ragged tiles
    dict(C=3, H=40, W=40, kind="weighted_l1", lam=1.0, ld=0.2, w=0.01),
    dict(C=7, H=33, W=20, kind="weighted_l1", lam=2.0, ld=0.0, w=0.1, nch=2),
    dict(C=1, H=64, W=48, kind="weighted_l2_quad", lam=1.5, ld=0.2, w=0.05),
    dict(C=1, H=30, W=31, kind="weighted_l1_quad", lam=1.0, ld=0.1, w=0.05),
    dict(C=1, H=16, W=16, kind="l2", lam=1.0, ld=0.0),
    dict(C=2, H=5, W=7, kind="l2", lam=1.0, ld=0.5),              # smaller than the window
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", LOSS_CASES)
def test_fused_loss_value_and_gradient_vs_oracle(case):
    from gftorf_b200 import train_ops as T
    g = torch.Generator().manual_seed(case["H"] * 100 + case["W"])
    img = torch.rand(case["C"], case["H"], case["W"], generator=g) + 0.05
    gt = (img + 0.2 * torch.randn(img.shape, generator=g)).clamp(0, 1.2)
    kw = dict(kind=case["kind"], lam=case["lam"], lambda_dssim=case["ld"], w=case.get("w", 0.0), nch=case.get("nch"))
    leaf = img.clone().requires_grad_(True)
    ref = orc.loss_term(leaf, gt, **kw)
    ref.backward()
    loss, grad = T.fused_loss(img.cuda(), gt.cuda(), **kw)
    assert abs(float(loss) - float(ref)) <= 2e-5 * max(abs(float(ref)), 1e-3)
    assert rel(grad, leaf.grad) < 3e-5
    # the value accumulates into a caller-supplied scalar
    acc = torch.full((1,), 2.0, device="cuda")
    T.fused_loss(img.cuda(), gt.cuda(), loss_out=acc, **kw)
    assert abs(float(acc) - 2.0 - float(ref)) <= 1e-4


@pytest.mark.gpu
def test_fused_loss_vs_reference_generated_golden():
    """The CUDA loss against values and gradients computed by the reference's own loss_utils.py."""
    from gftorf_b200 import train_ops as T
    for c in golden_loss_cases():
        loss, grad = T.fused_loss(c["img"].cuda(), c["gt"].cuda(), **c["kw"])
        assert abs(float(loss) - c["loss"]) <= 2e-5 * max(abs(c["loss"]), 1e-3), c["name"]
        assert rel(grad, c["grad"]) < 3e-5, c["name"]


@pytest.mark.gpu
def test_fused_loss_full_size_properties():
    """640x480 (BASELINE configs[1]): identical images give zero loss and zero gradient; the
    gradient of an L2-only term is exactly linear in the difference."""
    from gftorf_b200 import train_ops as T
    img = torch.rand(3, 480, 640, device="cuda")
    loss, grad = T.fused_loss(img, img.clone(), kind="l1", lambda_dssim=0.2)
    assert abs(float(loss)) < 1e-6 and float(grad.abs().max()) < 1e-9
    gt = torch.rand_like(img)
    _, g1 = T.fused_loss(img, gt, kind="l2", lambda_dssim=0.0)
    assert rel(g1, 2 * (img - gt) / img.numel()) < 1e-6


@pytest.mark.gpu
def test_flat_adam_vs_torch_optim_adam_and_oracle():
    from gftorf_b200 import train_ops as T
    torch.manual_seed(3)
    names = ["xyz", "f_dc_color", "f_rest_color", "opacity", "scaling", "rotation", "phase_offset"]
    shapes = [(2001, 3), (2001, 1, 3), (2001, 15, 3), (2001, 1), (2001, 3), (2001, 4), (1,)]
    lrs = [1.6e-4, 2.5e-3, 1.25e-4, 0.05, 5e-3, 1e-3, 0.0]
    init = [torch.randn(s) for s in shapes]
    tp = [p.clone().requires_grad_(True) for p in init]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(tp, lrs)], lr=0.0, eps=1e-15)
    fa = T.FlatAdam([(n, p.cuda(), lr) for n, p, lr in zip(names, init, lrs)])
    npp = [p.numpy().copy() for p in init]
    m = [np.zeros_like(a) for a in npp]
    v = [np.zeros_like(a) for a in npp]
    for step in range(1, 7):
        grads = [torch.randn(s) * 10.0 ** -(step % 3) for s in shapes]
        if step == 4:                      # a learning-rate change between steps (update_learning_rate)
            lrs[0] = 8e-5
            opt.param_groups[0]["lr"] = lrs[0]
            fa.set_lr("xyz", lrs[0])
        for p, g, n in zip(tp, grads, names):
            p.grad = g.clone()
            fa.params[n].grad.copy_(g.cuda())
        opt.step()
        fa.step(zero_grad=True)
        for a, g, mm, vv, lr in zip(npp, grads, m, v, lrs):
            orc.adam_step(a, g.numpy(), mm, vv, lr, step)
        assert float(fa.grad.abs().max()) == 0.0       # zero_grad fused into the step
    for n, p, a, lr in zip(names, tp, npp, lrs):
        ours = fa.params[n].detach().cpu()
        assert rel(ours, p.detach()) < 3e-7, n
        assert rel(ours, torch.from_numpy(a)) < 3e-7, n
        if lr == 0.0:
            assert torch.equal(ours, p.detach()), n


@pytest.mark.gpu
def test_flat_adam_full_size_is_a_pure_function_of_its_inputs():
    """P = 300k, 91 floats per Gaussian (BASELINE configs[1]): two independent runs of the same
    three steps give bit-identical parameters (no atomics, no ordering dependence)."""
    from gftorf_b200 import train_ops as T
    P = 300000
    res = []
    for _ in range(2):
        torch.manual_seed(11)
        groups = [("xyz", torch.randn(P, 3, device="cuda"), 1.6e-4), ("f_rest_color", torch.randn(P, 45, device="cuda"), 1e-4),
                  ("rest", torch.randn(P, 43, device="cuda"), 1e-3)]
        fa = T.FlatAdam(groups)
        for s in range(3):
            fa.grad.copy_(torch.randn(fa.grad.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(s)))
            fa.step()
        res.append(fa.flat.clone())
    assert torch.equal(res[0], res[1])


# ------------------------------------------------------------------------------------------------
# f2 — densification
# ------------------------------------------------------------------------------------------------
def make_model(P, isotropic=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    params = dict(xyz=r(P, 3), f_dc_color=r(P, 1, 3), f_rest_color=r(P, 15, 3) * 0.1, phase_f_dc=r(P, 1, 1),
                  phase_f_rest=r(P, 15, 1) * 0.1, amp_f_dc=r(P, 1, 1), amp_f_rest=r(P, 15, 1) * 0.1,
                  opacity=r(P, 1) * 2, scaling=r(P, 1 if isotropic else 3) * 1.2 - 3.0, rotation=r(P, 4),
                  f_seg_color=torch.rand(P, 1, generator=g))
    m = {k: r(*v.shape) * 0.01 for k, v in params.items()}
    v = {k: torch.rand(v_.shape, generator=g) * 1e-4 for k, v_ in params.items()}
    acc = torch.rand(P, 1, generator=g)
    den = torch.randint(0, 4, (P, 1), generator=g).float()      # zeros -> NaN gradients -> 0
    return params, m, v, acc, den


def test_densify_restatement_counts_and_order():
    """CPU pin of the oracle: sizes add up, survivors keep their relative order, new Gaussians get
    zero moments, statistics of the three parts match a by-hand classification."""
    P = 400
    params, m, v, acc, den = make_model(P, seed=3)
    kw = dict(max_grad=0.3, min_opacity=0.2, extent=4.0, percent_dense=0.01)
    z = lambda s: torch.normal(mean=torch.zeros_like(s), std=s, generator=torch.Generator().manual_seed(9))
    p2, m2, v2 = orc.densify_and_prune(params, m, v, acc.clone(), den, normal_fn=z, **kw)
    g = torch.nan_to_num(acc / den, nan=0.0).squeeze(-1)
    ms = torch.exp(params["scaling"]).max(dim=1).values
    op = torch.sigmoid(params["opacity"]).squeeze(-1)
    pr = (op < 0.2) | (ms > 0.05 * 4.0) | (ms < 0.001 * 4.0)
    s = (g >= 0.3) & (ms > 0.04)
    c = (g.abs() >= 0.3) & (ms <= 0.04)
    n_keep, n_clone = int((~s & ~pr).sum()), int((c & ~pr).sum())
    assert p2["xyz"].shape[0] >= n_keep + n_clone
    assert torch.equal(p2["f_dc_color"][:n_keep], params["f_dc_color"][~s & ~pr])
    assert torch.equal(m2["xyz"][:n_keep], m["xyz"][~s & ~pr])
    assert torch.equal(p2["rotation"][n_keep:n_keep + n_clone], params["rotation"][c & ~pr])
    assert float(m2["xyz"][n_keep:].abs().max()) == 0.0 and float(v2["opacity"][n_keep:].abs().max()) == 0.0
    assert (p2["xyz"].shape[0] - n_keep - n_clone) % 2 == 0


def golden_assemble_cases():
    z = np.load(os.path.join(GOLDEN, "train_assemble.npz"))
    out = []
    for n in sorted({k.split("/")[0] for k in z.files}):
        t = lambda key: torch.from_numpy(z[n + "/" + key])
        iso, st, dy = (bool(x) for x in z[n + "/meta"])
        c = dict(name=n, mask=t("mask"), iso=iso, regions=tuple(r for r, on in (("static", st), ("dynamic", dy)) if on))
        for pre, names in (("raw", tuple(RAW_KEYS)), ("graw", tuple(RAW_KEYS)), ("delta", DELTA_KEYS), ("gdelta", DELTA_KEYS),
                           ("out", OUT_KEYS), ("gout", OUT_KEYS)):
            c[pre] = {k: t(pre + "/" + k) for k in names}
        out.append(c)
    return out


RAW_KEYS = ("xyz", "opacity_raw", "scaling_raw", "rotation_raw", "f_dc_color", "f_rest_color", "f_dc_phase",
            "f_rest_phase", "f_dc_amp", "f_rest_amp")
DELTA_KEYS = ("d_xyz", "d_rot", "d_sh", "d_sh_p")
OUT_KEYS = ("means3D", "opacities", "scales", "rotations", "shs", "shs_p")


def test_assemble_restatement_matches_the_reference_render():
    """tests/golden/train_assemble.npz: the reference's OWN render() (gaussian_renderer/__init__.py)
    and model getters run on CPU with the rasterizer replaced by a recorder
    (tests/golden/make_assemble_golden.py): the six tensors it hands to the rasterizer, and the
    raw-parameter / deformation gradients autograd derives for a known gradient of those tensors."""
    for c in golden_assemble_cases():
        leaves = {k: v.clone().requires_grad_(True) for k, v in c["raw"].items()}
        dl = {k: v.clone().requires_grad_(True) for k, v in c["delta"].items()}
        o = orc.assemble(leaves, c["mask"], dl, c["iso"], c["regions"])
        for k in OUT_KEYS:
            assert float((o[k] - c["out"][k]).abs().max()) <= 1e-6, (c["name"], k)
        torch.autograd.backward([o[k] for k in OUT_KEYS], [c["gout"][k] for k in OUT_KEYS])
        for k in RAW_KEYS:
            assert rel(leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k]), c["graw"][k]) < 1e-5 \
                or float(c["graw"][k].abs().max()) == 0.0, (c["name"], k)
        for k in DELTA_KEYS:
            got = dl[k].grad if dl[k].grad is not None else torch.zeros_like(dl[k])
            assert float((got - c["gdelta"][k]).abs().max() if got.numel() else 0.0) <= 1e-6, (c["name"], k)


@pytest.mark.gpu
def test_assemble_vs_reference_generated_golden():
    """The CUDA assembly (forward and backward) against the reference's own render() + getters."""
    from gftorf_b200 import train_ops as T
    for c in golden_assemble_cases():
        raw = {k: v.cuda().requires_grad_(True) for k, v in c["raw"].items()}
        dl = {k: v.cuda().requires_grad_(True) for k, v in c["delta"].items()}
        dyn = T.dyn_index_from_mask(c["mask"].cuda())
        o = T.assemble_gaussians(raw, dyn, dl, c["iso"], c["regions"])
        for k in OUT_KEYS:
            assert float((o[k].cpu() - c["out"][k]).abs().max()) <= 2e-6, (c["name"], k)
        torch.autograd.backward([o[k] for k in OUT_KEYS], [c["gout"][k].cuda() for k in OUT_KEYS])
        for k in RAW_KEYS:
            assert float((raw[k].grad.cpu() - c["graw"][k]).abs().max()) <= 2e-5 * max(1.0, float(c["graw"][k].abs().max())), (c["name"], k)
        for k in DELTA_KEYS:
            if dl[k].numel():
                assert float((dl[k].grad.cpu() - c["gdelta"][k]).abs().max()) <= 2e-5 * max(1.0, float(c["gdelta"][k].abs().max())), (c["name"], k)


def golden_densify_cases():
    z = np.load(os.path.join(GOLDEN, "train_densify.npz"))
    out = []
    for n in sorted({k.split("/")[0] for k in z.files}):
        t = lambda key: torch.from_numpy(z[n + "/" + key])
        max_grad, min_op, extent, pd, size, iso = (float(x) for x in z[n + "/meta"])
        c = dict(name=n, acc=t("acc"), den=t("den"), samples=t("samples"),
                 kw=dict(max_grad=max_grad, min_opacity=min_op, extent=extent, percent_dense=pd,
                         size_prune=bool(size), isotropic=bool(iso)))
        for pre in ("in", "in_m", "in_v", "out", "out_m", "out_v"):
            c[pre] = {g: t(pre + "/" + g) for g in orc.GROUPS}
        out.append(c)
    return out


def test_densify_restatement_matches_the_reference_method():
    """tests/golden/train_densify.npz: inputs, recorded split noise and results of the reference's
    OWN GaussianModel.densify_and_prune run on CPU (tests/golden/make_densify_golden.py).  The
    restatement must reproduce the new Gaussian set — order, rows and Adam moments — exactly."""
    for c in golden_densify_cases():
        p, m, v = orc.densify_and_prune(c["in"], c["in_m"], c["in_v"], c["acc"].clone(), c["den"],
                                        normal_fn=lambda s, c=c: c["samples"], **c["kw"])
        for g in orc.GROUPS:
            assert p[g].shape == c["out"][g].shape, (c["name"], g)
            assert float((p[g] - c["out"][g]).abs().max() if p[g].numel() else 0.0) <= 1e-6, (c["name"], g)
            assert torch.equal(m[g], c["out_m"][g]) and torch.equal(v[g], c["out_v"][g]), (c["name"], g)


@pytest.mark.gpu
def test_densify_vs_reference_generated_golden():
    """The CUDA densification against the reference's own densify_and_prune (recorded noise replayed)."""
    from gftorf_b200 import train_ops as T
    cu = lambda d: {k: t.cuda() for k, t in d.items()}
    for c in golden_densify_cases():
        p, m, v, info = T.densify_and_prune(cu(c["in"]), cu(c["in_m"]), cu(c["in_v"]), c["acc"].cuda(), c["den"].cuda(),
                                            samples=c["samples"].cuda(), **c["kw"])
        assert info["P_new"] == c["out"]["xyz"].shape[0], c["name"]
        for g in orc.GROUPS:
            assert p[g].shape == c["out"][g].shape, (c["name"], g)
            assert float((p[g].cpu() - c["out"][g]).abs().max() if p[g].numel() else 0.0) <= 2e-6, (c["name"], g)
            assert torch.equal(m[g].cpu(), c["out_m"][g]) and torch.equal(v[g].cpu(), c["out_v"][g]), (c["name"], g)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [dict(P=1, iso=False, size=True), dict(P=999, iso=False, size=True),
                                  dict(P=5000, iso=False, size=False), dict(P=3001, iso=True, size=True),
                                  dict(P=70000, iso=False, size=True)])
def test_densify_and_prune_vs_oracle(case):
    """Same inputs, same torch.normal stream: the new Gaussian set must come out in the
    reference's order with identical rows (children's positions to 1e-6: bmm vs explicit dot)."""
    from gftorf_b200 import train_ops as T
    P = case["P"]
    params, m, v, acc, den = make_model(P, isotropic=case["iso"], seed=P)
    kw = dict(max_grad=0.3, min_opacity=0.2, extent=4.0, percent_dense=0.01)
    cu = lambda d: {k: t.cuda() for k, t in d.items()}
    gen_o = torch.Generator("cuda").manual_seed(5)
    gen_p = torch.Generator("cuda").manual_seed(5)
    z = lambda s: torch.normal(mean=torch.zeros_like(s), std=s, generator=gen_o)
    # the oracle on GPU tensors, so exp / log / sigmoid are the same device functions
    p2, m2, v2 = orc.densify_and_prune(cu(params), cu(m), cu(v), acc.cuda(), den.cuda(), size_prune=case["size"],
                                       isotropic=case["iso"], normal_fn=z, **kw)
    q2, n2, w2, info = T.densify_and_prune(cu(params), cu(m), cu(v), acc.cuda(), den.cuda(), size_prune=case["size"],
                                           isotropic=case["iso"], generator=gen_p, **kw)
    assert info["P_new"] == p2["xyz"].shape[0]
    for g in orc.GROUPS:
        assert q2[g].shape == p2[g].shape, g
        if g == "xyz":
            assert torch.equal(q2[g][:info["kept"] + info["clones"]], p2[g][:info["kept"] + info["clones"]])
            if q2[g].numel():
                assert float((q2[g] - p2[g]).abs().max()) <= 1e-5 * max(1.0, float(p2[g].abs().max()))
        elif g == "scaling":
            assert float((q2[g] - p2[g]).abs().max() if q2[g].numel() else 0.0) <= 1e-6
        else:
            assert torch.equal(q2[g], p2[g]), g
        assert torch.equal(n2[g], m2[g]) and torch.equal(w2[g], v2[g]), g
