"""TEST / BENCH INFRASTRUCTURE ONLY — the reference's autograd Function restated over a checker
module (the reference kernels in oracle/_ref, or the CPU port), so the reference's PyTorch
operator chain can be driven end to end around it (bench.py's c3_iter and cpu_pipeline blocks).

Restates _RasterizeGaussians (diff_gaussian_rasterization_w_tof/__init__.py:69-206): the same 27
forward / 36 backward arguments to the native layer, the same 11 outputs, the same routing of the
12 gradients.  `view` is a bench/test view dict (W, H, bg, viewmatrix, projmatrix, campos,
tanfovx, tanfovy, near_n, far_n, depth_range)."""
import torch


def make(mod):
    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, sh_p, opacities, scales, rotations, view):
            e = torch.Tensor([])
            out = mod.rasterize_gaussians(
                view["bg"], means3D, e, e, opacities, scales, rotations, 1.0, e, view["viewmatrix"],
                view["projmatrix"], view["tanfovx"], view["tanfovy"], view["H"], view["W"], sh, sh_p, 3,
                view["campos"], False, False, view["near_n"], view["far_n"], view["depth_range"], False, 0.0, 0.0)
            (R, color, phasor, depth, normal, acc, entropy, dd, ad, pixels, distribution, radii, geom, binning,
             img) = out
            ctx.view, ctx.R = view, R
            ctx.save_for_backward(means3D, scales, rotations, radii, sh, sh_p)
            ctx.buffers = (geom, binning, img)      # tensors (reference kernels) or handles (CPU port)
            ctx.mark_non_differentiable(pixels, radii)
            return color, phasor, depth, normal, acc, entropy, dd, ad, pixels, distribution, radii

        @staticmethod
        def backward(ctx, g_color, g_phasor, g_depth, g_normal, g_acc, g_entropy, g_dd, g_ad, g_pix, g_dist, _):
            v = ctx.view
            means3D, scales, rotations, radii, sh, sh_p = ctx.saved_tensors
            geom, binning, img = ctx.buffers
            e = torch.Tensor([])
            H, W = v["H"], v["W"]
            dev = means3D.device

            def z(t, c):
                return t if t is not None else torch.zeros((c, H, W), dtype=torch.float32, device=dev)
            b = mod.rasterize_gaussians_backward(
                v["bg"], means3D, radii, e, e, scales, rotations, 1.0, e, v["viewmatrix"], v["projmatrix"],
                v["tanfovx"], v["tanfovy"], z(g_color, 3), z(g_phasor, 7), z(g_depth, 1), z(g_normal, 3), z(g_acc, 1),
                z(g_entropy, 1), z(g_dd, 1), z(g_ad, 1), sh, sh_p, 3, v["campos"], geom, ctx.R, binning, img, False,
                v["near_n"], v["far_n"], v["depth_range"], False, 0.0, 0.0)
            # (means2D, colors_precomp, phasors_precomp, opacities, means3D, cov3D, sh, sh_p, scales, rotations, ...)
            return b[4], b[0], b[6], b[7], b[3], b[8], b[9], None

    def rasterize(asm, means2D, view):
        """asm: dict means3D, opacities, scales, rotations, shs, shs_p (the assembly's outputs)."""
        if means2D is None:
            means2D = torch.zeros_like(asm["means3D"])
        return _Fn.apply(asm["means3D"], means2D, asm["shs"], asm["shs_p"], asm["opacities"], asm["scales"],
                         asm["rotations"], view)
    return rasterize
