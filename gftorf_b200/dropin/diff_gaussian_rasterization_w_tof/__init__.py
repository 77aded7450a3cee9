"""Drop-in for the reference package of the same name
(submodules/diff-gaussian-rasterization-w-tof/diff_gaussian_rasterization_w_tof/__init__.py).
Put `<repo>/gftorf_b200/dropin` (and `<repo>`) on sys.path and the unchanged reference code
(`from diff_gaussian_rasterization_w_tof import GaussianRasterizationSettings, GaussianRasterizer`,
gaussian_renderer/__init__.py:14) runs on the B200-native library."""
from gftorf_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                    rasterize_gaussians, _RasterizeGaussians, _C,
                                    cpu_deep_copy_tuple)
