"""gftorf_b200 — B200-native RGB+ToF Gaussian rasterizer and 3-NN initialiser behind gftorf's
Python surface.  See DESIGN.md.  Importing this package does not load the CUDA library; the first
call does, and fails loudly if it is not built (there is no fallback path)."""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians,
                         _RasterizeGaussians)
from .knn import distCUDA2
from .views import ViewSpec, rasterize_views, forward_views, backward_views

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
           "_RasterizeGaussians", "distCUDA2", "ViewSpec", "rasterize_views", "forward_views",
           "backward_views"]
