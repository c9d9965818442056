"""Drop-in replacement for the reference package of the same name
(gaussians/diff_gaussian_rasterization_depth_alpha/diff_gaussian_rasterization_depth_alpha/__init__.py):
`from diff_gaussian_rasterization_depth_alpha import GaussianRasterizationSettings, GaussianRasterizer`
as done by the reference's gaussians/gaussian_renderer.py:14 keeps working unchanged; the work is
done by the sm_100a CUDA library of animatablegaussians_b200 (C ABI: include/agr_rasterizer.h).
"""
from animatablegaussians_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeGaussians,
    cpu_deep_copy_tuple,
    rasterize_gaussians,
)
