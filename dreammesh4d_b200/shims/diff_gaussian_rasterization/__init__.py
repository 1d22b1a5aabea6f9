"""Drop-in module named like the rasterizer the reference imports
(``from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`` —
custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:8-11,
.../diff_sugar_rasterizer_normal.py:8-11).  Put ``dreammesh4d_b200/shims`` on PYTHONPATH (or call
``dreammesh4d_b200.install_shim()``) and the plugin, launch.py and the YAML configs run unchanged on libdm4d.so."""
from dreammesh4d_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_batch  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_batch"]
