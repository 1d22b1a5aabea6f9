"""dreammesh4d_b200 — B200 (sm_100a) implementation of DreamMesh4D's dynamic-stage hot path
(sparse-control skinning -> surface-bound Gaussian update -> tile rasterizer forward/backward)
behind the reference's plugin surface.  See DESIGN.md."""
__version__ = "0.1.0"


def install_shim() -> None:
    """Registers dreammesh4d_b200's rasterizer under the module name the reference imports
    (``diff_gaussian_rasterization``) without touching the reference tree."""
    import importlib
    import sys
    from pathlib import Path
    shim = str(Path(__file__).resolve().parent / "shims")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    importlib.import_module("diff_gaussian_rasterization")
