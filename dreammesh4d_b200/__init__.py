"""dreammesh4d_b200 — B200 (sm_100a) implementation of DreamMesh4D's dynamic-stage hot path
(sparse-control skinning -> surface-bound Gaussian update -> tile rasterizer forward/backward)
behind the reference's plugin surface.  See DESIGN.md."""
__version__ = "0.1.0"
