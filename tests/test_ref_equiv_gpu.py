"""The reference-equivalent GPU baseline that bench.py times next to the product (bench_ref_equiv/)
must itself be correct, otherwise its time means nothing: one view forward + backward against the CPU oracle, at the
parity tolerances, and integer-identical radii / n_contrib to the product (both share the projection math)."""
import numpy as np
import pytest
import torch

from dreammesh4d_b200 import rasterizer as R
from tests import helpers as Hh
from tests.test_raster_parity_gpu import run_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P,H,W,seed", [(20000, 256, 256, 1), (3000, 100, 130, 0)])
def test_ref_equiv_matches_oracle(P, H, W, seed):
    from bench_ref_equiv.ref_equiv import RefEquivView
    means, scales, rots, opac, cols = Hh.random_scene(P, seed)
    V, PV, campos, tanx, tany = Hh.cameras(1, seed=seed + 10)
    bg = torch.ones(3)
    o = run_oracle(P, H, W, means, scales, rots, opac, cols, V[0], PV[0], tanx[0], tany[0], bg)
    ok = torch.from_numpy(~o.ambiguous)[None]
    d = lambda x: x.cuda().contiguous()
    vp = R.make_view_params(d(V), d(PV), d(campos), tanx, tany, d(bg[None]))
    g = torch.Generator().manual_seed(seed)
    gC, gD, gA = torch.randn(3, H, W, generator=g) * ok, torch.randn(1, H, W, generator=g) * ok, torch.randn(1, H, W, generator=g) * ok
    view = RefEquivView(P, H, W, torch.device("cuda"))
    n = view.forward_backward(d(means), d(scales), d(rots), d(opac), d(cols), vp[0].contiguous(), d(gC), d(gD), d(gA))
    torch.cuda.synchronize()
    assert n == o.num_rendered
    np.testing.assert_array_equal(view.radii.cpu().numpy(), o.radii)
    okn = ok.numpy()
    for name, got, ref in (("color", view.color, o.color), ("depth", view.depth, o.depth), ("alpha", view.alpha, o.alpha)):
        err = np.abs(got.cpu().numpy() - ref)[np.broadcast_to(okn, ref.shape)].max() / max(np.abs(ref).max(), 1e-30)
        assert err <= Hh.TOL_IMAGE, f"{name}: {err}"
    ref = o.backward(gC.numpy(), gD.numpy(), gA.numpy())
    for name in ("means3D", "means2D", "colors", "opacities", "scales", "rotations"):
        assert Hh.rel_linf(view.grads[name].cpu().numpy(), ref[name]) <= Hh.TOL_GRAD, name
