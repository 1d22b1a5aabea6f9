"""GPU: the zero-edit drop-in path (one GaussianRasterizer per view, called twice like the reference's renderer,
diff_sugar_rasterizer_temporal.py:144,169-178,202-211): plan re-use between the two calls, grow-only capacities with
deferred overflow verification, and the None / single-set conventions of the host bindings."""
import warnings

import numpy as np
import pytest
import torch

from dreammesh4d_b200 import rasterizer as R
from dreammesh4d_b200 import skinning, synthetic
from tests import helpers as Hh

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _settings(V, PV, campos, tanx, tany, H, W, v=0):
    return R.GaussianRasterizationSettings(H, W, float(tanx[v]), float(tany[v]), torch.ones(3, device=DEV), 1.0,
                                           V[v].to(DEV), PV[v].to(DEV), 0, campos[v].to(DEV), False, False)


def _scene(P=4000, seed=0):
    means, scales, rots, opac, cols = Hh.random_scene(P, seed=seed)
    return [t.to(DEV).requires_grad_(True) for t in (means, scales, rots, opac, cols)]


def test_second_call_reuses_the_plan_and_equals_a_fresh_call():
    H = W = 96
    cams = Hh.cameras(1, seed=3)
    means, scales, rots, opac, cols = _scene()
    normals = torch.nn.functional.normalize(torch.randn(means.shape[0], 3, device=DEV), dim=-1).requires_grad_(True)
    gC = torch.randn(3, H, W, device=DEV)
    gD, gA = torch.randn(1, H, W, device=DEV) * 0.1, torch.randn(1, H, W, device=DEV)

    def two_calls(fast):
        for t in (means, scales, rots, opac, cols, normals):
            t.grad = None
        r1 = R.GaussianRasterizer(_settings(*cams, H, W))
        m2d = torch.zeros_like(means, requires_grad=True)
        c1, rad1, d1, a1 = r1(means3D=means, means2D=m2d, opacities=opac, colors_precomp=cols, scales=scales, rotations=rots)
        r2 = r1 if fast else R.GaussianRasterizer(_settings(*cams, H, W))      # a fresh instance has no plan to re-use
        c2, rad2, d2, a2 = r2(means3D=means, means2D=torch.zeros_like(means), opacities=opac, colors_precomp=normals,
                              scales=scales, rotations=rots)
        hit = r2._plan_hit(means, scales, rots, opac) is not None and r2 is r1
        ((c1 * gC).sum() + (d1 * gD).sum() + (a1 * gA).sum() + (c2 * gC.flip(0)).sum() + (a2 * gA).sum() * 0.5).backward()
        return (c1, d1, a1, c2, d2, a2, rad1, rad2), [t.grad.clone() for t in (means, scales, rots, opac, cols, normals, m2d)], hit

    fast, gfast, hit = two_calls(True)
    slow, gslow, _ = two_calls(False)
    assert hit
    for a, b in zip(fast, slow):
        assert torch.equal(a, b)                      # same kernels on the same sorted stream: bit for bit
    for a, b, name in zip(gfast, gslow, ("means", "scales", "rots", "opac", "cols", "normals", "means2D")):
        assert Hh.rel_linf(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5, name      # atomics: summation order only


def test_plan_is_not_reused_for_other_tensors_or_after_an_in_place_update():
    H = W = 64
    cams = Hh.cameras(1, seed=4)
    means, scales, rots, opac, cols = _scene(2000, seed=1)
    r = R.GaussianRasterizer(_settings(*cams, H, W))
    args = dict(means2D=None, opacities=opac, colors_precomp=cols, scales=scales, rotations=rots)
    with torch.no_grad():
        c1, *_ = r(means3D=means, **args)
        assert r._plan_hit(means, scales, rots, opac) is not None
        moved = (means.detach() + 0.05).requires_grad_(True)
        assert r._plan_hit(moved, scales, rots, opac) is None                  # another tensor object
        c2, *_ = r(means3D=moved, **args)
        assert not torch.equal(c1, c2)
        means.data.add_(0.05)                                                  # .data edits do not bump the version ...
        means.add_(0.0)                                                        # ... a real in-place op does
        assert r._plan_hit(means, scales, rots, opac) is None


def test_capacity_book_goes_asynchronous_and_recovers_from_an_overflow():
    H = W = 64
    cams = Hh.cameras(1, seed=5)
    book = R.CAPACITY_BOOK
    small = _scene(1500, seed=2)
    key = (1500, H, W, 3, torch.cuda.current_device())
    book.high.pop(key, None)
    book.slack, slack0 = 16, book.slack          # the default 64k-instance slack would hide the jump on this tiny scene

    def call(scene, scale=1.0):
        m, s, q, o, c = scene
        r = R.GaussianRasterizer(_settings(*cams, H, W))
        with torch.no_grad():
            out = r(means3D=m, means2D=None, opacities=o, colors_precomp=c, scales=s * scale, rotations=q)
        return out[0], r._plan[0]

    c_exact, st0 = call(small)
    assert st0.pending is None and key in book.high                           # first call of a shape: exact sizing
    c_spec, st1 = call(small)
    assert st1.capacity > st0.capacity and torch.equal(c_exact, c_spec)        # speculative capacity, same image
    torch.cuda.synchronize()
    book.poll()
    assert st1.pending is None and not st1.overflowed
    # same shape, Gaussians 12x larger: the instance count jumps far beyond the margin -> detected, warned, recovered
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        _, st2 = call(small, scale=12.0)
        torch.cuda.synchronize()
        book.poll()
    assert st2.overflowed and any("speculative capacity" in str(x.message) for x in w)
    c_big, st3 = call(small, scale=12.0)
    assert st3.pending is None and not st3.overflowed                          # sized exactly again
    book.slack = slack0
    exact = R.rasterize_batch(small[0].detach(), small[3].detach(), small[1].detach() * 12.0, small[2].detach(), small[4].detach(),
                              R.GaussianRasterizer(_settings(*cams, H, W))._view_params(), H, W)[0][0]
    assert torch.equal(c_big, exact)


def test_single_set_attributes_broadcast_over_views():
    """[1,P,k] attributes are one shared set (advisor finding: they used to be indexed with the view's set number)."""
    H = W = 64
    B = 3
    V, PV, campos, tanx, tany = Hh.cameras(B, seed=6)
    means, scales, rots, opac, cols = [t.detach() for t in _scene(1500, seed=3)]
    vp = R.make_view_params(V.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, torch.ones(B, 3), set_index=torch.arange(B))
    sets = means[None].repeat(B, 1, 1).requires_grad_(True)                   # per-view means, everything else shared
    sh2 = [t.clone().requires_grad_(True) for t in (scales, rots, opac, cols)]
    sh3 = [t.clone()[None].requires_grad_(True) for t in (scales, rots, opac, cols)]
    g = torch.randn(B, 3, H, W, device=DEV)
    outs = []
    for (s, q, o, c) in (sh2, sh3):
        col, _, dep, alp = R.rasterize_batch(sets, o, s, q, c, vp, H, W, distinct_sets=True)
        (col * g).sum().backward()
        outs.append(col.detach())
    assert torch.equal(outs[0], outs[1])
    for a, b in zip(sh2, sh3):
        assert a.grad.shape == a.shape and b.grad.shape == b.shape
        assert Hh.rel_linf(b.grad[0].cpu().numpy(), a.grad.cpu().numpy()) <= 1e-5


@pytest.mark.parametrize("method", ["lbs", "dqs"])
def test_skinning_accepts_the_reference_none_conventions(method):
    """dynamic_sugar.py:144-145,401-404: scale is None for dqs (without d_scale), opacity is None unless hybrid."""
    scene = synthetic.make_sugar_scene(2_000, g=3)
    graph = synthetic.make_deform_graph(scene.verts, 32, 4)
    trans, rot, scale, opac = [t.to(DEV).requires_grad_(True) for t in synthetic.random_node_attrs(2, 32, seed=3)]
    d = lambda t: t.to(DEV)
    rq, _ = skinning.sugar_rest_frames(d(scene.verts), d(scene.faces.int()), d(scene.complex_rot), scene.g)
    common = (d(scene.verts), d(scene.faces.int()), d(graph.nbr_idx.int()), d(graph.nbr_w), d(scene.bary), rq)
    full = skinning.skin_gaussians(trans, rot, scale, opac, *common, method=method)
    lean = skinning.skin_gaussians(trans, rot, None if method == "dqs" else scale, None, *common, method=method)
    for a, b in zip(full, lean):
        assert torch.equal(a, b)
    (lean[0].sum() + lean[1].sum()).backward()
    assert trans.grad is not None and rot.grad is not None and opac.grad is None
    with pytest.raises(ValueError):
        skinning.skin_gaussians(trans, rot, scale, None, *common, method="hybrid")


def test_alpha_edited_in_place_is_refused_by_the_backward():
    H = W = 32
    cams = Hh.cameras(1, seed=7)
    means, scales, rots, opac, cols = _scene(500, seed=4)
    r = R.GaussianRasterizer(_settings(*cams, H, W))
    c, _, d, a = r(means3D=means, means2D=None, opacities=opac, colors_precomp=cols, scales=scales, rotations=rots)
    a.detach().mul_(0.5)
    with pytest.raises(RuntimeError, match="modified in place"):
        c.sum().backward()
