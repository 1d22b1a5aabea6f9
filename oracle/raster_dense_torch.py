"""Dense per-pixel PyTorch (fp64, CPU, autograd) formulation of the rasterizer forward.

ORACLE / TEST INFRASTRUCTURE ONLY — used to cross-check the hand-derived backward in
raster_oracle.c against torch autograd of the Appendix A.2
forward on small scenes (P <= a few hundred, images <= 64x64).  PARITY UNPINNED: the
reference's rasterizer (diff-gaussian-rasterization, ashawkey fork,
/root/reference/README.md:35) is not available; call sites anchoring the semantics:
custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:129-144,169-178.

Gradient conventions mirrored from A.3: straight-through min(0.99, .) clamp, no gradient
through skip/termination decisions, no gradient to bg, means2D gradient in NDC units.
"""
from __future__ import annotations

import torch


def quat_to_R(q: torch.Tensor) -> torch.Tensor:
    r, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y),
    ], dim=-1).reshape(*q.shape[:-1], 3, 3)


def project(means3D, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy, H, W, scale_modifier=1.0,
            means2D=None):
    """A.2 steps 1-6 with autograd. Returns dict(depth, xy, conic, valid)."""
    P = means3D.shape[0]
    V = viewmatrix.reshape(4, 4)          # transposed convention: p_view = [p,1] @ V
    PV = projmatrix.reshape(4, 4)
    ph = torch.cat([means3D, torch.ones(P, 1, dtype=means3D.dtype)], dim=1)
    t = ph @ V[:, :3]
    hom = ph @ PV
    p_w = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * p_w[:, None]
    if means2D is not None:
        ndc = ndc + means2D[:, :2]
    Rm = quat_to_R(rotations)
    L = Rm * (scale_modifier * scales)[:, None, :]
    Sigma = L @ L.transpose(1, 2)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    tz = t[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    cx = torch.clamp(t[:, 0] / tz, -limx, limx) * tz
    cy = torch.clamp(t[:, 1] / tz, -limy, limy) * tz
    J = torch.zeros(P, 2, 3, dtype=means3D.dtype)
    J[:, 0, 0] = fx / tz
    J[:, 0, 2] = -(fx * cx) / (tz * tz)
    J[:, 1, 1] = fy / tz
    J[:, 1, 2] = -(fy * cy) / (tz * tz)
    W3 = V[:3, :3].t()                    # W3[i][j] = V[i + 4 j] in flat indexing
    A = J @ W3
    cov = A @ Sigma @ A.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + 0.3
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], dim=-1)
    xy = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], dim=-1)
    return {"depth": tz, "xy": xy, "conic": conic, "valid": tz > 0.2}


def render_dense(means3D, scales, rotations, opacities, colors, viewmatrix, projmatrix, tanfovx, tanfovy, bg,
                 H, W, rect, radii, scale_modifier=1.0, means2D=None):
    """Dense compositing. ``rect`` [P,4] int (tile rect min.x,min.y,max.x,max.y) and ``radii`` [P]
    come from the C oracle (integer decisions are not differentiable and are shared).
    Returns color [C,H,W], depth [1,H,W], alpha [1,H,W], n_contrib-free."""
    g = project(means3D, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy, H, W, scale_modifier, means2D)
    dt = means3D.dtype
    C = colors.shape[1]
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pfx, pfy = xs.to(dt), ys.to(dt)
    tile_x, tile_y = xs // 16, ys // 16
    order = sorted(range(means3D.shape[0]), key=lambda i: (float(g["depth"][i].detach()), i))
    T = torch.ones(H, W, dtype=dt)
    done = torch.zeros(H, W, dtype=torch.bool)
    Cacc = torch.zeros(C, H, W, dtype=dt)
    D = torch.zeros(H, W, dtype=dt)
    Wgt = torch.zeros(H, W, dtype=dt)
    for i in order:
        if int(radii[i]) <= 0:
            continue
        r = [int(v) for v in rect[i]]
        in_rect = (tile_x >= r[0]) & (tile_x < r[2]) & (tile_y >= r[1]) & (tile_y < r[3])
        dx = g["xy"][i, 0] - pfx
        dy = g["xy"][i, 1] - pfy
        con = g["conic"][i]
        power = -0.5 * (con[0] * dx * dx + con[2] * dy * dy) - con[1] * dx * dy
        ea = opacities[i].reshape(()) * torch.exp(power)
        alpha = ea + (torch.clamp(ea, max=0.99) - ea).detach()      # straight-through clamp (A.3)
        live = in_rect & ~done & ~(power.detach() > 0) & ~(alpha.detach() < 1.0 / 255.0)
        test_T = T * (1 - alpha)
        stop = live & (test_T.detach() < 1e-4)
        done = done | stop
        live = live & ~stop
        w = torch.where(live, alpha * T, torch.zeros_like(T))
        Cacc = Cacc + colors[i][:, None, None] * w
        D = D + g["depth"][i] * w
        Wgt = Wgt + w
        T = torch.where(live, test_T, T)
    color = Cacc + T * bg.detach()[:, None, None]
    return color, D[None], Wgt[None]
