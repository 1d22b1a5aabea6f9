#!/usr/bin/env python
"""CPU simulation behind the render kernels' culling design: on one view of the C3 bench scene,
using the CPU oracle's projected Gaussians, sorted per-tile instance lists and n_contrib, counts

  * the (unit, instance) pairs per instance for 16x2 strips, 8x4 blocks and 4x4 / 4x2 / 2x2 cells (exact
    contribution test: power <= 0 and alpha >= 1/255 for some pixel of the unit), and
  * the backward's warp iterations per instance when every warp (an 8x4 block) runs one instance queue per 4x4 cell
    (the shipped design), per 4x2 cell or per 2x2 cell, with the queues rebuilt every CHUNK staged instances and the
    warp stepping max(queue lengths) times per chunk.

Test/analysis infrastructure (imports the oracle); usage: python scripts/sim_cell_queues.py [CHUNK=64]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from oracle.raster_oracle import RasterOracle  # noqa: E402


def main():
    CH = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    scene, graph, node = bench.build_scene(False)
    V, PV, campos, tanx, tany = bench.build_cameras(0)
    o = bench.gaussian_sets_oracle(scene, graph, node)
    P, H, W, v = scene.n_gaussians, 512, 512, 0
    scales = torch.cat([torch.full((P, 1), scene.thickness), scene.log_scales.exp()], -1).numpy()
    orc = RasterOracle(P, H, W, 3, "f32")
    orc.forward(o["means3D"][v].numpy(), scales, o["rotations"][v].numpy(), torch.sigmoid(scene.densities).numpy(),
                o["colors"].numpy(), V[v].numpy(), PV[v].numpy(), float(tanx[v]), float(tany[v]), np.ones(3, np.float32))
    xy, co = orc.xy.copy(), orc.conic_opacity.copy()
    ids, _ = orc.point_list()
    rg, nc = orc.ranges.astype(np.int64), orc.n_contrib
    gx = (W + 15) // 16
    yy, xx = np.mgrid[0:16, 0:16]
    inst = px = 0
    pairs = dict.fromkeys(("16x2", "8x4", "4x4", "4x2", "2x2"), 0)
    iters = dict.fromkeys(("4x4", "4x2", "2x2"), 0)
    for t in range(len(rg)):
        b, e = rg[t]
        if e <= b:
            continue
        g = ids[b:e].astype(np.int64)
        n = len(g)
        inst += n
        tx, ty = t % gx, t // gx
        dx = xy[g, 0][:, None, None] - (tx * 16 + xx)[None]
        dy = xy[g, 1][:, None, None] - (ty * 16 + yy)[None]
        c = co[g]
        power = -0.5 * (c[:, 0, None, None] * dx * dx + c[:, 2, None, None] * dy * dy) - c[:, 1, None, None] * dx * dy
        ok = (power <= 0) & (np.minimum(0.99, c[:, 3, None, None] * np.exp(power)) >= 1 / 255)
        px += ok.sum()
        pairs["16x2"] += ok.reshape(n, 8, 2, 16).any(axis=(2, 3)).sum()
        pairs["8x4"] += ok.reshape(n, 4, 4, 2, 8).any(axis=(2, 4)).sum()
        cells = {"4x4": ok.reshape(n, 4, 4, 4, 4).any(axis=(2, 4)), "4x2": ok.reshape(n, 8, 2, 4, 4).any(axis=(2, 4)),
                 "2x2": ok.reshape(n, 8, 2, 8, 2).any(axis=(2, 4))}
        for k, m in cells.items():
            pairs[k] += m.sum()
        last = nc[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16].astype(np.int64)
        for w in range(8):
            bx, by = w & 1, w >> 1
            nl = min(n, int(last[by * 4:(by + 1) * 4, bx * 8:(bx + 1) * 8].max()))     # warp_last
            if nl == 0:
                continue
            pad = (-nl) % CH
            per_chunk = lambda a: np.concatenate([a[:nl], np.zeros(pad, bool)]).reshape(-1, CH).sum(1)
            iters["4x4"] += np.max([per_chunk(cells["4x4"][:, by, 2 * bx + i]) for i in range(2)], axis=0).sum()
            iters["4x2"] += np.max([per_chunk(cells["4x2"][:, 2 * by + j, 2 * bx + i]) for j in range(2) for i in range(2)], axis=0).sum()
            iters["2x2"] += np.max([per_chunk(cells["2x2"][:, 2 * by + j, 4 * bx + i]) for j in range(2) for i in range(4)], axis=0).sum()
    print(f"view {v}: {inst} instances, {px / inst:.2f} contributing pixels per instance, queue window {CH}")
    for k, val in pairs.items():
        print(f"  pairs per instance, unit {k:>4}: {val / inst:.3f}")
    for k, val in iters.items():
        print(f"  backward warp iterations per instance, one queue per {k} cell: {val / inst:.3f}")


if __name__ == "__main__":
    main()
