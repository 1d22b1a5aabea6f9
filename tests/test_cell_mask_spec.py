"""Specification checks (CPU, numpy) of two pieces of index logic the render kernels rely on.  They restate the
formulas of dreammesh4d_b200/csrc/raster_binning.cu (`cell_mask`) and raster_render.cu (`PixelMap`,
`half_reduce_scatter10` / `slot_of10`) and check the PROPERTIES the kernels need: the 4x4-cell mask never drops a
contributing pixel and is tight; the pixel map is a bijection tile <-> (warp, lane) consistent with the mask bit order;
the 10-slot half-warp reduction leaves slot s's total in exactly one lane per half.  The kernels themselves are checked
against the oracle on the GPU (tests/test_raster_parity_gpu.py, bit-exact n_contrib); these tests pin the derivations."""
import numpy as np

f = np.float32


def cell_mask_spec(ax, ay, cx, cy, cz, thr, tile_x0, tile_y0, rows=4):
    """numpy restatement of cell_mask(): [n, 16/rows, 4] bool, index [row j, column i]; rows = DM4D_CELL_ROWS."""
    det = cx * cz - cy * cy
    icx, idet = f(1) / cx, f(1) / det
    X, Y = np.sqrt(thr * cz * idet), np.sqrt(thr * cx * idet)
    yX = -cy * X / cz
    tcx = thr * cx
    mask = np.zeros((len(ax), 16 // rows, 4), bool)
    for j in range(16 // rows):
        y0 = tile_y0 + f(rows * j) - ay - f(0.01)
        y1 = y0 + f(3.02 if rows == 4 else 1.02)
        live = ~((y0 > Y) | (y1 < -Y))
        ya, yb = np.clip(y0, -Y, Y), np.clip(y1, -Y, Y)
        da, db = np.sqrt(np.maximum(tcx - det * ya * ya, 0)), np.sqrt(np.maximum(tcx - det * yb * yb, 0))
        xr = np.maximum((da - cy * ya) * icx, (db - cy * yb) * icx)
        xl = np.minimum((-da - cy * ya) * icx, (-db - cy * yb) * icx)
        xr = np.where((yX >= y0) & (yX <= y1), X, xr)
        xl = np.where((-yX >= y0) & (-yX <= y1), -X, xl)
        for i in range(4):
            x0 = tile_x0 + f(4 * i) - ax - f(0.01)
            mask[:, j, i] = live & (xl <= x0 + f(3.02)) & (xr >= x0) & (thr > 0)
    return mask


def test_cell_mask_is_conservative_and_tight():
    g = np.random.default_rng(0)
    n = 60_000
    # random ellipses (conics) around random centres near a tile, incl. strongly elongated and sub-pixel ones
    ang = g.uniform(0, np.pi, n)
    s1, s2 = np.exp(g.uniform(np.log(0.55), np.log(9.0), n)), np.exp(g.uniform(np.log(0.55), np.log(9.0), n))   # sigmas in px
    c, s = np.cos(ang), np.sin(ang)
    a_ = c * c * s1 ** 2 + s * s * s2 ** 2
    b_ = c * s * (s1 ** 2 - s2 ** 2)
    c_ = s * s * s1 ** 2 + c * c * s2 ** 2
    det = a_ * c_ - b_ * b_
    cx, cy, cz = (c_ / det).astype(f), (-b_ / det).astype(f), (a_ / det).astype(f)
    op = g.uniform(0.01, 0.99, n).astype(f)
    ax, ay = g.uniform(-12, 28, n).astype(f), g.uniform(-12, 28, n).astype(f)
    t = np.log(f(255.0) * op)
    thr = np.where(t > 0, f(2.0) * (t + f(1e-4)) * f(1.002), f(0)).astype(f)          # raster_preprocess.cu
    yy, xx = np.mgrid[0:16, 0:16]
    dx, dy = ax[:, None, None] - xx[None].astype(f), ay[:, None, None] - yy[None].astype(f)
    power = f(-0.5) * (cx[:, None, None] * dx * dx + cz[:, None, None] * dy * dy) - cy[:, None, None] * dx * dy
    alpha = np.minimum(f(0.99), op[:, None, None] * np.exp(power))
    contributes = (power <= 0) & (alpha >= f(1 / 255))
    exact = contributes.reshape(n, 4, 4, 4, 4).any(axis=(2, 4))
    mask = cell_mask_spec(ax, ay, cx, cy, cz, thr, f(0), f(0))
    assert not (exact & ~mask).any(), "the cell mask dropped a contributing pixel"
    assert exact.sum() > 50_000
    assert mask.sum() <= 1.06 * exact.sum(), (mask.sum(), exact.sum())      # tight: only the 0.01 px / threshold margins
    # the experimental 4x2 cells (DM4D_CELL_ROWS=2): same test with 8 row strips
    exact2 = contributes.reshape(n, 8, 2, 4, 4).any(axis=(2, 4))
    mask2 = cell_mask_spec(ax, ay, cx, cy, cz, thr, f(0), f(0), rows=2)
    assert not (exact2 & ~mask2).any() and mask2.sum() <= 1.08 * exact2.sum()


def test_pixel_map_and_mask_bit_order():
    seen = set()
    for warp in range(8):
        for lane in range(32):
            half, li = lane >> 4, lane & 15
            px = ((((warp & 1) << 1) | half) << 2) + (li & 3)
            py = ((warp >> 1) << 2) + (li >> 2)
            assert 2 * warp + half == (py // 4) * 4 + px // 4          # cell_bit == 4 cy + cx
            seen.add((px, py))
    assert len(seen) == 256


def test_half_warp_reduction_slot_layout():
    g = np.random.default_rng(1)
    v = g.standard_normal((32, 10))
    li = np.arange(32) & 15
    b3, b2, b1, b0 = (li & 8) != 0, (li & 4) != 0, (li & 2) != 0, (li & 1) != 0
    shfl = lambda x, d: x[np.arange(32) ^ d]
    w, x, y = np.zeros((32, 6)), np.zeros((32, 4)), np.zeros((32, 2))
    for i in range(5):
        w[:, i] = np.where(b3, v[:, i + 5], v[:, i]) + shfl(np.where(b3, v[:, i], v[:, i + 5]), 8)
    for i in range(3):
        x[:, i] = np.where(b2, w[:, i + 3], w[:, i]) + shfl(np.where(b2, w[:, i], w[:, i + 3]), 4)
    for i in range(2):
        y[:, i] = np.where(b1, x[:, i + 2], x[:, i]) + shfl(np.where(b1, x[:, i], x[:, i + 2]), 2)
    tot = np.where(b0, y[:, 1], y[:, 0]) + shfl(np.where(b0, y[:, 0], y[:, 1]), 1)

    def slot_of10(l):
        t = l & 3
        u = (3 if l & 4 else 0) + t
        return ((5 if l & 8 else 0) + u) if (t <= 2 and u <= 4) else -1

    for half in range(2):
        owners = {}
        for l in range(16):
            s = slot_of10(l)
            if s >= 0:
                assert s not in owners
                owners[s] = l
                assert abs(tot[half * 16 + l] - v[half * 16:(half + 1) * 16, s].sum()) < 1e-12
        assert sorted(owners) == list(range(10))


# ---- EXPERIMENTAL 4x2-cell variant (DM4D_CELL_ROWS=2, quarter-warp queues; not the default build) -------------------
def test_experimental_4x2_pixel_map_and_reductions():
    seen = set()
    for warp in range(8):
        for lane in range(32):
            grp, li = lane >> 3, lane & 7
            cx, cy8 = ((warp & 1) << 1) | (grp & 1), ((warp >> 1) << 1) | (grp >> 1)
            px, py = (cx << 2) + (li & 3), (cy8 << 1) + (li >> 2)
            assert 4 * cy8 + cx == (py // 2) * 4 + px // 4 and 0 <= 4 * cy8 + cx < 32
            # same 8x4 block per warp as the default layout
            assert (px // 8, py // 4) == (warp & 1, warp >> 1)
            seen.add((px, py))
    assert len(seen) == 256
    g = np.random.default_rng(2)
    shfl = lambda x, d: x[np.arange(32) ^ d]
    li = np.arange(32) & 7
    b2, b1, b0 = (li & 4) != 0, (li & 2) != 0, (li & 1) != 0
    # 10 slots -> up to two per lane
    v = g.standard_normal((32, 10))
    w, x, out = np.zeros((32, 6)), np.zeros((32, 4)), np.zeros((32, 2))
    for i in range(5):
        w[:, i] = np.where(b2, v[:, i + 5], v[:, i]) + shfl(np.where(b2, v[:, i], v[:, i + 5]), 4)
    for i in range(3):
        x[:, i] = np.where(b1, w[:, i + 3], w[:, i]) + shfl(np.where(b1, w[:, i], w[:, i + 3]), 2)
    for i in range(2):
        out[:, i] = np.where(b0, x[:, i + 2], x[:, i]) + shfl(np.where(b0, x[:, i], x[:, i + 2]), 1)

    def quarter_slot10(l, k):
        t = (2 if l & 1 else 0) + k
        u = (3 if l & 2 else 0) + t
        return ((5 if l & 4 else 0) + u) if (t <= 2 and u <= 4) else -1

    for grp in range(4):
        owners = {}
        for l in range(8):
            for k in range(2):
                s = quarter_slot10(l, k)
                if s >= 0:
                    assert s not in owners
                    owners[s] = (l, k)
                    assert abs(out[grp * 8 + l, k] - v[grp * 8:(grp + 1) * 8, s].sum()) < 1e-12
        assert sorted(owners) == list(range(10))
    # 16 slots -> lane l holds slots 2l, 2l+1
    v = g.standard_normal((32, 16))
    w8, w4, out = np.zeros((32, 8)), np.zeros((32, 4)), np.zeros((32, 2))
    for i in range(8):
        w8[:, i] = np.where(b2, v[:, i + 8], v[:, i]) + shfl(np.where(b2, v[:, i], v[:, i + 8]), 4)
    for i in range(4):
        w4[:, i] = np.where(b1, w8[:, i + 4], w8[:, i]) + shfl(np.where(b1, w8[:, i], w8[:, i + 4]), 2)
    for i in range(2):
        out[:, i] = np.where(b0, w4[:, i + 2], w4[:, i]) + shfl(np.where(b0, w4[:, i], w4[:, i + 2]), 1)
    for lane in range(32):
        for k in range(2):
            assert abs(out[lane, k] - v[(lane >> 3) * 8:(lane >> 3) * 8 + 8, 2 * (lane & 7) + k].sum()) < 1e-12
