"""The camera window over the section grid (preparePointBuffer, main.cpp:459-618), CPU side:
the oracle's literal four-memcpy-loops restatement equals the closed form the CUDA kernel implements
(window(x, y) = section[carry_x][carry_y]((x + cx) % res, (y + cy) % res) at every level), and the host arithmetic of the
C ABI (hmrt_window_place, no device needed) equals the oracle's restatement of main.cpp:461-516."""
import numpy as np
import pytest

import oraclelib as ol


def _sections(coarse, levels, seed, with_colors=True):
    rng = np.random.default_rng(seed)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    pyr = [[rng.random(total, dtype=np.float32) * 100 + 10 * (2 * a + b) for b in range(2)] for a in range(2)]
    col = [[rng.integers(0, 256, (res[0], res[0], 3), dtype=np.uint8) for _ in range(2)] for _ in range(2)] if with_colors else None
    return res, idx, total, pyr, col


def closed_form(pyr, col, coarse, levels, cx, cy):
    res, idx, total = ol.pyramid_layout(coarse, levels)
    out = np.empty(total, np.float32)
    for l in range(levels):
        r = res[l]
        sx, sy = cx << (levels - 1 - l), cy << (levels - 1 - l)
        lv = [[pyr[a][b][idx[l]:idx[l] + r * r].reshape(r, r) for b in range(2)] for a in range(2)]
        # rows = y, columns = x; a big 2r x 2r mosaic [bottom rows first] and a window cut out of it
        mosaic = np.block([[lv[0][0], lv[1][0]], [lv[0][1], lv[1][1]]])
        out[idx[l]:idx[l] + r * r] = mosaic[sy:sy + r, sx:sx + r].ravel()
    out_c = None
    if col is not None:
        r = res[0]
        sx, sy = cx << (levels - 1), cy << (levels - 1)
        mosaic = np.concatenate([np.concatenate([col[0][0], col[1][0]], axis=1), np.concatenate([col[0][1], col[1][1]], axis=1)], axis=0)
        out_c = mosaic[sy:sy + r, sx:sx + r].copy()
    return out, out_c


@pytest.mark.parametrize("coarse,levels", [(8, 4), (4, 1), (3, 3), (5, 2), (16, 5)])
def test_oracle_loops_equal_closed_form(coarse, levels):
    res, idx, total, pyr, col = _sections(coarse, levels, seed=coarse * 10 + levels)
    for cx, cy in [(0, 0), (0, coarse - 1), (coarse - 1, 0), (coarse // 2, coarse // 3), (coarse - 1, coarse - 1), (1, 1)]:
        if cx >= coarse or cy >= coarse:
            continue
        got, got_c = ol.oracle_compose_window(pyr, col, coarse, levels, cx, cy)
        want, want_c = closed_form(pyr, col, coarse, levels, cx, cy)
        assert not np.isnan(got).any(), f"cell ({cx},{cy}): the four loops left window cells unwritten"
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), (cx, cy)
        assert (got_c == want_c).all(), (cx, cy)


def _grid_origins(cam, grid, coarse, levels):
    """initializeSections, main.cpp:276-288."""
    size = float(coarse << (levels - 1))
    return np.array([[[np.float32(cam[0]) + np.float32((i - grid / 2.0) * size), np.float32(cam[2]) + np.float32((j - grid / 2.0) * size)]
                      for j in range(grid)] for i in range(grid)], np.float32)


def test_window_place_equals_oracle_and_reference_rules():
    import hmrt

    coarse, levels, grid = 32, 8, 4
    size = coarse << (levels - 1)
    cam0 = (1234.5, 80.0, -321.25)
    org = _grid_origins(cam0, grid, coarse, levels)
    rng = np.random.default_rng(5)
    # cameras anywhere the reference's manageSections would leave them: inside the inner 2 x 2 sections
    cams = [cam0] + [(cam0[0] + float(rng.uniform(-size, size * 0.999)), 50.0, cam0[2] + float(rng.uniform(-size, size * 0.999))) for _ in range(200)]
    cams += [(float(org[1, 1, 0]), 1.0, float(org[1, 1, 1])), (float(org[2, 2, 0]), 1.0, float(org[2, 2, 1]) + 128.0)]  # on section / cell borders
    for cam in cams:
        rc, want = ol.oracle_window_place(cam, org, grid, coarse, levels)
        assert rc == 0
        got = hmrt.window_place(cam, org, grid, coarse, levels)
        for f in ("min_x", "min_y", "max_x", "max_y", "cell_x", "cell_y"):
            assert getattr(got, f) == getattr(want, f), (cam, f)
        assert np.array_equal(np.array(got.camera, np.float32).view(np.uint32), np.array(want.camera, np.float32).view(np.uint32))
        # the window is one section wide: it straddles at most two sections per axis, and starts in the lower-left one
        assert got.max_x - got.min_x in (0, 1) and got.max_y - got.min_y in (0, 1)
        assert (got.cell_x == 0) == (got.max_x == got.min_x) or got.cell_x == 0
        # camera_point_buffer: the camera sits half a window from the window origin, up to the coarse-cell snap (main.cpp:513-516)
        assert (coarse - 1) * 64.0 <= got.camera[0] < (coarse + 1) * 64.0 and got.camera[1] == np.float32(cam[1])
    # the camera of initializeSections itself: window = the four central sections, cut in the middle
    p = hmrt.window_place(cam0, org, grid, coarse, levels)
    assert (p.min_x, p.max_x, p.min_y, p.max_y) == (1, 2, 1, 2) and (p.cell_x, p.cell_y) == (coarse // 2, coarse // 2)
    # a window that leaves the grid is an error, not an out-of-bounds read
    with pytest.raises(hmrt.HmrtError):
        hmrt.window_place((cam0[0] - 2.6 * size, 0.0, cam0[2]), org, grid, coarse, levels)
    assert ol.oracle_window_place((cam0[0] - 2.6 * size, 0.0, cam0[2]), org, grid, coarse, levels)[0] != 0
