"""CPU test of the division-free tile cursor that every kernel uses to walk a
block (bri17_b200/csrc/modal_kernels.cu: TileCursor / make_geom), replayed on
the host through bri17_debug_walk_tiles: every tile is visited exactly once, by
exactly one CTA, with (row, chunk, a, b) consistent with the row-major
numbering of tests/test_bri17.cpp:62-64,71 / :76-79,88."""
import ctypes as C

import numpy as np
import pytest

import bri17_b200 as b
from bri17_b200 import _lib

CASES = [
    # shape, tile, max_ctas
    ((3, 4, 5), 512, 592), ((512, 512, 512), 512, 296), ((64, 64), 256, 736), ((4096, 4096), 256, 736),
    ((5, 7, 600), 512, 296), ((3, 2, 1025), 512, 7), ((2, 4097), 512, 3), ((300, 5), 128, 1000),
    ((1, 1, 1), 512, 296), ((40, 24, 16), 256, 100), ((6, 5, 513), 512, 296), ((1, 1, 100000), 512, 37),
    ((7, 3, 3000), 512, 4), ((1000, 3), 512, 592),
]


@pytest.mark.parametrize("shape,tile,max_ctas", CASES)
def test_cursor_covers_every_tile_once(shape, tile, max_ctas):
    dim = len(shape)
    grid_cls = b.CartesianGrid2f64 if dim == 2 else b.CartesianGrid3f64
    hooke = (b.Hooke2f64 if dim == 2 else b.Hooke3f64)(1.0, 0.3, grid_cls(shape, (1.0,) * dim))
    lib = _lib.load()
    n_inner = shape[-1]
    n_mid = shape[1] if dim == 3 else 1
    n_rows = shape[0] * n_mid
    cpr = -(-n_inner // tile)
    n_tiles = n_rows * cpr
    if n_tiles > 600000:          # sample CTAs on the big grids
        ctas = [0, 1, 2, max_ctas // 2, max_ctas - 1]
    else:
        ctas = range(max_ctas)
    cap = n_tiles // max(1, min(max_ctas, n_tiles) - cpr) + cpr + 8
    seen = np.zeros(n_tiles, dtype=np.int32) if n_tiles <= 600000 else None
    grid = C.c_int()
    for cta in ctas:
        out = np.zeros((cap, 5), dtype=np.int64)
        n = lib.bri17_debug_walk_tiles(hooke._plan, None, None, tile, max_ctas, cta,
                                       out.ctypes.data_as(C.POINTER(C.c_int64)), cap, C.byref(grid))
        assert 0 <= n <= cap, (n, cap)
        g = grid.value
        assert 1 <= g <= max(1, min(max_ctas, n_tiles))
        if cta >= g:
            assert n == 0
            continue
        t = out[:n]
        # the cursor's claim about each tile ...
        tiles, rows, chunks, a, bb = t.T
        # ... against plain integer arithmetic
        assert np.array_equal(tiles, cta + g * np.arange(n))
        assert np.array_equal(rows, tiles // cpr) and np.array_equal(chunks, tiles % cpr)
        assert np.array_equal(a, rows // n_mid) and np.array_equal(bb, rows % n_mid)
        assert n == len(range(cta, n_tiles, g))
        if seen is not None:
            seen[tiles] += 1
        # when the grid is a multiple of the tiles per row a CTA keeps its chunk (register-cached columns)
        if g % cpr == 0 and n:
            assert np.all(chunks == chunks[0])
    if seen is not None:
        assert np.all(seen == 1)


def test_cursor_with_slab_offsets():
    hooke = b.Hooke3f64(1.0, 0.3, b.CartesianGrid3f64((37, 11, 130), (1., 1., 1.)))
    lib = _lib.load()
    kb = np.array([5, 3, 7], dtype=np.intc)
    ls = np.array([9, 4, 100], dtype=np.intc)
    out = np.zeros((64, 5), dtype=np.int64)
    grid = C.c_int()
    total = 0
    for cta in range(36):
        n = lib.bri17_debug_walk_tiles(hooke._plan, kb.ctypes.data_as(C.POINTER(C.c_int)),
                                       ls.ctypes.data_as(C.POINTER(C.c_int)), 512, 36, cta,
                                       out.ctypes.data_as(C.POINTER(C.c_int64)), 64, C.byref(grid))
        total += n
        assert np.all(out[:n, 3] < 9) and np.all(out[:n, 4] < 4)      # local (a, b); k = k_begin + (a, b, col)
    assert total == 9 * 4 * 1 and grid.value == 36
