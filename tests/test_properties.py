"""Property tests (hypothesis) of the host-side logic and of the oracle: any
block of frequencies computed on its own equals the same slice of the
whole-grid result (the property every multi-GPU decomposition relies on), and
slab partitions tile the axis."""
import numpy as np
from hypothesis import given, settings, strategies as st

from bri17_b200 import slab
from oracle import oracle

shapes3 = st.tuples(st.integers(1, 7), st.integers(1, 6), st.integers(1, 9))
shapes2 = st.tuples(st.integers(1, 9), st.integers(1, 12))


def _sub_block(draw_ints, shape):
    kb, loc = [], []
    for n, (a, b) in zip(shape, draw_ints):
        lo = a % n
        kb.append(lo)
        loc.append(1 + b % (n - lo))
    return tuple(kb), tuple(loc)


@settings(max_examples=40, deadline=None)
@given(shape=st.one_of(shapes2, shapes3),
       cuts=st.lists(st.tuples(st.integers(0, 100), st.integers(0, 100)), min_size=3, max_size=3),
       seed=st.integers(0, 2**16))
def test_any_block_equals_the_slice_of_the_full_apply(shape, cuts, seed):
    dim = len(shape)
    L = tuple(0.7 * n + 0.1 * d for d, n in enumerate(shape))
    kb, loc = _sub_block(cuts[:dim], shape)
    u = oracle.synthetic_u_hat(dim, shape, seed)
    p = oracle.port()
    full = p.apply_modal_stiffness(shape, L, 5.6, 0.3, u)
    sl = (slice(None),) + tuple(slice(a, a + n) for a, n in zip(kb, loc))
    part = p.apply_modal_stiffness(shape, L, 5.6, 0.3, np.ascontiguousarray(u[sl]), k_begin=kb)
    assert np.array_equal(part, full[sl])
    k = oracle.freq_index_map(kb, loc)
    grid = np.stack(np.meshgrid(*[np.arange(a, a + n) for a, n in zip(kb, loc)], indexing="ij"), axis=-1)
    assert np.array_equal(k, grid.reshape(-1, dim).astype(np.int32))


@settings(max_examples=200, deadline=None)
@given(n0=st.integers(1, 5000), world=st.integers(1, 16))
def test_slab_ranges_tile_the_axis(n0, world):
    ranges = [slab.slab_range(n0, g, world) for g in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n0
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 0


@settings(max_examples=30, deadline=None)
@given(shape=shapes3, seed=st.integers(0, 2**16))
def test_stiffness_is_symmetric_positive_and_zero_at_origin(shape, seed):
    rng = np.random.default_rng(seed)
    L = tuple(float(x) for x in rng.uniform(0.5, 3.0, size=3))
    mu, nu = float(rng.uniform(0.1, 10)), float(rng.uniform(-0.5, 0.45))
    p = oracle.port()
    for k in np.ndindex(*shape):
        K = p.modal_stiffness(shape, L, mu, nu, k)
        assert np.all(K.imag == 0) and np.array_equal(K, K.T)
        if not any(k):
            assert np.all(K == 0)
        else:
            assert np.all(np.linalg.eigvalsh(K.real) > -1e-12 * np.abs(K).max())
