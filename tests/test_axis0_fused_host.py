"""CPU replay of the fused axis-0 kernel (bri17_b200/csrc/axis0_fused.cuh).

The kernel's phase code is __host__ __device__; ``bri17_debug_axis0_fused_host``
runs it thread by thread on host memory, so the index arithmetic of
FFT(axis 0) -> K^ -> inverse FFT(axis 0) (digit-reversed spectrum, swizzled
shared-memory layout, twiddles, column -> (k1, k2) mapping, Hermitian weights of
the Parseval sum) is checked here against numpy's FFT and the CPU oracle without
a GPU.  The GPU tests then run the same code on the device."""
import ctypes as C

import numpy as np
import pytest

import bri17_b200 as b
from bri17_b200 import _lib

MU, NU = 5.6, 0.3


def _tables(shape, L):
    dim = len(shape)
    grid = (b.CartesianGrid2f64 if dim == 2 else b.CartesianGrid3f64)(shape, L)
    hooke = (b.Hooke2f64 if dim == 2 else b.Hooke3f64)(MU, NU, grid)
    out = []
    for d in range(dim):
        t = hooke.tables(d)
        out.append(np.ascontiguousarray(np.concatenate([t["phi"], t["chi"], t["psi"]])))
    return out


def _run(shape, L, k1_begin, n1_loc, S2e, X, out_scale=1.0, herm=0, want_dot=False, k1_major=0):
    """X: (dim, N0, n1_loc[, S2e]) complex128 block of a grid `shape` whose fastest spectral
    extent is S2e (3-D) -- in place through the CPU replay."""
    lib = _lib.load_rs()
    dim = len(shape)
    tabs = _tables(shape, L)
    p = [t.ctypes.data_as(C.POINTER(C.c_double)) for t in tabs] + [None] * (3 - dim)
    S = n1_loc * (S2e if dim == 3 else 1)
    dot = C.c_double(0.0)
    rc = lib.bri17_debug_axis0_fused_host(dim, shape[0], S, S2e if dim == 3 else 1, k1_begin, shape[1],
                                          shape[2] if dim == 3 else 1, p[0], p[1], p[2], MU, NU,
                                          out_scale, herm, k1_major, X.ctypes.data,
                                          C.byref(dot) if want_dot else None)
    assert rc == 0, _lib.load().bri17_last_error()
    return dot.value


def _reference(oracle_mod, shape, L, k1_begin, X, out_scale):
    """numpy FFT along axis 0 + CPU oracle K^ on the block + unnormalised inverse."""
    dim = len(shape)
    Xh = np.fft.fft(X, axis=1)
    kb = (0, k1_begin) + (0,) * (dim - 2)
    Fh = oracle_mod.best().apply_modal_stiffness(shape, L, MU, NU, Xh, k_begin=kb) * out_scale
    return Xh, Fh, np.fft.ifft(Fh, axis=1) * shape[0]


@pytest.mark.parametrize("N0", [16, 32, 64, 128, 256, 512, 1024])
def test_fused_axis0_replay_3d(oracle_mod, N0):
    """Every supported length (radix 2/4/8/16 first stage; two and three stages), a k1 slab with an
    offset and a column count that is not a multiple of the tile width."""
    shape = (N0, 12, 10)
    L = (1.1 * N0, 13.2, 13.0)
    k1_begin, n1_loc, S2e = 5, 3, 10           # 30 columns: partial last tile for every W
    rng = np.random.default_rng(N0)
    X = rng.standard_normal((3, N0, n1_loc, S2e)) + 1j * rng.standard_normal((3, N0, n1_loc, S2e))
    _, _, ref = _reference(oracle_mod, shape, L, k1_begin, X, 0.37)
    got = X.copy()
    _run(shape, L, k1_begin, n1_loc, S2e, got, out_scale=0.37)
    assert np.abs(got - ref).max() <= 2e-14 * np.abs(ref).max()


@pytest.mark.parametrize("N0", [16, 64, 512])
def test_fused_axis0_replay_2d(oracle_mod, N0):
    shape = (N0, 21)
    L = (1.1 * N0, 25.2)
    k1_begin, n1_loc = 4, 17
    rng = np.random.default_rng(100 + N0)
    X = rng.standard_normal((2, N0, n1_loc)) + 1j * rng.standard_normal((2, N0, n1_loc))
    _, _, ref = _reference(oracle_mod, shape, L, k1_begin, X, 1.0)
    got = X.copy()
    _run(shape, L, k1_begin, n1_loc, 1, got)
    assert np.abs(got - ref).max() <= 2e-14 * np.abs(ref).max()


def test_fused_axis0_many_tiles_w4(oracle_mod):
    """W = 4 (N0 = 512): several tiles, columns straddling k1 rows of a 513-wide-like ragged
    half spectrum (S2e = 7 here), exercising the swizzled layout on full and partial tiles."""
    shape = (512, 9, 12)
    L = (563.2, 10.8, 15.6)
    S2e = 12 // 2 + 1
    k1_begin, n1_loc = 2, 5                    # 35 columns = 8 full tiles + 3
    rng = np.random.default_rng(7)
    X = rng.standard_normal((3, 512, n1_loc, S2e)) + 1j * rng.standard_normal((3, 512, n1_loc, S2e))
    Xh, Fh, ref = _reference(oracle_mod, shape, L, k1_begin, X, 1.0 / 3.0)
    got = X.copy()
    dot = _run(shape, L, k1_begin, n1_loc, S2e, got, out_scale=1.0 / 3.0, herm=12, want_dot=True)
    assert np.abs(got - ref).max() <= 2e-14 * np.abs(ref).max()
    # Parseval sum with Hermitian pair weights: k2 = 0 and k2 = N2/2 once, the others twice
    w = np.full(S2e, 2.0)
    w[0] = 1.0
    w[-1] = 1.0
    expect = float(np.sum(w * np.real(np.conj(Xh) * Fh)))
    assert abs(dot - expect) <= 1e-12 * abs(expect)


def test_fused_axis0_dot_full_spectrum(oracle_mod):
    """hermitian_n = 0: sum_k Re(u^_k^H f^_k) = N0 * <x, y> along axis 0 (Parseval)."""
    shape = (64, 6, 5)
    L = (70.4, 7.2, 6.5)
    rng = np.random.default_rng(11)
    X = rng.standard_normal((3, 64, 6, 5)) + 1j * rng.standard_normal((3, 64, 6, 5))
    Xh, Fh, ref = _reference(oracle_mod, shape, L, 0, X, 2.0)
    got = X.copy()
    dot = _run(shape, L, 0, 6, 5, got, out_scale=2.0, want_dot=True)
    expect = float(np.sum(np.real(np.conj(Xh) * Fh)))
    assert abs(dot - expect) <= 1e-12 * abs(expect)
    assert abs(dot - float(np.sum(np.real(np.conj(X) * got)))) <= 1e-12 * abs(expect)


@pytest.mark.parametrize("N0", [32, 512, 1024])
def test_fused_axis0_k1_major_layout(oracle_mod, N0):
    """The k1-major global layout [c][k1][n0][k2] (rows of a column S2e elements apart) gives the
    same result as the natural layout [c][n0][k1][k2]; tiles straddle k1 blocks (S2e = 7)."""
    shape = (N0, 9, 12)
    L = (1.1 * N0, 10.8, 15.6)
    S2e, k1_begin, n1_loc = 7, 3, 5
    rng = np.random.default_rng(N0 + 1)
    X = rng.standard_normal((3, N0, n1_loc, S2e)) + 1j * rng.standard_normal((3, N0, n1_loc, S2e))
    _, _, ref = _reference(oracle_mod, shape, L, k1_begin, X, 0.5)
    Xt = np.ascontiguousarray(X.transpose(0, 2, 1, 3))                 # [c][k1][n0][k2]
    dot = _run(shape, L, k1_begin, n1_loc, S2e, Xt, out_scale=0.5, herm=12, want_dot=True, k1_major=1)
    got = Xt.transpose(0, 2, 1, 3)
    assert np.abs(got - ref).max() <= 2e-14 * np.abs(ref).max()
    nat = X.copy()
    dot_nat = _run(shape, L, k1_begin, n1_loc, S2e, nat, out_scale=0.5, herm=12, want_dot=True)
    assert np.array_equal(nat, got) and dot == dot_nat


def test_fused_axis0_unsupported_length():
    lib = _lib.load_rs()
    X = np.zeros((3, 48, 4), dtype=np.complex128)
    t = np.zeros(3 * 48)
    tp = t.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib.bri17_debug_axis0_fused_host(3, 48, 4, 2, 0, 48, 48, tp, tp, tp, 1.0, 0.3, 1.0, 0, 0,
                                          X.ctypes.data, None)
    assert rc == _lib.ERR_UNSUPPORTED
