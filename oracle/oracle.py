"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Loads ``oracle/liboracle.so`` (plain-C restatement, ``bri17_oracle.c``) and,
when present, ``oracle/_ref/libbri17_ref.so`` (the unmodified reference header
compiled with the harness loop, ``ref_driver.cpp``).  Both expose the same
entry points, so every check can be run against either.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product
package ``bri17_b200`` never does.

Layout conventions (reference: ``tests/test_bri17.cpp:66-67,81-83``): fields
are planar by component, ``buf[c, k0, k1(, k2)]`` complex128 in C order, i.e.
element ``(c, i)`` at ``buf.ravel()[i + c*comp_stride]``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_i32p = C.POINTER(C.c_int)
_f64p = C.POINTER(C.c_double)


def build(verbose: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def _ints(v):
    a = np.ascontiguousarray(v, dtype=np.intc)
    return a, a.ctypes.data_as(_i32p)


def _dbls(v):
    a = np.ascontiguousarray(v, dtype=np.float64)
    return a, a.ctypes.data_as(_f64p)


class _Impl:
    """One implementation (the C port, or the compiled reference header)."""

    def __init__(self, path: str, prefix: str, kind: str):
        self.path, self.prefix, self.kind = path, prefix, kind
        self.lib = C.CDLL(path)
        f = self._fn("modal_stiffness")
        f.argtypes = [C.c_int, _i32p, _f64p, C.c_double, C.c_double, _i32p, _f64p]
        f.restype = None
        f = self._fn("modal_strain_displacement")
        f.argtypes = [C.c_int, _i32p, _f64p, _i32p, _f64p]
        f.restype = None
        f = self._fn("get_cell_nodes")
        f.argtypes = [C.c_int, _i32p, C.c_int, _i32p]
        f.restype = None
        f = self._fn("apply_modal_stiffness")
        f.argtypes = [C.c_int, _i32p, _f64p, C.c_double, C.c_double, _i32p, _i32p,
                      C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        f.restype = None
        f = self._fn("apply_strain_displacement")
        f.argtypes = [C.c_int, _i32p, _f64p, _i32p, _i32p, C.c_int64, C.c_int64,
                      C.c_void_p, C.c_void_p]
        f.restype = None
        f = self._fn("max_threads")
        f.argtypes = []
        f.restype = C.c_int

    def _fn(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    # -- per-mode API (bri17.hpp:247-292, :212-236) -------------------------
    def modal_stiffness(self, shape, L, mu, nu, k) -> np.ndarray:
        dim = len(shape)
        _, sp = _ints(shape); _, lp = _dbls(L); _, kp = _ints(k)
        K = np.empty(dim * dim, dtype=np.complex128)
        self._fn("modal_stiffness")(dim, sp, lp, mu, nu, kp, K.ctypes.data_as(_f64p))
        return K.reshape(dim, dim)

    def modal_strain_displacement(self, shape, L, k) -> np.ndarray:
        dim = len(shape)
        _, sp = _ints(shape); _, lp = _dbls(L); _, kp = _ints(k)
        B = np.empty(dim, dtype=np.complex128)
        self._fn("modal_strain_displacement")(dim, sp, lp, kp, B.ctypes.data_as(_f64p))
        return B

    def get_cell_nodes(self, shape, cell) -> np.ndarray:
        dim = len(shape)
        _, sp = _ints(shape)
        nodes = np.empty(1 << dim, dtype=np.intc)
        self._fn("get_cell_nodes")(dim, sp, int(cell), nodes.ctypes.data_as(_i32p))
        return nodes

    # -- whole-grid loops (tests/test_bri17.cpp:58-92, :194-235) ------------
    def apply_modal_stiffness(self, shape, L, mu, nu, u_hat, k_begin=None,
                              out=None, nthreads=1) -> np.ndarray:
        """f^ = K^ u^ on the block ``k_begin + [0, u_hat.shape[1:])``."""
        dim = len(shape)
        u_hat = np.ascontiguousarray(u_hat, dtype=np.complex128)
        assert u_hat.shape[0] == dim and u_hat.ndim == dim + 1
        local = u_hat.shape[1:]
        if k_begin is None:
            k_begin = (0,) * dim
        f_hat = np.empty_like(u_hat) if out is None else out
        _, sp = _ints(shape); _, lp = _dbls(L)
        _, kb = _ints(k_begin); _, ls = _ints(local)
        stride = int(np.prod(local, dtype=np.int64))
        self._fn("apply_modal_stiffness")(dim, sp, lp, mu, nu, kb, ls, stride,
                                          u_hat.ctypes.data, f_hat.ctypes.data,
                                          int(nthreads))
        return f_hat

    def apply_strain_displacement(self, shape, L, u_hat, k_begin=None) -> np.ndarray:
        """eps^ (Mandel order) = sym(B^ (x) u^) on a block of frequencies."""
        dim = len(shape)
        nsym = dim * (dim + 1) // 2
        u_hat = np.ascontiguousarray(u_hat, dtype=np.complex128)
        local = u_hat.shape[1:]
        if k_begin is None:
            k_begin = (0,) * dim
        eps = np.empty((nsym,) + tuple(local), dtype=np.complex128)
        _, sp = _ints(shape); _, lp = _dbls(L)
        _, kb = _ints(k_begin); _, ls = _ints(local)
        stride = int(np.prod(local, dtype=np.int64))
        self._fn("apply_strain_displacement")(dim, sp, lp, kb, ls, stride, stride,
                                              u_hat.ctypes.data, eps.ctypes.data)
        return eps

    def max_threads(self) -> int:
        return int(self._fn("max_threads")())


def _load(path, prefix, kind):
    if not os.path.exists(path):
        return None
    return _Impl(path, prefix, kind)


def port(fast: bool = False) -> _Impl:
    """The plain-C restatement (``bri17_oracle.c``)."""
    path = os.path.join(_HERE, "liboracle_fast.so" if fast else "liboracle.so")
    if not os.path.exists(path):
        build()
    return _Impl(path, "oracle", "port")


def ref(fast: bool = False):
    """The compiled reference header (``oracle/_ref``), or None if absent."""
    name = "libbri17_ref_fast.so" if fast else "libbri17_ref.so"
    return _load(os.path.join(_HERE, "_ref", name), "ref", "reference")


def best(fast: bool = False) -> _Impl:
    """``oracle/_ref`` when it exists (real reference arithmetic), else the port."""
    return ref(fast) or port(fast)


def apply_eigenstress(shape, L, mu, nu, tau_hat, k_begin=None):
    """Per-mode eigenstress solve over a block (python/demo.py:33-40 loop, C port
    only -- the compiled reference cannot instantiate the Eigen-based method):
    returns (eta_hat[nsym, *local], u_hat[dim, *local]).  PARITY UNPINNED."""
    lib = port().lib
    dim = len(shape)
    nsym = dim * (dim + 1) // 2
    tau_hat = np.ascontiguousarray(tau_hat, dtype=np.complex128)
    local = tau_hat.shape[1:]
    assert tau_hat.shape[0] == nsym
    if k_begin is None:
        k_begin = (0,) * dim
    eta = np.empty_like(tau_hat)
    u = np.empty((dim,) + tuple(local), dtype=np.complex128)
    _, sp = _ints(shape); _, lp = _dbls(L); _, kb = _ints(k_begin); _, ls = _ints(local)
    lib.oracle_apply_eigenstress.argtypes = [C.c_int, _i32p, _f64p, C.c_double, C.c_double, _i32p, _i32p,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_apply_eigenstress.restype = None
    lib.oracle_apply_eigenstress(dim, sp, lp, mu, nu, kb, ls, tau_hat.ctypes.data, eta.ctypes.data,
                                 u.ctypes.data)
    return eta, u


def freq_index_map(k_begin, local_shape) -> np.ndarray:
    """Multi-index of every linear element of a row-major block
    (tests/test_bri17.cpp:62-64,71 / :76-79,88), int32 ``[prod(local), dim]``."""
    lib = port().lib
    dim = len(local_shape)
    n = int(np.prod(local_shape, dtype=np.int64))
    out = np.empty((n, dim), dtype=np.int32)
    _, kb = _ints(k_begin); _, ls = _ints(local_shape)
    lib.oracle_freq_index_map.argtypes = [C.c_int, _i32p, _i32p, C.c_void_p]
    lib.oracle_freq_index_map.restype = None
    lib.oracle_freq_index_map(dim, kb, ls, out.ctypes.data)
    return out


def synthetic_u_hat(dim: int, local_shape, seed) -> np.ndarray:
    """Synthetic modal displacement: complex standard
    normal, planar by component, C order.  ``seed`` may be an int or a
    sequence (e.g. ``[seed, slab]``)."""
    rng = np.random.default_rng(seed)
    shp = (dim,) + tuple(local_shape)
    return rng.standard_normal(shp) + 1j * rng.standard_normal(shp)
