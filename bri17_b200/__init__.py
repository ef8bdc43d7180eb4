"""bri17_b200 -- B200-native modal operator path of bri17, Python host side.

Mirrors the reference's Python surface (``python/pybri17.cpp:98-107``):
``CartesianGrid2f64`` / ``CartesianGrid3f64`` / ``Hooke2f64`` / ``Hooke3f64``
with the same attribute and method names (``modal_stiffness_matrix``,
``modal_strain_displacement``), and adds :class:`ModalOperator`, the
whole-grid operator that the reference only has as a loop nest in its test
harness (``tests/test_bri17.cpp:56-107``).  Everything numerical goes through
the C ABI of ``libbri17_b200.so`` (``include/bri17_b200.h``); PyTorch is used
only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Bri17Error, check

__all__ = ["CartesianGrid2f64", "CartesianGrid3f64", "Hooke2f64", "Hooke3f64",
           "ModalOperator", "Bri17Error", "HOST_ONLY"]

__version__ = "0.1"          # metadata/version.txt:1
HOST_ONLY = -1               # device ordinal of a plan without a GPU (per-mode API only)


def _ints(v, n=None):
    a = np.ascontiguousarray(v, dtype=np.intc)
    if n is not None and a.shape != (n,):
        raise ValueError(f"expected {n} integers, got shape {a.shape}")
    return a


def _p_i32(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def _p_f64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---------------------------------------------------------------------------
# CartesianGrid -- include/bri17/bri17.hpp:34-159, python/pybri17.cpp:33-61
# ---------------------------------------------------------------------------
class _CartesianGrid:
    dim = 0
    dtype = np.dtype(np.float64)

    def __init__(self, shape, L):
        if len(shape) != self.dim or len(L) != self.dim:
            raise TypeError(f"expected {self.dim} entries in shape and L")
        self.shape = tuple(int(n) for n in shape)
        self.L = tuple(float(x) for x in L)
        # bri17.hpp:48,57-58: `int const size` -- the product is kept exact here
        self.size = int(np.prod(self.shape, dtype=np.int64))

    num_nodes_per_cell = property(lambda self: 1 << self.dim)   # bri17.hpp:39

    def __repr__(self):  # bri17.hpp:61-69 ("d" is typeid(double).name() with GCC)
        return ("CartesianGrid<d," + str(self.dim) + ">{shape={" +
                "".join(f"{n}," for n in self.shape) + "},L={" +
                "".join(f"{x:g}," for x in self.L) + "}}")

    def get_node_at(self, *ijk):  # bri17.hpp:78-93, row-major
        if len(ijk) != self.dim:
            raise TypeError(f"this method expects a {self.dim}D multi-index")
        node = 0
        for n, i in zip(self.shape, ijk):
            node = node * n + int(i)
        return node

    def get_cell_nodes(self, cell):  # bri17.hpp:127-158
        idx = np.unravel_index(int(cell), self.shape)
        nodes = []
        for local in range(1 << self.dim):
            ijk = [(int(idx[d]) + ((local >> (self.dim - 1 - d)) & 1)) % self.shape[d]
                   for d in range(self.dim)]
            nodes.append(self.get_node_at(*ijk))
        return nodes


class CartesianGrid2f64(_CartesianGrid):
    dim = 2


class CartesianGrid3f64(_CartesianGrid):
    dim = 3


def _check_contiguous(a, dtype, what):
    """python/pybri17.cpp:9-15: one-dimensional, contiguous, else invalid_argument."""
    if not isinstance(a, np.ndarray) or a.ndim != 1 or a.strides[0] != a.itemsize \
            or a.dtype != dtype:
        raise ValueError(f"expected one-dimensional, contiguous array ({what})")


def _freq(k, shape, wrap):
    """``k`` as the reference's binding takes it: pybind11's ``array_t<int>`` converts any integer
    sequence (python/pybri17.cpp:76-95) and the C++ methods never bound-check ``k``
    (bri17.hpp:212,247).  K^ is N-periodic in k (cos/sin of 2 pi k/N), so ``wrap=True`` maps k
    into [0, N); B^ uses half angles and is NOT N-periodic (k -> k+N flips signs), so it keeps
    the table range and rejects indices outside [0, N)."""
    a = np.asarray(k)
    if a.ndim != 1 or not np.issubdtype(a.dtype, np.integer) and a.dtype != np.bool_:
        raise ValueError("expected one-dimensional, contiguous array (k)")
    if a.shape[0] != len(shape):
        raise ValueError(f"expected {len(shape)} frequency indices")
    a = a.astype(np.int64)
    if wrap:
        a = np.mod(a, np.asarray(shape, dtype=np.int64))
    return np.ascontiguousarray(a, dtype=np.intc)


# ---------------------------------------------------------------------------
# Hooke -- include/bri17/bri17.hpp:176-356, python/pybri17.cpp:63-96
# ---------------------------------------------------------------------------
class _Hooke:
    dim = 0

    def __init__(self, mu, nu, grid, device=HOST_ONLY):
        if grid.dim != self.dim:
            raise TypeError(f"expected a {self.dim}D grid")
        self.mu, self.nu, self.grid = float(mu), float(nu), grid
        self._lib = _lib.load()
        self._plan = C.c_void_p()
        shape = _ints(grid.shape)
        L = np.ascontiguousarray(grid.L, dtype=np.float64)
        check(self._lib.bri17_plan_create(C.byref(self._plan), self.dim, _p_i32(shape),
                                          _p_f64(L), self.mu, self.nu, int(device)))
        self.device = int(device)

    def __del__(self):
        plan, self._plan = getattr(self, "_plan", None), None
        if plan:
            self._lib.bri17_plan_destroy(plan)

    def __repr__(self):  # bri17.hpp:197-202
        return f"Hooke<d,{self.dim}>{{mu={self.mu:g},nu={self.nu:g},grid={self.grid!r}\n"

    def modal_stiffness_matrix(self, k, K):
        """Hooke::modal_stiffness (bri17.hpp:247-292) under its Python name
        (python/pybri17.cpp:82): fills ``K`` (dim*dim complex128, row-major)."""
        k = _freq(k, self.grid.shape, wrap=True)
        _check_contiguous(K, np.dtype(np.complex128), "K")
        check(self._lib.bri17_modal_stiffness_mode_f64(self._plan, _p_i32(k), _p_f64(K)))

    def modal_strain_displacement(self, k, B):
        """Hooke::modal_strain_displacement (bri17.hpp:212-236)."""
        k = _freq(k, self.grid.shape, wrap=False)
        _check_contiguous(B, np.dtype(np.complex128), "B")
        check(self._lib.bri17_modal_strain_displacement_mode_f64(self._plan, _p_i32(k), _p_f64(B)))

    def modal_eigenstress_to_opposite_strain(self, k, tau, eta):
        """Hooke::modal_eigenstress_to_opposite_strain (bri17.hpp:308-355,
        python/pybri17.cpp:88-95): ``eta`` <- -strain induced by eigenstress ``tau``
        (Mandel notation, 3 or 6 complex128)."""
        k = _freq(k, self.grid.shape, wrap=False)
        _check_contiguous(tau, np.dtype(np.complex128), "tau")
        _check_contiguous(eta, np.dtype(np.complex128), "eta")
        check(self._lib.bri17_modal_eigenstress_to_opposite_strain_mode_f64(
            self._plan, _p_i32(k), _p_f64(tau), _p_f64(eta)))

    def tables(self, axis):
        """Host copy of the per-axis tables: dict phi/chi/psi/c/s."""
        n = self.grid.shape[axis]
        out = {name: np.empty(n) for name in ("phi", "chi", "psi", "c", "s")}
        check(self._lib.bri17_plan_get_tables(self._plan, axis, *[_p_f64(out[k]) for k in
                                                                  ("phi", "chi", "psi", "c", "s")]))
        return out

    def set_option(self, key, value):
        check(self._lib.bri17_plan_set_option(self._plan, key.encode(), int(value)))

    def info(self, key):
        v = C.c_int64()
        check(self._lib.bri17_plan_get_info(self._plan, key.encode(), C.byref(v)))
        return v.value


class Hooke2f64(_Hooke):
    dim = 2


class Hooke3f64(_Hooke):
    dim = 3


def _dev_ptr(t, device=None):
    """Device pointer of a torch tensor / cuda-array-interface object / int; ``device``: the
    ordinal the tensor must live on (the plan's)."""
    if isinstance(t, int):
        return t
    if hasattr(t, "data_ptr"):
        if not t.is_cuda:
            raise ValueError("expected a CUDA tensor (use the *_host entry point for host memory)")
        if device is not None and t.device.index != device:
            raise ValueError(f"tensor lives on {t.device}, the plan on cuda:{device}")
        if not t.is_contiguous():
            raise ValueError("expected a contiguous tensor")
        return t.data_ptr()
    if hasattr(t, "__cuda_array_interface__"):
        return t.__cuda_array_interface__["data"][0]
    raise TypeError(f"cannot take a device pointer from {type(t)}")


def _stream_ptr(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


# ---------------------------------------------------------------------------
# ModalOperator -- the whole-grid operator (tests/test_bri17.cpp:56-107)
# ---------------------------------------------------------------------------
class ModalOperator:
    """Block-diagonal modal operator on one CUDA device.

    ``u_hat`` / ``f_hat`` are ``complex128`` arrays of shape ``(dim, *local)``,
    planar by component (tests/test_bri17.cpp:66-67, :81-83); ``local`` is the
    whole grid or, with ``k_begin``, a block of it (a ``k0`` slab per GPU).
    """

    def __init__(self, shape, L, mu, nu, device=0):
        dim = len(shape)
        grid = (CartesianGrid2f64 if dim == 2 else CartesianGrid3f64)(shape, L)
        self.hooke = (Hooke2f64 if dim == 2 else Hooke3f64)(mu, nu, grid, device=device)
        self.dim, self.grid, self.device = dim, grid, int(device)
        self._lib, self._plan = self.hooke._lib, self.hooke._plan

    # -- helpers -------------------------------------------------------------
    def _block(self, arr_shape, ncomp, k_begin):
        if len(arr_shape) != self.dim + 1 or arr_shape[0] != ncomp:
            raise ValueError(f"expected shape ({ncomp}, *local{self.dim}d), got {tuple(arr_shape)}")
        local = _ints(arr_shape[1:], self.dim)
        kb = _ints(k_begin if k_begin is not None else (0,) * self.dim, self.dim)
        return kb, local

    def set_option(self, key, value):
        self.hooke.set_option(key, value)

    def info(self, key):
        return self.hooke.info(key)

    # -- K1/K2 ---------------------------------------------------------------
    def apply_modal_stiffness(self, u_hat, out=None, k_begin=None, out_scale=1.0, stream=None):
        """f^[k] = K^[k] u^[k] for every frequency of the block (device tensors)."""
        import torch
        kb, local = self._block(u_hat.shape, self.dim, k_begin)
        if u_hat.dtype != torch.complex128:
            raise ValueError("u_hat must be complex128")
        if out is None:
            out = torch.empty_like(u_hat)
        elif out.shape != u_hat.shape or out.dtype != u_hat.dtype:
            raise ValueError("out must match u_hat")
        check(self._lib.bri17_modal_stiffness_apply_f64(
            self._plan, _dev_ptr(u_hat, self.device), _dev_ptr(out, self.device), _p_i32(kb), _p_i32(local), 0,
            float(out_scale), _stream_ptr(stream)))
        return out

    def apply_modal_stiffness_dot(self, u_hat, out=None, k_begin=None, out_scale=1.0, hermitian_n=0,
                                  stream=None):
        """``(f^, sum_k w_k Re(u^_k^H f^_k))``: the apply plus the Parseval sum of the block
        (bri17_modal_stiffness_apply_dot_f64; ``hermitian_n`` = full length of the fastest axis
        when the block is the half spectrum of a real field)."""
        import torch
        kb, local = self._block(u_hat.shape, self.dim, k_begin)
        if u_hat.dtype != torch.complex128:
            raise ValueError("u_hat must be complex128")
        if out is None:
            out = torch.empty_like(u_hat)
        elif out.shape != u_hat.shape or out.dtype != u_hat.dtype:
            raise ValueError("out must match u_hat")
        scratch = torch.empty(1184 + 1, dtype=torch.float64, device=u_hat.device)
        check(self._lib.bri17_modal_stiffness_apply_dot_f64(
            self._plan, _dev_ptr(u_hat, self.device), _dev_ptr(out, self.device), _p_i32(kb), _p_i32(local), 0,
            float(out_scale), int(hermitian_n), scratch.data_ptr(), scratch.data_ptr() + 8, 1184,
            _stream_ptr(stream)))
        return out, float(scratch[0].item())

    def apply_modal_stiffness_host(self, u_hat, out=None, k_begin=None, out_scale=1.0):
        """Same on host memory (numpy arrays or pinned CPU torch tensors);
        copies are pipelined with the kernel inside the library."""
        kb, local = self._block(u_hat.shape, self.dim, k_begin)
        if hasattr(u_hat, "data_ptr"):       # torch CPU tensor (e.g. pinned)
            import torch
            if u_hat.is_cuda or u_hat.dtype != torch.complex128 or not u_hat.is_contiguous():
                raise ValueError("expected a contiguous complex128 CPU tensor")
            if out is None:
                out = torch.empty_like(u_hat)
            src, dst = u_hat.data_ptr(), out.data_ptr()
        else:
            u_hat = np.ascontiguousarray(u_hat, dtype=np.complex128)
            if out is None:
                out = np.empty_like(u_hat)
            if out.dtype != np.complex128 or not out.flags.c_contiguous or out.shape != u_hat.shape:
                raise ValueError("out must be a C-contiguous complex128 array like u_hat")
            src, dst = u_hat.ctypes.data, out.ctypes.data
        check(self._lib.bri17_modal_stiffness_apply_host_f64(
            self._plan, src, dst, _p_i32(kb), _p_i32(local), 0, float(out_scale)))
        return out

    # -- K6 ------------------------------------------------------------------
    def freq_index_map(self, local_shape=None, k_begin=None, stream=None):
        """int32 ``[prod(local), dim]``: the multi-index the kernels derive."""
        import torch
        local = _ints(local_shape if local_shape is not None else self.grid.shape, self.dim)
        kb = _ints(k_begin if k_begin is not None else (0,) * self.dim, self.dim)
        n = int(np.prod(local, dtype=np.int64))
        out = torch.empty((n, self.dim), dtype=torch.int32, device=f"cuda:{self.device}")
        check(self._lib.bri17_freq_index_map(self._plan, _dev_ptr(out), _p_i32(kb),
                                             _p_i32(local), _stream_ptr(stream)))
        return out

    # -- fields ----------------------------------------------------------------
    def modal_stiffness_field(self, local_shape=None, k_begin=None, stream=None):
        """K^[k] for every mode: complex128 ``[*local, dim, dim]``."""
        import torch
        local = _ints(local_shape if local_shape is not None else self.grid.shape, self.dim)
        kb = _ints(k_begin if k_begin is not None else (0,) * self.dim, self.dim)
        out = torch.empty(tuple(local) + (self.dim, self.dim), dtype=torch.complex128,
                          device=f"cuda:{self.device}")
        check(self._lib.bri17_modal_stiffness_field_f64(self._plan, _dev_ptr(out), _p_i32(kb),
                                                        _p_i32(local), _stream_ptr(stream)))
        return out

    def modal_strain_displacement_field(self, local_shape=None, k_begin=None, stream=None):
        """B^[k] for every mode: complex128 ``[*local, dim]``."""
        import torch
        local = _ints(local_shape if local_shape is not None else self.grid.shape, self.dim)
        kb = _ints(k_begin if k_begin is not None else (0,) * self.dim, self.dim)
        out = torch.empty(tuple(local) + (self.dim,), dtype=torch.complex128,
                          device=f"cuda:{self.device}")
        check(self._lib.bri17_modal_strain_displacement_field_f64(
            self._plan, _dev_ptr(out), _p_i32(kb), _p_i32(local), _stream_ptr(stream)))
        return out

    # -- per-mode direct solves (bri17.hpp:308-355 batched) ---------------------
    def _solve(self, fn_name, x, nin, nout, k_begin, mode_major, stream):
        import torch
        if x.dtype != torch.complex128:
            raise ValueError("expected complex128")
        if mode_major:          # (*local, ncomp): the layout of python/demo.py:21
            if x.dim() != self.dim + 1 or x.shape[-1] != nin:
                raise ValueError(f"expected shape (*local, {nin})")
            local = _ints(x.shape[:-1], self.dim)
            out = torch.empty(tuple(x.shape[:-1]) + (nout,), dtype=x.dtype, device=x.device)
            strides = (1, nin, 1, nout)
        else:
            kb_, local = self._block(x.shape, nin, k_begin)
            out = torch.empty((nout,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            strides = (0, 1, 0, 1)
        kb = _ints(k_begin if k_begin is not None else (0,) * self.dim, self.dim)
        fn = getattr(self._lib, fn_name)
        if fn_name in ("bri17_eigenstress_to_displacement_f64", "bri17_eigenstress_to_force_f64"):
            rc = fn(self._plan, _dev_ptr(x, self.device), _dev_ptr(out, self.device), _p_i32(kb), _p_i32(local), *strides,
                    _stream_ptr(stream))
        else:
            rc = fn(self._plan, _dev_ptr(x, self.device), _dev_ptr(out, self.device), _p_i32(kb), _p_i32(local), strides[0],
                    strides[1], _stream_ptr(stream))
        check(rc)
        return out

    def solve_modal_stiffness(self, f_hat, k_begin=None, mode_major=False, stream=None):
        """u^ = K^-1 f^ per mode, u^(0) = 0: the exact inverse of apply_modal_stiffness."""
        return self._solve("bri17_modal_stiffness_solve_f64", f_hat, self.dim, self.dim, k_begin,
                           mode_major, stream)

    def eigenstress_to_displacement(self, tau_hat, k_begin=None, mode_major=False, stream=None):
        """u^ = K^-1 (tau^ . conj(B^)) per mode (bri17.hpp:340-341)."""
        nsym = self.dim * (self.dim + 1) // 2
        return self._solve("bri17_eigenstress_to_displacement_f64", tau_hat, nsym, self.dim, k_begin,
                           mode_major, stream)

    def eigenstress_to_force(self, tau_hat, k_begin=None, mode_major=False, stream=None):
        """f^ = tau^ . conj(B^) per mode (bri17.hpp:340 without the solve): the modal force of the
        periodic inclusion problem (python/demo.py:11-23), right-hand side of the CG solve."""
        nsym = self.dim * (self.dim + 1) // 2
        return self._solve("bri17_eigenstress_to_force_f64", tau_hat, nsym, self.dim, k_begin,
                           mode_major, stream)

    def eigenstress_to_opposite_strain(self, tau_hat, k_begin=None, mode_major=False, stream=None):
        """eta^ = -eps^ induced by the eigenstress tau^ (bri17.hpp:308-355), Mandel notation."""
        nsym = self.dim * (self.dim + 1) // 2
        return self._solve("bri17_eigenstress_to_opposite_strain_f64", tau_hat, nsym, nsym, k_begin,
                           mode_major, stream)

    # -- K3 --------------------------------------------------------------------
    def apply_strain_displacement(self, u_hat, out=None, k_begin=None, out_scale=1.0, stream=None):
        """eps^ = sym(B^ (x) u^) in Mandel order: ``(nsym, *local)``."""
        import torch
        kb, local = self._block(u_hat.shape, self.dim, k_begin)
        nsym = self.dim * (self.dim + 1) // 2
        if u_hat.dtype != torch.complex128:
            raise ValueError("u_hat must be complex128")
        want = (nsym,) + tuple(u_hat.shape[1:])
        if out is None:
            out = torch.empty(want, dtype=u_hat.dtype, device=u_hat.device)
        elif tuple(out.shape) != want or out.dtype != u_hat.dtype:
            raise ValueError(f"out must be complex128 {want}")
        check(self._lib.bri17_strain_displacement_apply_f64(
            self._plan, _dev_ptr(u_hat, self.device), _dev_ptr(out, self.device), _p_i32(kb), _p_i32(local),
            0, 0, float(out_scale), _stream_ptr(stream)))
        return out
