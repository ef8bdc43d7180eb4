"""Real-space operator F = (|h|/|N|) iDFT(K^ DFT(u)) and CG, Python host side.

Mirrors ``StiffnessMatrixFactory::compute_Ku`` of the reference harness
(``tests/test_bri17.cpp:56-107``) over the C ABI of ``libbri17_b200_rs.so``
(``include/bri17_b200_realspace.h``): cuFFT locally, slab decomposition with
one exchange per direction across the ranks of a ``torch.distributed`` group.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from . import _dev_ptr, _stream_ptr

PHASES = ("fft_local_fwd", "exchange_fwd", "fft_axis0_fwd", "modal", "fft_axis0_inv",
          "exchange_bwd", "fft_local_inv", "total")
EXCHANGE_NCCL, EXCHANGE_PEER_STORE = 0, 1


class RealSpaceOperator:
    """One rank of the (possibly distributed) real-space operator.

    Real-space fields: ``complex128 (dim, n0_count, N1[, N2])`` (this rank's slab
    of axis 0; imaginary part zero for real data, as in the reference).
    Fourier-space fields: ``(dim, N0, k1_count[, N2])`` (slab of axis 1).
    """

    def __init__(self, shape, L, mu, nu, device=0, rank=0, world=1, unique_id: bytes | None = None,
                 exchange_mode=EXCHANGE_PEER_STORE):
        self._lib = _lib.load_rs()
        self.shape, self.L = tuple(int(n) for n in shape), tuple(float(x) for x in L)
        self.dim, self.device, self.rank, self.world = len(shape), int(device), int(rank), int(world)
        self.mu, self.nu = float(mu), float(nu)
        self._plan = C.c_void_p()
        sh = np.ascontiguousarray(self.shape, dtype=np.intc)
        ll = np.ascontiguousarray(self.L, dtype=np.float64)
        uid = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        check(self._lib.bri17_rs_plan_create(
            C.byref(self._plan), self.dim, sh.ctypes.data_as(C.POINTER(C.c_int)),
            ll.ctypes.data_as(C.POINTER(C.c_double)), self.mu, self.nu, self.device, self.rank,
            self.world, uid, int(exchange_mode)))
        v = [C.c_int() for _ in range(4)]
        check(self._lib.bri17_rs_plan_local(self._plan, *[C.byref(x) for x in v]))
        self.n0_begin, self.n0_count, self.k1_begin, self.k1_count = (x.value for x in v)
        self.real_shape = (self.dim, self.n0_count) + self.shape[1:]
        self.fourier_shape = (self.dim, self.shape[0], self.k1_count) + self.shape[2:]
        self.exchange_bytes = int(self._lib.bri17_rs_plan_exchange_bytes(self._plan, 0))
        self.exchange_bytes_real = int(self._lib.bri17_rs_plan_exchange_bytes(self._plan, 1))

    @classmethod
    def from_process_group(cls, shape, L, mu, nu, device, exchange_mode=EXCHANGE_PEER_STORE):
        """Collective constructor: rank 0 draws the NCCL unique id and broadcasts
        it through the default torch.distributed group."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return cls(shape, L, mu, nu, device=device, exchange_mode=exchange_mode)
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [None]
        if rank == 0:
            buf = C.create_string_buffer(128)
            check(_lib.load_rs().bri17_rs_unique_id(buf))
            box[0] = buf.raw
        dist.broadcast_object_list(box, src=0)
        return cls(shape, L, mu, nu, device=device, rank=rank, world=world, unique_id=box[0],
                   exchange_mode=exchange_mode)

    def set_option(self, key, value):
        check(self._lib.bri17_rs_plan_set_option(self._plan, key.encode(), int(value)))

    def info(self, key):
        """"fused_axis0", "k1_major", "k1_major_real", "fused_launches", "pipeline", "exchange_mode",
        "exchange_chunks", "exchange_chunks_real", "fft_chunk_mib", "fft_chunk_planes", "barriers"."""
        v = C.c_int64()
        check(self._lib.bri17_rs_plan_get_info(self._plan, key.encode(), C.byref(v)))
        return v.value

    def close(self):
        plan, self._plan = getattr(self, "_plan", None), None
        if plan:
            self._lib.bri17_rs_plan_destroy(plan)

    __del__ = close

    def _check(self, t, shape, what, dtype=None):
        import torch
        dtype = dtype or torch.complex128
        if tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            raise ValueError(f"{what}: expected {dtype} {tuple(shape)}, got {t.dtype} {tuple(t.shape)}")
        if not t.is_cuda or t.device.index != self.device:
            raise ValueError(f"{what}: expected a tensor on cuda:{self.device}, got {t.device}")
        if not t.is_contiguous():
            raise ValueError(f"{what}: expected a contiguous tensor")

    def apply(self, u, out=None, stream=None):
        """F = (|h|/|N|) iDFT(K^ DFT(u)) on this rank's slab (collective)."""
        import torch
        self._check(u, self.real_shape, "u")
        if out is None:
            out = torch.empty_like(u)
        self._check(out, self.real_shape, "out")
        check(self._lib.bri17_real_space_apply_f64(self._plan, _dev_ptr(u), _dev_ptr(out),
                                                   _stream_ptr(stream)))
        return out

    def apply_real(self, u, out=None, stream=None):
        """Same operator on a REAL field, float64 ``(dim, n0_count, N1[, N2])``: r2c
        half-spectrum path (half the FFT, exchange and modal work)."""
        import torch
        self._check(u, self.real_shape, "u", torch.float64)
        if out is None:
            out = torch.empty_like(u)
        self._check(out, self.real_shape, "out", torch.float64)
        check(self._lib.bri17_real_space_apply_real_f64(self._plan, _dev_ptr(u), _dev_ptr(out),
                                                        _stream_ptr(stream)))
        return out

    def apply_with_dot(self, u, out=None, stream=None):
        """``(F, <u, F>)``: the operator and the global scalar product in one go (the Parseval
        sum is accumulated by the kernel that applies K^).  ``u`` complex128 or float64."""
        import torch
        real = u.dtype == torch.float64
        self._check(u, self.real_shape, "u", u.dtype if real else torch.complex128)
        if out is None:
            out = torch.empty_like(u)
        self._check(out, self.real_shape, "out", u.dtype)
        dot = C.c_double()
        check(self._lib.bri17_real_space_apply_dot_f64(self._plan, _dev_ptr(u), _dev_ptr(out), int(real),
                                                       C.byref(dot), _stream_ptr(stream)))
        return out, dot.value

    def cg_solve_real(self, b, rtol=1e-8, max_iter=1000, check_every=10, stream=None):
        """CG on real float64 fields (see cg_solve)."""
        import torch
        self._check(b, self.real_shape, "b", torch.float64)
        x = torch.empty_like(b)
        it, res = C.c_int(), C.c_double()
        check(self._lib.bri17_cg_solve_real_f64(self._plan, _dev_ptr(b), _dev_ptr(x), float(rtol),
                                                int(max_iter), int(check_every), C.byref(it),
                                                C.byref(res), _stream_ptr(stream)))
        return x, it.value, res.value

    def forward_fft(self, x, stream=None):
        """DFT (sign -1, unnormalised) of a real-space slab -> Fourier-space slab."""
        import torch
        ncomp = x.shape[0]
        self._check(x, (ncomp,) + self.real_shape[1:], "x")
        out = torch.empty((ncomp,) + self.fourier_shape[1:], dtype=x.dtype, device=x.device)
        check(self._lib.bri17_rs_forward_fft_f64(self._plan, _dev_ptr(x), _dev_ptr(out), ncomp,
                                                 _stream_ptr(stream)))
        return out

    def inverse_fft(self, x_hat, scale=None, stream=None):
        """Inverse DFT (x_hat is destroyed); ``scale`` defaults to 1/|N|."""
        import torch
        ncomp = x_hat.shape[0]
        self._check(x_hat, (ncomp,) + self.fourier_shape[1:], "x_hat")
        if scale is None:
            scale = 1.0 / float(np.prod(self.shape, dtype=np.float64))
        out = torch.empty((ncomp,) + self.real_shape[1:], dtype=x_hat.dtype, device=x_hat.device)
        check(self._lib.bri17_rs_inverse_fft_f64(self._plan, _dev_ptr(x_hat), _dev_ptr(out), ncomp,
                                                 float(scale), _stream_ptr(stream)))
        return out

    def timings(self):
        """Milliseconds per phase of the last apply (synchronises the stream)."""
        ms = (C.c_double * 8)()
        check(self._lib.bri17_rs_plan_last_timings(self._plan, ms, 8))
        return dict(zip(PHASES, ms))

    def cg_solve(self, b, rtol=1e-8, max_iter=1000, check_every=10, stream=None):
        """Solve A x = b (zero-mean b) by conjugate gradients; returns (x, iterations, rel_residual)."""
        import torch
        self._check(b, self.real_shape, "b")
        x = torch.empty_like(b)
        it, res = C.c_int(), C.c_double()
        check(self._lib.bri17_cg_solve_f64(self._plan, _dev_ptr(b), _dev_ptr(x), float(rtol),
                                           int(max_iter), int(check_every), C.byref(it), C.byref(res),
                                           _stream_ptr(stream)))
        return x, it.value, res.value
