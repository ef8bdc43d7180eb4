"""Correctness checks of the DISTRIBUTED real-space operator and CG, callable from a torchrun
worker (tests/dist_gpu_worker.py) and, as in-run parity checks, from bench.py at N > 1.

Checker code: uses the CPU oracle (tests/realspace_ref.py, oracle/) to judge the CUDA path; the
product never imports it.  Every function is collective over the default process group and
returns the MAX over ranks, so rank 0 can report it.

  small_grids       non-divisible, zero-slab and power-of-two grids against the numpy restatement
                    of tests/test_bri17.cpp:56-107 (complex and real fields)
  dense_kat         the reference's dense-matrix known-answer test (tests/test_bri17.cpp:130-150,
                    :335-361, :363-536) with compute_Ku running on all ranks
  vs_single_gpu     distributed apply against the same ranks' slabs of a single-GPU apply
  plane_waves       size-independent property at the FULL benchmark size: a superposition of a
                    few Fourier modes must come back as |h| K^(k) a e^{i phi}, K^(k) from the oracle
  inclusion_problem the periodic inclusion problem of python/demo.py:11-23 (BASELINE config 5):
                    right-hand side tau^ . conj(B^), CG to rtol, against the per-mode direct solve
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SPACING = (1.1, 1.2, 1.3)


def spacing_L(shape):
    return tuple(float(n) * h for n, h in zip(shape, SPACING))


def _max_over_ranks(value, device):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _make(shape, L, mu, nu, local_rank, mode, pipeline=None, fused=None, chunks=None):
    """fused: 1 = fused axis-0 pass on the k1-major layout (default), 2 = fused pass on the natural
    layout, 0 = cuFFT + modal kernel + cuFFT.  chunks: sub-slabs per component of the pipelined
    apply (None = the plan's own choice by size)."""
    from bri17_b200.realspace import RealSpaceOperator
    op = RealSpaceOperator.from_process_group(shape, L, mu, nu, device=local_rank, exchange_mode=mode)
    if pipeline is not None:
        op.set_option("pipeline", pipeline)
    if fused is not None:
        op.set_option("fused_axis0", 1 if fused else 0)
        op.set_option("k1_major", 1 if fused == 1 else 0)
    if chunks is not None:
        op.set_option("exchange_chunks", chunks)
    return op


def small_grids(local_rank, mode=1, pipeline=1, fused=1, mu=5.6, nu=0.3,
                shapes=((66, 60, 50), (3, 4, 5), (32, 32, 32), (64, 20, 18), (12, 10), (64, 48)), chunks=None):
    """Distributed apply (complex + real fields, applied twice: buffer reuse) against the numpy
    restatement.  (66, 60, 50) divides by no rank count; (3, 4, 5) leaves ranks without rows when
    world > 3 (and without k1 columns when world > 4)."""
    import torch
    from oracle import oracle
    from realspace_ref import real_space_apply_ref
    dev = torch.device("cuda", local_rank)
    o = oracle.best()
    worst = 0.0
    for shape in shapes:
        dim = len(shape)
        L = spacing_L(shape)
        rng = np.random.default_rng(100 + sum(shape))
        u = rng.standard_normal((dim,) + shape)
        ref = real_space_apply_ref(o, shape, L, mu, nu, u + 0j)
        scale = np.abs(ref).max()
        op = _make(shape, L, mu, nu, local_rank, mode, pipeline, fused, chunks)
        a0, a1 = op.n0_begin, op.n0_begin + op.n0_count
        uc = torch.from_numpy(np.ascontiguousarray(u[:, a0:a1]) + 0j).to(dev)
        ur = torch.from_numpy(np.ascontiguousarray(u[:, a0:a1])).to(dev)
        err = 0.0
        for _ in range(2):
            F = op.apply(uc).cpu().numpy()
            Fr = op.apply_real(ur).cpu().numpy()
        if F.size:
            err = max(np.abs(F - ref[:, a0:a1]).max(), np.abs(Fr - ref[:, a0:a1].real).max()) / scale
        # <u, A u> from the operator itself against the plain scalar product (all ranks)
        _, dot = op.apply_with_dot(ur)
        expect = float(np.sum(u * ref.real))
        err = max(err, abs(dot - expect) / abs(expect) * 1e-1)       # 1e-12 on the 1e-13 scale
        worst = max(worst, err)
        op.close()
    return _max_over_ranks(worst, dev)


def dense_kat(local_rank, world, mode=1, pipeline=1, real=False, chunks=None):
    """Column j of the dense stiffness matrix = distributed real_space_apply(e_j), compared entry
    by entry with the classical FE assembly of the Maxima element matrix at the reference's
    tolerance 1e-15*|e| + 1e-14; the reference's own grid (3, 4, 5) -- zero-row slabs for
    world > 3 -- and a grid whose first two extents divide by the rank count."""
    import torch
    from oracle import kat
    dev = torch.device("cuda", local_rank)
    E = kat.load_elements()
    worst = 0.0
    n = max(4, world)
    for shape in ((3, 4, 5), (n, n, 5)):
        dim = 3
        L = tuple(float(m) * h for m, h in zip(shape, kat.SPACING[3]))
        op = _make(shape, L, kat.MU, kat.NU, local_rank, mode, pipeline, chunks=chunks)
        a0, cnt = op.n0_begin, op.n0_count
        size = int(np.prod(shape))
        plane = shape[1] * shape[2]
        K_loc = np.zeros((dim, cnt * plane, size * dim))
        dtype = torch.float64 if real else torch.complex128
        u = torch.zeros((dim, cnt) + shape[1:], dtype=dtype, device=dev)
        max_imag = 0.0
        for j in range(size * dim):
            c, i = divmod(j, size)
            row, rest = divmod(i, plane)
            mine = a0 <= row < a0 + cnt
            if mine:
                u[c].view(-1)[(row - a0) * plane + rest] = 1.0
            Ku = (op.apply_real(u) if real else op.apply(u)).cpu().numpy()
            if mine:
                u[c].view(-1)[(row - a0) * plane + rest] = 0.0
            if not real and Ku.size:
                max_imag = max(max_imag, float(np.abs(Ku.imag).max()))       # :140-144
            K_loc[:, :, j] = Ku.real.reshape(dim, cnt * plane)
        expected = kat.assemble_expected_stiffness(shape, E["Ke3"])
        err = 0.0
        for c in range(dim):                                                    # DOF = node + |N|*component
            rows = slice(c * size + a0 * plane, c * size + (a0 + cnt) * plane)
            e, a = expected[rows], K_loc[c]
            if e.size:
                viol = np.abs(a - e) - (kat.RTOL * np.abs(e) + kat.ATOL)         # :17
                err = max(err, float(viol.max()))
        worst = max(worst, err, max_imag - kat.IMAG_TOL)
        op.close()
    return _max_over_ranks(worst, dev)      # <= 0 means every entry is within the reference tolerance


def vs_single_gpu(local_rank, edge=256, mode=1, mu=5.6, nu=0.3, chunks=None):
    """Distributed apply on an edge^3 grid against this rank's slab of a SINGLE-GPU apply of the
    same field (generated from one seed on every rank)."""
    import torch
    from bri17_b200.realspace import RealSpaceOperator
    dev = torch.device("cuda", local_rank)
    shape = (edge,) * 3
    L = spacing_L(shape)
    g = torch.Generator(device=dev).manual_seed(256)
    u = torch.randn((3,) + shape, dtype=torch.float64, device=dev, generator=g)
    single = RealSpaceOperator(shape, L, mu, nu, device=local_rank)
    F1 = single.apply_real(u)
    single.close()
    op = _make(shape, L, mu, nu, local_rank, mode, chunks=chunks)
    a0, a1 = op.n0_begin, op.n0_begin + op.n0_count
    us = u[:, a0:a1].contiguous()
    Fr = op.apply_real(us)
    Fc = op.apply(us + 0j)
    scale = float(F1.abs().max())
    err = max(float((Fr - F1[:, a0:a1]).abs().max()), float((Fc.real - F1[:, a0:a1]).abs().max()),
              float(Fc.imag.abs().max())) / scale if us.numel() else 0.0
    op.close()
    return _max_over_ranks(err, dev)


def _wave(shape, k, n0_begin, n0_count, dev):
    """e^{+2 pi i (k . n / N)} on this rank's slab, from exactly reduced 1-D phases."""
    import torch
    parts = []
    for d, (kd, nd) in enumerate(zip(k, shape)):
        n = torch.arange(n0_begin, n0_begin + n0_count, device=dev) if d == 0 else torch.arange(nd, device=dev)
        ph = ((n * int(kd)) % nd).to(torch.float64) * (2.0 * math.pi / nd)
        parts.append(torch.complex(torch.cos(ph), torch.sin(ph)))
    return parts[0][:, None, None] * parts[1][None, :, None] * parts[2][None, None, :]


def plane_waves(op, real, nwaves=4, seed=5):
    """u = sum_m a_m e^{i phi_m} (real part for real fields) must come back as
    sum_m |h| K^(k_m) a_m e^{i phi_m}: one K^ per mode from the CPU oracle, everything else
    (local FFTs, exchange, axis-0 pass, inverse) at the operator's full size on the GPUs."""
    import torch
    from oracle import oracle
    o = oracle.best()
    dev = torch.device("cuda", op.device)
    shape, L = op.shape, op.L
    rng = np.random.default_rng(seed)
    ks = [tuple(int(rng.integers(0, n)) for n in shape) for _ in range(nwaves)]
    ks[0] = (1, shape[1] - 1, shape[2] // 2)                      # low / wrapped / Nyquist indices
    h_vol = float(np.prod([l / n for l, n in zip(L, shape)]))
    amps = [rng.standard_normal(3) + 1j * rng.standard_normal(3) for _ in ks]
    Kas = [(o.modal_stiffness(shape, L, op.mu, op.nu, k).real @ a) * h_vol for k, a in zip(ks, amps)]
    # memory-lean on purpose (the fields are tens of GiB at 1024^3): waves are regenerated instead
    # of kept, the expectation is formed one component at a time
    u = torch.zeros(op.real_shape, dtype=torch.float64 if real else torch.complex128, device=dev)
    for k, a in zip(ks, amps):
        w = _wave(shape, k, op.n0_begin, op.n0_count, dev)
        for c in range(3):
            if real:
                u[c].add_(w.real, alpha=float(a[c].real)).add_(w.imag, alpha=-float(a[c].imag))
            else:
                u[c].add_(w, alpha=complex(a[c]))
        del w
    F = op.apply_real(u) if real else op.apply(u)
    del u
    err, scale = 0.0, 0.0
    for c in range(3):
        E = torch.zeros(op.real_shape[1:], dtype=F.dtype, device=dev)
        for k, Ka in zip(ks, Kas):
            w = _wave(shape, k, op.n0_begin, op.n0_count, dev)
            if real:
                E.add_(w.real, alpha=float(Ka[c].real)).add_(w.imag, alpha=-float(Ka[c].imag))
            else:
                E.add_(w, alpha=complex(Ka[c]))
            del w
        if E.numel():
            scale = max(scale, float(E.abs().max()))
            err = max(err, float((F[c] - E).abs().max()))
        del E
    scale = _max_over_ranks(scale, dev)
    return _max_over_ranks(err / scale if scale > 0 else 0.0, dev)


def inclusion_problem(op, rtol=1e-8, max_iter=20000, check_every=25, mu=None, nu=None):
    """python/demo.py:11-23 in 3-D on op's grid: eigenstress patch [0, N/8)^3 with tau_in = unit
    last Mandel component, b = (|h|/|N|) iDFT(tau^ . conj(B^)) (bri17.hpp:340, theory.rst:151-157),
    CG on real fields, solution compared with the one-pass direct solve u^ = K^-1 (tau^ . conj B^)
    (bri17.hpp:341).  Returns a dict; times the CG call with CUDA events."""
    import torch
    import bri17_b200 as b
    dev = torch.device("cuda", op.device)
    shape, L = op.shape, op.L
    dim = len(shape)
    nsym = dim * (dim + 1) // 2
    modal = b.ModalOperator(shape, L, op.mu, op.nu, device=op.device)
    patch = [max(1, n // 8) for n in shape]
    tau = torch.zeros((nsym, op.n0_count) + shape[1:], dtype=torch.complex128, device=dev)
    lo, hi = op.n0_begin, min(op.n0_begin + op.n0_count, patch[0])
    if hi > lo:
        tau[(-1, slice(0, hi - lo)) + tuple(slice(0, p) for p in patch[1:])] = 1.0
    tau_hat = op.forward_fft(tau)
    del tau
    kb = (0, op.k1_begin) + (0,) * (dim - 2)
    total = float(np.prod(shape, dtype=np.float64))
    h_vol = float(np.prod([l / n for l, n in zip(L, shape)]))
    if tau_hat.numel():
        f_hat = modal.eigenstress_to_force(tau_hat, k_begin=kb)
        u_hat = modal.eigenstress_to_displacement(tau_hat, k_begin=kb)
    else:
        f_hat = torch.empty((dim,) + tuple(tau_hat.shape[1:]), dtype=tau_hat.dtype, device=dev)
        u_hat = torch.empty_like(f_hat)
    del tau_hat
    b_real = op.inverse_fft(f_hat, scale=h_vol / total).real.contiguous()
    u_direct = op.inverse_fft(u_hat).real.contiguous()
    del f_hat, u_hat
    torch.cuda.synchronize()
    op.cg_solve_real(b_real, rtol=0.0, max_iter=2, check_every=0)        # allocates the work vectors
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream()
    e0.record(stream)
    x, iters, res = op.cg_solve_real(b_real, rtol=rtol, max_iter=max_iter, check_every=check_every)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = _max_over_ranks(e0.elapsed_time(e1), dev)
    scale = _max_over_ranks(float(u_direct.abs().max()) if u_direct.numel() else 0.0, dev)
    err = _max_over_ranks(float((x - u_direct).abs().max()) / scale if x.numel() else 0.0, dev)
    return {"iterations": int(iters), "rel_residual": float(res), "rtol": rtol, "ms_total": ms,
            "iterations_per_s": iters / (ms * 1e-3) if ms > 0 else None,
            "max_err_vs_direct_solve": err,
            "patch": patch, "converged": bool(res <= rtol)}
