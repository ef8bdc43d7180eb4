"""Per-mode direct solves: Hooke::modal_eigenstress_to_opposite_strain
(bri17.hpp:308-355) on the host and batched on the GPU.

PARITY UNPINNED against the reference: the method needs Eigen's LLT (absent,
unversioned) and no reference test calls it.  What is checked: (1) the oracle's
restatement against the method's mathematical definition evaluated with
numpy.linalg, (2) host C ABI and C++ header bit for bit against the oracle,
(3) GPU kernels against the oracle at 1e-12 per mode."""
import numpy as np
import pytest

import bri17_b200 as b

MU, NU = 1.0, 0.3   # python/demo.py:13-14


def _definition(o, shape, L, k, tau):
    dim = len(shape)
    s2 = np.sqrt(2.0)
    pairs = [(0, 0), (1, 1), (0, 1)] if dim == 2 else [(0, 0), (1, 1), (2, 2), (1, 2), (2, 0), (0, 1)]
    K = o.modal_stiffness(shape, L, MU, NU, k).real
    B = o.modal_strain_displacement(shape, L, k)
    T = np.zeros((dim, dim), dtype=complex)
    for s, (p, q) in enumerate(pairs):
        T[p, q] = T[q, p] = tau[s] if p == q else tau[s] / s2
    u = np.linalg.solve(K, T @ B.conj())
    E = 0.5 * (np.outer(B, u) + np.outer(u, B))
    return np.array([E[p, q] if p == q else s2 * E[p, q] for p, q in pairs]), u


CASES = [(2, (6, 8), (1.0, 1.5)), (3, (6, 8, 5), (1.0, 1.5, 2.0)), (2, (16, 16), (1.0, 1.0))]


@pytest.mark.parametrize("dim,shape,L", CASES)
def test_oracle_restatement_matches_definition(oracle_mod, dim, shape, L):
    nsym = dim * (dim + 1) // 2
    rng = np.random.default_rng(5)
    tau = rng.standard_normal((nsym,) + shape) + 1j * rng.standard_normal((nsym,) + shape)
    eta, u = oracle_mod.apply_eigenstress(shape, L, MU, NU, tau)
    o = oracle_mod.best()
    for k in np.ndindex(*shape):
        sl = (slice(None),) + k
        if not any(k):
            assert np.all(eta[sl] == 0) and np.all(u[sl] == 0)      # bri17.hpp:336-339
            continue
        e_def, u_def = _definition(o, shape, L, k, tau[sl])
        assert np.abs(eta[sl] - e_def).max() <= 1e-12 * np.abs(e_def).max()
        assert np.abs(u[sl] - u_def).max() <= 1e-12 * np.abs(u_def).max()


@pytest.mark.parametrize("dim,shape,L", CASES)
def test_host_per_mode_api_bitwise_vs_oracle(oracle_mod, dim, shape, L):
    nsym = dim * (dim + 1) // 2
    grid = (b.CartesianGrid2f64 if dim == 2 else b.CartesianGrid3f64)(shape, L)
    hooke = (b.Hooke2f64 if dim == 2 else b.Hooke3f64)(MU, NU, grid)
    rng = np.random.default_rng(6)
    tau = rng.standard_normal((nsym,) + shape) + 1j * rng.standard_normal((nsym,) + shape)
    eta_o, _ = oracle_mod.apply_eigenstress(shape, L, MU, NU, tau)
    e = np.empty(nsym, dtype=np.complex128)
    for k in np.ndindex(*shape):
        sl = (slice(None),) + k
        hooke.modal_eigenstress_to_opposite_strain(np.array(k, dtype=np.intc),
                                                   np.ascontiguousarray(tau[sl]), e)
        assert np.array_equal(e, eta_o[sl]), k


@pytest.mark.gpu
@pytest.mark.parametrize("dim,shape,L", CASES + [(3, (5, 4, 300), (1.0, 1.0, 7.0)), (2, (3, 700), (1.0, 9.0))])
def test_gpu_batched_solves_vs_oracle(oracle_mod, dim, shape, L):
    import torch
    nsym = dim * (dim + 1) // 2
    rng = np.random.default_rng(8)
    tau = rng.standard_normal((nsym,) + shape) + 1j * rng.standard_normal((nsym,) + shape)
    eta_o, u_o = oracle_mod.apply_eigenstress(shape, L, MU, NU, tau)
    op = b.ModalOperator(shape, L, MU, NU)
    td = torch.from_numpy(tau).cuda()

    def rel(a, ref):
        num, den = np.abs(a - ref).max(axis=0), np.abs(ref).max(axis=0)
        assert np.all(num[den == 0] == 0)
        return float((num[den > 0] / den[den > 0]).max())

    eta = op.eigenstress_to_opposite_strain(td).cpu().numpy()
    u = op.eigenstress_to_displacement(td).cpu().numpy()
    assert rel(eta, eta_o) <= 1e-12 and rel(u, u_o) <= 1e-12
    # mode-major layout of python/demo.py:21,37-38: tau[k0, k1(, k2), sym]
    tm = torch.from_numpy(np.ascontiguousarray(np.moveaxis(tau, 0, -1))).cuda()
    eta_m = op.eigenstress_to_opposite_strain(tm, mode_major=True).cpu().numpy()
    u_m = op.eigenstress_to_displacement(tm, mode_major=True).cpu().numpy()
    assert np.array_equal(np.moveaxis(eta_m, -1, 0), eta) and np.array_equal(np.moveaxis(u_m, -1, 0), u)
    # slabs (k_begin offsets) reproduce the full result
    a0 = shape[0] // 2
    part = op.eigenstress_to_opposite_strain(td[:, a0:].contiguous(),
                                             k_begin=(a0,) + (0,) * (dim - 1)).cpu().numpy()
    assert np.array_equal(part, eta[:, a0:])
    # K^-1 is the exact inverse of the stiffness apply away from k = 0
    v = torch.from_numpy(rng.standard_normal((dim,) + shape) + 1j * rng.standard_normal((dim,) + shape)).cuda()
    back = op.solve_modal_stiffness(op.apply_modal_stiffness(v)).cpu().numpy()
    vn = v.cpu().numpy().copy()
    vn[(slice(None),) + (0,) * dim] = 0
    assert np.abs(back - vn).max() <= 1e-11 * np.abs(vn).max()


@pytest.mark.gpu
def test_gpu_inclusion_problem_direct_vs_cg(oracle_mod):
    """The periodic inclusion problem of python/demo.py:11-23 (eigenstress patch
    [0, N/8)^dim, tau_in = last Mandel component): the one-pass direct solve in
    Fourier space against matrix-free CG in real space."""
    import torch
    from bri17_b200.realspace import RealSpaceOperator
    shape, L = (32, 32, 32), (1.0, 1.0, 1.0)
    dim, nsym = 3, 6
    op = b.ModalOperator(shape, L, MU, NU)
    rs = RealSpaceOperator(shape, L, MU, NU)
    tau = torch.zeros((nsym,) + shape, dtype=torch.complex128, device="cuda")
    tau[-1, :4, :4, :4] = 1.0                                   # python/demo.py:16-23
    tau_hat = rs.forward_fft(tau)
    u_hat = op.eigenstress_to_displacement(tau_hat)             # K^ u^ = tau^ . conj(B^)
    u_direct = rs.inverse_fft(u_hat.clone())
    assert float(u_direct.imag.abs().max()) <= 1e-13 * float(u_direct.real.abs().max())
    # the same displacement from CG: A u = |h| iDFT(rhs^), rhs^ = K^ u^
    rhs_hat = op.apply_modal_stiffness(u_hat)
    h_vol = float(np.prod([l / n for l, n in zip(L, shape)]))
    b_real = rs.inverse_fft(rhs_hat, scale=h_vol / float(np.prod(shape)))
    x, iters, res = rs.cg_solve(b_real, rtol=1e-11, max_iter=2000, check_every=5)
    assert res <= 1e-11
    assert float((x - u_direct).abs().max()) <= 1e-7 * float(u_direct.abs().max())
