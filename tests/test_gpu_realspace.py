"""GPU tests of the end-to-end real-space operator (cuFFT -> modal kernel ->
cuFFT) and of CG, single GPU; the multi-GPU variants run under torchrun from
test_multi_gpu_realspace (needs >= 2 GPUs)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from bri17_b200.realspace import RealSpaceOperator  # noqa: E402
from oracle import kat  # noqa: E402
from realspace_ref import direct_solve_ref, real_space_apply_ref  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MU, NU = 5.6, 0.3


def spacing_L(shape):
    return tuple(float(n) * h for n, h in zip(shape, (1.1, 1.2, 1.3)))


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_kat_end_to_end_on_gpu(dim):
    """The reference's "Global assembly tests" (tests/test_bri17.cpp:335-361,
    :363-536) with the WHOLE compute_Ku on the GPU: column j of the dense
    stiffness matrix = real_space_apply(e_j)."""
    E = kat.load_elements()
    shape, L = kat.SHAPE[dim], kat.grid_L(dim)
    op = RealSpaceOperator(shape, L, kat.MU, kat.NU)
    size = int(np.prod(shape))
    K = np.zeros((size * dim, size * dim))
    u = torch.zeros((dim,) + shape, dtype=torch.complex128, device="cuda")
    max_imag = 0.0
    for j in range(size * dim):
        u.view(-1)[j] = 1.0
        Ku = op.apply(u).cpu().numpy()
        u.view(-1)[j] = 0.0
        max_imag = max(max_imag, float(np.abs(Ku.imag).max()))
        K[:, j] = Ku.real.ravel()
    assert max_imag <= kat.IMAG_TOL                                   # :140-144
    kat.assert_equal(kat.assemble_expected_stiffness(shape, E[f"Ke{dim}"]), K)   # :360, :535


@pytest.mark.parametrize("shape", [(16, 12, 10), (64, 64), (5, 7), (8, 8, 33), (32, 32, 32)])
def test_real_space_apply_vs_numpy_restatement(oracle_mod, shape):
    dim = len(shape)
    L = spacing_L(shape)
    rng = np.random.default_rng(41)
    u = rng.standard_normal((dim,) + shape) + 0j
    ref = real_space_apply_ref(oracle_mod.best(), shape, L, MU, NU, u)
    op = RealSpaceOperator(shape, L, MU, NU)
    F = op.apply(torch.from_numpy(u).cuda()).cpu().numpy()
    assert np.abs(F - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.abs(F.imag).max() <= 1e-13 * np.abs(ref).max()
    t = op.timings()
    assert t["total"] > 0 and set(t) >= {"modal", "exchange_fwd"}


@pytest.mark.parametrize("shape", [(16, 12, 10), (64, 64), (5, 7), (8, 8, 33), (9, 6, 7), (32, 32, 32)])
def test_real_half_spectrum_path_vs_numpy_restatement(oracle_mod, shape):
    """r2c / c2r path on real float64 fields against the
    c2c restatement of the reference harness: same operator, half the work."""
    dim = len(shape)
    L = spacing_L(shape)
    rng = np.random.default_rng(43)
    u = rng.standard_normal((dim,) + shape)
    ref = real_space_apply_ref(oracle_mod.best(), shape, L, MU, NU, u + 0j).real
    op = RealSpaceOperator(shape, L, MU, NU)
    ud = torch.from_numpy(u).cuda()
    F = op.apply_real(ud).cpu().numpy()
    assert np.abs(F - ref).max() <= 1e-13 * np.abs(ref).max()
    assert torch.equal(ud.cpu(), torch.from_numpy(u))                 # input preserved
    Fc = op.apply(torch.from_numpy(u + 0j).cuda()).cpu().numpy()      # and the c2c path agrees
    assert np.abs(F - Fc.real).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_kat_through_real_path(dim):
    """The reference's dense-matrix known-answer tests through the r2c path."""
    E = kat.load_elements()
    shape, L = kat.SHAPE[dim], kat.grid_L(dim)
    op = RealSpaceOperator(shape, L, kat.MU, kat.NU)
    size = int(np.prod(shape))
    K = np.zeros((size * dim, size * dim))
    u = torch.zeros((dim,) + shape, dtype=torch.float64, device="cuda")
    for j in range(size * dim):
        u.view(-1)[j] = 1.0
        K[:, j] = op.apply_real(u).cpu().numpy().ravel()
        u.view(-1)[j] = 0.0
    kat.assert_equal(kat.assemble_expected_stiffness(shape, E[f"Ke{dim}"]), K)


def test_cg_real_path(oracle_mod):
    shape = (16, 12, 10)
    L = spacing_L(shape)
    rng = np.random.default_rng(5)
    x_true = rng.standard_normal((3,) + shape)
    x_true -= x_true.mean(axis=(1, 2, 3), keepdims=True)
    b = real_space_apply_ref(oracle_mod.best(), shape, L, MU, NU, x_true + 0j).real
    op = RealSpaceOperator(shape, L, MU, NU)
    x, iters, res = op.cg_solve_real(torch.from_numpy(np.ascontiguousarray(b)).cuda(), rtol=1e-11,
                                     max_iter=2000, check_every=5)
    assert res <= 1e-11 and np.abs(x.cpu().numpy() - x_true).max() <= 1e-7 * np.abs(x_true).max()


def test_forward_inverse_fft_conventions():
    """theory.rst:60 (sign -1, unnormalised) and :72 (1/|N| on the inverse)."""
    shape = (12, 10, 9)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3,) + shape) + 1j * rng.standard_normal((3,) + shape)
    op = RealSpaceOperator(shape, spacing_L(shape), MU, NU)
    xd = torch.from_numpy(x).cuda()
    xh = op.forward_fft(xd)
    ref = np.fft.fftn(x, axes=(1, 2, 3))
    assert np.abs(xh.cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()
    assert torch.equal(xd.cpu(), torch.from_numpy(x))                 # input preserved
    back = op.inverse_fft(xh).cpu().numpy()
    assert np.abs(back - x).max() <= 1e-13 * np.abs(x).max()
    x6 = rng.standard_normal((6,) + shape) + 0j                        # strain-like field, 6 components
    xh6 = op.forward_fft(torch.from_numpy(x6).cuda()).cpu().numpy()
    assert np.abs(xh6 - np.fft.fftn(x6, axes=(1, 2, 3))).max() <= 1e-12


@pytest.mark.parametrize("shape", [(16, 12, 10), (24, 20)])
def test_cg_solves_the_periodic_problem(oracle_mod, shape):
    dim = len(shape)
    L = spacing_L(shape)
    rng = np.random.default_rng(3)
    x_true = rng.standard_normal((dim,) + shape)
    x_true -= x_true.mean(axis=tuple(range(1, dim + 1)), keepdims=True)   # u^(0) = 0 (theory.rst:208-212)
    o = oracle_mod.best()
    b = real_space_apply_ref(o, shape, L, MU, NU, x_true + 0j)
    op = RealSpaceOperator(shape, L, MU, NU)
    x, iters, res = op.cg_solve(torch.from_numpy(b).cuda(), rtol=1e-11, max_iter=2000, check_every=5)
    x = x.cpu().numpy()
    assert res <= 1e-11 and 0 < iters < 2000
    assert np.abs(x - x_true).max() <= 1e-7 * np.abs(x_true).max()
    direct = direct_solve_ref(o, shape, L, MU, NU, b)
    assert np.abs(x - direct).max() <= 1e-7 * np.abs(direct).max()


@pytest.mark.parametrize("shape", [(16, 12, 10), (32, 32, 32), (64, 12, 10), (128, 6, 5), (256, 4, 6),
                                   (512, 3, 4), (1024, 2, 3), (64, 64), (16, 9)])
def test_fused_axis0_pass_equals_cufft_path(oracle_mod, shape):
    """The one-kernel FFT(axis 0) -> K^ -> iFFT(axis 0) pass (every supported length, both
    field types) against the cuFFT + modal kernel + cuFFT path and the numpy restatement."""
    dim = len(shape)
    L = spacing_L(shape)
    rng = np.random.default_rng(17)
    u = rng.standard_normal((dim,) + shape)
    ref = real_space_apply_ref(oracle_mod.best(), shape, L, MU, NU, u + 0j)
    op = RealSpaceOperator(shape, L, MU, NU)
    assert op.info("fused_axis0") == 1
    uc, ur = torch.from_numpy(u + 0j).cuda(), torch.from_numpy(u).cuda()
    Fc, Fr = op.apply(uc).cpu().numpy(), op.apply_real(ur).cpu().numpy()
    assert op.info("fused_launches") == 2
    # default layout policy on one GPU: k1-major for complex fields (3-D), natural for real fields
    assert op.info("k1_major") == (dim == 3) and op.info("k1_major_real") == 0
    op.set_option("k1_major", 0)                           # fused pass on the natural layout
    Hc = op.apply(uc).cpu().numpy()
    op.set_option("k1_major", 1)                           # ... and forced k1-major for real fields too
    assert op.info("k1_major_real") == (dim == 3)
    Hr = op.apply_real(ur).cpu().numpy()
    assert op.info("fused_launches") == 4
    op.set_option("k1_major", 0)
    op.set_option("fused_axis0", 0)
    assert op.info("fused_axis0") == 0
    Gc, Gr = op.apply(uc).cpu().numpy(), op.apply_real(ur).cpu().numpy()
    assert op.info("fused_launches") == 4
    scale = np.abs(ref).max()
    for got in (Fc, Gc, Hc):
        assert np.abs(got - ref).max() <= 1e-13 * scale
    for got in (Fr, Gr, Hr):
        assert np.abs(got - ref.real).max() <= 1e-13 * scale
    assert np.abs(Fc - Gc).max() <= 1e-13 * scale and np.abs(Fr - Gr).max() <= 1e-13 * scale
    assert torch.equal(uc.cpu(), torch.from_numpy(u + 0j)) and torch.equal(ur.cpu(), torch.from_numpy(u))


@pytest.mark.parametrize("shape", [(16, 12, 10), (9, 6, 7), (64, 33), (12, 10)])
@pytest.mark.parametrize("fused", [1, 0])
def test_apply_with_dot(oracle_mod, shape, fused):
    """F = A u and <u, A u> in one go (Parseval sum inside the K^ kernel), complex and real
    fields, fused and cuFFT axis-0 paths, against the plain scalar product of the fields."""
    dim = len(shape)
    L = spacing_L(shape)
    rng = np.random.default_rng(23)
    u = rng.standard_normal((dim,) + shape)
    op = RealSpaceOperator(shape, L, MU, NU)
    op.set_option("fused_axis0", fused)
    ur = torch.from_numpy(u).cuda()
    F, dot = op.apply_with_dot(ur)
    assert torch.equal(F, op.apply_real(ur))
    expect = float((ur * F).sum())
    assert expect > 0 and abs(dot - expect) <= 1e-12 * expect
    v = rng.standard_normal((dim,) + shape) + 1j * rng.standard_normal((dim,) + shape)
    vc = torch.from_numpy(v).cuda()
    G, dotc = op.apply_with_dot(vc)
    assert torch.equal(G, op.apply(vc))
    expect = float((vc.conj() * G).sum().real)
    assert abs(dotc - expect) <= 1e-12 * expect


@pytest.mark.parametrize("shape", [(16, 12, 10), (7, 5, 3), (9, 6, 7), (12, 10), (32, 8, 6)])
@pytest.mark.parametrize("k1_major", [0, 1])
def test_local_transforms_in_chunks_of_planes(oracle_mod, shape, k1_major):
    """Option "fft_chunk_mib": the local transforms run plane range by plane range through the
    per-batch-size plan cache (the machinery the multi-GPU sub-slab pipeline uses).  With chunks of a
    single plane (two when the real planes hold an odd number of doubles: 16-byte alignment) the
    result must equal the whole-slab transform bit for bit, complex and real fields, natural and
    k1-major spectral layouts."""
    dim = len(shape)
    L = spacing_L(shape)
    rng = np.random.default_rng(7)
    u = torch.from_numpy(rng.standard_normal((dim,) + shape)).cuda()
    whole = RealSpaceOperator(shape, L, MU, NU)
    whole.set_option("k1_major", k1_major)
    chunked = RealSpaceOperator(shape, L, MU, NU)
    chunked.set_option("k1_major", k1_major)
    chunked.set_option("fft_chunk_planes", 1)          # "fft_chunk_mib" gives the same in MiB
    Fw, Fc = whole.apply(u + 0j), chunked.apply(u + 0j)
    # (a 1-D local transform -- 2-D grids -- is a single cuFFT kernel and is never split)
    assert chunked.info("fft_chunk_planes") == (1 if dim == 3 else shape[0])
    assert whole.info("fft_chunk_planes") == shape[0]
    assert torch.equal(Fw, Fc)
    Rw, Rc = whole.apply_real(u), chunked.apply_real(u)
    assert torch.equal(Rw, Rc)
    ref = real_space_apply_ref(oracle_mod.best(), shape, L, MU, NU, u.cpu().numpy() + 0j)
    assert np.abs(Rc.cpu().numpy() - ref.real).max() <= 1e-13 * np.abs(ref).max()


def test_cg_projects_out_the_null_space(oracle_mod):
    """K^(0) = 0 (bri17.hpp:336-339, theory.rst:208-212): a right-hand side with a non-zero
    mean is projected, CG converges to the zero-mean solution instead of drifting."""
    shape = (16, 12, 10)
    L = spacing_L(shape)
    rng = np.random.default_rng(9)
    x_true = rng.standard_normal((3,) + shape)
    x_true -= x_true.mean(axis=(1, 2, 3), keepdims=True)
    b = real_space_apply_ref(oracle_mod.best(), shape, L, MU, NU, x_true + 0j).real
    shift = np.array([0.3, -1.7, 2.5]).reshape(3, 1, 1, 1) * np.abs(b).max()
    op = RealSpaceOperator(shape, L, MU, NU)
    x, iters, res = op.cg_solve_real(torch.from_numpy(np.ascontiguousarray(b + shift)).cuda(), rtol=1e-11,
                                     max_iter=2000, check_every=5)
    assert res <= 1e-11 and 0 < iters < 2000
    assert np.abs(x.cpu().numpy() - x_true).max() <= 1e-7 * np.abs(x_true).max()
    xc, _, resc = op.cg_solve(torch.from_numpy(b + shift + 1j * shift).cuda(), rtol=1e-11, max_iter=2000,
                              check_every=5)
    assert resc <= 1e-11
    assert np.abs(xc.cpu().numpy() - x_true).max() <= 1e-7 * np.abs(x_true).max()
    # a constant right-hand side is entirely in the null space: zero iterations, x = 0
    x0, it0, res0 = op.cg_solve_real(torch.ones((3,) + shape, dtype=torch.float64, device="cuda"))
    assert it0 == 0 and res0 == 0.0 and float(x0.abs().max()) == 0.0


def test_inclusion_problem_cg_vs_direct_solve(oracle_mod):
    """BASELINE config 5 in small: the periodic inclusion problem of python/demo.py:11-23
    (eigenstress patch [0, N/8)^3, tau_in = last Mandel component).  Right-hand side from
    tau^ . conj(B^) (bri17.hpp:340), matrix-free CG on real fields to rtol 1e-10, against the
    one-pass per-mode direct solve (bri17.hpp:341)."""
    import bri17_b200 as b
    shape, L = (64, 64, 64), (1.0, 1.0, 1.0)
    mu, nu = 1.0, 0.3                                              # python/demo.py:13-14
    rs = RealSpaceOperator(shape, L, mu, nu)
    op = b.ModalOperator(shape, L, mu, nu)
    tau = torch.zeros((6,) + shape, dtype=torch.complex128, device="cuda")
    tau[-1, :8, :8, :8] = 1.0                                      # python/demo.py:16-23
    tau_hat = rs.forward_fft(tau)
    f_hat = op.eigenstress_to_force(tau_hat)
    u_direct = rs.inverse_fft(op.eigenstress_to_displacement(tau_hat)).real
    h_vol = float(np.prod([l / n for l, n in zip(L, shape)]))
    b_real = rs.inverse_fft(f_hat, scale=h_vol / float(np.prod(shape))).real.contiguous()
    x, iters, res = rs.cg_solve_real(b_real, rtol=1e-10, max_iter=5000, check_every=10)
    assert res <= 1e-10 and iters < 5000
    assert float((x - u_direct).abs().max()) <= 1e-6 * float(u_direct.abs().max())


def test_multi_gpu_realspace():
    """Every-GPU run of tests/dist_gpu_worker.py (both exchange modes)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = torch.cuda.device_count()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert out.returncode == 0
    assert f"DIST_REALSPACE_OK {world}" in out.stdout
