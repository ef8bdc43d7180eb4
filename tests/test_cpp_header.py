"""The drop-in C++ header include/bri17/bri17.hpp: its per-mode API against the
oracle (bit for bit), and -- on the GPU -- bri17::ModalOperator against the
reference-style loop nest written with the header's own Hooke::modal_stiffness."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


@pytest.fixture(scope="module")
def hdr(oracle_mod):
    out = subprocess.run(["make", "-C", CPP, "libheader_shim.so"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    impl = oracle_mod._Impl.__new__(oracle_mod._Impl)
    impl.path, impl.prefix, impl.kind = os.path.join(CPP, "libheader_shim.so"), "hdr", "header"
    impl.lib = C.CDLL(impl.path)
    i32p, f64p = C.POINTER(C.c_int), C.POINTER(C.c_double)
    impl.lib.hdr_modal_stiffness.argtypes = [C.c_int, i32p, f64p, C.c_double, C.c_double, i32p, f64p]
    impl.lib.hdr_modal_strain_displacement.argtypes = [C.c_int, i32p, f64p, i32p, f64p]
    impl.lib.hdr_get_cell_nodes.argtypes = [C.c_int, i32p, C.c_int, i32p]
    impl.lib.hdr_eigenstress_to_opposite_strain.argtypes = [C.c_int, i32p, f64p, C.c_double, C.c_double,
                                                            i32p, f64p, f64p]
    impl.lib.hdr_repr.argtypes = [C.c_int, i32p, f64p, C.c_double, C.c_double, C.c_int, C.c_char_p, C.c_int]
    for name in ("hdr_modal_stiffness", "hdr_modal_strain_displacement", "hdr_get_cell_nodes",
                 "hdr_eigenstress_to_opposite_strain"):
        getattr(impl.lib, name).restype = None
    return impl


@pytest.mark.parametrize("shape,L", [((3, 4), (3.3, 4.8)), ((3, 4, 5), (3.3, 4.8, 6.5)),
                                      ((64, 64), (70.4, 76.8)), ((7, 6, 33), (1., 2., 3.)),
                                      ((512, 2, 2), (1., 1., 1.))])
def test_per_mode_methods_bitwise(oracle_mod, hdr, shape, L):
    o = oracle_mod.best()
    for k in np.ndindex(*shape):
        assert np.array_equal(hdr.modal_stiffness(shape, L, 5.6, 0.3, k),
                              o.modal_stiffness(shape, L, 5.6, 0.3, k)), k
        assert np.array_equal(hdr.modal_strain_displacement(shape, L, k),
                              o.modal_strain_displacement(shape, L, k)), k
    for cell in range(int(np.prod(shape))):
        assert np.array_equal(hdr.get_cell_nodes(shape, cell), o.get_cell_nodes(shape, cell))


def test_repr_and_float_instantiation(oracle_mod, hdr):
    buf = C.create_string_buffer(512)
    sh = (C.c_int * 3)(3, 4, 5)
    L = (C.c_double * 3)(3.3, 4.8, 6.5)
    hdr.lib.hdr_repr(3, sh, L, 5.6, 0.3, 0, buf, 512)
    assert buf.value.decode() == "CartesianGrid<d,3>{shape={3,4,5,},L={3.3,4.8,6.5,}}"
    hdr.lib.hdr_repr(3, sh, L, 5.6, 0.3, 1, buf, 512)
    assert buf.value.decode() == ("Hooke<d,3>{mu=5.6,nu=0.3,grid=CartesianGrid<d,3>"
                                  "{shape={3,4,5,},L={3.3,4.8,6.5,}}\n")
    r = oracle_mod.ref()
    if r is not None:
        buf2 = C.create_string_buffer(512)
        r.lib.ref_repr.argtypes = hdr.lib.hdr_repr.argtypes
        for which in (0, 1):
            r.lib.ref_repr(3, sh, L, 5.6, 0.3, which, buf2, 512)
            hdr.lib.hdr_repr(3, sh, L, 5.6, 0.3, which, buf, 512)
            assert buf.value == buf2.value
    assert hdr.lib.hdr_float_instantiates() == 1


@pytest.mark.parametrize("dim", [2, 3])
def test_eigenstress_to_opposite_strain(oracle_mod, hdr, dim):
    """bri17.hpp:308-355.  No reference test pins this method and Eigen is not
    available, so it is checked against its definition: K^ u = tau . conj(B^),
    eta = sym(B^ (x) u) in Mandel notation, zero at k = 0 (parity unpinned)."""
    shape = (6, 8) if dim == 2 else (6, 8, 5)
    L = (1.0, 1.5) if dim == 2 else (1.0, 1.5, 2.0)
    nsym = dim * (dim + 1) // 2
    o = oracle_mod.best()
    rng = np.random.default_rng(4)
    i32p, f64p = C.POINTER(C.c_int), C.POINTER(C.c_double)
    sh = np.array(shape, dtype=np.intc)
    ll = np.array(L, dtype=np.float64)
    s2 = np.sqrt(2.0)
    pairs = [(0, 0), (1, 1), (0, 1)] if dim == 2 else [(0, 0), (1, 1), (2, 2), (1, 2), (2, 0), (0, 1)]
    tau_all = rng.standard_normal((nsym,) + shape) + 1j * rng.standard_normal((nsym,) + shape)
    eta_oracle, _ = oracle_mod.apply_eigenstress(shape, L, 1.0, 0.3, tau_all)
    for k in list(np.ndindex(*shape))[::3]:
        tau = np.ascontiguousarray(tau_all[(slice(None),) + k])
        eta = np.empty(nsym, dtype=np.complex128)
        kk = np.array(k, dtype=np.intc)
        hdr.lib.hdr_eigenstress_to_opposite_strain(dim, sh.ctypes.data_as(i32p), ll.ctypes.data_as(f64p),
                                                   1.0, 0.3, kk.ctypes.data_as(i32p),
                                                   tau.ctypes.data_as(f64p), eta.ctypes.data_as(f64p))
        # same Cholesky, same order: header == oracle restatement bit for bit
        assert np.array_equal(eta, eta_oracle[(slice(None),) + k])
        if not any(k):
            assert np.all(eta == 0)
            continue
        K = o.modal_stiffness(shape, L, 1.0, 0.3, k).real
        B = o.modal_strain_displacement(shape, L, k)
        T = np.zeros((dim, dim), dtype=complex)
        for s, (p, q) in enumerate(pairs):
            T[p, q] = T[q, p] = tau[s] if p == q else tau[s] / s2
        u = np.linalg.solve(K, T @ B.conj())
        E = 0.5 * (np.outer(B, u) + np.outer(u, B))
        expected = np.array([E[p, q] if p == q else s2 * E[p, q] for p, q in pairs])
        assert np.abs(eta - expected).max() <= 1e-12 * max(np.abs(expected).max(), 1e-300)


@pytest.mark.gpu
def test_modal_operator_cpp_program():
    """tests/cpp/test_modal_operator.cpp: C++ host -> drop-in header -> C ABI -> kernels."""
    exe = os.path.join(CPP, "test_modal_operator")
    if not os.path.exists(exe):
        out = subprocess.run(["make", "-C", CPP, "test_modal_operator"], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("max_abs_diff 0.000e+00 invalid_argument_thrown 1") == 2


@pytest.mark.gpu
def test_real_space_cpp_program():
    """tests/cpp/test_real_space.cpp: the reference's global-assembly check written in
    C++ against bri17::RealSpaceOperator (c2c and r2c paths, CG)."""
    exe = os.path.join(CPP, "test_real_space")
    if not os.path.exists(exe):
        out = subprocess.run(["make", "-C", CPP, "test_real_space"], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count(" OK") == 2
