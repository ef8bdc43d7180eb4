"""CPU tests of the host side: the C-ABI library loads and exports what
include/bri17_b200.h declares, the Python mirror of the reference interface
behaves like the reference (names, argument checks, errors), and the per-mode
host API reproduces the oracle.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

import bri17_b200 as b
from bri17_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            src = open(os.path.join(inc, fn)).read()
            names += re.findall(r"BRI17_API\s+[\w \*]+?\b(bri17_\w+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert len(declared) >= 15
    core = ctypes.CDLL(_lib.LIB_PATH)
    rs = ctypes.CDLL(_lib.RS_LIB_PATH)          # needs the core library loaded first
    for name in declared:
        assert hasattr(core, name) or hasattr(rs, name), \
            f"{name} declared in include/*.h but not exported"
    # and the ctypes tables cover the headers one to one
    assert sorted(list(_lib.SIGNATURES) + list(_lib.RS_SIGNATURES)) == declared
    assert _lib.load().bri17_version() == 100
    _lib.load_rs()


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_grid_mirror(oracle_mod):
    g2 = b.CartesianGrid2f64((3, 4), (3.3, 4.8))
    g3 = b.CartesianGrid3f64((3, 4, 5), (3.3, 4.8, 6.5))
    assert (g2.dim, g2.size, g2.shape, g2.L, g2.num_nodes_per_cell) == (2, 12, (3, 4), (3.3, 4.8), 4)
    assert (g3.dim, g3.size, g3.num_nodes_per_cell) == (3, 60, 8)
    assert g3.dtype == np.float64
    assert g2.get_node_at(2, 3) == 11 and g3.get_node_at(1, 2, 3) == (1 * 4 + 2) * 5 + 3
    with pytest.raises(TypeError):
        g2.get_node_at(1, 2, 3)          # static_assert in the reference (bri17.hpp:79)
    p = oracle_mod.port()
    for g in (g2, g3):
        for cell in range(g.size):
            assert g.get_cell_nodes(cell) == list(p.get_cell_nodes(g.shape, cell))
    assert repr(g3) == "CartesianGrid<d,3>{shape={3,4,5,},L={3.3,4.8,6.5,}}"


def test_repr_matches_compiled_reference(oracle_mod):
    r = oracle_mod.ref()
    if r is None:
        pytest.skip("oracle/_ref not built")
    fn = r.lib.ref_repr
    fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double),
                   ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    buf = ctypes.create_string_buffer(512)
    for shape, L in (((3, 4), (3.3, 4.8)), ((64, 8, 2), (1.0, 0.5, 1e-3))):
        dim = len(shape)
        grid = (b.CartesianGrid2f64 if dim == 2 else b.CartesianGrid3f64)(shape, L)
        hooke = (b.Hooke2f64 if dim == 2 else b.Hooke3f64)(5.6, 0.3, grid)
        sh = (ctypes.c_int * dim)(*shape)
        ll = (ctypes.c_double * dim)(*L)
        fn(dim, sh, ll, 5.6, 0.3, 0, buf, 512)
        assert repr(grid) == buf.value.decode()
        fn(dim, sh, ll, 5.6, 0.3, 1, buf, 512)
        assert repr(hooke) == buf.value.decode()


@pytest.mark.parametrize("dim,shape,L", [(2, (3, 4), (3.3, 4.8)), (3, (3, 4, 5), (3.3, 4.8, 6.5)),
                                          (2, (64, 64), (70.4, 76.8)), (3, (6, 5, 33), (1., 1., 1.))])
def test_per_mode_api_bitwise_vs_oracle(oracle_mod, dim, shape, L):
    """Hooke{2,3}f64.modal_stiffness_matrix / modal_strain_displacement
    (python/pybri17.cpp:76-87) against the oracle, every mode, bit for bit."""
    o = oracle_mod.best()
    grid = (b.CartesianGrid2f64 if dim == 2 else b.CartesianGrid3f64)(shape, L)
    hooke = (b.Hooke2f64 if dim == 2 else b.Hooke3f64)(5.6, 0.3, grid)
    K = np.empty(dim * dim, dtype=np.complex128)
    B = np.empty(dim, dtype=np.complex128)
    for k in np.ndindex(*shape):
        kk = np.array(k, dtype=np.intc)
        hooke.modal_stiffness_matrix(kk, K)
        hooke.modal_strain_displacement(kk, B)
        assert np.array_equal(K.reshape(dim, dim), o.modal_stiffness(shape, L, 5.6, 0.3, k)), k
        assert np.array_equal(B, o.modal_strain_displacement(shape, L, k)), k


def test_per_mode_api_vs_golden(golden):
    z, meta = golden
    for n, m in enumerate(meta):
        if f"K_{n}" not in z:
            continue
        dim, shape = m["dim"], tuple(m["shape"])
        grid = (b.CartesianGrid2f64 if dim == 2 else b.CartesianGrid3f64)(shape, m["L"])
        hooke = (b.Hooke2f64 if dim == 2 else b.Hooke3f64)(m["mu"], m["nu"], grid)
        K = np.empty(dim * dim, dtype=np.complex128)
        B = np.empty(dim, dtype=np.complex128)
        Kall, Ball = z[f"K_{n}"], z[f"B_{n}"]
        for i, k in enumerate(np.ndindex(*shape)):
            kk = np.array(k, dtype=np.intc)
            hooke.modal_stiffness_matrix(kk, K)
            hooke.modal_strain_displacement(kk, B)
            Kg, Bg = Kall[i], Ball[i]
            assert np.abs(K.reshape(dim, dim) - Kg).max() <= 1e-12 * max(np.abs(Kg).max(), 1e-300)
            assert np.abs(B - Bg).max() <= 1e-12 * max(np.abs(Bg).max(), 1e-300) + 1e-15


def test_tables_follow_reference_formulas():
    """bri17.hpp:259-263, :218-221 evaluated with numpy's libm-backed functions."""
    grid = b.CartesianGrid3f64((8, 12, 10), (8.8, 14.4, 13.0))
    hooke = b.Hooke3f64(5.6, 0.3, grid)
    for axis, (n, L) in enumerate(zip(grid.shape, grid.L)):
        t = hooke.tables(axis)
        k = np.arange(n)
        h = L / n
        beta = 2 * np.pi * k / n
        assert np.allclose(t["phi"], 2 * (1 - np.cos(beta)) / h / h, rtol=1e-14, atol=1e-15)
        assert np.allclose(t["chi"], (2 + np.cos(beta)) / 3, rtol=1e-15)
        assert np.allclose(t["psi"], np.sin(beta) / h, rtol=1e-14, atol=1e-15)
        assert np.allclose(t["c"], np.cos(np.pi * k / n), rtol=1e-15, atol=1e-16)
        assert np.allclose(t["s"], np.sin(np.pi * k / n) * n / L, rtol=1e-15, atol=1e-16)
        assert t["phi"][0] == 0.0 and t["psi"][0] == 0.0      # K^(0) = 0 exactly


def test_error_behaviour():
    grid = b.CartesianGrid2f64((4, 4), (1., 1.))
    hooke = b.Hooke2f64(1.0, 0.3, grid)
    K = np.empty(4, dtype=np.complex128)
    k = np.zeros(2, dtype=np.intc)
    # python/pybri17.cpp:9-15: non-contiguous / not 1-D -> invalid_argument
    with pytest.raises(ValueError):
        hooke.modal_stiffness_matrix(k, np.empty(8, dtype=np.complex128)[::2])
    with pytest.raises(ValueError):
        hooke.modal_stiffness_matrix(k, K.reshape(2, 2))
    # the reference never bound-checks k (bri17.hpp:247) and K^ is N-periodic in k: indices outside
    # [0, N) are accepted and wrapped; any integer sequence converts like pybind11's array_t<int>
    K2 = np.empty(4, dtype=np.complex128)
    hooke.modal_stiffness_matrix(np.array([1, 3], dtype=np.intc), K)
    hooke.modal_stiffness_matrix([5, -1], K2)
    assert np.array_equal(K, K2)
    hooke.modal_stiffness_matrix(np.array([1, 3], dtype=np.int64), K2)
    assert np.array_equal(K, K2)
    with pytest.raises(ValueError):          # k must be one-dimensional, of length dim
        hooke.modal_stiffness_matrix(np.zeros((2, 1), dtype=np.intc), K)
    with pytest.raises(ValueError):
        hooke.modal_stiffness_matrix(np.array([0.5, 1.0]), K)
    with pytest.raises(ValueError):          # B^ uses half angles, not N-periodic: outside [0, N) is refused
        hooke.modal_strain_displacement(np.array([0, 4], dtype=np.intc), np.empty(2, dtype=np.complex128))
    with pytest.raises(TypeError):           # 3-D grid into a 2-D Hooke (template mismatch)
        b.Hooke2f64(1.0, 0.3, b.CartesianGrid3f64((2, 2, 2), (1., 1., 1.)))
    lib = _lib.load()
    plan = ctypes.c_void_p()
    shape = (ctypes.c_int * 4)(2, 2, 2, 2)
    L = (ctypes.c_double * 4)(1, 1, 1, 1)
    assert lib.bri17_plan_create(ctypes.byref(plan), 4, shape, L, 1.0, 0.3, -1) == _lib.ERR_INVALID_ARG
    assert b"dim must be 2 or 3" in lib.bri17_last_error()       # bri17.hpp:35-36
    bad = (ctypes.c_int * 2)(4, 0)
    assert lib.bri17_plan_create(ctypes.byref(plan), 2, bad, L, 1.0, 0.3, -1) == _lib.ERR_INVALID_ARG


def test_no_cpu_fallback_for_whole_grid_operators():
    """A plan without a device refuses every whole-grid entry point."""
    lib = _lib.load()
    grid = b.CartesianGrid2f64((4, 4), (1., 1.))
    hooke = b.Hooke2f64(1.0, 0.3, grid)          # HOST_ONLY plan
    u = np.zeros((2, 4, 4), dtype=np.complex128)
    rc = lib.bri17_modal_stiffness_apply_host_f64(hooke._plan, u.ctypes.data, u.ctypes.data,
                                                  None, None, 0, 1.0)
    assert rc == _lib.ERR_CUDA
    assert b"no CPU fallback" in lib.bri17_last_error()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under bri17_b200/ or include/
    may reference it."""
    for base in ("bri17_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".h", ".hpp", ".cpp")) or fn == "Makefile":
                    src = open(os.path.join(dirpath, fn), errors="ignore").read()
                    hit = re.search(r"(from|import)\s+oracle|liboracle|oracle/|_ref/|libbri17_ref", src)
                    assert hit is None, (os.path.join(dirpath, fn), hit.group(0))


def test_pybri17_module_name():
    """`import pybri17` exposes the reference's Python names (python/pybri17.cpp:98-107)."""
    import pybri17
    for name in ("CartesianGrid2f64", "CartesianGrid3f64", "Hooke2f64", "Hooke3f64"):
        assert hasattr(pybri17, name)
    grid = pybri17.CartesianGrid2f64((4, 4), (1.0, 1.0))
    hooke = pybri17.Hooke2f64(1.0, 0.3, grid)
    assert (hooke.mu, hooke.nu, hooke.grid) == (1.0, 0.3, grid) and grid.dim == 2
    for method in ("modal_strain_displacement", "modal_stiffness_matrix",
                   "modal_eigenstress_to_opposite_strain"):
        assert callable(getattr(hooke, method))
    assert pybri17.__version__ == "0.1"
