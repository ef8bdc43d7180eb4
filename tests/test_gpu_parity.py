"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on identical input bytes.

Gates:
  * frequency index map: bit-exact (int32 equality);
  * f^ = K^ u^: the kernels use the host-built tables and unfused fp64
    arithmetic in the reference's order, so the result is expected to be
    bit-identical to the oracle's parity build; the asserted bound is the
    north-star one, max_k |f_gpu - f_ref|_inf / |f_ref|_inf <= 1e-12 per mode;
  * f^(0) = 0 exactly;
  * the four dense-matrix known-answer tests of the reference through the GPU
    operator at the reference tolerance 1e-15*|e| + 1e-14.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import bri17_b200 as b  # noqa: E402
from oracle import kat  # noqa: E402

MU, NU = 5.6, 0.3
TOL = 1e-12   # BASELINE.json north_star: 1e-12 relative error per mode (max-norm)


def spacing_L(shape):
    h = (1.1, 1.2, 1.3)
    return tuple(float(n) * h[d] for d, n in enumerate(shape))


def per_mode_rel_err(f, ref):
    """max over modes of max_c|f - ref| / max_c|ref| (modes with ref == 0 must match exactly)."""
    num = np.abs(f - ref).max(axis=0)
    den = np.abs(ref).max(axis=0)
    zero = den == 0
    assert np.all(num[zero] == 0), "non-zero output where the reference is exactly zero"
    return float((num[~zero] / den[~zero]).max()) if (~zero).any() else 0.0


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


SHAPES = [
    (3, 4), (3, 4, 5),                      # the reference's own test grids
    (64, 64),                               # BASELINE config 1
    (16, 12, 10), (8, 8, 33), (1, 1, 1), (1, 1), (2, 3, 1),
    (5, 7, 600), (3, 2, 1025),              # rows longer than one tile, ragged tails
    (4, 2000), (300, 5), (2, 4097),
    (40, 24, 16),                           # more tiles than CTAs in a wave? (rows=960)
]


@pytest.mark.parametrize("shape", SHAPES)
def test_freq_index_map_bit_exact(oracle_mod, shape):
    op = b.ModalOperator(shape, spacing_L(shape), MU, NU)
    k = op.freq_index_map().cpu().numpy()
    assert k.dtype == np.int32
    assert np.array_equal(k, oracle_mod.freq_index_map((0,) * len(shape), shape))


def test_freq_index_map_slabs_bit_exact(oracle_mod):
    shape = (37, 11, 130)
    op = b.ModalOperator(shape, spacing_L(shape), MU, NU)
    for kb, local in (((5, 0, 0), (9, 11, 130)), ((36, 10, 129), (1, 1, 1)),
                      ((0, 3, 7), (37, 4, 100)), ((30, 0, 0), (7, 11, 130))):
        k = op.freq_index_map(local, kb).cpu().numpy()
        assert np.array_equal(k, oracle_mod.freq_index_map(kb, local))


@pytest.mark.parametrize("shape", SHAPES)
def test_apply_matches_oracle(oracle_mod, shape):
    dim = len(shape)
    L = spacing_L(shape)
    o = oracle_mod.best()
    u = oracle_mod.synthetic_u_hat(dim, shape, seed=len(shape) * 1000 + shape[-1])
    ref = o.apply_modal_stiffness(shape, L, MU, NU, u)
    op = b.ModalOperator(shape, L, MU, NU)
    f = op.apply_modal_stiffness(to_dev(u)).cpu().numpy()
    assert per_mode_rel_err(f, ref) <= TOL
    assert np.all(f[(slice(None),) + (0,) * dim] == 0)          # f^(0) = 0 exactly
    assert np.array_equal(f, ref), "expected bit-identical output (unfused fp64, host tables)"


@pytest.mark.parametrize("mapping", [1, 2])
@pytest.mark.parametrize("shape", SHAPES + [(6, 5, 513), (3, 257), (7, 3, 300)])
def test_row_and_flat_mappings_match_oracle(oracle_mod, shape, mapping):
    """Both tile mappings (row tiles with register-cached columns; flat tiles for
    ragged rows such as a 513-wide half spectrum) on every shape: same bits, and
    the index map exposes the mapping in use."""
    dim = len(shape)
    L = spacing_L(shape)
    u = oracle_mod.synthetic_u_hat(dim, shape, seed=123)
    ref = oracle_mod.best().apply_modal_stiffness(shape, L, MU, NU, u)
    op = b.ModalOperator(shape, L, MU, NU)
    op.set_option("mapping", mapping)
    f = op.apply_modal_stiffness(to_dev(u)).cpu().numpy()
    assert op.info("last_flat") == (mapping == 2)
    assert np.array_equal(f, ref)
    k = op.freq_index_map().cpu().numpy()
    assert np.array_equal(k, oracle_mod.freq_index_map((0,) * dim, shape))
    # a slab with offsets on every axis
    if all(n >= 2 for n in shape):
        kb = tuple(n // 2 for n in shape)
        local = tuple(n - n // 2 for n in shape)
        sl = (slice(None),) + tuple(slice(a, a + n) for a, n in zip(kb, local))
        fs = op.apply_modal_stiffness(to_dev(u[sl]), k_begin=kb).cpu().numpy()
        assert np.array_equal(fs, ref[sl])


@pytest.mark.parametrize("dim", [2, 3])
def test_every_kernel_variant_matches_oracle(oracle_mod, dim):
    shape = (6, 1100) if dim == 2 else (6, 5, 1100)
    L = spacing_L(shape)
    u = oracle_mod.synthetic_u_hat(dim, shape, seed=77)
    ref = oracle_mod.best().apply_modal_stiffness(shape, L, MU, NU, u)
    op = b.ModalOperator(shape, L, MU, NU)
    ud = to_dev(u)
    op.set_option("mapping", 1)        # the variants are row-tile kernels
    for v in range(op.info("num_variants")):
        op.set_option("apply_variant", v)
        out = torch.full_like(ud, float("nan"))
        f = op.apply_modal_stiffness(ud, out=out).cpu().numpy()
        assert np.array_equal(f, ref), f"variant {v}"


def test_golden_reference_vectors(oracle_mod, golden):
    """Committed outputs of the compiled reference header (tests/golden)."""
    z, meta = golden
    for n, m in enumerate(meta):
        shape = tuple(m["shape"])
        u = oracle_mod.synthetic_u_hat(m["dim"], shape, m["seed"])
        op = b.ModalOperator(shape, m["L"], m["mu"], m["nu"])
        f = op.apply_modal_stiffness(to_dev(u)).cpu().numpy()
        assert per_mode_rel_err(f, z[f"f_hat_{n}"]) <= TOL
        e = op.apply_strain_displacement(to_dev(u)).cpu().numpy()
        assert per_mode_rel_err(e, z[f"eps_hat_{n}"]) <= TOL
        if f"K_{n}" in z:
            K = op.modal_stiffness_field().cpu().numpy().reshape(-1, m["dim"], m["dim"])
            Kg = z[f"K_{n}"]
            assert np.abs(K - Kg).max() <= TOL * np.abs(Kg).max()
            B = op.modal_strain_displacement_field().cpu().numpy().reshape(-1, m["dim"])
            Bg = z[f"B_{n}"]
            assert np.abs(B - Bg).max() <= TOL * np.abs(Bg).max()


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_known_answer_tests_through_gpu(dim):
    """tests/test_bri17.cpp:335-361, :363-536, :538-559, :561-605 with the GPU
    operator in the middle of the FFT sandwich."""
    E = kat.load_elements()
    shape, L = kat.SHAPE[dim], kat.grid_L(dim)
    op = b.ModalOperator(shape, L, kat.MU, kat.NU)
    K, max_imag = kat.actual_stiffness(
        shape, L, lambda u: op.apply_modal_stiffness(to_dev(u)).cpu().numpy())
    assert max_imag <= kat.IMAG_TOL
    kat.assert_equal(kat.assemble_expected_stiffness(shape, E[f"Ke{dim}"]), K)
    B, max_imag = kat.actual_strain_displacement(
        shape, L, lambda u: op.apply_strain_displacement(to_dev(u)).cpu().numpy())
    assert max_imag <= kat.IMAG_TOL
    kat.assert_equal(kat.assemble_expected_strain_displacement(shape, E[f"Be{dim}"]), B)


def test_slab_blocks_and_strides(oracle_mod):
    """k_begin offsets (multi-GPU slabs) and the raw C-ABI stride argument."""
    import ctypes as C
    shape = (12, 9, 70)
    L = spacing_L(shape)
    u = oracle_mod.synthetic_u_hat(3, shape, seed=5)
    ref = oracle_mod.best().apply_modal_stiffness(shape, L, MU, NU, u)
    op = b.ModalOperator(shape, L, MU, NU)
    for a0, a1 in ((0, 3), (3, 4), (4, 12)):
        f = op.apply_modal_stiffness(to_dev(u[:, a0:a1]), k_begin=(a0, 0, 0)).cpu().numpy()
        assert np.array_equal(f, ref[:, a0:a1])
    # slab addressed inside the full planar buffer: comp_stride = full grid size
    ud = to_dev(u)
    fd = torch.zeros_like(ud)
    lib, plan = op._lib, op._plan
    kb = (C.c_int * 3)(5, 0, 0)
    loc = (C.c_int * 3)(4, 9, 70)
    off = 5 * 9 * 70 * 16
    rc = lib.bri17_modal_stiffness_apply_f64(plan, ud.data_ptr() + off, fd.data_ptr() + off, kb, loc,
                                             12 * 9 * 70, 1.0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    f = fd.cpu().numpy()
    assert np.array_equal(f[:, 5:9], ref[:, 5:9])
    assert np.all(f[:, :5] == 0) and np.all(f[:, 9:] == 0)      # nothing outside the slab is touched


def test_in_place_and_out_scale(oracle_mod):
    shape = (10, 6, 40)
    L = spacing_L(shape)
    u = oracle_mod.synthetic_u_hat(3, shape, seed=9)
    ref = oracle_mod.best().apply_modal_stiffness(shape, L, MU, NU, u)
    op = b.ModalOperator(shape, L, MU, NU)
    ud = to_dev(u)
    op.apply_modal_stiffness(ud, out=ud)                          # input and output alias
    assert np.array_equal(ud.cpu().numpy(), ref)
    scale = float(np.prod(L) / np.prod(shape) / np.prod(shape))   # |h|/|N|, tests/test_bri17.cpp:98
    f = op.apply_modal_stiffness(to_dev(u), out_scale=scale).cpu().numpy()
    assert np.array_equal(f, ref * scale)


def test_host_buffer_path_matches_device_path(oracle_mod):
    shape = (23, 16, 48)
    L = spacing_L(shape)
    u = oracle_mod.synthetic_u_hat(3, shape, seed=21)
    ref = oracle_mod.best().apply_modal_stiffness(shape, L, MU, NU, u)
    op = b.ModalOperator(shape, L, MU, NU)
    for rows in (0, 1, 5, 23, 100):
        op.set_option("host_chunk_rows", rows)
        assert np.array_equal(op.apply_modal_stiffness_host(u), ref)
    pinned = torch.from_numpy(u).pin_memory()
    out = torch.empty_like(pinned).pin_memory()
    op.set_option("host_chunk_rows", 4)
    op.apply_modal_stiffness_host(pinned, out=out, k_begin=None)
    assert np.array_equal(out.numpy(), ref)
    shape2 = (50, 300)
    u2 = oracle_mod.synthetic_u_hat(2, shape2, seed=22)
    op2 = b.ModalOperator(shape2, spacing_L(shape2), MU, NU)
    op2.set_option("host_chunk_rows", 7)
    assert np.array_equal(op2.apply_modal_stiffness_host(u2),
                          oracle_mod.best().apply_modal_stiffness(shape2, spacing_L(shape2), MU, NU, u2))


@pytest.mark.parametrize("shape", [(3, 4), (3, 4, 5), (16, 12, 10), (6, 700), (4, 3, 520)])
def test_strain_and_field_kernels(oracle_mod, shape):
    dim = len(shape)
    L = spacing_L(shape)
    o = oracle_mod.best()
    op = b.ModalOperator(shape, L, MU, NU)
    ks = np.stack(np.unravel_index(np.arange(int(np.prod(shape))), shape), axis=1)
    K = op.modal_stiffness_field().cpu().numpy().reshape(-1, dim, dim)
    Kref = np.stack([o.modal_stiffness(shape, L, MU, NU, k) for k in ks])
    assert np.array_equal(K, Kref)                      # table-driven: bit-identical
    B = op.modal_strain_displacement_field().cpu().numpy().reshape(-1, dim)
    Bref = np.stack([o.modal_strain_displacement(shape, L, k) for k in ks])
    # the prefactor uses the device sincos (<= 2 ulp) on the reference's argument
    assert np.abs(B - Bref).max() <= TOL * np.abs(Bref).max()
    u = oracle_mod.synthetic_u_hat(dim, shape, seed=31)
    e = op.apply_strain_displacement(to_dev(u)).cpu().numpy()
    eref = o.apply_strain_displacement(shape, L, u)
    assert per_mode_rel_err(e, eref) <= TOL


def test_argument_errors():
    op = b.ModalOperator((4, 4, 4), (1., 1., 1.), MU, NU)
    u = torch.zeros((3, 4, 4, 4), dtype=torch.complex128, device="cuda")
    with pytest.raises(ValueError):      # block outside the grid
        op.apply_modal_stiffness(u, k_begin=(1, 0, 0))
    with pytest.raises(ValueError):      # wrong dtype
        op.apply_modal_stiffness(u.to(torch.complex64))
    with pytest.raises(ValueError):      # wrong component count
        op.apply_modal_stiffness(u[:2])
    with pytest.raises(ValueError):      # host tensor into the device entry point
        op.apply_modal_stiffness(u.cpu())
    import ctypes as C
    rc = op._lib.bri17_modal_stiffness_apply_f64(op._plan, u.data_ptr() + 8, u.data_ptr(), None, None,
                                                 0, 1.0, None)
    assert rc == 1 and b"16-byte aligned" in op._lib.bri17_last_error()


def test_large_2d_4096_sampled_rows(oracle_mod):
    """BASELINE config 2 (2-D 4096^2): full GPU apply, oracle on sampled k0 rows."""
    shape = (4096, 4096)
    L = spacing_L(shape)
    op = b.ModalOperator(shape, L, MU, NU)
    rng = np.random.default_rng(2)
    u = torch.from_numpy(rng.standard_normal((2,) + shape + (2,))).cuda().view(torch.float64)
    u = torch.view_as_complex(u.reshape((2,) + shape + (2,)))
    f = op.apply_modal_stiffness(u)
    assert bool((f[:, 0, 0] == 0).all())
    o = oracle_mod.best()
    worst = 0.0
    for k0 in (0, 1, 2, 1023, 2047, 2048, 2049, 4094, 4095):
        ref = o.apply_modal_stiffness(shape, L, MU, NU, u[:, k0:k0 + 1].cpu().numpy(), k_begin=(k0, 0))
        worst = max(worst, per_mode_rel_err(f[:, k0:k0 + 1].cpu().numpy(), ref))
    assert worst <= TOL


def test_large_3d_512_sampled_slabs(oracle_mod):
    """BASELINE config 3 (3-D 512^3, 12 GiB of fields): full GPU apply, oracle
    on sampled k0 planes (the operator is per-mode independent, so this is an
    exact check of those planes), plus size-independent properties."""
    shape = (512, 512, 512)
    L = spacing_L(shape)
    op = b.ModalOperator(shape, L, MU, NU)
    g = torch.Generator(device="cuda").manual_seed(3)
    u = torch.view_as_complex(torch.randn((3,) + shape + (2,), dtype=torch.float64, device="cuda",
                                          generator=g))
    f = op.apply_modal_stiffness(u)
    assert bool((f[:, 0, 0, 0] == 0).all())
    o = oracle_mod.best()
    worst = 0.0
    for k0 in (0, 1, 255, 256, 257, 511):
        ref = o.apply_modal_stiffness(shape, L, MU, NU, u[:, k0:k0 + 1].cpu().numpy(), k_begin=(k0, 0, 0))
        worst = max(worst, per_mode_rel_err(f[:, k0:k0 + 1].cpu().numpy(), ref))
    assert worst <= TOL
    # K^ is real symmetric: <v, K u> = <K v, u> (Hermitian form), checked on a slab
    v = torch.view_as_complex(torch.randn((3, 8, 512, 512, 2), dtype=torch.float64, device="cuda",
                                          generator=g))
    Kv = op.apply_modal_stiffness(v, k_begin=(100, 0, 0))
    us, fs = u[:, 100:108], f[:, 100:108]
    lhs = torch.sum(torch.conj(v) * fs)
    rhs = torch.sum(torch.conj(Kv) * us)
    assert abs(complex(lhs - rhs)) <= 1e-11 * abs(complex(lhs))
    # index map on 512^3 slabs: a k0 slab (modal apply at N>1) and a k1 slab (the Fourier-side
    # block of the distributed real-space apply), bit-exact against the row-major loop nest
    for kb, loc in (((504, 0, 0), (8, 512, 512)), ((0, 448, 0), (16, 64, 512)), ((100, 200, 0), (3, 5, 257))):
        k = op.freq_index_map(loc, kb).cpu().numpy()
        assert np.array_equal(k, oracle_mod.freq_index_map(kb, loc))
    # slabs reproduce the full-grid result bit for bit (what each GPU computes at N>1)
    fs2 = op.apply_modal_stiffness(u[:, 448:512].contiguous(), k_begin=(448, 0, 0))
    assert torch.equal(torch.view_as_real(fs2), torch.view_as_real(f[:, 448:512]))


# ---------------------------------------------------------------------------
# round 2: inclusion right-hand side, Parseval sum, reentrancy, 64-bit offsets
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(6, 8), (6, 8, 5), (5, 4, 300), (3, 700)])
def test_eigenstress_to_force_vs_definition(oracle_mod, shape):
    """f^ = tau^ . conj(B^) (bri17.hpp:324-332, :340) per mode against numpy on the oracle's B^."""
    dim = len(shape)
    nsym = dim * (dim + 1) // 2
    L = spacing_L(shape)
    o = oracle_mod.best()
    rng = np.random.default_rng(12)
    tau = rng.standard_normal((nsym,) + shape) + 1j * rng.standard_normal((nsym,) + shape)
    op = b.ModalOperator(shape, L, MU, NU)
    f = op.eigenstress_to_force(to_dev(tau)).cpu().numpy()
    pairs = [(0, 0), (1, 1), (0, 1)] if dim == 2 else [(0, 0), (1, 1), (2, 2), (1, 2), (2, 0), (0, 1)]
    ref = np.zeros((dim,) + shape, dtype=complex)
    for k in np.ndindex(*shape):
        if not any(k):
            continue                                        # :336-339: zero at the null frequency
        sl = (slice(None),) + k
        B = o.modal_strain_displacement(shape, L, k)
        T = np.zeros((dim, dim), dtype=complex)
        for s, (p, q) in enumerate(pairs):
            T[p, q] = T[q, p] = tau[sl][s] if p == q else tau[sl][s] / np.sqrt(2.0)
        ref[sl] = T @ B.conj()
    scale = np.abs(ref).max(axis=0)
    err = np.abs(f - ref).max(axis=0)
    assert np.all(f[(slice(None),) + (0,) * dim] == 0)
    assert (err <= 1e-12 * np.maximum(scale, np.abs(tau).max(axis=0) * 1e-3)).all()
    # and K^-1 of it is the displacement of the one-pass direct solve
    u1 = op.solve_modal_stiffness(op.eigenstress_to_force(to_dev(tau))).cpu().numpy()
    u2 = op.eigenstress_to_displacement(to_dev(tau)).cpu().numpy()
    assert np.abs(u1 - u2).max() <= 1e-11 * np.abs(u2).max()


@pytest.mark.parametrize("shape,herm", [((16, 12, 10), 0), ((8, 8, 33), 64), ((5, 7, 600), 0), ((64, 33), 64),
                                        ((3, 4, 5), 8), ((2, 4097), 0)])
def test_apply_with_parseval_sum(oracle_mod, shape, herm):
    """bri17_modal_stiffness_apply_dot_f64: same f^ bit for bit, plus sum_k w_k Re(u^H f^)."""
    dim = len(shape)
    full = shape[:-1] + ((herm or shape[-1]),)             # the grid; the block keeps k_last < shape[-1]
    L = spacing_L(full)
    u = oracle_mod.synthetic_u_hat(dim, shape, seed=5)
    op = b.ModalOperator(full, L, MU, NU)
    ud = to_dev(u)
    f_plain = op.apply_modal_stiffness(ud, out_scale=0.25)
    f, dot = op.apply_modal_stiffness_dot(ud, out_scale=0.25, hermitian_n=herm)
    assert torch.equal(f, f_plain)
    fn = f.cpu().numpy()
    w = np.ones(shape[-1])
    if herm:
        k = np.arange(shape[-1])
        w = np.where((k == 0) | (2 * k == herm), 1.0, 2.0)
    expect = float(np.sum(w * np.real(np.conj(u) * fn)))
    assert abs(dot - expect) <= 1e-12 * abs(expect)
    _, dot2 = op.apply_modal_stiffness_dot(ud, out_scale=0.25, hermitian_n=herm)
    assert dot2 == dot                                      # deterministic summation order


def test_plan_is_reentrant_two_threads(oracle_mod):
    """plans are immutable after creation -> two host threads launch on ONE plan
    concurrently (own streams, own buffers); results stay bit-identical to the oracle.  The
    host-buffer entry point shares one staging set per plan and serialises internally."""
    import threading
    shape = (24, 20, 130)
    L = spacing_L(shape)
    op = b.ModalOperator(shape, L, MU, NU)
    o = oracle_mod.best()
    us = [oracle_mod.synthetic_u_hat(3, shape, seed=70 + t) for t in range(2)]
    refs = [o.apply_modal_stiffness(shape, L, MU, NU, u) for u in us]
    errors = []

    def device_worker(t):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                ud = to_dev(us[t])
                out = torch.empty_like(ud)
                for _ in range(200):
                    op.apply_modal_stiffness(ud, out=out, stream=s)
                s.synchronize()
                if not np.array_equal(out.cpu().numpy(), refs[t]):
                    errors.append(f"device thread {t}: result differs")
        except Exception as e:       # noqa: BLE001
            errors.append(repr(e))

    def host_worker(t):
        try:
            for _ in range(10):
                f = op.apply_modal_stiffness_host(us[t])
                if not np.array_equal(f, refs[t]):
                    errors.append(f"host thread {t}: result differs")
        except Exception as e:       # noqa: BLE001
            errors.append(repr(e))

    for worker in (device_worker, host_worker):
        threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
    assert not errors, errors
    assert op.info("launches") >= 400


def _free_gib():
    free, _ = torch.cuda.mem_get_info()
    return free / 2**30


def _check_planes(oracle_mod, op, shape, L, u_planes, f, planes):
    """Sampled k0 planes of an (in-place) result against the compiled reference."""
    o = oracle_mod.best()
    dim = len(shape)
    for a, u in zip(planes, u_planes):
        ref = o.apply_modal_stiffness(shape, L, MU, NU, u, k_begin=(a,) + (0,) * (dim - 1))
        got = f[:, a:a + 1].cpu().numpy()
        assert per_mode_rel_err(got, ref) <= TOL, a
        assert np.array_equal(got, ref), a


def test_1024_cubed_in_place_offsets_beyond_2_31(oracle_mod):
    """BASELINE config 4's grid on ONE GPU, in place (48 GiB): element offsets of component 2
    exceed 2^31 (the reference's `int` indexing overflows there, tests/test_bri17.cpp:83,87);
    both mappings; index map of slabs at the far end; sampled planes against oracle/_ref."""
    if _free_gib() < 60:
        pytest.skip("needs 60 GiB of free device memory")
    shape = (1024, 1024, 1024)
    L = spacing_L(shape)
    op = b.ModalOperator(shape, L, MU, NU)
    planes = [0, 511, 1023]
    for mapping in (1, 2):
        op.set_option("mapping", mapping)
        g = torch.Generator(device="cuda").manual_seed(1024 + mapping)
        u = torch.view_as_complex(torch.randn((3,) + shape + (2,), dtype=torch.float64, device="cuda",
                                              generator=g))
        keep = [u[:, a:a + 1].cpu().numpy() for a in planes]
        op.apply_modal_stiffness(u, out=u)                      # in place
        assert op.info("last_flat") == (mapping == 2)
        _check_planes(oracle_mod, op, shape, L, keep, u, planes)
        for kb, local in (((1020, 0, 0), (4, 1024, 1024)), ((1023, 1000, 0), (1, 24, 1024))):
            k = op.freq_index_map(local, kb).cpu().numpy()
            assert np.array_equal(k, oracle_mod.freq_index_map(kb, local))
        del u
        torch.cuda.empty_cache()


@pytest.mark.parametrize("shape", [(2049, 1024, 1024), (65536, 32769)])
def test_block_larger_than_2_31_modes(oracle_mod, shape):
    """More than 2^31 - 1 modes in ONE block: the flat mapping's 64-bit index branch
    (flat_index) and 64-bit tile counters, in place (96 / 64 GiB).  Sampled planes against
    oracle/_ref and the index map across the 2^31 boundary."""
    dim = len(shape)
    modes = int(np.prod(shape, dtype=np.int64))
    assert modes > 2**31
    need = modes * 16 * dim / 2**30
    if _free_gib() < need + 30:
        pytest.skip(f"needs {need + 30:.0f} GiB of free device memory")
    L = spacing_L(shape)
    op = b.ModalOperator(shape, L, MU, NU)
    op.set_option("mapping", 2)
    g = torch.Generator(device="cuda").manual_seed(31)
    u = torch.view_as_complex(torch.randn((dim,) + shape + (2,), dtype=torch.float64, device="cuda",
                                          generator=g))
    planes = [0, shape[0] // 2, shape[0] - 1]
    keep = [u[:, a:a + 1].cpu().numpy() for a in planes]
    op.apply_modal_stiffness(u, out=u)
    assert op.info("last_flat") == 1
    _check_planes(oracle_mod, op, shape, L, keep, u, planes)
    del u
    torch.cuda.empty_cache()
    # index map of the whole block (int32 [modes][dim]): windows at both ends and across 2^31
    k = op.freq_index_map()
    plane = modes // shape[0]
    for a in (0, 2**31 // plane - 1, shape[0] - 2):
        local = (2,) + shape[1:]
        kb = (a,) + (0,) * (dim - 1)
        got = k[a * plane:(a + 2) * plane].cpu().numpy()
        assert np.array_equal(got, oracle_mod.freq_index_map(kb, local)), a
    del k
    torch.cuda.empty_cache()


def test_field_writers_large(oracle_mod):
    """Mode-major B^ / K^ fields (smem-staged writers) on rows that span several tiles and on
    rows shorter than a tile; per-mode comparison with the oracle."""
    for shape in ((4, 3, 700), (6, 5, 33), (5, 1000), (7, 3)):
        dim = len(shape)
        L = spacing_L(shape)
        o = oracle_mod.best()
        op = b.ModalOperator(shape, L, MU, NU)
        Bf = op.modal_strain_displacement_field().cpu().numpy()
        Kf = op.modal_stiffness_field().cpu().numpy()
        for k in list(np.ndindex(*shape))[::7]:
            Bref = o.modal_strain_displacement(shape, L, k)
            assert np.abs(Bf[k] - Bref).max() <= 1e-12 * max(np.abs(Bref).max(), 1e-300) + 1e-15
            assert np.array_equal(Kf[k], o.modal_stiffness(shape, L, MU, NU, k))
