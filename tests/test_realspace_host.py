"""CPU checks of the real-space C ABI that need no device: argument validation
happens before any CUDA call, and the exchange-size helper is pure geometry."""
import ctypes as C

from bri17_b200 import _lib


def test_rs_plan_argument_validation():
    lib = _lib.load_rs()
    plan = C.c_void_p()
    shape = (C.c_int * 3)(8, 8, 8)
    L = (C.c_double * 3)(1., 1., 1.)
    cases = [
        (dict(dim=4), b"dim must be 2 or 3"),
        (dict(nranks=17), b"at most 16 ranks"),
        (dict(rank=2, nranks=2), b"bad rank"),
        (dict(nranks=2), b"nccl_unique_id is NULL"),
        (dict(mode=3), b"exchange_mode must be 0 or 1"),
    ]
    for kw, msg in cases:
        a = dict(dim=3, rank=0, nranks=1, mode=0)
        a.update(kw)
        rc = lib.bri17_rs_plan_create(C.byref(plan), a["dim"], shape, L, 1.0, 0.3, 0, a["rank"], a["nranks"],
                                      None, a["mode"])
        assert rc == _lib.ERR_INVALID_ARG and msg in _lib.load().bri17_last_error(), (kw, _lib.load().bri17_last_error())
    bad = (C.c_int * 3)(8, 0, 8)
    assert lib.bri17_rs_plan_create(C.byref(plan), 3, bad, L, 1.0, 0.3, 0, 0, 1, None, 0) == _lib.ERR_INVALID_ARG
    assert lib.bri17_rs_plan_destroy(None) == 0
    buf = (C.c_double * 8)()
    assert lib.bri17_rs_plan_last_timings(None, buf, 8) == _lib.ERR_INVALID_ARG


def _slabs(n, P):
    return [(q * n) // P for q in range(P + 1)]


def _exchange(shape, P, real, k1_major, nchunks, direction, blocks):
    """bri17_debug_exchange_host on numpy blocks; returns the output blocks of every virtual rank."""
    import numpy as np
    lib = _lib.load_rs()
    dim = len(shape)
    N0, S1 = shape[0], (shape[1] if dim == 3 else (shape[1] // 2 + 1 if real else shape[1]))
    S2e = 1 if dim == 2 else (shape[2] // 2 + 1 if real else shape[2])
    n0b, k1b = _slabs(N0, P), _slabs(S1, P)
    outs = []
    for q in range(P):
        if direction == 0:
            n1q = k1b[q + 1] - k1b[q]
            shp = (dim, n1q, N0, S2e) if (k1_major and dim == 3) else (dim, N0, n1q, S2e)
        else:
            shp = (dim, n0b[q + 1] - n0b[q], S1, S2e)
        outs.append(np.full(shp, np.nan + 0j, dtype=np.complex128))
    ins = [np.ascontiguousarray(b) for b in blocks]
    pin = (C.c_void_p * P)(*[a.ctypes.data for a in ins])
    pout = (C.c_void_p * P)(*[a.ctypes.data for a in outs])
    rc = lib.bri17_debug_exchange_host(dim, (C.c_int * dim)(*shape), P, int(real), int(k1_major), nchunks,
                                       direction, pin, pout)
    assert rc == 0, _lib.load().bri17_last_error()
    return outs


def test_exchange_index_arithmetic_on_virtual_ranks():
    """The copy plans of the fused all-to-all (exchange_forward / exchange_backward in realspace.cu: the
    ones the GPU kernels execute) replayed on the CPU for 1..8 virtual ranks: uneven and empty slabs,
    complex and half-spectrum layouts, natural and k1-major Fourier-side layouts, 1..4 sub-slabs per
    component.  Forward must be the axis-0-slab -> axis-1-slab transposition of a global array, backward its
    inverse; every output element must be written exactly once (NaN canaries)."""
    import numpy as np
    rng = np.random.default_rng(3)
    for shape in ((8, 8, 6), (9, 7, 5), (3, 4, 5), (12, 10), (5, 16)):
        dim = len(shape)
        for real in (False, True):
            S1 = shape[1] if dim == 3 else (shape[1] // 2 + 1 if real else shape[1])
            S2e = 1 if dim == 2 else (shape[2] // 2 + 1 if real else shape[2])
            A = rng.standard_normal((dim, shape[0], S1, S2e)) + 1j * rng.standard_normal((dim, shape[0], S1, S2e))
            for P in (1, 2, 3, 4, 8):
                n0b, k1b = _slabs(shape[0], P), _slabs(S1, P)
                T = [A[:, n0b[r]:n0b[r + 1]] for r in range(P)]
                for k1_major in ((0, 1) if dim == 3 else (0,)):
                    for nchunks in (1, 3, 4):
                        X = _exchange(shape, P, real, k1_major, nchunks, 0, T)
                        for q in range(P):
                            want = A[:, :, k1b[q]:k1b[q + 1]]
                            if k1_major:
                                want = want.transpose(0, 2, 1, 3)
                            assert np.array_equal(X[q], want), (shape, real, P, k1_major, nchunks, q)
                        D = _exchange(shape, P, real, k1_major, nchunks, 1, X)
                        for r in range(P):
                            assert np.array_equal(D[r], T[r]), (shape, real, P, k1_major, nchunks, r)
