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
