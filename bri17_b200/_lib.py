"""ctypes binding of the C ABI declared in include/bri17_b200.h.

There is deliberately no fallback: if the shared library is missing or a call
fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libbri17_b200.so")

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NCCL, ERR_UNSUPPORTED, ERR_BREAKDOWN = range(6)

_i32p = C.POINTER(C.c_int)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/bri17_b200.h one to one
SIGNATURES = {
    "bri17_last_error": (C.c_char_p, []),
    "bri17_version": (C.c_int, []),
    "bri17_set_last_error": (None, [C.c_char_p]),
    "bri17_plan_create": (C.c_int, [C.POINTER(_vp), C.c_int, _i32p, _f64p, C.c_double, C.c_double, C.c_int]),
    "bri17_plan_destroy": (C.c_int, [_vp]),
    "bri17_plan_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "bri17_plan_get_info": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int64)]),
    "bri17_plan_get_tables": (C.c_int, [_vp, C.c_int, _f64p, _f64p, _f64p, _f64p, _f64p]),
    "bri17_modal_stiffness_mode_f64": (C.c_int, [_vp, _i32p, _f64p]),
    "bri17_modal_strain_displacement_mode_f64": (C.c_int, [_vp, _i32p, _f64p]),
    "bri17_modal_stiffness_apply_f64": (C.c_int, [_vp, _vp, _vp, _i32p, _i32p, C.c_int64, C.c_double, _vp]),
    "bri17_modal_stiffness_apply_host_f64": (C.c_int, [_vp, _vp, _vp, _i32p, _i32p, C.c_int64, C.c_double]),
    "bri17_modal_stiffness_field_f64": (C.c_int, [_vp, _vp, _i32p, _i32p, _vp]),
    "bri17_modal_strain_displacement_field_f64": (C.c_int, [_vp, _vp, _i32p, _i32p, _vp]),
    "bri17_strain_displacement_apply_f64": (C.c_int, [_vp, _vp, _vp, _i32p, _i32p, C.c_int64, C.c_int64, C.c_double, _vp]),
    "bri17_freq_index_map": (C.c_int, [_vp, _vp, _i32p, _i32p, _vp]),
    "bri17_debug_walk_tiles": (C.c_int, [_vp, _i32p, _i32p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64),
                                         C.c_int, _i32p]),
    "bri17_modal_eigenstress_to_opposite_strain_mode_f64": (C.c_int, [_vp, _i32p, _f64p, _f64p]),
    "bri17_modal_stiffness_solve_f64": (C.c_int, [_vp, _vp, _vp, _i32p, _i32p, C.c_int64, C.c_int64, _vp]),
    "bri17_eigenstress_to_displacement_f64": (C.c_int, [_vp, _vp, _vp, _i32p, _i32p, C.c_int64, C.c_int64,
                                                        C.c_int64, C.c_int64, _vp]),
    "bri17_eigenstress_to_opposite_strain_f64": (C.c_int, [_vp, _vp, _vp, _i32p, _i32p, C.c_int64,
                                                           C.c_int64, _vp]),
    "bri17_eigenstress_to_force_f64": (C.c_int, [_vp, _vp, _vp, _i32p, _i32p, C.c_int64, C.c_int64,
                                                 C.c_int64, C.c_int64, _vp]),
    "bri17_modal_stiffness_apply_dot_f64": (C.c_int, [_vp, _vp, _vp, _i32p, _i32p, C.c_int64, C.c_double,
                                                      C.c_int, _vp, _vp, C.c_int, _vp]),
}

# include/bri17_b200_realspace.h (libbri17_b200_rs.so)
RS_LIB_PATH = os.path.join(HERE, "lib", "libbri17_b200_rs.so")
RS_SIGNATURES = {
    "bri17_rs_unique_id": (C.c_int, [_vp]),
    "bri17_rs_plan_create": (C.c_int, [C.POINTER(_vp), C.c_int, _i32p, _f64p, C.c_double, C.c_double,
                                       C.c_int, C.c_int, C.c_int, _vp, C.c_int]),
    "bri17_rs_plan_destroy": (C.c_int, [_vp]),
    "bri17_rs_plan_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "bri17_rs_plan_local": (C.c_int, [_vp, _i32p, _i32p, _i32p, _i32p]),
    "bri17_rs_plan_real_count": (C.c_int64, [_vp]),
    "bri17_rs_plan_fourier_count": (C.c_int64, [_vp]),
    "bri17_real_space_apply_f64": (C.c_int, [_vp, _vp, _vp, _vp]),
    "bri17_rs_forward_fft_f64": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "bri17_rs_inverse_fft_f64": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_double, _vp]),
    "bri17_rs_plan_modal": (_vp, [_vp]),
    "bri17_rs_plan_last_timings": (C.c_int, [_vp, _f64p, C.c_int]),
    "bri17_rs_plan_exchange_bytes": (C.c_int64, [_vp, C.c_int]),
    "bri17_cg_solve_f64": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_int, C.c_int, _i32p, _f64p, _vp]),
    "bri17_real_space_apply_real_f64": (C.c_int, [_vp, _vp, _vp, _vp]),
    "bri17_cg_solve_real_f64": (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_int, C.c_int, _i32p, _f64p, _vp]),
    "bri17_rs_plan_get_info": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int64)]),
    "bri17_real_space_apply_dot_f64": (C.c_int, [_vp, _vp, _vp, C.c_int, _f64p, _vp]),
    "bri17_debug_axis0_fused_host": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                               _f64p, _f64p, _f64p, C.c_double, C.c_double, C.c_double,
                                               C.c_int, C.c_int, _vp, _f64p]),
    "bri17_debug_exchange_host": (C.c_int, [C.c_int, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
}

_lib = None
_rs_lib = None


def _preload_wheel_libs() -> None:
    """libbri17_b200_rs.so needs libnccl.so.2 and libcufft.so.11.  When the PyTorch wheels are
    installed, load THEIR copies first (what `import torch` would do): the dynamic loader keeps
    one library per soname, and torch does not import against the older system NCCL."""
    import importlib.util
    for pkg, name in (("nvidia.nccl", "libnccl.so.2"), ("nvidia.cufft", "libcufft.so.11")):
        try:
            spec = importlib.util.find_spec(pkg)
        except (ImportError, ValueError):
            spec = None
        for base in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
            path = os.path.join(base, "lib", name)
            if os.path.exists(path):
                try:
                    C.CDLL(path, mode=C.RTLD_GLOBAL)
                except OSError:
                    pass
                break


def load_rs() -> C.CDLL:
    """Load libbri17_b200_rs.so (cuFFT + NCCL layer); no fallback."""
    global _rs_lib
    if _rs_lib is not None:
        return _rs_lib
    load()
    _preload_wheel_libs()
    if not os.path.exists(RS_LIB_PATH):
        raise RuntimeError(f"{RS_LIB_PATH} is missing: build it with `python -m bri17_b200.build`")
    lib = C.CDLL(RS_LIB_PATH)
    for name, (restype, argtypes) in RS_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _rs_lib = lib
    return lib


def load() -> C.CDLL:
    """Load libbri17_b200.so and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m bri17_b200.build` "
            "(nvcc, sm_100a). bri17_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class Bri17Error(RuntimeError):
    pass


def check(rc: int) -> None:
    """Translate a status code the way the C++ wrapper does (bri17.hpp errors):
    invalid argument -> ValueError (std::invalid_argument), else RuntimeError."""
    if rc == OK:
        return
    msg = load().bri17_last_error().decode()
    if rc == ERR_INVALID_ARG:
        raise ValueError(msg)
    raise Bri17Error(f"[bri17_b200 error {rc}] {msg}")
