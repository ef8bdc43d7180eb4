"""Drop-in module name of the reference's Python binding (python/pybri17.cpp:98-107).

``import pybri17`` gives the same four classes with the same constructors,
attributes and method names as the reference's pybind11 module, implemented on
top of the C ABI of libbri17_b200.so (see bri17_b200/__init__.py), plus the
batched operators that replace per-mode Python loops.
"""
from bri17_b200 import (CartesianGrid2f64, CartesianGrid3f64, Hooke2f64, Hooke3f64,  # noqa: F401
                        ModalOperator, __version__)

__author__ = "bri17-b200 contributors (API after S. Brisard's pybri17)"
__all__ = ["CartesianGrid2f64", "CartesianGrid3f64", "Hooke2f64", "Hooke3f64", "ModalOperator"]
