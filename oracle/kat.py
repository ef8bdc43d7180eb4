"""Known-answer tests of the reference, restated (TEST INFRASTRUCTURE ONLY).

Restates the "Global assembly tests" of ``tests/test_bri17.cpp:318-606``:
the dense global stiffness (resp. strain-displacement) matrix is built column
by column through the FFT sandwich and compared with the classical FE assembly
of a Maxima-derived element matrix.  FFTW (absent here) is replaced by
``numpy.fft``; the element matrices are read from ``tests/golden/kat_elements.json``
(parsed out of the reference test file by ``tests/golden/make_golden.py``).

The per-frequency operator is pluggable (``apply_K`` / ``apply_B`` callables)
so the same driver checks the CPU oracle *and* the CUDA path.
"""
from __future__ import annotations

import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden", "kat_elements.json")

# tests/test_bri17.cpp:320-333
MU, NU = 5.6, 0.3
SHAPE = {2: (3, 4), 3: (3, 4, 5)}
SPACING = {2: (1.1, 1.2), 3: (1.1, 1.2, 1.3)}
RTOL, ATOL = 1e-15, 1e-14          # :360, :535, :558, :604
IMAG_TOL = 1e-14                   # :140-144, :282


def grid_L(dim):
    # std::transform(shape, spacing, multiplies) :326-327, :332-333
    return tuple(float(n) * h for n, h in zip(SHAPE[dim], SPACING[dim]))


def load_elements():
    with open(GOLDEN) as f:
        g = json.load(f)
    return {k: np.array(v["values"], dtype=np.float64).reshape(v["shape"])
            for k, v in g["matrices"].items()}


def cell_nodes(shape, cell):
    """CartesianGrid::get_cell_nodes, bri17.hpp:127-158."""
    dim = len(shape)
    idx = np.unravel_index(cell, shape)
    nodes = []
    for local in range(1 << dim):
        # last axis fastest within the cell (bri17.hpp:134-137, :146-153)
        off = [(local >> (dim - 1 - d)) & 1 for d in range(dim)]
        ijk = [(idx[d] + off[d]) % shape[d] for d in range(dim)]
        nodes.append(int(np.ravel_multi_index(ijk, shape)))
    return nodes


def assemble_expected_stiffness(shape, Ke):
    """tests/test_bri17.cpp:153-174."""
    dim = len(shape)
    size = int(np.prod(shape))
    nn = 1 << dim
    K = np.zeros((size * dim, size * dim))
    for cell in range(size):
        nodes = cell_nodes(shape, cell)
        for ie in range(nn * dim):
            i = nodes[ie % nn] + size * (ie // nn)
            for je in range(nn * dim):
                j = nodes[je % nn] + size * (je // nn)
                K[i, j] += Ke[ie, je]
    return K


def assemble_expected_strain_displacement(shape, Be):
    """tests/test_bri17.cpp:295-316."""
    dim = len(shape)
    nsym = dim * (dim + 1) // 2
    size = int(np.prod(shape))
    nn = 1 << dim
    B = np.zeros((size * nsym, size * dim))
    for cell in range(size):
        nodes = cell_nodes(shape, cell)
        for il in range(nsym):
            i = il * size + cell
            for jl in range(nn * dim):
                j = nodes[jl % nn] + size * (jl // nn)
                B[i, j] += Be[il, jl]
    return B


def actual_stiffness(shape, L, apply_K):
    """StiffnessMatrixFactory::run + compute_Ku, tests/test_bri17.cpp:56-150.

    ``apply_K(u_hat[dim, *shape]) -> f_hat[dim, *shape]`` is the modal apply."""
    dim = len(shape)
    size = int(np.prod(shape))
    axes = tuple(range(1, dim + 1))
    cell_volume = 1.0
    for d in range(dim):
        cell_volume *= L[d] / shape[d]                 # :96
    correction = cell_volume / size                    # :98
    K = np.zeros((size * dim, size * dim))
    max_imag = 0.0
    u = np.zeros((dim,) + tuple(shape), dtype=np.complex128)
    for j in range(size * dim):
        u.ravel()[j] = 1.0                             # :136
        u_hat = np.fft.fftn(u, axes=axes)              # :57  FFTW_FORWARD
        f_hat = apply_K(u_hat)                         # :58-92
        # :95 FFTW_BACKWARD is unnormalised; numpy's ifftn carries 1/|N|
        Ku = np.fft.ifftn(f_hat, axes=axes) * size * correction
        max_imag = max(max_imag, float(np.abs(Ku.imag).max()))
        K[:, j] = Ku.real.ravel()                      # :145
        u.ravel()[j] = 0.0                             # :147
    return K, max_imag


def actual_strain_displacement(shape, L, apply_B):
    """StrainDisplacementMatrixFactory::run + compute_Bu, :194-292.

    ``apply_B(u_hat[dim, *shape]) -> eps_hat[nsym, *shape]``."""
    dim = len(shape)
    nsym = dim * (dim + 1) // 2
    size = int(np.prod(shape))
    axes = tuple(range(1, dim + 1))
    B = np.zeros((size * nsym, size * dim))
    max_imag = 0.0
    u = np.zeros((dim,) + tuple(shape), dtype=np.complex128)
    for j in range(size * dim):
        u.ravel()[j] = 1.0
        u_hat = np.fft.fftn(u, axes=axes)
        e_hat = apply_B(u_hat)
        Bu = np.fft.ifftn(e_hat, axes=axes)            # :236-244 (1/|N|)
        max_imag = max(max_imag, float(np.abs(Bu.imag).max()))
        B[:, j] = Bu.real.ravel()
        u.ravel()[j] = 0.0
    return B, max_imag


def assert_equal(expected, actual, rtol=RTOL, atol=ATOL):
    """tests/test_bri17.cpp:9-28: |a-e| <= rtol*|e| + atol entry-wise."""
    assert expected.shape == actual.shape
    err = np.abs(actual - expected)
    tol = rtol * np.abs(expected) + atol
    bad = np.argwhere(err > tol)
    if bad.size:
        i, j = bad[0]
        raise AssertionError(
            f"[{i}, {j}]: expected = {expected[i, j]!r}, actual = {actual[i, j]!r} "
            f"({len(bad)} entries out of tolerance, max err {err.max():.3e})")
    return float(err.max())
