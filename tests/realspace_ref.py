"""numpy restatement of the real-space operator of the reference harness
(tests/test_bri17.cpp:56-107) on top of the CPU oracle -- test helper."""
import numpy as np


def real_space_apply_ref(oracle_impl, shape, L, mu, nu, u):
    """F = (|h|/|N|) * iDFT_unnormalised(K^ DFT(u)); u: (dim, *shape) complex."""
    dim = len(shape)
    axes = tuple(range(1, dim + 1))
    u_hat = np.fft.fftn(u, axes=axes)                                   # :57
    f_hat = oracle_impl.apply_modal_stiffness(shape, L, mu, nu, u_hat)  # :58-92
    cell_volume = float(np.prod([l / n for l, n in zip(L, shape)]))     # :96
    return np.fft.ifftn(f_hat, axes=axes) * cell_volume                 # :95-106 (ifftn carries 1/|N|)


def direct_solve_ref(oracle_impl, shape, L, mu, nu, b):
    """Zero-mean solution of A x = b through the per-mode inverse (k != 0)."""
    dim = len(shape)
    axes = tuple(range(1, dim + 1))
    cell_volume = float(np.prod([l / n for l, n in zip(L, shape)]))
    b_hat = np.fft.fftn(b, axes=axes) / cell_volume
    x_hat = np.zeros_like(b_hat)
    for k in np.ndindex(*shape):
        if not any(k):
            continue
        K = oracle_impl.modal_stiffness(shape, L, mu, nu, k).real
        x_hat[(slice(None),) + k] = np.linalg.solve(K, b_hat[(slice(None),) + k])
    return np.fft.ifftn(x_hat, axes=axes)
