#!/usr/bin/env python
"""The eigenstress-patch ("inclusion") problem of the reference's demo, batched.

The reference solves it with a Python double loop that calls
``Hooke2f64.modal_eigenstress_to_opposite_strain`` once per frequency
(python/demo.py:33-40 of bri17: 65 536 interpreter crossings on a 256x256 grid).
Here the same per-mode map runs once over the whole spectrum on the GPU
(``ModalOperator.eigenstress_to_opposite_strain``), with the mode-major layout
``tau[k0, k1, sym]`` that the reference demo uses, and a handful of modes are
cross-checked against the per-mode host API (same names as pybri17).

    python examples/demo_inclusion.py [N]          # needs a CUDA device
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pybri17  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dim, sym = 2, 3
shape, L = (N, N), (1.0, 1.0)
mu, nu = 1.0, 0.3
grid = pybri17.CartesianGrid2f64(shape, L)
hooke = pybri17.Hooke2f64(mu, nu, grid)
print(grid, hooke, sep="\n", end="")

# eigenstress: unit shear (last Mandel component) inside the patch [0, N/8)^2, zero outside
patch = N // 8
tau = torch.zeros(shape + (sym,), dtype=torch.complex128, device="cuda")
tau[:patch, :patch, -1] = 1.0
tau_hat = torch.fft.fftn(tau, dim=(0, 1)).contiguous()     # mode-major (k0, k1, sym), C order

op = pybri17.ModalOperator(shape, L, mu, nu)
torch.cuda.synchronize()
t0 = time.perf_counter()
eta_hat = op.eigenstress_to_opposite_strain(tau_hat, mode_major=True)      # every frequency at once
torch.cuda.synchronize()
dt = time.perf_counter() - t0
eps = -torch.fft.ifftn(eta_hat, dim=(0, 1)).real                            # cell-averaged strain field
print(f"{N}x{N}: {N * N} modes in {dt * 1e3:.3f} ms on the GPU; "
      f"mean strain {eps.mean(dim=(0, 1)).tolist()}, max |eps_xy| {float(eps[..., 2].abs().max()):.6f}")

# cross-check a few frequencies against the per-mode API (the reference's loop body)
eta_k = np.empty(sym, dtype=np.complex128)
worst = 0.0
for k in ((0, 0), (1, 0), (3, 5), (N // 2, N // 2), (N - 1, N - 1)):
    kk = np.array(k, dtype=np.intc)
    hooke.modal_eigenstress_to_opposite_strain(kk, tau_hat[k].cpu().numpy().copy(), eta_k)
    got = eta_hat[k].cpu().numpy()
    worst = max(worst, float(np.abs(got - eta_k).max() / max(np.abs(eta_k).max(), 1e-300)) if any(k) else
                float(np.abs(got).max()))
print(f"max relative difference to the per-mode host API on 5 modes: {worst:.2e}")
assert worst <= 1e-12
