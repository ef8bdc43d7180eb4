#!/usr/bin/env python
"""The periodic inclusion problem in 3-D, solved two ways on 1..8 GPUs (BASELINE config 5).

An eigenstress patch [0, N/8)^3 with a unit last Mandel component (the 3-D version of the reference's
python/demo.py:11-23) loads a periodic, homogeneous body.  The nodal displacement is

  * the solution of  K u = b,  b = (|h|/|N|) iDFT(tau^ . conj(B^))   (bri17.hpp:340, theory.rst:151-157)
    by matrix-free CG on the real-space operator  F = (|h|/|N|) iDFT(K^ DFT(u))  -- what a heterogeneous
    problem would have to do, and the workload of BASELINE config 5;
  * the one-pass direct solve  u^ = K^-1 (tau^ . conj(B^))  per frequency (bri17.hpp:341), the map the
    reference's demo evaluates mode by mode in a Python loop.

Both must agree; the script prints the iteration count, the time per iteration and the difference.

    python examples/inclusion_cg.py [--edge 256]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/inclusion_cg.py --edge 512                                         # slabs over 8 GPUs
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bri17_b200 as b  # noqa: E402
from bri17_b200.realspace import RealSpaceOperator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--edge", type=int, default=256)
ap.add_argument("--rtol", type=float, default=1e-8)
ap.add_argument("--max-iter", type=int, default=20000)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank() if world > 1 else 0

shape, L, mu, nu = (args.edge,) * 3, (1.0, 1.0, 1.0), 1.0, 0.3      # python/demo.py:9,13-14
op = (RealSpaceOperator.from_process_group(shape, L, mu, nu, device=local) if world > 1
      else RealSpaceOperator(shape, L, mu, nu, device=local))
modal = b.ModalOperator(shape, L, mu, nu, device=local)

# eigenstress on this rank's n0 slab: unit last Mandel component inside the patch
patch = max(1, args.edge // 8)
tau = torch.zeros((6, op.n0_count) + shape[1:], dtype=torch.complex128, device=dev)
lo, hi = op.n0_begin, min(op.n0_begin + op.n0_count, patch)
if hi > lo:
    tau[-1, :hi - lo, :patch, :patch] = 1.0
tau_hat = op.forward_fft(tau)                                      # Fourier side: this rank's k1 slab
del tau
kb = (0, op.k1_begin, 0)
f_hat = modal.eigenstress_to_force(tau_hat, k_begin=kb)            # tau^ . conj(B^), every mode at once
u_hat = modal.eigenstress_to_displacement(tau_hat, k_begin=kb)     # K^-1 of it, every mode at once
del tau_hat
h_vol = float(np.prod([l / n for l, n in zip(L, shape)]))
b_real = op.inverse_fft(f_hat, scale=h_vol / float(np.prod(shape, dtype=np.float64))).real.contiguous()
u_direct = op.inverse_fft(u_hat).real.contiguous()
del f_hat, u_hat

op.cg_solve_real(b_real, rtol=0.0, max_iter=2, check_every=0)      # allocates the work vectors
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
x, iterations, residual = op.cg_solve_real(b_real, rtol=args.rtol, max_iter=args.max_iter, check_every=25)
e1.record()
torch.cuda.synchronize()

stats = torch.tensor([float((x - u_direct).abs().max()) if x.numel() else 0.0,
                      float(u_direct.abs().max()) if x.numel() else 0.0, e0.elapsed_time(e1)],
                     dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
err, scale, ms = stats.tolist()
if rank == 0:
    print(f"{args.edge}^3 on {world} GPU(s): CG converged to {residual:.2e} in {iterations} iterations, "
          f"{ms / max(iterations, 1):.3f} ms per iteration ({iterations / (ms * 1e-3):.1f} iterations/s); "
          f"max |u_CG - u_direct| / max |u_direct| = {err / scale:.2e}")
op.close()
if world > 1:
    dist.destroy_process_group()
