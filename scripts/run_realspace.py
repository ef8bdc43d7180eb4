#!/usr/bin/env python
"""A few single-GPU real-space applies (for ncu captures and quick timings).

    python scripts/run_realspace.py [--edge 512] [--applies 3] [--real] [--no-fused] [--cg 0]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bri17_b200.realspace import RealSpaceOperator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--edge", type=int, default=512)
ap.add_argument("--shape", type=str, default="", help="N0,N1,N2 (overrides --edge)")
ap.add_argument("--applies", type=int, default=3)
ap.add_argument("--real", action="store_true")
ap.add_argument("--no-fused", action="store_true")
ap.add_argument("--no-k1-major", action="store_true")
ap.add_argument("--k1-major", action="store_true")
ap.add_argument("--fft-chunk-mib", type=int, default=-1, help="chunk of the local 2-D transforms (0 = whole slab)")
ap.add_argument("--cg", type=int, default=0, help="also time this many CG iterations")
args = ap.parse_args()

shape = tuple(int(v) for v in args.shape.split(",")) if args.shape else (args.edge,) * 3
L = tuple(n * h for n, h in zip(shape, (1.1, 1.2, 1.3)))
op = RealSpaceOperator(shape, L, 5.6, 0.3)
if args.no_fused:
    op.set_option("fused_axis0", 0)
if args.no_k1_major:
    op.set_option("k1_major", 0)
if args.k1_major:
    op.set_option("k1_major", 1)
if args.fft_chunk_mib >= 0:
    op.set_option("fft_chunk_mib", args.fft_chunk_mib)
if args.real:
    u = torch.randn(op.real_shape, dtype=torch.float64, device="cuda")
    fn = op.apply_real
else:
    u = torch.view_as_complex(torch.randn(op.real_shape + (2,), dtype=torch.float64, device="cuda"))
    fn = op.apply
F = torch.empty_like(u)
for _ in range(2):
    fn(u, out=F)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.applies):
    fn(u, out=F)
e1.record()
torch.cuda.synchronize()
out = {"shape": list(shape), "real": args.real, "fused_axis0": bool(op.info("fused_axis0")),
       "k1_major": bool(op.info("k1_major_real" if args.real else "k1_major")),
       "fft_chunk_mib": op.info("fft_chunk_mib"), "fft_chunk_planes": op.info("fft_chunk_planes"),
       "ms_per_apply": e0.elapsed_time(e1) / args.applies, "phases_ms": op.timings()}
if args.cg:
    b = fn(u).clone()
    cg = op.cg_solve_real if args.real else op.cg_solve
    cg(b, rtol=0.0, max_iter=2, check_every=0)
    torch.cuda.synchronize()
    e0.record()
    _, its, res = cg(b, rtol=0.0, max_iter=args.cg, check_every=0)
    e1.record()
    torch.cuda.synchronize()
    out["cg_ms_per_iteration"] = e0.elapsed_time(e1) / max(its, 1)
print(json.dumps(out))
