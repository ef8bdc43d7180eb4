#!/usr/bin/env python
"""Timings of the secondary kernels on 512^3 (1 GPU): strain recovery (K3),
K^/B^ field writers, index map.  CUDA events, 20 launches after 3 warm-ups."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bri17_b200 as b  # noqa: E402

ONCE = "--once" in sys.argv                     # one launch per kernel (for ncu captures)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
edge = int(args[0]) if args else 512
shape = (edge,) * 3
L = tuple(n * h for n, h in zip(shape, (1.1, 1.2, 1.3)))
op = b.ModalOperator(shape, L, 5.6, 0.3)
M = edge ** 3
u = torch.view_as_complex(torch.randn((3,) + shape + (2,), dtype=torch.float64, device="cuda"))
eps = torch.empty((6,) + shape, dtype=torch.complex128, device="cuda")


def timed(fn, n=20):
    if ONCE:
        n = 1
    for _ in range(0 if ONCE else 3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
ms = timed(lambda: op.apply_strain_displacement(u, out=eps))
out["strain_apply"] = {"ms": ms, "bytes_per_mode": 144, "gbs": 144 * M / ms / 1e6}
del eps
tau = torch.view_as_complex(torch.randn((6,) + shape + (2,), dtype=torch.float64, device="cuda"))
ms = timed(lambda: op.eigenstress_to_opposite_strain(tau), 10)
out["eigenstress_to_opposite_strain"] = {"ms": ms, "bytes_per_mode": 192, "gbs": 192 * M / ms / 1e6}
for variant in (0, 1):      # 0: round-1 shapes (2 CTAs / SM, 128 registers); 1: one mode per thread, 3 CTAs / SM
    op.set_option("solve_variant", variant)
    ms = timed(lambda: op.eigenstress_to_displacement(tau), 10)
    out[f"eigenstress_to_displacement_v{variant}"] = {"ms": ms, "bytes_per_mode": 144, "gbs": 144 * M / ms / 1e6}
del tau
for variant in (0, 1):
    op.set_option("solve_variant", variant)
    ms = timed(lambda: op.solve_modal_stiffness(u), 10)
    out[f"modal_stiffness_solve_v{variant}"] = {"ms": ms, "bytes_per_mode": 96, "gbs": 96 * M / ms / 1e6}
slab = (edge, edge, edge)    # field writers on the whole grid (K^: 144 B/mode = 19.3 GB at 512^3)
ms = timed(lambda: op.modal_stiffness_field(slab, (0, 0, 0)), 10)
out["stiffness_field"] = {"ms": ms, "bytes_per_mode": 144, "gbs": 144 * M / ms / 1e6}
ms = timed(lambda: op.modal_strain_displacement_field(slab, (0, 0, 0)), 10)
out["strain_field"] = {"ms": ms, "bytes_per_mode": 48, "gbs": 48 * M / ms / 1e6}
ms = timed(lambda: op.freq_index_map(slab, (0, 0, 0)), 10)
out["index_map"] = {"ms": ms, "bytes_per_mode": 12, "gbs": 12 * M / ms / 1e6}
# ragged fastest axis: the 513-wide half spectrum of a 1024^3 r2c block, row tiles vs flat tiles
del u
rag = (256, 256, 513)
opr = b.ModalOperator((256, 256, 1024), (1.0, 1.0, 1.0), 5.6, 0.3)
ur = torch.view_as_complex(torch.randn((3,) + rag + (2,), dtype=torch.float64, device="cuda"))
fr = torch.empty_like(ur)
for name, mapping in (("rows", 1), ("flat", 2)):
    opr.set_option("mapping", mapping)
    ms = timed(lambda: opr.apply_modal_stiffness(ur, out=fr))
    out[f"apply_256x256x513_{name}"] = {"ms": ms, "gbs": 96 * ur[0].numel() / ms / 1e6}
opf = b.ModalOperator(shape, L, 5.6, 0.3)
uf = torch.view_as_complex(torch.randn((3,) + shape + (2,), dtype=torch.float64, device="cuda"))
ff = torch.empty_like(uf)
for name, mapping in (("rows", 1), ("flat", 2)):
    opf.set_option("mapping", mapping)
    ms = timed(lambda: opf.apply_modal_stiffness(uf, out=ff))
    out[f"apply_512^3_{name}"] = {"ms": ms, "gbs": 96 * M / ms / 1e6}
# strong-scaling shards of the headline kernel on ONE GPU: what a rank of N = 2, 4, 8 runs (fixed costs only)
for planes in (256, 128, 64):
    us, fs = uf[:, :planes].contiguous(), ff[:, :planes].contiguous()
    opf.set_option("mapping", 0)
    ms = timed(lambda: opf.apply_modal_stiffness(us, out=fs, k_begin=(192, 0, 0)), 50)
    out[f"apply_512^3_shard_{planes}_planes"] = {"ms": ms, "gbs": 96 * planes * edge * edge / ms / 1e6}
print(json.dumps(out))
