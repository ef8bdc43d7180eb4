#!/usr/bin/env python
"""bench_realspace.py -- end-to-end real-space apply and CG (BASELINE configs[3], configs[4]).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        bench_realspace.py --edge 1024 --mode 1 [--cg-iters 20]

Reported separately from bench.py (the headline metric is the modal apply).
A step = one F = (|h|/|N|) iDFT(K^ DFT(u)) on an edge^3 grid sharded in n0
slabs over the N GPUs (tests/test_bri17.cpp:56-107 of the reference); timed
with CUDA events on the launch stream, max over ranks; per-phase times come
from events recorded inside the library.  Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MU, NU, SPACING = 5.6, 0.3, (1.1, 1.2, 1.3)
PEER_GBS = 770.0   # measured peer copy bandwidth per direction (B200_PROFILING.md)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=512)
    ap.add_argument("--mode", type=int, default=1, help="0 = NCCL send/recv + pack, 1 = fused peer-store kernel")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cg-iters", type=int, default=0, help="also time this many CG iterations")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="serialise the phases (gives the per-phase breakdown; default overlaps exchange and FFTs)")
    ap.add_argument("--real", action="store_true",
                    help="real float64 fields through the r2c half-spectrum path (default: complex, like the reference)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from bri17_b200 import slab
    from bri17_b200.realspace import PHASES, RealSpaceOperator

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shape = (args.edge,) * 3
    L = tuple(n * h for n, h in zip(shape, SPACING))
    op = RealSpaceOperator.from_process_group(shape, L, MU, NU, device=local_rank, exchange_mode=args.mode)
    if args.no_pipeline:
        op.set_option("pipeline", 0)
    pipelined = world > 1 and args.mode == 1 and not args.no_pipeline
    gen = torch.Generator(device=dev).manual_seed(4000 + rank)
    if args.real:
        u = torch.randn(op.real_shape, dtype=torch.float64, device=dev, generator=gen)
    else:
        u = torch.zeros(op.real_shape + (2,), dtype=torch.float64, device=dev)
        u[..., 0].normal_(generator=gen)                   # real field carried as complex (imag = 0)
        u = torch.view_as_complex(u)
    F = torch.empty_like(u)
    apply = op.apply_real if args.real else op.apply
    cg = op.cg_solve_real if args.real else op.cg_solve
    stream = torch.cuda.current_stream()
    rdev = dev if world > 1 else None

    for _ in range(args.warmup):
        apply(u, out=F)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        apply(u, out=F)
    e1.record(stream)
    barrier()
    ms = slab.max_over_ranks(e0.elapsed_time(e1) / args.steps, rdev)
    phases = {k: slab.max_over_ranks(v, rdev) for k, v in op.timings().items()}
    modes = args.edge ** 3
    xbytes = op.exchange_bytes_real if args.real else op.exchange_bytes
    spectrum = (args.edge // 2 + 1) / args.edge if args.real else 1.0
    line = {
        "metric": "real-space apply (FFT -> modal K -> iFFT), applies/s", "value": 1e3 / ms, "unit": "applies/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "effective_gmodes_per_s": modes / (ms * 1e-3) / 1e9, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D Q8 {args.edge}^3 real-space apply, n0 slabs over {world} GPU(s)",
                   "fields": "real float64, r2c half spectrum" if args.real else "complex128 (c2c, as the reference)",
                   "exchange": ["nccl send/recv + pack/unpack kernels", "fused peer-store kernel (CUDA IPC over NVLink)"][args.mode]
                   if world > 1 else "none (single GPU)",
                   "schedule": "exchange of component c overlapped with the FFTs of the other components (2 streams)"
                   if pipelined else "phases serialised"},
        "phases_ms_last_apply_max_over_ranks": phases,
        "pipelined": pipelined,
        "exchange": None if (world == 1 or pipelined) else {
            "bytes_sent_per_gpu_per_direction": xbytes,
            "fwd_gbs_per_gpu": xbytes / (phases["exchange_fwd"] * 1e-3) / 1e9,
            "bwd_gbs_per_gpu": xbytes / (phases["exchange_bwd"] * 1e-3) / 1e9,
            "frac_of_measured_peer_copy_770": xbytes / (phases["exchange_fwd"] * 1e-3) / 1e9 / PEER_GBS},
        "modal_gbs": 96 * modes * spectrum / world / (phases["modal"] * 1e-3) / 1e9,
    }

    if args.cg_iters > 0:
        # periodic inclusion-like right-hand side: b = A(u0) with zero-mean u0 (in the range of A)
        b = apply(u - 0.0).clone()
        barrier()
        cg(b, rtol=0.0, max_iter=2, check_every=0)                      # warm-up, allocates work vectors
        barrier()
        e0.record(stream)
        x, iters, res = cg(b, rtol=0.0, max_iter=args.cg_iters, check_every=0)
        e1.record(stream)
        barrier()
        cg_ms = slab.max_over_ranks(e0.elapsed_time(e1), rdev)
        line["cg"] = {"iterations": iters, "ms_total": cg_ms, "iterations_per_s": iters / (cg_ms * 1e-3),
                      "rel_residual_after": res,
                      "per_iteration": "1 real-space apply + fused (x,r update + <r,r>) + <p,Ap> + direction update; "
                                       "scalars stay on device, 2 one-element all-reduces"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    op.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
