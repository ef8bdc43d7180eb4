#!/usr/bin/env python
"""torchrun worker: the distributed real-space apply at full size, timed for each sub-slab count of the
pipelined schedule, with parity checks in the same run (checker code from tests/).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        scripts/run_realspace_dist.py [--edge 1024] [--applies 5]
Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import realspace_checks as rc  # noqa: E402
from bri17_b200 import slab  # noqa: E402
from bri17_b200.realspace import RealSpaceOperator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--edge", type=int, default=1024)
ap.add_argument("--applies", type=int, default=5)
args = ap.parse_args()

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
shape = (args.edge,) * 3
L = tuple(n * h for n, h in zip(shape, (1.1, 1.2, 1.3)))
out = {"world": world, "edge": args.edge}
out["small_grids_vs_numpy_restatement"] = rc.small_grids(local, 1, 1, 1)
op = RealSpaceOperator.from_process_group(shape, L, 5.6, 0.3, device=local, exchange_mode=1)
stream = torch.cuda.current_stream()
for real in (False, True):
    if real:
        u = torch.randn(op.real_shape, dtype=torch.float64, device=dev)
    else:
        u = torch.zeros(op.real_shape + (2,), dtype=torch.float64, device=dev)
        u[..., 0].normal_()
        u = torch.view_as_complex(u)
    F = torch.empty_like(u)
    fn = op.apply_real if real else op.apply
    for J, ctas in ((1, 592), (2, 592), (4, 592), (4, 296), (4, 148), (4, 96), (4, 64)):
        op.set_option("exchange_chunks", J)
        op.set_option("copy_ctas", ctas)       # grid cap of the exchange kernel (SMs left to the transforms)
        for _ in range(2):
            fn(u, out=F)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.applies):
            fn(u, out=F)
        e1.record(stream)
        dist.barrier()
        torch.cuda.synchronize()
        ms = slab.max_over_ranks(e0.elapsed_time(e1) / args.applies, dev)
        phases = {k: round(slab.max_over_ranks(v, dev), 3) for k, v in op.timings().items()}
        key = f"{'real' if real else 'complex'}_J{J}_ctas{ctas}"
        out[key] = {"ms_per_apply": ms, "phases_ms": phases}
    op.set_option("copy_ctas", 0)          # back to the default policy
    del u, F
    torch.cuda.empty_cache()
    op.set_option("exchange_chunks", 0)
    out[f"plane_waves_{'real' if real else 'complex'}"] = rc.plane_waves(op, real)
    torch.cuda.empty_cache()
op.close()
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
