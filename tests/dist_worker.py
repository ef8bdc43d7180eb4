"""World-size-N worker for tests/test_multi_rank.py (gloo, CPU).

Exercises the host logic of the multi-GPU path: slab partition, k_begin
offsets, max-over-ranks timing reduction, and that slabs computed
independently (here by the CPU oracle standing in for the kernel, which needs
a GPU) assemble into the full-grid result with no data-path collective."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bri17_b200 import slab  # noqa: E402
from oracle import oracle    # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    shape, L, mu, nu = (11, 6, 7), (12.1, 7.2, 9.1), 5.6, 0.3   # 11 planes: uneven split
    kb, local = slab.rank_block(shape, rank, world)
    u_full = oracle.synthetic_u_hat(3, shape, seed=99)           # same bytes on every rank
    a0, a1 = kb[0], kb[0] + local[0]
    f_loc = oracle.port().apply_modal_stiffness(shape, L, mu, nu, u_full[:, a0:a1], k_begin=kb)
    k_loc = oracle.freq_index_map(kb, local)
    parts = [None] * world
    dist.all_gather_object(parts, (a0, a1, f_loc, k_loc))
    t = slab.max_over_ranks(10.0 + rank)
    ok = True
    if rank == 0:
        ranges = sorted((p[0], p[1]) for p in parts)
        ok &= ranges[0][0] == 0 and ranges[-1][1] == shape[0]
        ok &= all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        ok &= max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1
        f = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])], axis=1)
        k = np.concatenate([p[3] for p in sorted(parts, key=lambda p: p[0])], axis=0)
        full = oracle.port().apply_modal_stiffness(shape, L, mu, nu, u_full)
        ok &= np.array_equal(f, full)
        ok &= np.array_equal(k, oracle.freq_index_map((0, 0, 0), shape))
        ok &= t == 10.0 + world - 1
        print("MULTIRANK_OK" if ok else "MULTIRANK_FAIL", world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
