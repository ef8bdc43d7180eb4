"""Frequency-slab decomposition for multi-GPU runs.

The modal apply is independent per frequency, so rank ``g`` of ``P`` simply
owns the slab ``k0 in [start_g, stop_g)`` of the slowest axis and runs the same
kernel with ``k_begin[0] = start_g``: no collective is needed on the data path.
In the planar layout a slab is ``dim`` contiguous chunks of
``(stop_g-start_g) * prod(shape[1:])`` complex numbers.
"""
from __future__ import annotations


def slab_range(n0: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of ``range(n0)``: sizes differ by at most 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (rank * n0) // world, ((rank + 1) * n0) // world


def rank_block(shape, rank: int, world: int):
    """``(k_begin, local_shape)`` of the block owned by ``rank``."""
    start, stop = slab_range(int(shape[0]), rank, world)
    k_begin = (start,) + (0,) * (len(shape) - 1)
    local = (stop - start,) + tuple(int(n) for n in shape[1:])
    return k_begin, local


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar over the default process group (identity
    when torch.distributed is not initialised): multi-GPU times are reported
    as the slowest rank's."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
