"""Multi-rank host logic on CPU (gloo): the N>1 path of bench.py shards k0
slabs across ranks with no data-path collective."""
import os
import socket
import subprocess
import sys

import pytest

from bri17_b200 import slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_properties():
    for n0 in (1, 2, 3, 11, 512, 1000):
        for world in (1, 2, 3, 4, 8):
            r = [slab.slab_range(n0, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n0
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert slab.rank_block((512, 512, 512), 3, 8) == ((192, 0, 0), (64, 512, 512))
    assert slab.rank_block((4096, 4096), 1, 2) == ((2048, 0), (2048, 4096))
    with pytest.raises(ValueError):
        slab.slab_range(8, 8, 8)
    assert slab.max_over_ranks(1.5) == 1.5          # no process group: identity


@pytest.mark.parametrize("world", [2, 3])
def test_world_size_n_gloo(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300,
                         env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"MULTIRANK_OK {world}" in out.stdout
