"""Build libbri17_b200.so (sm_100a) in-tree: ``python -m bri17_b200.build``."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libbri17_b200.so")


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile every CUDA source with nvcc for sm_100a; returns the .so path."""
    args = ["make", "-C", os.path.join(HERE, "csrc")]
    if force:
        args.append("-B")
    out = subprocess.run(args, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("libbri17_b200 build failed:\n" + out.stdout + out.stderr)
    if verbose:
        sys.stdout.write(out.stdout)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="-B" in sys.argv))
