#!/bin/bash
# One-shot 8-GPU session: modal apply bench (both arms), real-space apply and CG.
N=${1:-8}
bash scripts/gpu_multi.sh $N
bash scripts/gpu_realspace_bench.sh $N
nproc > gpurun_out/nproc_n$N.txt
