#!/bin/bash
# Round 2, short 8-GPU session: 1024^3 real-space apply for each sub-slab count, parity in the same run.
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    scripts/run_realspace_dist.py > gpurun_out/r02_realspace_dist_n${N}.log 2>&1; echo "rc=$?"
grep "^{" gpurun_out/r02_realspace_dist_n${N}.log | tail -c 6000
grep -i "error\|Traceback" gpurun_out/r02_realspace_dist_n${N}.log | head -5
