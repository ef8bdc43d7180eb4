#!/bin/bash
# Round 2: quick 2-GPU check of the final build (distributed correctness worker, --quick).
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    tests/dist_gpu_worker.py --quick > gpurun_out/r02_dist_worker_quick_n2.log 2>&1; echo "dist worker rc=$?"
grep -v "^W\|^\[W\|OMP_NUM" gpurun_out/r02_dist_worker_quick_n2.log | tail -12
