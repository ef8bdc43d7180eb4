#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
: > gpurun_out/realspace_r2c_n$N.log
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29530 bench_realspace.py "$@" 2>&1 | grep "^{" | tee -a gpurun_out/realspace_r2c_n$N.log | cut -c1-1200
}
run --edge 1024 --mode 1 --real --steps 5
run --edge 512 --mode 1 --real --cg-iters 20
run --edge 512 --mode 1 --cg-iters 20
