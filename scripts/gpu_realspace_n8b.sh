#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
: > gpurun_out/realspace_n${N}_v2.log
run() {
  local edge=$1 mode=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29520 bench_realspace.py --edge $edge --mode $mode "$@" 2>&1 | grep "^{" | tee -a gpurun_out/realspace_n${N}_v2.log
}
run 1024 1 --steps 5
run 512 1 --cg-iters 20
run 1024 0 --steps 5
