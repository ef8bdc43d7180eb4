#!/bin/bash
# Usage: gpu_realspace_bench.sh N  -- end-to-end real-space apply + CG on N GPUs
N=${1:-2}
mkdir -p gpurun_out
run() {  # edge mode extra...
  local edge=$1 mode=$2; shift 2
  if [ "$N" = 1 ]; then
    timeout 600 python bench_realspace.py --edge $edge --mode $mode "$@"
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29520 bench_realspace.py --edge $edge --mode $mode "$@"
  fi 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/realspace_n$N.log
}
: > gpurun_out/realspace_n$N.log
run 512 1 --cg-iters 20
if [ "$N" != 1 ]; then run 512 0 --cg-iters 20; run 1024 1 --steps 5; run 1024 0 --steps 5; fi
