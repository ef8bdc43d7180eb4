#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_realspace.py -m gpu -x -q -s > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"
grep -E "mode|DIST|passed|failed|Error|error" gpurun_out/pytest_r2c.log | tail -25
: > gpurun_out/realspace_r2c_n$N.log
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29530 bench_realspace.py "$@" 2>&1 | grep "^{" | tee -a gpurun_out/realspace_r2c_n$N.log | cut -c1-900
}
run --edge 512 --mode 1 --real --cg-iters 20
run --edge 512 --mode 0 --real
run --edge 1024 --mode 1 --real --steps 5
