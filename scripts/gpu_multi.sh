#!/bin/bash
# Multi-GPU bench (torchrun, one rank per GPU) + the C++ GPU test.  Usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 300 python -m pytest tests/test_cpp_header.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_n$N.log 2>&1; echo "bench N=$N rc=$?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.log | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref N=$N rc=$?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/bench_ref_n$N.log | tail -2 | cut -c1-600
