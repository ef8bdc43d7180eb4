#!/bin/bash
# Round 2, 8-GPU session: distributed correctness worker (full) + bench.py --gpus 8 (strong-scaled headline,
# weak companion, 1024^3 real-space apply and 512^3 CG records with in-run parity).  No reference arm here
# (CPU only; the driver runs it).
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n${N}.txt 2>&1
PORT=29521
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    tests/dist_gpu_worker.py $2 > gpurun_out/r02_dist_worker_n${N}.log 2>&1; echo "dist worker rc=$?"
grep -v "^W\|^\[W\|OMP_NUM" gpurun_out/r02_dist_worker_n${N}.log | tail -12
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}.log 2>&1; echo "bench rc=$?"
grep "^{" gpurun_out/r02_bench_n${N}.log | tail -c 9000
