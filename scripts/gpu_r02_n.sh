#!/bin/bash
# Round 2, multi-GPU session: bash scripts/gpu_r02_n.sh N [quick]
# distributed correctness worker (all exchange modes / layouts / KAT / inclusion) + bench.py on N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n${N}.txt 2>&1
PORT=29511
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    tests/dist_gpu_worker.py $2 > gpurun_out/r02_dist_worker_n${N}.log 2>&1; echo "dist worker rc=$?"
grep -v "^W\|^\[W\|OMP_NUM" gpurun_out/r02_dist_worker_n${N}.log | tail -45
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}.log 2>&1; echo "bench rc=$?"
grep "^{" gpurun_out/r02_bench_n${N}.log | tail -c 7000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+2)) \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/r02_bench_ref_n${N}.log 2>&1; echo "ref rc=$?"
grep "^{" gpurun_out/r02_bench_ref_n${N}.log | cut -c1-400
