#!/bin/bash
# First-contact GPU run: system info, smoke, GPU tests, kernel-variant sweep.
mkdir -p gpurun_out
(nproc; lscpu | head -25; free -g | head -2; nvidia-smi; nvidia-smi topo -m) > gpurun_out/sysinfo.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -5 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --sweep --steps 50 --warmup 5 > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"
cat gpurun_out/bench1.log
