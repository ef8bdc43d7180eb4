#!/bin/bash
# Final single-GPU validation of the round: what the driver will run, plus a fresh launch list.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
timeout 300 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; grep "^{" gpurun_out/bench_final.log | cut -c1-400
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; echo "ref rc=$?"; grep "^{" gpurun_out/bench_ref_final.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1; echo "launch list rc=$?"
