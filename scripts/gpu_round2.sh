#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu2.log
timeout 300 python scripts/bench_aux.py | tee gpurun_out/bench_aux2.json
