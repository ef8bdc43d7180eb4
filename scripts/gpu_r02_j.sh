#!/bin/bash
# Round 2: whole GPU test suite of the final build (1 GPU).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_j_pytest_gpu.log
