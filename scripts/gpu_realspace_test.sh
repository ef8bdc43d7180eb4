#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_realspace.py -m gpu -x -q -s > gpurun_out/pytest_realspace.log 2>&1; echo "rc=$?"
tail -40 gpurun_out/pytest_realspace.log
