#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_realspace.py -m gpu -x -q -s > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"
grep -E "mode|DIST|passed|failed|Error|error" gpurun_out/pytest_r2c.log | tail -25
