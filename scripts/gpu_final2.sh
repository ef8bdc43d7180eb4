#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; grep "^{" gpurun_out/bench_final.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'])"
