#!/bin/bash
# Round 2, single-GPU check of a build: smoke, the whole GPU test suite, secondary kernels, the bench line.
mkdir -p gpurun_out
TAG=${1:-f}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_${TAG}_pytest_gpu.log
timeout 300 python scripts/bench_aux.py 2>&1 | grep "^{" | tee gpurun_out/r02_${TAG}_bench_aux.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_${TAG}_bench_n1.log 2>&1; echo "bench rc=$?"; grep "^{" gpurun_out/r02_${TAG}_bench_n1.log | tail -c 7000
