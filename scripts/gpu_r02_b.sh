#!/bin/bash
# Round 2: fused axis-0 kernel iteration (1 GPU): real-space tests, timings fused vs cuFFT path, ncu capture.
mkdir -p gpurun_out
TAG=${1:-b}
timeout 900 python -m pytest tests/test_gpu_realspace.py -m gpu -x -q > gpurun_out/r02_${TAG}_pytest_rs.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_${TAG}_pytest_rs.log
for extra in "" "--real" "--no-k1-major" "--no-k1-major --real" "--no-fused"; do
  timeout 300 python scripts/run_realspace.py --edge 512 --applies 10 --cg 10 $extra 2>&1 | grep "^{" | tee -a gpurun_out/r02_${TAG}_realspace_n1.jsonl
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:axis0_fused -s 2 -c 1 \
    -o gpurun_out/r02_${TAG}_prof_axis0_fused -f python scripts/run_realspace.py --edge 512 --applies 1 > gpurun_out/r02_${TAG}_prof_axis0.log 2>&1; echo "ncu fused rc=$?"
