#!/bin/bash
# Round 2: fused axis-0 kernel iteration (1 GPU): real-space tests, 512^3 timings (complex / real / CG),
# the N0 = 1024 configuration on a 1024x256x512 grid (same volume), ncu capture of the 512 and 1024 kernels.
mkdir -p gpurun_out
TAG=${1:-d}
timeout 900 python -m pytest tests/test_gpu_realspace.py -m gpu -x -q > gpurun_out/r02_${TAG}_pytest_rs.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_${TAG}_pytest_rs.log
for extra in "" "--real"; do
  timeout 300 python scripts/run_realspace.py --edge 512 --applies 10 --cg 10 $extra 2>&1 | grep "^{" | tee -a gpurun_out/r02_${TAG}_realspace_n1.jsonl
done
for extra in "" "--real"; do
  timeout 300 python scripts/run_realspace.py --shape 1024,256,512 --applies 10 $extra 2>&1 | grep "^{" | tee -a gpurun_out/r02_${TAG}_realspace_n1.jsonl
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:axis0_fused -s 2 -c 1 \
    -o gpurun_out/r02_${TAG}_prof_axis0_fused -f python scripts/run_realspace.py --edge 512 --applies 1 > gpurun_out/r02_${TAG}_prof_axis0.log 2>&1; echo "ncu fused rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:axis0_fused -s 2 -c 1 \
    -o gpurun_out/r02_${TAG}_prof_axis0_fused_1024 -f python scripts/run_realspace.py --shape 1024,256,512 --applies 1 > gpurun_out/r02_${TAG}_prof_axis0_1024.log 2>&1; echo "ncu fused 1024 rc=$?"
