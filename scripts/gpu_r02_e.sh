#!/bin/bash
# Round 2: secondary kernels on one GPU -- timings (full 512^3 fields) and one ncu --set full launch of each.
mkdir -p gpurun_out
TAG=${1:-e}
./tests/cpp/mix_bw | tee gpurun_out/r02_${TAG}_mix_bw.json
timeout 300 python scripts/bench_aux.py 2>&1 | grep "^{" | tee gpurun_out/r02_${TAG}_bench_aux.json
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"strain_displacement_kernel|modal_solve_kernel|modal_field_kernel" -c 8 \
    -o gpurun_out/r02_${TAG}_prof_aux -f python scripts/bench_aux.py --once > gpurun_out/r02_${TAG}_prof_aux.log 2>&1; echo "ncu aux rc=$?"
