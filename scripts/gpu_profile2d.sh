#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:modal_stiffness_apply -s 3 -c 2 \
    -o gpurun_out/prof_apply2d_default -f python bench.py --dim 2 --edge 4096 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof2d_default.log 2>&1; echo "rc=$?"
