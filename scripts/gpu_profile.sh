#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe) + a 2-D sweep.
mkdir -p gpurun_out
BENCH="python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:modal_stiffness_apply -s 3 -c 3 \
    -o gpurun_out/prof_apply3d -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof3d.log 2>&1; echo "full 3d rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:modal_stiffness_apply -s 3 -c 2 \
    -o gpurun_out/prof_apply2d -f python bench.py --dim 2 --edge 4096 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof2d.log 2>&1; echo "full 2d rc=$?"
timeout 300 python bench.py --dim 2 --edge 4096 --sweep --steps 200 --warmup 20 --no-e2e > gpurun_out/bench2d.log 2>&1; echo "bench2d rc=$?"
cat gpurun_out/bench2d.log
timeout 300 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench default rc=$?"
cat gpurun_out/bench_default.log
ls -la gpurun_out
