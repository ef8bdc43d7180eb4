#!/bin/bash
# Round 2, last single-GPU look at the final build: smoke, 2-D 4096^2 line (BASELINE configs[1]), quick headline line.
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_k_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_k_smoke.log
timeout 200 python bench.py --dim 2 --edge 4096 --no-e2e --no-cpu-baseline 2>&1 | grep "^{" | tee gpurun_out/r02_k_bench2d.json | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 --no-realspace --no-cpu-baseline --e2e-steps 2 2>&1 | grep "^{" | tee gpurun_out/r02_k_bench3d_quick.json | cut -c1-300
