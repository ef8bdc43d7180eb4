#!/bin/bash
# Round 2, first single-GPU session: smoke, full GPU test suite, bench line, secondary kernels,
# real-space apply fused vs cuFFT axis-0 path, ncu capture of the fused kernel.
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02_pytest_gpu.log
for extra in "" "--real" "--no-fused" "--no-fused --real"; do
  timeout 300 python scripts/run_realspace.py --edge 512 --applies 5 --cg 10 $extra 2>&1 | grep "^{" | tee -a gpurun_out/r02_realspace_n1.jsonl
done
timeout 300 python scripts/bench_aux.py 2>&1 | grep "^{" | tee gpurun_out/r02_bench_aux.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.log 2>&1; echo "bench rc=$?"; tail -c 6000 gpurun_out/r02_bench_n1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:axis0_fused -s 2 -c 2 \
    -o gpurun_out/r02_prof_axis0_fused -f python scripts/run_realspace.py --edge 512 --applies 1 > gpurun_out/r02_prof_axis0.log 2>&1; echo "ncu fused rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/r02_launches_realspace.csv python scripts/run_realspace.py --edge 512 --applies 2 > gpurun_out/r02_launches_realspace.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out | tail -15
