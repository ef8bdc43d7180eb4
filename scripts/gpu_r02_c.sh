#!/bin/bash
# Round 2: L2-sized chunks for the local 2-D transforms (BatchFft) -- sweep of the chunk size at 512^3 on one
# GPU, complex and real fields, then the real-space GPU tests with the default chunk.
mkdir -p gpurun_out
TAG=${1:-c}
for mib in 0 8 16 32 48 64 96; do
  for extra in "" "--real"; do
    timeout 300 python scripts/run_realspace.py --edge 512 --applies 10 --fft-chunk-mib $mib $extra 2>&1 | grep "^{" | tee -a gpurun_out/r02_${TAG}_chunk_sweep.jsonl
  done
done
timeout 300 python scripts/run_realspace.py --edge 512 --applies 10 --cg 10 2>&1 | grep "^{" | tee -a gpurun_out/r02_${TAG}_chunk_sweep.jsonl
timeout 300 python scripts/run_realspace.py --edge 512 --applies 10 --cg 10 --real 2>&1 | grep "^{" | tee -a gpurun_out/r02_${TAG}_chunk_sweep.jsonl
timeout 900 python -m pytest tests/test_gpu_realspace.py -m gpu -x -q > gpurun_out/r02_${TAG}_pytest_rs.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_${TAG}_pytest_rs.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_${TAG}_launches_chunked.csv python scripts/run_realspace.py --edge 512 --applies 1 > gpurun_out/r02_${TAG}_launches_chunked.log 2>&1; echo "launch list rc=$?"
