#!/bin/bash
# Round 2, single-GPU check of the final build: smoke, whole GPU test suite, secondary kernels (solve variants,
# strong-scaling shards), the bench guards (--rs-fault) and the default bench line.
mkdir -p gpurun_out
TAG=${1:-i}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_${TAG}_pytest_gpu.log
timeout 300 python scripts/bench_aux.py 2>&1 | grep "^{" | tee gpurun_out/r02_${TAG}_bench_aux.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --rs-fault raise > gpurun_out/r02_${TAG}_fault_raise.log 2>&1; echo "fault raise rc=$?"; grep "^{" gpurun_out/r02_${TAG}_fault_raise.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['realspace'], d['cg'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --rs-fault hang --rs-timeout 5 > gpurun_out/r02_${TAG}_fault_hang.log 2>&1; echo "fault hang rc=$?"; grep "^{" gpurun_out/r02_${TAG}_fault_hang.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['realspace'], d['cg'])"
timeout 900 python bench.py > gpurun_out/r02_${TAG}_bench_n1.log 2>&1; echo "bench rc=$?"; grep "^{" gpurun_out/r02_${TAG}_bench_n1.log | cut -c1-400
