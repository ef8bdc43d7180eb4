#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
: > gpurun_out/realspace_pipe_n$N.log
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29530 bench_realspace.py "$@" 2>&1 | grep "^{" | tee -a gpurun_out/realspace_pipe_n$N.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][6:12], d['config'].get('fields','')[:7], 'pipelined' if d.get('pipelined') else 'serial', '| ms', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['phases_ms_last_apply_max_over_ranks'].items()}, 'cg', d.get('cg',{}).get('iterations_per_s'))
"
}
run --edge 1024 --mode 1 --real --steps 5
run --edge 1024 --mode 1 --real --steps 5 --no-pipeline
run --edge 1024 --mode 1 --steps 5
run --edge 512 --mode 1 --real --cg-iters 20
run --edge 512 --mode 1 --cg-iters 20
