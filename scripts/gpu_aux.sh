#!/bin/bash
# Single-GPU auxiliary evidence: FP64 peak, secondary kernels, sanitizer, real-space 1-GPU bench.
mkdir -p gpurun_out
./tests/cpp/fp64_peak | tee gpurun_out/fp64_peak.json
timeout 300 python scripts/bench_aux.py | tee gpurun_out/bench_aux.json
cat > /tmp/san.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import bri17_b200 as b
from bri17_b200.realspace import RealSpaceOperator
for shape in ((3, 4, 5), (64, 64), (5, 7, 600), (4, 2000)):
    L = tuple(float(n) for n in shape)
    op = b.ModalOperator(shape, L, 5.6, 0.3)
    u = torch.view_as_complex(torch.randn((len(shape),) + shape + (2,), dtype=torch.float64, device="cuda"))
    for v in range(op.info("num_variants")):
        op.set_option("apply_variant", v)
        op.apply_modal_stiffness(u)
    op.apply_strain_displacement(u); op.freq_index_map(); op.modal_stiffness_field(); op.modal_strain_displacement_field()
    op.apply_modal_stiffness_host(u.cpu().numpy())
    rs = RealSpaceOperator(shape, L, 5.6, 0.3)
    F = rs.apply(u); rs.cg_solve(F - F.mean(dim=tuple(range(1, len(shape) + 1)), keepdim=True), max_iter=5)
torch.cuda.synchronize()
print("sanitizer workload done")
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/sanitizer_racecheck.log
timeout 600 python bench_realspace.py --edge 512 --cg-iters 20 2>&1 | grep "^{" | tee gpurun_out/realspace_n1.log
timeout 300 python bench.py --dim 2 --edge 4096 --no-e2e 2>&1 | grep "^{" | tee gpurun_out/bench2d_default.log
