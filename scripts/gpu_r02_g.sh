#!/bin/bash
# Round 2 evidence session (1 GPU): ncu launch list of the default bench command, ncu --set full of the headline
# kernel (roofline traffic), compute-sanitizer memcheck + racecheck over the round-2 kernels on small grids.
mkdir -p gpurun_out
TAG=${1:-g}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-realspace > gpurun_out/r02_${TAG}_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:modal_stiffness_apply -s 3 -c 1 \
    -o gpurun_out/r02_${TAG}_prof_apply3d -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-realspace > gpurun_out/r02_${TAG}_prof_apply3d.log 2>&1; echo "ncu apply rc=$?"
cat > /tmp/san.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import bri17_b200 as b
from bri17_b200.realspace import RealSpaceOperator
for shape in ((3, 4, 5), (64, 64), (5, 7, 600), (4, 2000)):
    L = tuple(float(n) for n in shape)
    dim = len(shape)
    op = b.ModalOperator(shape, L, 5.6, 0.3)
    u = torch.view_as_complex(torch.randn((dim,) + shape + (2,), dtype=torch.float64, device="cuda"))
    op.apply_modal_stiffness(u)
    op.apply_strain_displacement(u); op.freq_index_map(); op.modal_stiffness_field(); op.modal_strain_displacement_field()
    tau = torch.view_as_complex(torch.randn((dim * (dim + 1) // 2,) + shape + (2,), dtype=torch.float64, device="cuda"))
    op.eigenstress_to_displacement(tau); op.eigenstress_to_opposite_strain(tau); op.eigenstress_to_force(tau); op.solve_modal_stiffness(u)
# fused axis-0 pass: every supported N0 (two- and three-stage plans), natural and k1-major layouts, real and complex, CG
for shape in ((16, 6, 5), (32, 5, 6), (64, 4, 6), (128, 4, 5), (256, 3, 4), (512, 2, 3), (1024, 2, 2), (64, 12), (512, 6)):
    L = tuple(float(n) for n in shape)
    dim = len(shape)
    for k1 in (0, 1):
        rs = RealSpaceOperator(shape, L, 5.6, 0.3)
        rs.set_option("k1_major", k1)
        u = torch.view_as_complex(torch.randn((dim,) + shape + (2,), dtype=torch.float64, device="cuda"))
        F = rs.apply(u)
        Fr, dot = rs.apply_with_dot(u.real.contiguous())
        rs.cg_solve_real(Fr, max_iter=3)
        rs.close()
torch.cuda.synchronize()
print("sanitizer workload done")
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/r02_${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r02_${TAG}_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san.py > gpurun_out/r02_${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r02_${TAG}_sanitizer_racecheck.log
