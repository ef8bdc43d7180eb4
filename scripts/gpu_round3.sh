#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_e2e_sweep.py 2>&1 | grep "^{" | tee gpurun_out/e2e_sweep.jsonl
cat > /tmp/prof_aux.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bri17_b200 as b
shape = (256, 256, 512); L = (1., 1., 1.)
op = b.ModalOperator(shape, L, 5.6, 0.3)
u = torch.view_as_complex(torch.randn((3,) + shape + (2,), dtype=torch.float64, device="cuda"))
tau = torch.view_as_complex(torch.randn((6,) + shape + (2,), dtype=torch.float64, device="cuda"))
for _ in range(3):
    op.apply_strain_displacement(u); op.eigenstress_to_opposite_strain(tau); op.solve_modal_stiffness(u)
    op.set_option("mapping", 2); op.apply_modal_stiffness(u); op.set_option("mapping", 0)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'strain_displacement_kernel|modal_solve_kernel|apply_flat_kernel' -s 6 -c 4 \
    -o gpurun_out/prof_aux -f python /tmp/prof_aux.py > gpurun_out/prof_aux.log 2>&1; echo "ncu aux rc=$?"
