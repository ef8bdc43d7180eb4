#!/usr/bin/env python
"""Host-buffer (e2e) path tuning on 512^3: chunk size x stream count, and zero-copy."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bri17_b200 as b
shape = (512, 512, 512)
L = tuple(n * h for n, h in zip(shape, (1.1, 1.2, 1.3)))
op = b.ModalOperator(shape, L, 5.6, 0.3)
hu = torch.empty((3,) + shape, dtype=torch.complex128, pin_memory=True)
hf = torch.empty_like(hu).pin_memory()
torch.view_as_real(hu).normal_()
M = 512 ** 3
def run(label):
    op.apply_modal_stiffness_host(hu, out=hf)
    t0 = time.perf_counter()
    for _ in range(2):
        op.apply_modal_stiffness_host(hu, out=hf)
    dt = (time.perf_counter() - t0) / 2
    print(json.dumps({"config": label, "ms": dt * 1e3, "gmodes_s": M / dt / 1e9, "pcie_gbs_each_way": 48 * M / dt / 1e9}), flush=True)
for streams in (2, 3, 4, 6):
    for rows in (2, 4, 8, 16, 32):
        op.set_option("host_streams", streams)
        op.set_option("host_chunk_rows", rows)
        run(f"staged streams={streams} chunk_planes={rows}")
op.set_option("host_zero_copy", 1)
run("zero_copy")
ref = hf.clone()
op.set_option("host_zero_copy", 0)
op.set_option("host_streams", 3); op.set_option("host_chunk_rows", 8)
op.apply_modal_stiffness_host(hu, out=hf)
print(json.dumps({"zero_copy_equals_staged": bool(torch.equal(torch.view_as_real(ref), torch.view_as_real(hf)))}))
