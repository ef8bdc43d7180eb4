#!/usr/bin/env python
"""Stall samples of an ncu --set full --import-source on report, split at the CTA barriers of the
kernel (= the phases of axis0_fused_kernel), plus the hottest SASS lines.

    python scripts/ncu_phases.py gpurun_out/x.ncu-rep [top]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) > 10]
ix = {h: i for i, h in enumerate(hdr)}
S = lambda r, k: int(r[ix[k]] or 0)
tot = sum(S(r, "# Samples") for r in data)
print(f"{len(data)} SASS instructions, {tot} samples")
kinds = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stalls:", {k[6:]: sum(S(r, k) for r in data) for k in kinds if sum(S(r, k) for r in data) > tot * 0.01})
seg, segs = 0, {}
for n, r in enumerate(data):
    d = segs.setdefault(seg, {"samples": 0, "sass": 0, "executed": 0, "fp64": 0})
    d["samples"] += S(r, "# Samples"); d["sass"] += 1; d["executed"] += S(r, "Instructions Executed")
    src = r[ix["Source"]]
    if any(m in src for m in (" DADD", " DMUL", " DFMA")):
        d["fp64"] += S(r, "Instructions Executed")
    if "BAR.SYNC" in src or "WARPSYNC" in src or "DEPBAR" in src:
        print(f"  sync at {n}: {src.strip()[:50]} ({S(r, '# Samples')} samples)")
        seg += 1
for k, v in segs.items():
    print(f"segment {k}: {v} share {v['samples'] / tot:.3f}")
top = sorted(range(len(data)), key=lambda i: -S(data[i], "# Samples"))[:top_n]
for i in sorted(top):
    r = data[i]
    print(i, r[ix["Source"]].strip()[:58], S(r, "# Samples"),
          {k[6:]: S(r, k) for k in kinds if S(r, k) > 0.1 * S(r, "# Samples")})
