#!/usr/bin/env python
"""bench.py -- fp64 modal stiffness apply (bri17 hot path) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path -- f^[k] = K^[k] u^[k] for every
frequency (tests/test_bri17.cpp:58-92 of the reference) -- over one synthetic
3-D 512^3 field per GPU (BASELINE.json configs[2]; 2^27 modes, 6 GiB in +
6 GiB out, far larger than the 126 MB L2, so no flush is needed between
iterations).  With N > 1 every rank owns one k0 slab (no communication); the
headline line is weak-scaled (512^3 modes per GPU, global grid (512N)x512x512)
and the strong-scaled 512^3 figure of BASELINE config 3 is reported beside it.

Rank 0 prints ONE JSON line; see DESIGN.md section 6 for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gmodes/s (fp64 modal stiffness apply, 3D 512^3)"
UNIT = "Gmodes/s"
MU, NU = 5.6, 0.3                       # tests/test_bri17.cpp:320-321
SPACING = (1.1, 1.2, 1.3)               # tests/test_bri17.cpp:330
BYTES_PER_MODE = {2: 64, 3: 96}         # DIM complex in + DIM complex out


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--edge", type=int, default=512, help="grid edge (512 = the BASELINE workload)")
    ap.add_argument("--dim", type=int, default=3, choices=[2, 3],
                    help="3 = the BASELINE metric; 2 with --edge 4096 = BASELINE configs[1] (tuning only)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=-1, help="kernel variant (tuning)")
    ap.add_argument("--sweep", action="store_true", help="time every kernel variant (tuning aid)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed
    region runs (nvidia-smi's 200 ms period is too coarse for a 0.5 s region)."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80}

    def __init__(self, index, uuid=None):
        self.samples, self.reasons, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:        # the CUDA ordinal is not the NVML index under CUDA_VISIBLE_DEVICES: go by UUID
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report it, do not fail the bench
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reasons |= nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        names = [n for n, bit in {**self.BAD, **self.NOTE}.items() if self.reasons & bit]
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": names, "samples": len(self.samples)}


def workload_name(dim, edge, world):
    shape = (edge * world,) + (edge,) * (dim - 1)
    return (f"{dim}D Q{1 << dim} {edge}^{dim} modal stiffness apply per GPU (BASELINE configs[{dim - 1}]); "
            f"global grid {'x'.join(map(str, shape))}, one k0 slab per GPU, no comm")


def cpu_reference_leg(edge, steps, warmup, target_step_s, threads=None, world=1):
    """Times the reference's CPU implementation of the path (oracle/_ref: the
    unmodified reference header in the harness loop, OpenMP over k0 on all host
    threads; the C port if _ref is absent) on a bounded k0-slab sample of the
    same workload.  Returns (value Gmodes/s, s/step, descr dict)."""
    from oracle import oracle                      # checker/baseline only (see oracle/)
    impl = oracle.best(fast=True)
    # every host thread the process may use; not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    threads = threads or len(os.sched_getaffinity(0))
    shape = (edge * world, edge, edge)               # the same global grid as the GPU arm at this N
    L = tuple(n * h for n, h in zip(shape, SPACING))
    plane = edge * edge
    # calibrate on a few planes, then size the sample for ~target_step_s per step
    probe = max(1, min(edge, threads))
    u = oracle.synthetic_u_hat(3, (probe, edge, edge), seed=[3, 0])
    out = np.empty_like(u)
    impl.apply_modal_stiffness(shape, L, MU, NU, u, k_begin=(0, 0, 0), out=out, nthreads=threads)
    t0 = time.perf_counter()
    impl.apply_modal_stiffness(shape, L, MU, NU, u, k_begin=(0, 0, 0), out=out, nthreads=threads)
    rate = probe * plane / (time.perf_counter() - t0)
    planes = int(min(edge, max(threads, round(target_step_s * rate / plane))))
    k0 = (edge - planes) // 2                       # mid-spectrum slab
    u = oracle.synthetic_u_hat(3, (planes, edge, edge), seed=[3, 1])
    out = np.empty_like(u)
    for _ in range(warmup):
        impl.apply_modal_stiffness(shape, L, MU, NU, u, k_begin=(k0, 0, 0), out=out, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        impl.apply_modal_stiffness(shape, L, MU, NU, u, k_begin=(k0, 0, 0), out=out, nthreads=threads)
    dt = (time.perf_counter() - t0) / steps
    modes = planes * plane
    descr = {"value": modes / dt / 1e9, "unit": UNIT, "cores": int(threads), "kind": impl.kind,
             "sample": f"{planes} of {shape[0]} k0-planes of the {'x'.join(map(str, shape))} grid ({modes} modes) per step, "
                       f"{steps} steps, -O3 x86-64-v3 OpenMP build of {os.path.basename(impl.path)}"}
    return modes / dt / 1e9, dt, descr


def run_reference(args, rank, world):
    if rank != 0:
        return
    workload = workload_name(3, args.edge, args.gpus)
    value, dt, descr = cpu_reference_leg(args.edge, max(1, args.steps), max(0, args.warmup),
                                         target_step_s=0.25, world=args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "sample": descr["sample"],
                       "implementation": "reference CPU path: tests/test_bri17.cpp:76-91 loop over the unmodified "
                                         "Hooke::modal_stiffness (oracle/_ref), OpenMP over k0, host cores only"},
            "cpu_baseline": descr,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs {args.gpus} ranks (torchrun), found WORLD_SIZE={world}")

    import torch
    import torch.distributed as dist
    import bri17_b200 as b
    from bri17_b200 import slab

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bri17_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    edge, dim = args.edge, args.dim
    peak, peak_src = peaks()

    # ---- weak-scaled workload: one edge^dim slab per GPU of a (edge*N, edge[, edge]) grid ----
    shape = (edge * world,) + (edge,) * (dim - 1)
    L = tuple(n * h for n, h in zip(shape, SPACING))
    op = b.ModalOperator(shape, L, MU, NU, device=local_rank)
    if args.variant >= 0:
        op.set_option("apply_variant", args.variant)
    k_begin, local = slab.rank_block(shape, rank, world)
    modes_rank = int(np.prod(local, dtype=np.int64))
    gen = torch.Generator(device=dev).manual_seed(3000 + rank)
    u = torch.view_as_complex(torch.randn((dim,) + local + (2,), dtype=torch.float64, device=dev,
                                          generator=gen))
    f = torch.empty_like(u)
    stream = torch.cuda.current_stream()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.__enter__()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        if sampler:
            sampler.__exit__()
        return slab.max_over_ranks(e0.elapsed_time(e1) / steps, dev if world > 1 else None)  # ms

    if args.sweep and rank == 0:
        for v in range(op.info("num_variants")):
            op.set_option("apply_variant", v)
            ms = timed(lambda: op.apply_modal_stiffness(u, out=f, k_begin=k_begin), 20, 3)
            gbs = BYTES_PER_MODE[dim] * modes_rank / ms / 1e6
            print(f"# variant {v:2d} grid {op.info('last_grid')} block {op.info('last_block')}: "
                  f"{ms:.4f} ms  {gbs:.0f} GB/s  {gbs / peak:.3f} of peak", flush=True)
        op.set_option("apply_variant", args.variant)

    launches0 = op.info("launches")
    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(local_rank), "uuid", None))
    ms = timed(lambda: op.apply_modal_stiffness(u, out=f, k_begin=k_begin), args.steps, args.warmup,
               sampler)
    launches = op.info("launches") - launches0 - args.warmup
    value = modes_rank * world / (ms * 1e-3) / 1e9
    achieved = BYTES_PER_MODE[dim] * modes_rank / (ms * 1e-3) / 1e9      # GB/s per GPU

    # ---- cpu_baseline leg, part 1 (N=1, rank 0 only): the CPU reference doubles as the checker --
    # three k0 planes of the field just benchmarked are recomputed by it and compared (outside
    # the timed region).  This is the only place besides the timing below where oracle/ is used.
    cpu_leg = rank == 0 and world == 1 and dim == 3 and not args.no_cpu_baseline
    parity = None
    if cpu_leg:
        try:
            from oracle import oracle                  # checker / baseline only
            o = oracle.best()
            worst = 0.0
            for a in sorted({0, local[0] // 2, local[0] - 1}):
                k0 = k_begin[0] + a
                ref = o.apply_modal_stiffness(shape, L, MU, NU, u[:, a:a + 1].cpu().numpy(),
                                              k_begin=(k0,) + (0,) * (dim - 1))
                got = f[:, a:a + 1].cpu().numpy()
                den = np.abs(ref).max(axis=0)
                num = np.abs(got - ref).max(axis=0)
                nz = den > 0
                worst = max(worst, float((num[nz] / den[nz]).max()), float(num[~nz].max(initial=0.0)))
            parity = {"max_rel_err_per_mode": worst, "planes_checked": 3, "gate": 1e-12,
                      "oracle": o.kind}
        except Exception as e:
            parity = {"error": repr(e)}

    # ---- strong-scaled companion: the edge^3 grid of BASELINE config 3 split over N GPUs ----
    strong = None
    if world > 1:
        s_shape = (edge,) * dim
        s_L = tuple(n * h for n, h in zip(s_shape, SPACING))
        s_op = b.ModalOperator(s_shape, s_L, MU, NU, device=local_rank)
        s_kb, s_local = slab.rank_block(s_shape, rank, world)
        su, sf = u[:, :s_local[0]].contiguous(), f[:, :s_local[0]].contiguous()
        s_ms = timed(lambda: s_op.apply_modal_stiffness(su, out=sf, k_begin=s_kb), args.steps, args.warmup)
        strong = {"workload": f"3D {edge}^3 split into {world} k0 slabs", "ms_per_step": s_ms,
                  "value": edge ** dim / (s_ms * 1e-3) / 1e9, "unit": UNIT, "scaling": "strong"}
        del su, sf

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hu = torch.empty(u.shape, dtype=u.dtype, pin_memory=True)
        hf = torch.empty(u.shape, dtype=u.dtype, pin_memory=True)
        hu.copy_(u)
        torch.cuda.synchronize()
        del f                                       # make room for the staging buffers
        op.apply_modal_stiffness_host(hu, out=hf, k_begin=k_begin)          # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            op.apply_modal_stiffness_host(hu, out=hf, k_begin=k_begin)       # synchronous
        torch.cuda.synchronize()
        e_s = (time.perf_counter() - t0) / args.e2e_steps
        e_s = slab.max_over_ranks(e_s, dev if world > 1 else None)
        nbytes = u.numel() * 16
        e2e = {"value": modes_rank * world / e_s / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "ms_per_step": e_s * 1e3, "steps": args.e2e_steps,
               "api": "bri17_modal_stiffness_apply_host_f64 (pinned host buffers, chunked H2D/kernel/D2H)",
               "pcie_gbs_each_way": nbytes / e_s / 1e9}
        del hu, hf

    if rank == 0:
        traffic = None            # DRAM bytes per launch from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath) and edge == {3: 512, 2: 4096}[dim]:
            traffic = json.load(open(tpath))[str(dim)]["dram_bytes_per_launch"] / 1e9
        cpu = None
        if cpu_leg:                                   # cpu_baseline leg, part 2: timing
            try:
                _, _, cpu = cpu_reference_leg(edge, steps=5, warmup=1, target_step_s=3.0)
                # the reference as shipped is single-threaded: same build, one thread, smaller sample
                v1, _, d1 = cpu_reference_leg(edge, steps=2, warmup=1, target_step_s=1.0, threads=1)
                cpu["single_thread"] = {"value": v1, "unit": UNIT, "sample": d1["sample"]}
                cpu["parity_of_gpu_result_vs_this_baseline"] = parity
            except Exception as e:
                cpu = {"error": repr(e)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(dim, edge, world),
                       "modes_per_gpu": modes_rank, "mu": MU, "nu": NU, "spacing": SPACING,
                       "l2": f"inputs ({modes_rank * 16 * dim / 2**30:.2f} GiB) and outputs (same) per GPU vs the 126 MB L2: "
                             + ("far larger, no flush between iterations" if modes_rank * 16 * dim > 2**30
                                else "NOT larger than L2 (non-default size: timing is L2-assisted)"),
                       "kernel_variant": op.info("apply_variant"), "grid": op.info("last_grid"),
                       "block": op.info("last_block"), "smem": op.info("last_smem")},
            "gdof_per_s": value * dim,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_unit": "GB per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "algorithmic_gb_per_launch": BYTES_PER_MODE[dim] * modes_rank / 1e9,
                         "peak_source": peak_src,
                         "bytes_per_mode": BYTES_PER_MODE[dim], "modes_per_launch": modes_rank,
                         "kernel": f"modal_stiffness_apply_kernel<{dim},...>",
                         "timing": "CUDA events on the launch stream around the timed steps / steps"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "strong_scaling": strong,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
