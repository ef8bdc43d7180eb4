#!/usr/bin/env python
"""bench.py -- fp64 modal stiffness apply (bri17 hot path) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path -- f^[k] = K^[k] u^[k] for every
frequency (tests/test_bri17.cpp:58-92 of the reference) -- over one synthetic
3-D 512^3 field (BASELINE.json configs[2]; 2^27 modes, 6 GiB in + 6 GiB out;
even an eighth of it is far larger than the 126 MB L2, so no flush is needed between
iterations).  With N > 1 the SAME 512^3 grid is split into N k0 slabs, one per
rank, no communication (BASELINE configs[2]: "512^3 at 1/2/4/8 B200"): the
headline is strong-scaled, so value_N / (N value_1) is the ">= 7x on 8 GPUs"
claim itself; the weak-scaled figure (512^3 per GPU) is the companion key
`weak_scaling`.  The line also carries `realspace` (BASELINE configs[3]: FFT ->
K^ -> iFFT, 1024^3 with NVLink all-to-all transposes for N > 1, 512^3 at N = 1)
and `cg` (configs[4]: 512^3 matrix-free CG on the periodic inclusion problem),
each with in-run parity checks.

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
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaled companion figure (N > 1)")
    ap.add_argument("--no-realspace", action="store_true",
                    help="skip the real-space apply / CG records (BASELINE configs[3], [4]) and their parity checks")
    ap.add_argument("--rs-edge", type=int, default=0, help="real-space apply grid edge (default 1024 for N > 1, 512 at N = 1)")
    ap.add_argument("--rs-steps", type=int, default=5)
    ap.add_argument("--cg-edge", type=int, default=512)
    ap.add_argument("--cg-max-iter", type=int, default=20000)
    ap.add_argument("--rs-fault", default="", choices=["", "raise", "hang"],
                    help="test aid: make the real-space records fail / hang, to check that the headline line still prints")
    ap.add_argument("--rs-timeout", type=float, default=900.0,
                    help="seconds after which the real-space / CG records are abandoned and the headline line is printed")
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
            time.sleep(0.001)

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
    """Both arms (and every N) name the workload identically: BASELINE configs[2], strong-scaled."""
    shape = "x".join([str(edge)] * dim)
    return (f"{dim}D Q{1 << dim} {edge}^{dim} modal stiffness apply (BASELINE configs[{dim - 1}]): "
            f"the {shape} grid split into N k0 slabs, one per GPU, no comm")


def cpu_reference_leg(edge, steps, warmup, target_step_s, threads=None):
    """Times the reference's CPU implementation of the path (oracle/_ref: the
    unmodified reference header in the harness loop, OpenMP over k0 on all host
    threads; the C port if _ref is absent) on a bounded k0-slab sample of the
    same workload.  Returns (value Gmodes/s, s/step, descr dict)."""
    from oracle import oracle                      # checker/baseline only (see oracle/)
    impl = oracle.best(fast=True)
    # every host thread the process may use; not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    threads = threads or len(os.sched_getaffinity(0))
    shape = (edge, edge, edge)
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
                                         target_step_s=0.25)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "sample": descr["sample"],
                       "implementation": "reference CPU path: tests/test_bri17.cpp:76-91 loop over the unmodified "
                                         "Hooke::modal_stiffness (oracle/_ref), OpenMP over k0, host cores only"},
            "cpu_baseline": descr,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def pcie_probe(torch, dev, reduce_max, nbytes=1 << 30, reps=3):
    """Code-independent host-link probe: plain cudaMemcpyAsync (torch copy_) between per-rank
    pinned buffers and the GPU, H2D alone, D2H alone and both at once, all ranks together.
    Returns GB/s per GPU per direction (from the slowest rank's time)."""
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_in.zero_()
    h_out.zero_()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(h2d, d2h):
        best = None
        for _ in range(reps + 1):
            reduce_max(0.0)                      # line the ranks up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = reduce_max(time.perf_counter() - t0)
            best = dt if best is None else min(best, dt)
        return nbytes / best / 1e9

    out = {"h2d_only_gbs_per_gpu": run(True, False), "d2h_only_gbs_per_gpu": run(False, True),
           "duplex_gbs_per_gpu_each_way": run(True, True), "bytes": nbytes,
           "method": "torch copy_(non_blocking) = cudaMemcpyAsync between cudaHostAlloc'ed buffers and HBM, "
                     "one H2D and one D2H stream per rank, all ranks concurrently, slowest rank's wall time"}
    del h_in, h_out, d_in, d_out
    return out


def realspace_records(args, torch, dist, local_rank, rank, world, dev):
    """BASELINE configs[3] (real-space apply FFT -> K^ -> iFFT, NVLink all-to-all transposes) and
    configs[4] (matrix-free CG on the periodic inclusion problem), each WITH in-run parity checks
    (tests/realspace_checks.py: checker code, may use the CPU oracle).  Reported as extra keys of
    the bench line; the headline metric is untouched."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import realspace_checks as rc
    from bri17_b200 import slab
    from bri17_b200.realspace import PHASES, RealSpaceOperator

    rdev = dev if world > 1 else None
    stream = torch.cuda.current_stream()
    out = {"realspace": None, "cg": None}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_apply(op, real, steps, warmup):
        gen = torch.Generator(device=dev).manual_seed(4000 + rank)
        if real:
            u = torch.randn(op.real_shape, dtype=torch.float64, device=dev, generator=gen)
        else:
            u = torch.zeros(op.real_shape + (2,), dtype=torch.float64, device=dev)
            u[..., 0].normal_(generator=gen)               # real data carried as complex (tests/test_bri17.cpp:133-136)
            u = torch.view_as_complex(u)
        F = torch.empty_like(u)
        fn = op.apply_real if real else op.apply
        for _ in range(warmup):
            fn(u, out=F)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn(u, out=F)
        e1.record(stream)
        barrier()
        ms = slab.max_over_ranks(e0.elapsed_time(e1) / steps, rdev)
        phases = {k: slab.max_over_ranks(v, rdev) for k, v in op.timings().items()}
        del u, F
        return ms, phases

    # ---- configs[3]: real-space apply ----
    try:
        edge = args.rs_edge or (1024 if world > 1 else 512)
        shape = (edge,) * 3
        L = tuple(n * h for n, h in zip(shape, SPACING))
        rec = {"workload": f"3D Q8 {edge}^3 real-space apply F = (|h|/|N|) iDFT(K^ DFT(u)) (BASELINE configs[3]), "
                           f"n0 slabs over {world} GPU(s)",
               "exchange": "fused peer-store kernel over NVLink (CUDA IPC) + flag barriers in peer memory"
                           if world > 1 else "none (single GPU)"}
        parity = {}
        parity["small_grids_vs_numpy_restatement"] = rc.small_grids(local_rank, 1, 1, 1)
        parity["small_grids_gate"] = 1e-13
        if world > 1:
            parity["grid_256_distributed_vs_single_gpu"] = rc.vs_single_gpu(local_rank, edge=256)
            parity["dense_kat_worst_violation_of_reference_tolerance"] = rc.dense_kat(local_rank, world)
        torch.cuda.empty_cache()
        op = RealSpaceOperator.from_process_group(shape, L, MU, NU, device=local_rank, exchange_mode=1)
        rec["fused_axis0_kernel"] = bool(op.info("fused_axis0"))
        rec["pipelined"] = bool(op.info("pipeline"))
        rec["sub_slabs_per_component"] = {"complex": op.info("exchange_chunks"), "real": op.info("exchange_chunks_real")}
        for real in (False, True):
            key = "real_fields_r2c" if real else "complex_fields_c2c"
            ms, phases = timed_apply(op, real, args.rs_steps, 2)
            torch.cuda.empty_cache()
            xbytes = op.exchange_bytes_real if real else op.exchange_bytes
            rec[key] = {"ms_per_apply": ms, "applies_per_s": 1e3 / ms,
                        "effective_gmodes_per_s": edge ** 3 / (ms * 1e-3) / 1e9,
                        "phases_ms_last_apply_max_over_ranks": phases,
                        "exchange_bytes_per_gpu_per_direction": xbytes,
                        "exchange_gbs_if_fully_exposed": (2 * xbytes / (ms * 1e-3) / 1e9) if world > 1 else None,
                        "nvlink_peer_copy_gbs_measured": 770.0 if world > 1 else None}
            parity[f"plane_waves_{edge}_{'real' if real else 'complex'}"] = rc.plane_waves(op, real)
            torch.cuda.empty_cache()
        parity["plane_waves_gate"] = 1e-12
        op.close()
        del op
        torch.cuda.empty_cache()
        rec["parity"] = parity
        rec["parity_ok"] = bool(parity["small_grids_vs_numpy_restatement"] <= 1e-13
                                and parity.get("grid_256_distributed_vs_single_gpu", 0.0) <= 1e-13
                                and parity.get("dense_kat_worst_violation_of_reference_tolerance", 0.0) <= 0.0
                                and all(v <= 1e-12 for k, v in parity.items() if k.startswith("plane_waves_")
                                        and k != "plane_waves_gate"))
        out["realspace"] = rec
    except Exception as e:                                  # the headline line must still print
        import traceback
        out["realspace"] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
        try:                                                # a sticky CUDA error: nothing else can run in this process
            torch.cuda.synchronize()
        except Exception:
            out["cg"] = {"error": "skipped: the CUDA context is unusable after the real-space record failed"}
            return out

    # ---- configs[4]: CG on the periodic inclusion problem ----
    try:
        torch.cuda.empty_cache()
        edge = args.cg_edge
        shape = (edge,) * 3
        L = (1.0, 1.0, 1.0)                                 # python/demo.py:9, mu = 1, nu = 0.3 (:13-14)
        op = RealSpaceOperator.from_process_group(shape, L, 1.0, 0.3, device=local_rank, exchange_mode=1)
        inc = rc.inclusion_problem(op, rtol=1e-8, max_iter=args.cg_max_iter, check_every=25)
        rec = {"workload": f"3D Q8 {edge}^3 matrix-free CG on the periodic inclusion problem (BASELINE configs[4]; "
                           f"python/demo.py:11-23 in 3-D: eigenstress patch [0, N/8)^3, unit last Mandel component; "
                           f"mu = 1, nu = 0.3, L = 1), real fields, {world} GPU(s)",
               "per_iteration": "1 real-space apply (returns <p, A p> from the K^ kernel) + r -= alpha A p with <r, r> "
                                "+ (x += alpha p, p = r + beta p) = 8 vector passes, 2 scalar all-reduces, scalars on device",
               "fused_axis0_kernel": bool(op.info("fused_axis0")), **inc,
               "gate_err_vs_direct_solve": 1e-6,
               "parity_ok": bool(inc["converged"] and inc["max_err_vs_direct_solve"] <= 1e-6)}
        # complex fields as the reference carries them: fixed iteration count, timing only
        gen = torch.Generator(device=dev).manual_seed(5000 + rank)
        bz = torch.zeros(op.real_shape + (2,), dtype=torch.float64, device=dev)
        bz[..., 0].normal_(generator=gen)
        bz = torch.view_as_complex(bz)
        op.cg_solve(bz, rtol=0.0, max_iter=2, check_every=0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _, its, _ = op.cg_solve(bz, rtol=0.0, max_iter=20, check_every=0)
        e1.record(stream)
        barrier()
        cms = slab.max_over_ranks(e0.elapsed_time(e1), rdev)
        rec["complex_fields_iterations_per_s"] = its / (cms * 1e-3)
        rec["complex_fields_ms_per_iteration"] = cms / max(its, 1)
        rec["ms_per_iteration"] = inc["ms_total"] / max(inc["iterations"], 1)
        del bz
        op.close()
        out["cg"] = rec
    except Exception as e:
        import traceback
        out["cg"] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
    torch.cuda.empty_cache()
    return out


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
    rdev = dev if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    edge, dim = args.edge, args.dim
    peak, peak_src = peaks()
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
        return slab.max_over_ranks(e0.elapsed_time(e1) / steps, rdev)  # ms

    def make_field(local, seed):
        gen = torch.Generator(device=dev).manual_seed(seed)
        return torch.view_as_complex(torch.randn((dim,) + local + (2,), dtype=torch.float64, device=dev,
                                                 generator=gen))

    # ---- headline = BASELINE configs[2]: the edge^dim grid split into N k0 slabs (strong scaling) ----
    shape = (edge,) * dim
    L = tuple(n * h for n, h in zip(shape, SPACING))
    op = b.ModalOperator(shape, L, MU, NU, device=local_rank)
    if args.variant >= 0:
        op.set_option("apply_variant", args.variant)
    k_begin, local = slab.rank_block(shape, rank, world)
    modes_rank = int(np.prod(local, dtype=np.int64))
    modes_total = int(np.prod(shape, dtype=np.int64))
    u = make_field(local, 3000 + rank)
    f = torch.empty_like(u)

    if args.sweep:                                    # tuning aid; every rank takes part in the collectives
        for v in range(op.info("num_variants")):
            op.set_option("apply_variant", v)
            ms = timed(lambda: op.apply_modal_stiffness(u, out=f, k_begin=k_begin), 20, 3)
            gbs = BYTES_PER_MODE[dim] * modes_rank / ms / 1e6
            if rank == 0:
                print(f"# variant {v:2d} grid {op.info('last_grid')} block {op.info('last_block')}: "
                      f"{ms:.4f} ms  {gbs:.0f} GB/s  {gbs / peak:.3f} of peak", flush=True)
        op.set_option("apply_variant", args.variant)

    launches0 = op.info("launches")
    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(local_rank), "uuid", None))
    ms = timed(lambda: op.apply_modal_stiffness(u, out=f, k_begin=k_begin), args.steps, args.warmup,
               sampler)
    launches = op.info("launches") - launches0 - args.warmup
    launch_cfg = {"kernel_variant": op.info("apply_variant"), "grid": op.info("last_grid"),
                  "block": op.info("last_block"), "smem": op.info("last_smem")}   # of the TIMED launches
    value = modes_total / (ms * 1e-3) / 1e9
    achieved = BYTES_PER_MODE[dim] * modes_rank / (ms * 1e-3) / 1e9      # GB/s per GPU (slowest rank's time)

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

    # ---- weak-scaled companion (N > 1): one edge^dim slab per GPU of an (edge*N, edge, edge) grid ----
    weak = None
    if world > 1 and not args.no_weak:
        del u, f
        torch.cuda.empty_cache()
        w_shape = (edge * world,) + (edge,) * (dim - 1)
        w_L = tuple(n * h for n, h in zip(w_shape, SPACING))
        w_op = b.ModalOperator(w_shape, w_L, MU, NU, device=local_rank)
        w_kb, w_local = slab.rank_block(w_shape, rank, world)
        wu = make_field(w_local, 3100 + rank)
        wf = torch.empty_like(wu)
        w_ms = timed(lambda: w_op.apply_modal_stiffness(wu, out=wf, k_begin=w_kb), args.steps, args.warmup)
        w_modes = int(np.prod(w_local, dtype=np.int64))
        weak = {"workload": f"{'x'.join(map(str, w_shape))} grid, one {edge}^{dim} k0 slab per GPU", "ms_per_step": w_ms,
                "value": w_modes * world / (w_ms * 1e-3) / 1e9, "unit": UNIT, "scaling": "weak",
                "gbs_per_gpu": BYTES_PER_MODE[dim] * w_modes / (w_ms * 1e-3) / 1e9}
        del wu, wf, w_op
        torch.cuda.empty_cache()
        u = make_field(local, 3000 + rank)
        f = None

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hu = torch.empty(u.shape, dtype=u.dtype, pin_memory=True)
        hf = torch.empty(u.shape, dtype=u.dtype, pin_memory=True)
        hu.copy_(u)
        torch.cuda.synchronize()
        f = None                                    # make room for the staging buffers
        op.apply_modal_stiffness_host(hu, out=hf, k_begin=k_begin)          # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            op.apply_modal_stiffness_host(hu, out=hf, k_begin=k_begin)       # synchronous
        torch.cuda.synchronize()
        e_s = (time.perf_counter() - t0) / args.e2e_steps
        e_s = slab.max_over_ranks(e_s, rdev)
        nbytes = u.numel() * 16
        e2e = {"value": modes_total / e_s / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "bytes_are": "per rank (each rank moves its own k0 slab through its own host buffers)",
               "ms_per_step": e_s * 1e3, "steps": args.e2e_steps,
               "api": "bri17_modal_stiffness_apply_host_f64 (pinned host buffers, chunked H2D/kernel/D2H)",
               "pcie_gbs_each_way": nbytes / e_s / 1e9}
        del hu, hf
        try:                                        # what the host link gives to plain copies, same moment, same ranks
            e2e["pcie_probe"] = pcie_probe(torch, dev, lambda v: slab.max_over_ranks(v, rdev))
        except Exception as ex:
            e2e["pcie_probe"] = {"error": repr(ex)}
    del u
    f = None
    torch.cuda.empty_cache()

    # ---- the headline line is complete BEFORE the real-space / CG records run: whatever happens in
    # them (an exception, a poisoned CUDA context, a hang) rank 0 still prints it ----
    line = None
    if rank == 0:
        traffic = None            # DRAM bytes per launch from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath) and edge == {3: 512, 2: 4096}[dim] and world == 1:
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
        ws = modes_rank * 16 * dim
        line = {
            "metric": METRIC if (dim, edge) == (3, 512) else
            f"Gmodes/s (fp64 modal stiffness apply, {dim}D {edge}^{dim}; NOT the BASELINE metric)",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(dim, edge, world),
                       "modes_total": modes_total, "modes_per_gpu": modes_rank, "mu": MU, "nu": NU, "spacing": SPACING,
                       "l2": f"inputs ({ws / 2**30:.2f} GiB) and outputs (same) per GPU vs the 126 MB L2: "
                             + ("far larger, no flush between iterations" if ws > 2**29
                                else "NOT larger than L2 (non-default size: timing is L2-assisted)"),
                       **launch_cfg},
            "gdof_per_s": value * dim,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_unit": "GB per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_source": "committed ncu --set full capture of the same kernel and size "
                                           "(profiles/), not re-measured in this run" if traffic else None,
                         "algorithmic_gb_per_launch": BYTES_PER_MODE[dim] * modes_rank / 1e9,
                         "peak_source": peak_src,
                         "bytes_per_mode": BYTES_PER_MODE[dim], "modes_per_launch": modes_rank,
                         "kernel": f"modal_stiffness_apply_kernel<{dim},...>",
                         "timing": "CUDA events on the launch stream around the timed steps / steps, max over ranks"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "weak_scaling": weak,
            "realspace": None, "cg": None,
        }

    printed = threading.Lock()

    def emit(extra):
        """Rank 0 prints the line exactly once (other ranks: no-op)."""
        if rank == 0 and printed.acquire(blocking=False):
            line.update(extra)
            print(json.dumps(line), flush=True)

    if not args.no_realspace and dim == 3 and edge == 512:
        def give_up():                                # the records hang (a peer died, a flag never arrives)
            emit({"realspace": {"error": f"no result after {args.rs_timeout} s; records abandoned"}, "cg": None})
            os._exit(0)
        watchdog = threading.Timer(args.rs_timeout, give_up)
        watchdog.daemon = True
        watchdog.start()
        try:
            if args.rs_fault == "raise":
                raise RuntimeError("injected fault (--rs-fault raise)")
            if args.rs_fault == "hang":
                time.sleep(10 ** 6)
            extra = realspace_records(args, torch, dist, local_rank, rank, world, dev)
        except BaseException as e:                    # e.g. a sticky CUDA error surfacing outside the records' own handlers
            extra = {"realspace": {"error": repr(e)}, "cg": None}
        watchdog.cancel()
        emit(extra)
        poisoned = any(isinstance(v, dict) and "error" in v for v in extra.values())
        if poisoned:                                  # the CUDA context may be unusable: no collective teardown
            sys.stdout.flush()
            os._exit(0)
    else:
        emit({})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
