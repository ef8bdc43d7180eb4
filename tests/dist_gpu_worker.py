"""torchrun worker (NCCL, one rank per GPU, any world size up to 16): the distributed real-space
operator and CG against the numpy restatement and the reference's known-answer test, for both
exchange modes, serial and pipelined schedules, fused and cuFFT axis-0 paths.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tests/dist_gpu_worker.py [--quick]

Prints DIST_REALSPACE_OK <world> on success (rank 0)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import realspace_checks as rc  # noqa: E402
from bri17_b200.realspace import RealSpaceOperator  # noqa: E402
from oracle import oracle  # noqa: E402
from realspace_ref import real_space_apply_ref  # noqa: E402

MU, NU = 5.6, 0.3


def transforms_and_cg(local, mode, pipeline, shape):
    """Layout of the forward/inverse transforms (axis-1 slabs) and CG on b = A u."""
    dim = len(shape)
    L = rc.spacing_L(shape)
    o = oracle.best()
    rng = np.random.default_rng(100 + dim)
    u = rng.standard_normal((dim,) + shape)
    u -= u.mean(axis=tuple(range(1, dim + 1)), keepdims=True)
    ref = real_space_apply_ref(o, shape, L, MU, NU, u + 0j)
    op = RealSpaceOperator.from_process_group(shape, L, MU, NU, device=local, exchange_mode=mode)
    op.set_option("pipeline", pipeline)
    a0, a1 = op.n0_begin, op.n0_begin + op.n0_count
    ud = torch.from_numpy(np.ascontiguousarray(u[:, a0:a1]) + 0j).cuda()
    err = 0.0
    # forward transform lands in the axis-1 slab layout; immediately followed by a pipelined apply
    # (no host synchronisation in between: the exchange buffer hand-over must be ordered on-stream)
    xh_d = op.forward_fft(ud)
    F = op.apply(ud).cpu().numpy()
    xh = xh_d.cpu().numpy()
    k0, k1 = op.k1_begin, op.k1_begin + op.k1_count
    href = np.fft.fftn(u, axes=tuple(range(1, dim + 1)))[:, :, k0:k1]
    if xh.size:
        err = max(err, np.abs(xh - href).max() / np.abs(href).max())
    if F.size:
        err = max(err, np.abs(F - ref[:, a0:a1]).max() / np.abs(ref).max())
    back = op.inverse_fft(torch.from_numpy(np.ascontiguousarray(href)).cuda()).cpu().numpy()
    if back.size:
        err = max(err, np.abs(back - u[:, a0:a1]).max())
    # CG on b = A u recovers u (zero mean), real and complex fields; a shifted right-hand side
    # (non-zero mean) must give the same solution (null-space projection)
    shift = 0.5 * np.abs(ref).max()
    br = torch.from_numpy(np.ascontiguousarray(ref[:, a0:a1].real) + shift).cuda()
    xr, it_r, res_r = op.cg_solve_real(br, rtol=1e-11, max_iter=3000, check_every=5)
    bc = torch.from_numpy(np.ascontiguousarray(ref[:, a0:a1])).cuda()
    xc, it_c, res_c = op.cg_solve(bc, rtol=1e-11, max_iter=3000, check_every=5)
    cg_err = 0.0
    if xr.numel():
        cg_err = max(np.abs(xr.cpu().numpy() - u[:, a0:a1]).max(), np.abs(xc.cpu().numpy() - u[:, a0:a1]).max()) \
            / np.abs(u).max()
    t = torch.tensor([err, cg_err, res_r, res_c], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    op.close()
    err, cg_err, res_r, res_c = t.tolist()
    return err, cg_err, max(res_r, res_c), it_r


def main():
    quick = "--quick" in sys.argv
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True

    def report(name, value, bound, fmt="{:.2e}"):
        nonlocal ok
        good = value <= bound
        ok &= good
        if rank == 0:
            print(f"{name}: {fmt.format(value)} (bound {bound:g}) {'ok' if good else 'FAIL'}", flush=True)

    # (exchange mode, pipeline, axis-0 path: 1 fused + k1-major layout, 2 fused + natural layout, 0 cuFFT)
    configs = ((1, 1, 1), (1, 0, 1), (0, 0, 1), (1, 1, 0), (1, 1, 2))
    for mode, pipeline, fused in configs[:2] if quick else configs:
        tag = f"mode {mode} pipeline {pipeline} fused {fused}"
        report(f"{tag}: small grids vs numpy restatement", rc.small_grids(local, mode, pipeline, fused), 1e-13)
        for shape in ((9, 7, 5), (16, 12, 10), (12, 10)) + (() if quick else ((32, 32, 32),)):
            err, cg_err, res, iters = transforms_and_cg(local, mode, pipeline, shape)
            report(f"{tag} {shape}: fft layout/apply", err, 1e-13)
            report(f"{tag} {shape}: CG ({iters} it) solution error", cg_err, 1e-7)
            report(f"{tag} {shape}: CG residual", res, 1e-11)
    # sub-slab pipelining forced to 3 and 4 pieces per component on slabs of a few planes (empty and
    # uneven sub-slabs included), fused pass (power-of-two N0) and cuFFT axis-0 path
    for chunks in (3, 4):
        report(f"pipelined apply in {chunks} sub-slabs per component: small grids",
               rc.small_grids(local, 1, 1, 1, chunks=chunks), 1e-13)
    report("pipelined apply in 4 sub-slabs: dense KAT, real fields",
           rc.dense_kat(local, world, mode=1, pipeline=1, real=True, chunks=4), 0.0)
    report("pipelined apply in 3 sub-slabs: 128^3 distributed vs single-GPU slabs",
           rc.vs_single_gpu(local, edge=128, chunks=3), 1e-13)
    # the reference's dense-matrix known-answer test with compute_Ku on all ranks
    # (tests/test_bri17.cpp:130-150); value = worst violation of 1e-15*|e| + 1e-14, 0 = none
    report("dense KAT (3,4,5) + rank-divisible grid, complex fields, fused exchange",
           rc.dense_kat(local, world, mode=1, pipeline=1, real=False), 0.0)
    report("dense KAT, real fields (r2c path)", rc.dense_kat(local, world, mode=1, pipeline=1, real=True), 0.0)
    if not quick:
        report("dense KAT, NCCL exchange", rc.dense_kat(local, world, mode=0, pipeline=0, real=False), 0.0)
    report("256^3 distributed vs single-GPU slabs" if not quick else "128^3 distributed vs single-GPU slabs",
           rc.vs_single_gpu(local, edge=128 if quick else 256), 1e-13)
    # plane waves at a moderate size, both field types
    shape = (128, 96, 80)
    op = RealSpaceOperator.from_process_group(shape, rc.spacing_L(shape), MU, NU, device=local, exchange_mode=1)
    report("plane waves (128,96,80) complex", rc.plane_waves(op, real=False), 1e-12)
    report("plane waves (128,96,80) real", rc.plane_waves(op, real=True), 1e-12)
    op.close()
    # BASELINE config 5 in small: the inclusion problem, CG vs the per-mode direct solve
    shape = (64, 64, 64)
    op = RealSpaceOperator.from_process_group(shape, (1.0, 1.0, 1.0), 1.0, 0.3, device=local, exchange_mode=1)
    inc = rc.inclusion_problem(op, rtol=1e-10, check_every=10)
    op.close()
    report(f"inclusion 64^3: CG ({inc['iterations']} it) vs direct solve", inc["max_err_vs_direct_solve"], 1e-6)
    report("inclusion 64^3: CG residual", inc["rel_residual"], 1e-10)
    if rank == 0:
        print(("DIST_REALSPACE_OK" if ok else "DIST_REALSPACE_FAIL"), world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
