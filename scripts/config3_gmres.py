"""BASELINE.json configs[2] / SURVEY.md 8d config 3: restarted GMRES(m) / FGMRES on 3-D convection-diffusion
(non-symmetric upwind rows), synthetic tetrahedral box mesh, 1 GPU or METIS-partitioned over N GPUs.

    python scripts/config3_gmres.py [--axis 119] [--solver gmres|fgmres|bicgstab|idrs|tfqmr] [--m 50] [--steps 100]
    torchrun --nproc-per-node N scripts/config3_gmres.py ...

The solver is StormRuler's own GmresSolver template instantiated on Storm::DeviceVector (the C++23 drop-in,
stormruler_b200/host): every vector statement is a CUDA kernel, every dot/norm a fused reduction (plus, at N > 1, the
in-kernel NVLink all-reduce), the operator apply carries the halo exchange. Tolerances are 0 so exactly the requested
iterations run (an iteration = one inner Arnoldi step). Timing: --repeats pairs of solves of K = --steps and 3K
iterations, iterations/s = 2K / (min t_3K - min t_K): the per-solve setup (the generic template allocates and frees
its m + 1 basis vectors, 4 GB at 10 M cells, inside every solve() -- about 0.4 s of cudaMalloc/cudaFree) cancels, the
minima remove host hiccups; `solve_seconds_3K` keeps the whole-solve wall time next to it. Operator: -nu lap u + div(beta u), beta = (1, 0.5, 0.25),
nu chosen for a cell Peclet number |beta| h / nu = 2. Prints one JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import stormruler_b200 as sb  # noqa: E402
from stormruler_b200 import capi, dropin  # noqa: E402
from stormruler_b200.mesh import CELL_TET, Mesh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--axis", "--n", dest="n", type=int, default=119, help="hexes per axis (use --axis under torchrun: --n is ambiguous to its parser)")
    ap.add_argument("--solver", default="gmres")
    ap.add_argument("--m", type=int, default=50, help="restart length (Solver.hpp:159 default 50)")
    ap.add_argument("--steps", type=int, default=100, help="K: timed solves run K and 3K iterations")
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--path", default="fused", choices=["fused", "generic"],
                    help="fused: sb_gmres_solve (device-resident Arnoldi); generic: the reference template on DeviceVector")
    ap.add_argument("--converge", action="store_true", help="also solve to rel 1e-8 and report iterations + true residual")
    ap.add_argument("--precond", default="none", choices=["none", "jacobi", "chebyshev"],
                    help="generic path: preconditioner in the reference's pre_op slot (FGMRES differs from GMRES only with one: "
                         "SolverGmres.hpp:149-156,233-248)")
    ap.add_argument("--pre-side", default="right", choices=["left", "right"])
    ap.add_argument("--cheb-degree", type=int, default=4)
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    mesh = Mesh.box(CELL_TET, args.n, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)
    mesh.renumber_rcm()
    beta = (1.0, 0.5, 0.25)
    h = 1.0 / args.n
    nu = float(np.linalg.norm(beta)) * h / 2.0
    fu, bu = mesh.face_flux(beta)
    c = mesh.cell_centers()
    x_star = np.sin(np.pi * c[:, 0]) * np.sin(np.pi * c[:, 1]) * np.sin(np.pi * c[:, 2])
    mg = dist = None
    if world > 1:
        from stormruler_b200 import multigpu as mg
        dist = mg.init_process_group(cuda=True)
        part = mg.partition_mesh(mesh, world, capi.PART_METIS)
        loc = part.local(rank)
        ctx = mg.DistContext(local_rank, rank, world, part.info.vec_capacity, n_vectors=2 * args.m + 16)
        op = mg.DistConvDiffOperator(ctx, loc, nu, fu, bu)
        n_loc, owned = loc.n_owned, loc.owned_global
    else:
        ctx = sb.Context(local_rank)
        op = sb.ConvDiffOperator(ctx, mesh, nu, fu, bu)
        n_loc, owned = mesh.n_cells, slice(None)
    xs = ctx.vector(x_star[owned])
    b = ctx.zeros(n_loc)
    op.mul(b, xs)

    def run(iters, rel_tol):
        x = ctx.zeros(n_loc)
        ctx.sync()
        if dist:
            dist.barrier()
        t = time.perf_counter()
        if args.path == "fused" and args.solver in ("gmres", "fgmres") and args.precond == "none":
            g = sb.GmresSolver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=rel_tol,
                               num_inner_iterations=args.m, record=False)
            conv = g.solve(x, b, op)
            r = dropin.Result(conv, g.iteration, g.absolute_error, g.relative_error, g.history, g.trace, -1)
        else:
            r = dropin.solve(args.solver, op, x, b, num_iterations=iters, abs_tol=0.0, rel_tol=rel_tol, num_inner=args.m,
                             precond=None if args.precond == "none" else args.precond, pre_side=args.pre_side,
                             cheb_degree=args.cheb_degree, trace_cap=64)
        ctx.sync()
        dt_ = time.perf_counter() - t
        if dist:
            dt_ = mg.max_over_ranks(dt_)
        return r, x, dt_

    run(min(args.steps, 20), 0.0)                  # warm-up: allocations, kernel modules
    K = args.steps
    t1 = t3 = float("inf")
    for _ in range(args.repeats):
        r, x, t = run(K, 0.0)
        assert r.iterations == K
        t1 = min(t1, t)
        r, x, t = run(3 * K, 0.0)
        assert r.iterations == 3 * K
        t3 = min(t3, t)
    secs = max(t3 - t1, 1e-9)
    alg_apply = float(op.info.algorithmic_bytes_per_apply)
    if dist:
        alg_apply = mg.sum_over_ranks(alg_apply)
    N = mesh.n_cells
    # contract figure: inner step k moves B_apply + (4k+6) V
    ks = np.arange(K, 3 * K) % args.m              # the inner indices of the 2K timed iterations
    alg = float(np.sum(alg_apply + (4 * ks + 6) * 8.0 * N)) if args.solver in ("gmres", "fgmres") else None
    line = {"config": "config 3: convection-diffusion (upwind, non-symmetric), Peclet 2", "solver": args.solver,
            "restart": args.m, "cells": int(N), "n_gpus": world, "steps": 2 * K, "seconds": secs,
            "iterations_per_sec": 2 * K / secs, "solve_seconds_K": t1, "solve_seconds_3K": t3, "repeats": args.repeats,
            "applies": int(r.n_apply),
            "algorithmic_gbs": (alg / secs / 1e9) if alg else None,
            "frac_of_nominal_8TBs": (alg / secs / (8e12 * world)) if alg else None,
            "residual_after_steps": r.abs_err, "preconditioner": args.precond, "pre_side": args.pre_side,
            "path": "sb_gmres_solve (fused: device-resident Arnoldi, host Givens)" if (args.path == "fused" and args.solver in ("gmres", "fgmres") and args.precond == "none")
            else "reference solver template on Storm::DeviceVector (generic drop-in)"}
    if args.converge:
        rc, xc, sc = run(5000, 1e-8)
        xg = mg.gather_global(loc, xc.numpy(), N) if dist else xc.numpy()
        line.update({"converged": bool(rc.converged), "iterations_to_1e-8": int(rc.iterations), "solve_seconds": sc,
                     "rel_error_vs_exact": float(np.linalg.norm(xg - x_star) / np.linalg.norm(x_star))})
    if rank == 0:
        print(json.dumps(line), flush=True)
    del op, xs, b, x
    ctx.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
