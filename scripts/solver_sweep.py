"""Every solver of the reference on the device vector, at full size, against the contract bytes of SURVEY.md 8d.

    python scripts/solver_sweep.py [--axis 119] [--steps 40] [--solvers cg,bicgstab,...] [--out profiles/...json]

One 3-D Poisson problem (BASELINE.json configs[1]: jittered Kuhn-tetrahedral box mesh, Dirichlet mirror ghosts,
shuffled then RCM-renumbered, b = A x*), then for each solver: tolerances 0 so the iteration count is exact, one
warm-up solve, then `--repeats` pairs of timed solves of K and 3K iterations; iterations/s = 2K / (min t_3K - min t_K),
which cancels the initialisation (initial residual, allocations, IDR(s)'s host-generated shadow vectors: 0.3 s) and,
through the minima, the hiccups of a shared host (a first version with single samples and K = 40 was off by up to 10x
for individual solvers: 50 ms of signal against 0.4 s stalls). Wall clock around the C-ABI calls, stream drained
on both sides: the generic path is host-driven, so this is what an application sees. Solvers:

  * generic drop-in: StormRuler's own solver templates instantiated on Storm::DeviceVector
    (cg cgs bicgstab bicgstabl gmres fgmres tfqmr tfqmr1 idrs richardson) -- one kernel per vector statement, one
    fused reduction + host read-back per dot/norm, exactly the reference's statement sequence;
  * fused: fused_cg fused_bicgstab fused_gmres (sb_cg_solve / sb_bicgstab_solve / sb_gmres_solve);
  * grouped: grouped_idrs grouped_bicgstabl (Storm/B200/GroupedSolvers.hpp: the reference algorithms with their statements
    issued as sb_eval_group launches; host scalars).

Contract bytes per iteration (V = 8 N, B = algorithmic bytes of one apply): fused-minimum schedules
for CG (B + 9V), BiCGStab (2B + 15V), GMRES inner step k (B + (4k+6)V); as-written pass counts for the rest
(CGS 2B + 24V, TFQMR 2B + 40V, TFQMR1 2B + 34V, BiCGStab(2) (4B + 63V)/2, IDR(4) (5B + 173V)/4, Richardson B + 7V).
Prints one JSON line per solver and writes a summary.
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
from stormruler_b200 import dropin  # noqa: E402
from stormruler_b200.mesh import CELL_HEX, CELL_TET, Mesh  # noqa: E402

ALL = ["fused_cg", "fused_bicgstab", "fused_gmres", "cg", "bicgstab", "cgs", "bicgstabl", "tfqmr", "tfqmr1", "idrs",
       "gmres", "fgmres", "richardson", "grouped_idrs", "grouped_bicgstabl"]


def contract_bytes(solver: str, B: float, V: float, steps: int, m: int) -> float:
    """Algorithmic bytes of `steps` iterations."""
    # grouped_*: quoted against the same as-written bytes as the reference template they replace, so the two rows
    # compare as iterations/s; what the grouped solver itself moves (31 V / 25.5 V) is in DESIGN.md section 5
    base = solver.replace("fused_", "").replace("grouped_", "")
    if base in ("gmres", "fgmres"):
        ks = np.arange(steps) % m
        return float(np.sum(B + (4 * ks + 6) * V))
    per_it = {"cg": B + 9 * V, "bicgstab": 2 * B + 15 * V, "cgs": 2 * B + 24 * V, "tfqmr": 2 * B + 40 * V,
              "tfqmr1": 2 * B + 34 * V, "bicgstabl": (4 * B + 63 * V) / 2, "idrs": (5 * B + 173 * V) / 4,
              "richardson": B + 7 * V}[base]
    return per_it * steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--axis", type=int, default=119)
    ap.add_argument("--cell", default="tet", choices=["tet", "hex"])
    ap.add_argument("--steps", type=int, default=100, help="K: the timed solves run K and 3K iterations")
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--m", type=int, default=50, help="GMRES restart length")
    ap.add_argument("--solvers", default=",".join(ALL))
    ap.add_argument("--out", default="")
    ap.add_argument("--grouping", nargs="?", const=1, default=0, type=int,
                    help="generic solvers with Storm::B200::set_statement_grouping(true): chain-shaped statements queued "
                         "and launched as one sb_eval_group with the reduction behind them; --grouping 2: + dependency-"
                         "aware scheduling (consumers launch only the statements they depend on)")
    args = ap.parse_args()
    if args.grouping:
        dropin.set_statement_grouping(args.grouping)
    peak = 6550.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    t0 = time.time()
    mesh = Mesh.box(CELL_TET if args.cell == "tet" else CELL_HEX, args.axis, jitter=0.2, seed_jitter=42, shuffle=True,
                    seed_shuffle=43)
    mesh.renumber_rcm()
    n = mesh.n_cells
    c = mesh.cell_centers()
    x_star = np.sin(np.pi * c[:, 0]) * np.sin(np.pi * c[:, 1]) * np.sin(np.pi * c[:, 2])
    ctx = sb.Context(0)
    op = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
    b = ctx.zeros(n)
    op.mul(b, ctx.vector(x_star))
    B, V = float(op.info.algorithmic_bytes_per_apply), 8.0 * n
    print(f"[sweep] {n} {args.cell} cells, setup {time.time() - t0:.1f}s", file=sys.stderr, flush=True)

    def run(solver, iters):
        x = ctx.zeros(n)
        ctx.sync()
        t = time.perf_counter()
        if solver == "fused_cg" or solver == "fused_bicgstab":
            S = sb.CgSolver if solver == "fused_cg" else sb.BiCgStabSolver
            s = S(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, use_graph=True,
                  record=False)
            s.solve(x, b, op)
            it, err = s.iteration, s.absolute_error
        elif solver == "fused_gmres":
            s = sb.GmresSolver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=0.0,
                               num_inner_iterations=args.m, record=False)
            s.solve(x, b, op)
            it, err = s.iteration, s.absolute_error
        else:
            r = dropin.solve(solver, op, x, b, num_iterations=iters, abs_tol=0.0, rel_tol=0.0,
                             num_inner=args.m if solver in ("gmres", "fgmres") else 0, trace_cap=64)
            it, err = r.iterations, r.abs_err
        ctx.sync()
        assert it == iters, (solver, it, iters)
        return time.perf_counter() - t, err

    K = args.steps
    points = []
    for solver in args.solvers.split(","):
        try:
            run(solver, min(K, 10))
            t1 = t3 = float("inf")
            for _ in range(args.repeats):
                t1 = min(t1, run(solver, K)[0])
                t, err = run(solver, 3 * K)
                t3 = min(t3, t)
        except Exception as e:  # keep the sweep going: one solver's failure is a data point, not the end
            pt = {"solver": solver, "error": str(e)[:300]}
            print(json.dumps(pt), flush=True)
            points.append(pt)
            continue
        secs = max(t3 - t1, 1e-9)
        alg = contract_bytes(solver, B, V, 3 * K, args.m) - contract_bytes(solver, B, V, K, args.m)
        pt = {"solver": solver, "cells": int(n), "iterations_timed": 2 * K, "iterations_per_sec": 2 * K / secs,
              "ms_per_iteration": 1e3 * secs / (2 * K), "contract_bytes_per_iteration": alg / (2 * K),
              "contract_gbs": alg / secs / 1e9, "frac_of_measured_peak": alg / secs / 1e9 / peak,
              "frac_of_nominal_8TBs": alg / secs / 8e12, "residual_after_3K": err,
              "path": "fused" if solver.startswith("fused_") else
                      ("statement groups (Storm::B200 grouped solver)" if solver.startswith("grouped_") else
                       "reference template on Storm::DeviceVector")}
        print(json.dumps(pt), flush=True)
        points.append(pt)
        if args.out:  # rewritten after every solver: a cut-off run keeps what it measured
            with open(args.out, "w") as f:
                json.dump({"what": "solver sweep (SURVEY.md 8d contract bytes)", "statement_grouping": int(args.grouping), "cell": args.cell, "cells": int(n),
                           "K": K, "repeats": args.repeats, "restart": args.m, "peak_gbs": peak, "points": points}, f, indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
