"""Time to a converged solution at full size (VERDICT r01 task 7).

    python scripts/time_to_solution.py [--axis 119] [--rel-tol 1e-8] [--cpu] [--out profiles/...json]

The 3-D Poisson problem of bench.py (BASELINE.json configs[1]: 119^3 x 6 jittered Kuhn tetrahedra = 10 110 954 cells,
Dirichlet mirror ghosts, shuffled then RCM-renumbered, b = A x*, x0 = 0), solved to a RELATIVE residual of --rel-tol by

  * the fused CG and BiCGStab (sb_cg_solve / sb_bicgstab_solve),
  * the reference's own CgSolver / BiCgStabSolver templates on Storm::DeviceVector with the Chebyshev polynomial
    preconditioner in the reference's pre_op slot (Storm::ChebyshevPreconditioner, several degrees),
  * with --cpu: the reference's CgSolver header on a host vector + the face-loop operator, one thread, to the SAME
    tolerance (minutes at full size: this is the CPU side of the time-to-solution ratio).

For every run: iterations, operator applies, wall seconds around the call (stream drained on both sides), the TRUE
relative residual ||b - A x|| / ||b|| recomputed from the returned x, and the relative error against x*. One JSON
object per run on stdout, the collection in --out.
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
from stormruler_b200.mesh import CELL_TET, Mesh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--axis", type=int, default=119)
    ap.add_argument("--rel-tol", type=float, default=1e-8)
    ap.add_argument("--max-iterations", type=int, default=20000)
    ap.add_argument("--degrees", default="4,8,16")
    ap.add_argument("--cpu", action="store_true", help="also run the reference CPU CG to the same tolerance (minutes)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    t0 = time.time()
    mesh = Mesh.box(CELL_TET, args.axis, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)
    mesh.renumber_rcm()
    n = mesh.n_cells
    c = mesh.cell_centers()
    x_star = np.sin(np.pi * c[:, 0]) * np.sin(np.pi * c[:, 1]) * np.sin(np.pi * c[:, 2])
    ctx = sb.Context(0)
    op = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
    b = ctx.zeros(n)
    op.mul(b, ctx.vector(x_star))
    bh = b.numpy()
    b_norm = float(np.linalg.norm(bh))
    print(f"[tts] {n} cells, setup {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
    runs = []

    def finish(name, x, iters, converged, secs, extra):
        y = ctx.zeros(n)
        op.mul(y, x)
        true_res = float(np.linalg.norm(bh - y.numpy()) / b_norm)
        err = float(np.linalg.norm(x.numpy() - x_star) / np.linalg.norm(x_star))
        rec = {"run": name, "cells": int(n), "rel_tol": args.rel_tol, "converged": bool(converged), "iterations": int(iters),
               "seconds": secs, "true_relative_residual": true_res, "relative_error_vs_exact": err, **extra}
        print(json.dumps(rec), flush=True)
        runs.append(rec)

    for name, Solver, applies in (("fused_cg", sb.CgSolver, 1), ("fused_bicgstab", sb.BiCgStabSolver, 2)):
        for rep in range(2):   # the second run is the warm one
            s = Solver(num_iterations=args.max_iterations, absolute_error_tolerance=0.0, relative_error_tolerance=args.rel_tol,
                       use_graph=True, record=False)
            x = ctx.zeros(n)
            ctx.sync()
            t = time.perf_counter()
            conv = s.solve(x, b, op)
            ctx.sync()
            secs = time.perf_counter() - t
        finish(name, x, s.iteration, conv, secs, {"applies": applies * s.iteration + 1, "device_ms": s.solve_ms,
                                                    "path": "fused solver (sb_*_solve)"})
    for solver in ("cg", "bicgstab"):
        for deg in [0] + [int(d) for d in args.degrees.split(",") if d]:
            for rep in range(2):
                x = ctx.zeros(n)
                ctx.sync()
                t = time.perf_counter()
                r = dropin.solve(solver, op, x, b, num_iterations=args.max_iterations, abs_tol=0.0, rel_tol=args.rel_tol,
                                 precond="chebyshev" if deg else None, cheb_degree=deg, trace_cap=64)
                ctx.sync()
                secs = time.perf_counter() - t
            finish(f"template_{solver}" + (f"_chebyshev{deg}" if deg else ""), x, r.iterations, r.converged, secs,
                   {"applies": int(r.n_apply), "path": "reference template on Storm::DeviceVector" +
                    (f" + Storm::ChebyshevPreconditioner(degree {deg}) in the pre_op slot" if deg else "")})
    if args.cpu:
        from oracle import orc
        fm = orc.FaceMesh(mesh.n_cells, mesh.face_cell, mesh.face_area, mesh.face_dist, mesh.cell_vol, mesh.bface_cell,
                          mesh.bface_area, mesh.bface_dist)
        cpu = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
        bc = cpu.apply(x_star)
        t = time.perf_counter()
        r = orc.ref_solve("cg", cpu, bc, num_iterations=args.max_iterations, abs_tol=0.0, rel_tol=args.rel_tol, trace_cap=16)
        secs = time.perf_counter() - t
        true_res = float(np.linalg.norm(bc - cpu.apply(r.x)) / np.linalg.norm(bc))
        rec = {"run": "reference_cpu_cg", "cells": int(n), "rel_tol": args.rel_tol, "converged": bool(r.converged),
               "iterations": int(r.iterations), "seconds": secs, "true_relative_residual": true_res,
               "relative_error_vs_exact": float(np.linalg.norm(r.x - x_star) / np.linalg.norm(x_star)),
               "applies": int(r.n_apply), "cores": 1,
               "path": "the reference's CgSolver header on a host vector + face-loop operator, g++ -O2, one thread"}
        print(json.dumps(rec), flush=True)
        runs.append(rec)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"what": "time to a converged solution (SURVEY.md 8d config 2)", "axis": args.axis, "runs": runs}, f, indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
