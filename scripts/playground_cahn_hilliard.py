"""The reference's one runnable application on the device path: `main` + `cahn_hilliard_solve` of
source_apps/playground/Playground.cpp (:248-255 mesh load, :176-210 time loop), with the three type/spelling
changes of INTEGRATION.md section 8 and nothing else:

    read the 2-D Triangle mesh (.node/.edge/.ele)                   sb_mesh_read_tetgen_2d  (= read_mesh_from_tetgen + assign)
    c[cell] = rand() / RAND_MAX   (glibc, default seed)             host, uploaded once
    for time = 0, 1, ...:
        if time != 0:  cahn_hilliard_step(mesh, c, c_hat, w_hat); swap(c, c_hat)     dropin_cahn_hilliard_step (C++23 drop-in:
                                                                    the reference's CgSolver template on DeviceVector,
                                                                    B200::map for dF/dc, B200::div_grad for stormDivGrad)
        printf("time = %f\\n", total_time); save_vtk(mesh, "out/fields-<time>.vtk", {{"c", 0, &c}})   sb_mesh_write_vtk

    python scripts/playground_cahn_hilliard.py --mesh /path/to/step.1 [--steps 3] [--out out]
    python scripts/playground_cahn_hilliard.py --generate 96 64 [--steps 3]      (writes its own Triangle files first)
    python scripts/playground_cahn_hilliard.py --box 119 --steps 1 --max-iterations 100 --no-vtk [--grouping]
        the same time step on a synthetic 3-D mesh (119^3 x 6 = 10.1 M jittered Kuhn tetrahedra, RCM-renumbered): what the
        playground's caller costs at the size the Krylov benchmarks run at (per CG iteration: 2 face-ordered
        accumulating applies, 2 + 3 element-wise statements, 2 dots)

Without --renumber the cell and face order is the file's, and the iterates are those of the reference with the GPU
reduction tree (bit-identical to the oracle, tests/test_gpu_playground.py); --renumber applies RCM first (better
gather locality on large meshes; the face order, hence the last bits, change). The reference's CG does not converge on
this affine operator within its 2000-iteration cap (relative residual ~3e-3 on square_nb.1): that is reproduced, not
repaired; --max-iterations bounds the run time of a demonstration. --uniformed hands the operator to the reference's own
solve_non_uniform (Solver.hpp:271-292) instead, which is made for affine operators: CG then converges in about 50
iterations per step (bit-identical to the reference's run of that variant, tests/golden/cahn_hilliard_uniformed_*.npz).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import stormruler_b200 as sb  # noqa: E402
from stormruler_b200 import dropin  # noqa: E402
from stormruler_b200.mesh import Mesh  # noqa: E402


def write_triangle_files(prefix: str, nx: int, ny: int, lx: float = 2.0, ly: float = 2.0) -> None:
    """A structured triangulation of [0,lx]x[0,ly] in the file grammar the reference reads (IoTetgen.hpp:44-235):
    nx*ny squares split along the (i,j)-(i+1,j+1) diagonal into two counter-clockwise triangles; every edge listed once,
    label 0 inside, 1..4 on the bottom / right / top / left side."""
    node = lambda i, j: j * (nx + 1) + i  # noqa: E731
    with open(prefix + ".node", "w") as f:
        f.write(f"{(nx + 1) * (ny + 1)}  2  0  1\n")
        for j in range(ny + 1):
            for i in range(nx + 1):
                on_boundary = int(i in (0, nx) or j in (0, ny))
                f.write(f"{node(i, j)}  {lx * i / nx!r}  {ly * j / ny!r}  {on_boundary}\n")
    edges = []
    for j in range(ny + 1):
        for i in range(nx):
            edges.append((node(i, j), node(i + 1, j), 1 if j == 0 else (3 if j == ny else 0)))
    for j in range(ny):
        for i in range(nx + 1):
            edges.append((node(i, j), node(i, j + 1), 4 if i == 0 else (2 if i == nx else 0)))
    for j in range(ny):
        for i in range(nx):
            edges.append((node(i, j), node(i + 1, j + 1), 0))
    with open(prefix + ".edge", "w") as f:
        f.write(f"{len(edges)}  1\n")
        for k, (a, b, label) in enumerate(edges):
            f.write(f"{k}  {a}  {b}  {label}\n")
    with open(prefix + ".ele", "w") as f:
        f.write(f"{2 * nx * ny}  3  0\n")
        k = 0
        for j in range(ny):
            for i in range(nx):
                f.write(f"{k}  {node(i, j)}  {node(i + 1, j)}  {node(i + 1, j + 1)}\n")
                f.write(f"{k + 1}  {node(i, j)}  {node(i + 1, j + 1)}  {node(i, j + 1)}\n")
                k += 2


def initial_condition(n: int) -> np.ndarray:
    """c[cell] = (1.0 * rand()) / RAND_MAX in cell order (Playground.cpp:182-184): glibc's rand() from its default state."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    rand_max = 2147483647
    return np.array([(1.0 * libc.rand()) / rand_max for _ in range(n)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="", help="prefix of <prefix>.node/.edge/.ele (e.g. tests/_data/mesh/step.1)")
    ap.add_argument("--generate", type=int, nargs=2, metavar=("NX", "NY"), help="write a structured Triangle mesh first")
    ap.add_argument("--box", type=int, default=0, help="synthetic 3-D tetrahedral box mesh with this many hexes per axis")
    ap.add_argument("--grouping", nargs="?", const=1, default=0, type=int,
                    help="Storm::B200::set_statement_grouping(true); 2: + dependency-aware scheduling")
    ap.add_argument("--steps", type=int, default=3, help="time steps after the initial output (the playground runs 200000)")
    ap.add_argument("--out", default="out")
    ap.add_argument("--renumber", action="store_true", help="RCM-renumber the cells before the upload")
    ap.add_argument("--max-iterations", type=int, default=0, help="CG iteration cap per step (0 = the reference's 2000)")
    ap.add_argument("--uniformed", action="store_true",
                    help="hand the (affine) operator to the reference's solve_non_uniform instead of solve<CgSolver>: "
                         "not what the playground does, but what makes its CG converge (~50 iterations per step)")
    ap.add_argument("--no-vtk", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    prefix = args.mesh
    if args.generate:
        prefix = os.path.join(args.out, f"generated_{args.generate[0]}x{args.generate[1]}.1")
        write_triangle_files(prefix, *args.generate)
    if not prefix and not args.box:
        ap.error("give --mesh PREFIX, --generate NX NY or --box N")
    t0 = time.time()
    if args.box:
        from stormruler_b200.mesh import CELL_TET
        mesh = Mesh.box(CELL_TET, args.box, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)
        mesh.renumber_rcm()
        n = mesh.n_cells
        c_host = np.random.default_rng(1).random(n)
    else:
        mesh = Mesh.read_tetgen_2d(prefix[:-1] if prefix.endswith(".") else prefix)
        n = mesh.n_cells
        c_host = initial_condition(n)
        if args.renumber:
            perm = mesh.renumber_rcm()
            c_host = c_host[perm]
    if args.grouping:
        dropin.set_statement_grouping(args.grouping)
    print(f"mesh has {mesh.n_faces + mesh.n_bfaces} faces\nmesh has {n} cells\nmesh loaded ({time.time() - t0:.2f} s)", flush=True)

    ctx = sb.Context(0)
    faces = sb.FvmOperator(ctx, mesh, prefill=0, dt=0.0, form=sb.FORM_FAITHFUL)   # interior faces: homogeneous Neumann
    c, c_hat, w_hat = ctx.vector(c_host), ctx.zeros(n), ctx.zeros(n)
    total_time, records = 0.0, []
    for step in range(args.steps + 1):
        if step != 0:
            ctx.sync()
            start = time.perf_counter()
            res = dropin.cahn_hilliard_step(faces, c, c_hat, w_hat, num_iterations=args.max_iterations,
                                            uniformed=args.uniformed)
            ctx.sync()
            total_time += time.perf_counter() - start
            c, c_hat = c_hat, c                                                   # std::swap(c, c_hat)
            records.append({"step": step, "cg_iterations": res.iterations, "converged": res.converged,
                            "abs_err": res.abs_err, "rel_err": res.rel_err, "operator_evaluations": res.n_apply})
        print("time = %f" % total_time, flush=True)
        if not args.no_vtk:
            mesh.write_vtk(os.path.join(args.out, f"fields-{step:05d}.vtk"), {"c": c.numpy()})
    final = c.numpy()
    cg_its = sum(r["cg_iterations"] for r in records)
    print(json.dumps({"app": "playground cahn_hilliard_solve on the device path", "cells": n, "steps": args.steps,
                      "statement_grouping": int(args.grouping), "seconds": total_time,
                      "cg_iterations_per_sec": cg_its / total_time if total_time > 0 else None, "c_min": float(final.min()), "c_max": float(final.max()),
                      "c_mean": float(final.mean()), "per_step": records}), flush=True)
    del faces, c, c_hat, w_hat
    ctx.close()


if __name__ == "__main__":
    main()
