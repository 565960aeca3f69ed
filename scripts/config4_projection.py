"""BASELINE.json configs[3] / SURVEY.md 8d config 4 and 8f rank 4: a time-step driver around repeated solves --
the pressure-Poisson projection step of an incompressible flow solver: pure-Neumann (singular, consistent) Poisson
problem on a hexahedral mesh, fused CG warm-started from the previous step's pressure, 10 synthetic time steps.

    python scripts/config4_projection.py [--axis 160] [--steps 10] [--rel-tol 1e-8]
    torchrun --nproc-per-node N scripts/config4_projection.py --axis 368        (49.8 M hexes: the full config)
    torchrun --nproc-per-node N scripts/config4_projection.py --axis 368 --lattice
        the same mesh in lattice cell order, split into contiguous slabs; every rank builds ITS OWN local mesh from
        the lattice arithmetic (mesh.HexLatticeSlab: seconds and memory proportional to the slab, bit-identical to
        partitioning the global mesh with SB_PART_SLAB) instead of rank 0 building, renumbering and METIS-
        partitioning 49.8 M cells (minutes, 29 GB)

The reference has no incompressible Navier-Stokes code (README.md:29 claims it, the tree does not contain it), so the
right-hand sides are synthetic: at step k the pressure p*_k(x) = cos(pi x) cos(pi y) cos(pi z) cos(t_k)
+ 0.3 cos(2 pi x) cos(pi y) sin(t_k) (homogeneous Neumann on the unit box) is sampled at the cell centres, b_k = A p*_k
is formed on the device and projected onto the range of A (sum_i V_i b_i = 0: the rows are scaled by 1/V_i, so the
left null vector is the cell volumes). The mesh is NOT jittered: with equal cell volumes the row-scaled operator is
exactly symmetric, which CG needs. What is timed per step is what an application pays: upload of the new right-hand
side, projection (one fused dot + one update), the CG solve from the previous pressure, and the download of the result.
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
from stormruler_b200 import capi  # noqa: E402
from stormruler_b200.mesh import CELL_HEX, Mesh  # noqa: E402


def p_star(c, t):
    return (np.cos(np.pi * c[:, 0]) * np.cos(np.pi * c[:, 1]) * np.cos(np.pi * c[:, 2]) * np.cos(t)
            + 0.3 * np.cos(2 * np.pi * c[:, 0]) * np.cos(np.pi * c[:, 1]) * np.sin(t))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--axis", type=int, default=160, help="hexes per axis (368 -> 49.8 M cells, the full config)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--rel-tol", type=float, default=1e-8)
    ap.add_argument("--max-iterations", type=int, default=20000)
    ap.add_argument("--lattice", action="store_true", help="lattice cell order + slab partition, rank-local mesh build")
    ap.add_argument("--vtk", default="", help="prefix: write <prefix>-<step, 5 digits>.vtk (pressure and exact pressure) "
                                              "after every step, outside the timed region (rank 0)")
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    t0 = time.time()

    def build():
        m = Mesh.box(CELL_HEX, args.axis, jitter=0.0, seed_jitter=42, shuffle=True, seed_shuffle=43)
        m.renumber_rcm()
        return m

    mg = dist = None
    lattice_centers = None
    if args.lattice:
        from stormruler_b200.mesh import HexLattice, HexLatticeSlab
        assert not args.vtk, "--vtk needs the node-based mesh handle"
        ax = args.axis
        N = ax ** 3
        ids = np.arange(N)
        lattice_centers = lambda: np.stack([(ids % ax + 0.5) / ax, ((ids // ax) % ax + 0.5) / ax,  # noqa: E731
                                            (ids // (ax * ax) + 0.5) / ax], axis=1)
    if args.lattice and world > 1:
        from stormruler_b200 import multigpu as mg
        dist = mg.init_process_group(cuda=True)
        slab = HexLatticeSlab(ax, ax, ax, rank, world)
        loc, pinfo = slab.local, slab.info()
        mesh = None
        t_mesh = time.time() - t0
        ctx = mg.DistContext(local_rank, rank, world, pinfo["vec_capacity"], n_vectors=14)
        op = mg.DistOperator(ctx, loc, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=False)
        n_loc, centers_loc, vol_loc = loc.n_owned, slab.owned_centers(), np.full(loc.n_owned, slab.cell_vol_value)
    elif args.lattice:
        mesh = HexLattice(ax)
        t_mesh = time.time() - t0
        ctx = sb.Context(local_rank)
        op = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=False)
        n_loc, centers_loc, vol_loc = N, mesh.cell_centers(), np.asarray(mesh.cell_vol)
    elif world > 1:
        # rank 0 alone builds and partitions the global mesh and ships the local meshes (at 49.8 M cells the global
        # build peaks at 29 GB per process: the other ranks never hold it)
        from stormruler_b200 import multigpu as mg
        dist = mg.init_process_group(cuda=True)
        loc, pinfo, fields = mg.scatter_mesh(build, world, capi.PART_METIS,
                                             cell_fields=lambda m: {"centers": m.cell_centers(), "vol": np.asarray(m.cell_vol)})
        mesh = pinfo["mesh"]                            # rank 0 only
        N = pinfo["n_cells"]
        t_mesh = time.time() - t0
        ctx = mg.DistContext(local_rank, rank, world, pinfo["vec_capacity"], n_vectors=14)
        op = mg.DistOperator(ctx, loc, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=False)
        n_loc, centers_loc, vol_loc = loc.n_owned, fields["centers"], fields["vol"]
    else:
        mesh = build()
        N = mesh.n_cells
        t_mesh = time.time() - t0
        ctx = sb.Context(local_rank)
        op = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=False)
        n_loc, centers_loc, vol_loc = N, mesh.cell_centers(), np.asarray(mesh.cell_vol)
    centers = (lattice_centers() if args.lattice else mesh.cell_centers()) if rank == 0 else None
    vol = ctx.vector(vol_loc)
    ones = ctx.zeros(n_loc).fill(1.0)
    vol_total = ctx.dot(vol, ones)                     # global sum (all-reduced when distributed)
    p, b, ps = ctx.zeros(n_loc), ctx.zeros(n_loc), ctx.zeros(n_loc)
    v = sb.expr.v
    records, total_it, total_s = [], 0, 0.0
    for k in range(args.steps):
        t_k = 0.1 * k
        exact = p_star(centers_loc, t_k)
        ctx.sync()
        if dist:
            dist.barrier()
        t = time.perf_counter()
        ps.upload(exact)                               # the application's new data for this step (H2D)
        op.mul(b, ps)                                  # synthetic right-hand side b_k = A p*_k
        shift = ctx.dot(vol, b) / vol_total            # consistency: sum_i V_i b_i = 0
        (v(b) - shift * v(ones)).assign_to(b)
        s = sb.CgSolver(num_iterations=args.max_iterations, absolute_error_tolerance=0.0,
                        relative_error_tolerance=args.rel_tol, use_graph=True, record=False)
        conv = s.solve(p, b, op)                       # warm start: p holds the previous step's pressure
        result = p.numpy()                             # D2H of the step's result
        dt_ = time.perf_counter() - t
        if dist:
            dt_ = mg.max_over_ranks(dt_)
        # error against p*_k up to the constant the singular problem leaves free
        full = mg.gather_global(loc, result, N) if dist else result
        err = None
        if rank == 0:
            ex_full = p_star(centers, t_k)
            d = (full - full.mean()) - (ex_full - ex_full.mean())
            err = float(np.linalg.norm(d) / np.linalg.norm(ex_full - ex_full.mean()))
            if args.vtk:   # the playground's output step (Playground.cpp:205-208: save_vtk after the step)
                mesh.write_vtk(f"{args.vtk}-{k:05d}.vtk", {"p": full, "p_exact": ex_full})
        records.append({"step": k, "iterations": int(s.iteration), "converged": bool(conv), "seconds": dt_,
                        "solve_ms_device": float(s.solve_ms), "rel_residual": float(s.relative_error),
                        "rel_error_vs_p_star": err, "rhs_shift": float(shift)})
        total_it += int(s.iteration)
        total_s += dt_
    line = {"config": "config 4: pressure-Poisson projection (pure Neumann, CG, warm start), hexahedra", "cells": int(N),
            "mesh": "lattice order, slab partition, rank-local build" if args.lattice else "shuffled + RCM, METIS",
            "n_gpus": world, "steps": args.steps, "rel_tol": args.rel_tol, "total_iterations": total_it,
            "total_seconds": total_s, "iterations_per_sec_e2e": total_it / max(total_s, 1e-12),
            "mesh_build_s": round(t_mesh, 1), "per_step": records}
    if rank == 0:
        print(json.dumps(line), flush=True)
    del op, vol, ones, p, b, ps
    ctx.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
