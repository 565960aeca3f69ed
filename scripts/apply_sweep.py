"""BASELINE.json configs[4] / SURVEY.md 8d config 5: operator-apply (SpMV) bandwidth sweep against the HBM roofline.

    python scripts/apply_sweep.py [--cells tet,hex,poly] [--sizes 1e5,3e5,1e6,3e6,1e7,3e7] [--out profiles/...json]
    torchrun --nproc-per-node N scripts/apply_sweep.py ...      (N > 1: METIS-partitioned, halo exchange inside the apply)

For every (cell kind, size): synthetic jittered box mesh of about that many cells, shuffled then RCM-renumbered,
3-D Poisson with Dirichlet mirror ghosts, coefficient form. Timed: 40 chained applies y = A x, z = A y, ... (the
input of every apply is freshly written, as inside a Krylov iteration) bracketed by stream synchronisation, after 8
warm-up applies; GB/s = algorithmic bytes (24 N + 12 entries) / time. One JSON line per point and a
summary object at the end. `poly` = truncated octahedra (the Voronoi cells of a body-centred cubic lattice: 14 faces
per cell, F ~ 7 N, the 14-wide instantiation of the apply kernel), ingested as a face list and RCM-renumbered.
`hexlat` = uniform hexahedra in lattice order, face list generated directly (no node matching, no shuffle / RCM): the
cheap way to the 1e8 and 2e8-cell points (about 15 s of host time per 1e8 cells). Under torchrun every rank builds only
its own slab of the lattice (mesh.HexLatticeSlab: no global mesh, no partitioner), so 2e8 cells on 8 GPUs cost each rank
what 2.5e7 cells cost one.
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
from stormruler_b200.mesh import CELL_HEX, CELL_TET, HexLattice, HexLatticeSlab, Mesh, PolyMesh  # noqa: E402


def axis_for(kind, cells):
    per = {"tet": 6, "hex": 1, "poly": 2, "hexlat": 1}[kind]
    return max(2, round((cells / per) ** (1.0 / 3.0)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", default="tet,hex")
    ap.add_argument("--sizes", default="1e5,3e5,1e6,3e6,1e7,3e7")
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    peak = 6550.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    mg = dist = None
    if world > 1:
        from stormruler_b200 import multigpu as mg
        dist = mg.init_process_group(cuda=True)
    points = []
    for kind in args.cells.split(","):
        for size in (float(s) for s in args.sizes.split(",")):
            n_axis = axis_for(kind, size)
            t0 = time.time()
            if kind == "poly":
                mesh = PolyMesh.bcc(n_axis, stretch=(1.0, 1.3, 0.7)).to_mesh()   # face-list handle (sb_mesh_from_faces)
                mesh.renumber_rcm()
            elif kind == "hexlat" and world > 1:
                mesh = None                                # rank-local: nobody holds the global mesh
                slab = HexLatticeSlab(n_axis, n_axis, n_axis, rank, world)
            elif kind == "hexlat":
                mesh = HexLattice(n_axis)                  # lattice order (bandwidth n^2), generated directly
            else:
                mesh = Mesh.box(CELL_TET if kind == "tet" else CELL_HEX, n_axis, jitter=0.2, seed_jitter=42, shuffle=True,
                                seed_shuffle=43)
                mesh.renumber_rcm()
            t_mesh = time.time() - t0
            if world > 1:
                if mesh is None:
                    loc, vec_capacity = slab.local, slab.info()["vec_capacity"]
                    n_global, n_faces_global = n_axis ** 3, 3 * n_axis * n_axis * (n_axis - 1)
                else:
                    part = mg.partition_mesh(mesh, world, capi.PART_METIS)
                    loc, vec_capacity = part.local(rank), part.info.vec_capacity
                    n_global, n_faces_global = mesh.n_cells, mesh.n_faces
                ctx = mg.DistContext(local_rank, rank, world, vec_capacity, n_vectors=6)
                op = mg.DistOperator(ctx, loc, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
                n_loc = loc.n_owned
            else:
                ctx = sb.Context(local_rank)
                op = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
                n_loc = n_global = mesh.n_cells
                n_faces_global = mesh.n_faces
            rng = np.random.default_rng(rank)
            x, y = ctx.vector(rng.standard_normal(n_loc) * 1e-3), ctx.zeros(n_loc)

            def chain(k):
                a, b = x, y
                for _ in range(k):
                    op.mul(b, a)
                    a, b = b, a

            chain(8)
            ctx.sync()
            if dist:
                dist.barrier()
            t = time.perf_counter()
            chain(args.reps)
            ctx.sync()
            dt_ = (time.perf_counter() - t) / args.reps
            alg = float(op.info.algorithmic_bytes_per_apply)
            if dist:
                dt_ = mg.max_over_ranks(dt_)
                alg = mg.sum_over_ranks(alg)
            pt = {"cell": kind, "cells": int(n_global), "n_axis": int(n_axis), "n_gpus": world,
                  "faces_per_cell": round(2.0 * n_faces_global / n_global, 3), "width": int(op.info.width),
                  "us_per_apply": dt_ * 1e6, "applies_per_sec": 1.0 / dt_, "algorithmic_bytes": alg,
                  "gbs": alg / dt_ / 1e9, "frac_of_measured_peak": alg / dt_ / 1e9 / (peak * world),
                  "frac_of_nominal_8TBs": alg / dt_ / (8e12 * world), "mesh_build_s": round(t_mesh, 1)}
            if rank == 0:
                print(json.dumps(pt), flush=True)
            points.append(pt)
            del op, x, y
            ctx.close()
            del mesh
    if rank == 0:
        summary = {"what": "operator-apply bandwidth sweep (SURVEY.md 8d config 5)", "peak_gbs_per_gpu": peak,
                   "n_gpus": world, "reps": args.reps, "points": points}
        if args.out:
            with open(args.out, "w") as f:
                json.dump(summary, f, indent=1)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
