#!/usr/bin/env python
"""scale_ab.py -- A/B of the fused solvers' schedules and tuning bits on ONE problem build.

    python scripts/scale_ab.py --axis 59 [--solver bicgstab,cg] [--variants default,off,...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/scale_ab.py --axis 119

Building, renumbering and partitioning the 10 M-cell mesh costs far more box time than the solves, so every variant
(schedule x SB_TUNE_* bits) is timed in the same process on the same operator: W warm-up + K timed iterations in graph
replay (device time of the iteration loop, max over ranks), then the same K iterations with per-kernel events for the
slot times and the in-kernel waits (halo flags, other ranks' sums). One JSON line per (solver, variant) on rank 0,
plus a table on stderr. Every variant's residual after K iterations must equal the first one's bit for bit.
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

from stormruler_b200 import capi  # noqa: E402

T = capi
VARIANTS = {
    # name: (schedule, tuning)
    "default": (T.SCHEDULE_AUTO, 0),
    "off": (T.SCHEDULE_STEPWISE, T.TUNE_OFF),
    "noack": (T.SCHEDULE_STEPWISE, T.TUNE_NO_ACK),
    "push": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE),
    "stream": (T.SCHEDULE_STEPWISE, T.TUNE_STREAM_OPERATOR),
    "pdlf": (T.SCHEDULE_STEPWISE, T.TUNE_PDL_FINAL),
    "pdlfa": (T.SCHEDULE_STEPWISE, T.TUNE_PDL_FINAL | T.TUNE_PDL_AFTER_FINAL),
    "pdlall": (T.SCHEDULE_STEPWISE, T.TUNE_PDL_FINAL | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
    "push+stream": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_STREAM_OPERATOR),
    "push+pdlfa": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PDL_FINAL | T.TUNE_PDL_AFTER_FINAL),
    "push+pdlall": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PDL_FINAL | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
    "push+stream+pdlfa": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_STREAM_OPERATOR | T.TUNE_PDL_FINAL | T.TUNE_PDL_AFTER_FINAL),
    "push+stream+pdlall": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_STREAM_OPERATOR | T.TUNE_PDL_FINAL
                           | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
    "stream+pdlfa": (T.SCHEDULE_STEPWISE, T.TUNE_STREAM_OPERATOR | T.TUNE_PDL_FINAL | T.TUNE_PDL_AFTER_FINAL),
    "stream+pdlall": (T.SCHEDULE_STEPWISE, T.TUNE_STREAM_OPERATOR | T.TUNE_PDL_FINAL | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
    "red": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER),
    "red+stream": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_STREAM_OPERATOR),
    "red+pdla": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_PDL_AFTER_FINAL),
    "red+pdlall": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
    "red+stream+pdlall": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_STREAM_OPERATOR | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
    "red+push": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_PUSH_ON_PRODUCE),
    "red+push+stream": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_PUSH_ON_PRODUCE | T.TUNE_STREAM_OPERATOR),
    "red+push+pdlall": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
    "red+push+stream+pdlall": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_PUSH_ON_PRODUCE | T.TUNE_STREAM_OPERATOR
                               | T.TUNE_PDL_AFTER_FINAL | T.TUNE_PDL_APPLY),
    "red+noack": (T.SCHEDULE_STEPWISE, T.TUNE_IN_KERNEL_REDUCER | T.TUNE_NO_ACK),
    "stream+noack": (T.SCHEDULE_STEPWISE, T.TUNE_STREAM_OPERATOR | T.TUNE_NO_ACK),
    "lazy": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PUSH_LAZY),
    "lazy+stream": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PUSH_LAZY | T.TUNE_STREAM_OPERATOR),
    "lazy+stream+red": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PUSH_LAZY | T.TUNE_STREAM_OPERATOR | T.TUNE_IN_KERNEL_REDUCER),
    "lazy+stream+pdlfa": (T.SCHEDULE_STEPWISE, T.TUNE_PUSH_ON_PRODUCE | T.TUNE_PUSH_LAZY | T.TUNE_STREAM_OPERATOR | T.TUNE_PDL_FINAL
                          | T.TUNE_PDL_AFTER_FINAL),
    "folded": (T.SCHEDULE_FOLDED, T.TUNE_OFF),
    "folded+stream": (T.SCHEDULE_FOLDED, T.TUNE_STREAM_OPERATOR),
    "persistent": (T.SCHEDULE_PERSISTENT, 0),
}
SLOTS = {"bicgstab": ["direction", "apply+dot", "half_update", "apply+2dots", "final_update+2dots"],
         "cg": ["apply+dot", "update+dot", "direction"]}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--axis", type=int, default=119)
    ap.add_argument("--cell", default="tet", choices=["tet", "hex"])
    ap.add_argument("--solver", default="bicgstab,cg")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=2, help="timed solves per variant; the best is reported")
    ap.add_argument("--variants", default="off,stream,pdlfa,pdlall,red,red+stream,red+pdla,red+pdlall,red+stream+pdlall,noack,push,red+push,"
                                           "red+push+stream,red+push+pdlall,red+push+stream+pdlall,folded,persistent")
    ap.add_argument("--partition", default="metis", choices=["metis", "slab"])
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    args.n = args.axis
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))

    import bench as bench_mod
    import stormruler_b200 as sb
    mesh, x_star = bench_mod.build_problem(args)
    if world > 1:
        import torch
        from stormruler_b200 import multigpu as mg
        dist = mg.init_process_group(cuda=True)
        t0 = time.time()
        part = mg.partition_mesh(mesh, world, capi.PART_SLAB if args.partition == "slab" else capi.PART_METIS)
        loc = part.local(rank)
        log(f"[ab] rank {rank}: partition {time.time() - t0:.1f}s, owned {loc.n_owned}, halo {loc.n_halo}, nbrs {loc.n_nbr}")
        ctx = mg.DistContext(mg.local_device(), rank, world, part.info.vec_capacity, n_vectors=12, mode=capi.COMM_P2P)
        op = mg.DistOperator(ctx, loc, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
        n = loc.n_owned
        xs = ctx.vector(x_star[loc.owned_global])
        mx, sm = mg.max_over_ranks, mg.sum_over_ranks

        def fence():
            dist.barrier()
            torch.cuda.synchronize()
    else:
        ctx = sb.Context(int(os.environ.get("LOCAL_RANK", "0")))
        op = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
        n = mesh.n_cells
        xs = ctx.vector(x_star)
        mx = sm = float

        def fence():
            ctx.sync()
    b = ctx.zeros(n)
    op.mul(b, xs)
    results = []
    for solver in args.solver.split(","):
        Solver = sb.BiCgStabSolver if solver == "bicgstab" else sb.CgSolver
        first_res = None

        def solve(iters, **kw):
            s = Solver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, record=False, **kw)
            x = ctx.zeros(n)
            fence()
            s.solve(x, b, op)
            fence()
            assert s.iteration == iters, (s.iteration, iters)
            return s

        for name in args.variants.split(","):
            schedule, tuning = VARIANTS[name]
            if world == 1 and (tuning & (T.TUNE_PUSH_ON_PRODUCE | T.TUNE_NO_ACK)):
                continue   # multi-GPU mechanisms
            try:
                solve(max(args.warmup, 3), use_graph=True, schedule=schedule, tuning=tuning)
                best = None
                for _ in range(args.repeats):
                    s = solve(args.steps, use_graph=True, schedule=schedule, tuning=tuning)
                    ms = mx(s.iter_ms)
                    best = ms if best is None else min(best, ms)
                res = float(s.absolute_error)
                row = {"solver": solver, "variant": name, "n_gpus": world, "cells": int(mesh.n_cells), "schedule_used": int(s.schedule_used),
                       "tuning": int(tuning), "it_per_s": args.steps / (best * 1e-3), "us_per_it": 1e3 * best / args.steps,
                       "launches": int(sm(s.launches)), "residual_after_steps": res}
                if first_res is None:
                    first_res = res
                row["same_bits_as_first_variant"] = bool(res == first_res)
                if not args.no_profile and schedule in (T.SCHEDULE_AUTO, T.SCHEDULE_STEPWISE, T.SCHEDULE_FOLDED):
                    sp = solve(args.steps, profile=True, schedule=schedule, tuning=tuning)
                    k = args.steps * 1e-3
                    row["slot_us"] = {nm: mx(sp.kernel_ms[i]) / k for i, nm in enumerate(SLOTS[solver])}
                    row["halo_or_fold_wait_us"] = {nm: mx(sp.wait_ms[i]) / k for i, nm in enumerate(SLOTS[solver])}
                    row["allreduce_wait_us"] = {nm: mx(sp.ar_wait_ms[i]) / k for i, nm in enumerate(SLOTS[solver])}
            except Exception as e:  # noqa: BLE001 -- a variant that is not available (persistent + profile ...) must not end the run
                row = {"solver": solver, "variant": name, "error": f"{type(e).__name__}: {e}"[:300]}
            results.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
                if "error" in row:
                    log(f"[ab] {solver:9s} {name:20s} ERROR {row['error']}")
                else:
                    extra = ""
                    if "slot_us" in row:
                        extra = "  slots " + " ".join(f"{v:5.1f}" for v in row["slot_us"].values()) + \
                                "  halo-wait " + " ".join(f"{v:4.1f}" for v in row["halo_or_fold_wait_us"].values()) + \
                                "  ar-wait " + " ".join(f"{v:4.1f}" for v in row["allreduce_wait_us"].values())
                    log(f"[ab] {solver:9s} {name:20s} {row['it_per_s']:9.0f} it/s {row['us_per_it']:7.1f} us/it "
                        f"same_bits={row['same_bits_as_first_variant']}{extra}")
    if rank == 0 and args.out:
        with open(args.out, "w") as f:
            json.dump({"axis": args.axis, "cell": args.cell, "n_gpus": world, "steps": args.steps, "partition": args.partition,
                       "results": results}, f, indent=1)
    if world > 1:
        err = ctx.status()
        assert err == 0, f"comm error word {err:#x}"
        del op, xs, b
        ctx.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
