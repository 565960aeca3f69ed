#!/usr/bin/env python
"""bench.py -- Krylov iterations/s and HBM roofline fraction of the StormRuler hot path on B200.

Metric (BASELINE.json): Krylov iterations/s & HBM GB/s (% of roofline) at 10 M cells.
Workload at N=1 (BASELINE.json configs[1]): BiCGStab on a 3-D FVM Poisson
problem, synthetic jittered tetrahedral box mesh (119^3 hexes x 6 Kuhn tets = 10 110 954 cells),
homogeneous-Dirichlet mirror ghosts, cells shuffled then RCM-renumbered, fp64.
A "step" is one BiCGStab iteration = 2 operator applies + 5 fused reductions + 5 vector updates.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--solver cg|bicgstab]
                    [--n N_PER_AXIS] [--cell tet|hex]

Own arm: W untimed iterations, then exactly K iterations timed with CUDA events on the solver's
stream (tolerances disabled so the count is fixed; the per-iteration working set of ~1.2 GB is far
larger than the 126 MB L2, so no flush is needed between iterations). Prints ONE JSON line.
`--impl reference` times the reference's own CPU solver (oracle/_ref: the unmodified StormRuler
headers; falls back to the oracle's C port) on the same mesh and right-hand side on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------
def build_problem(args):
    """Mesh (host) + exact solution + names. Shared by both arms so they see the same arrays."""
    from stormruler_b200.mesh import CELL_HEX, CELL_TET, Mesh
    t0 = time.time()
    kind = CELL_TET if args.cell == "tet" else CELL_HEX
    mesh = Mesh.box(kind, args.n, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)
    t1 = time.time()
    bw0 = mesh.bandwidth
    mesh.renumber_rcm()
    t2 = time.time()
    log(f"[bench] mesh: {mesh.n_cells} {args.cell} cells, {mesh.n_faces} interior faces, {mesh.n_bfaces} boundary "
        f"faces; generate {t1 - t0:.1f}s, RCM {t2 - t1:.1f}s (bandwidth {bw0} -> {mesh.bandwidth})")
    c = mesh.cell_centers()
    x_star = np.sin(np.pi * c[:, 0]) * np.sin(np.pi * c[:, 1]) * np.sin(np.pi * c[:, 2])
    return mesh, x_star


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms. Started before the operator upload (nvidia-smi needs
    about a second before its first line) and stopped after the end-to-end leg, so the samples cover the timed
    regions; `sm_mhz` is the median over the samples taken under load (clock above 30 % of max)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, reasons, sm_max = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                sm_max = float(r[1])
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        busy = [v for v in sm if sm_max and v > 0.3 * sm_max] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": sm_max,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(busy) if sm_max else 0}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_rate(mesh, x_star, solver, budget_s=20.0, max_iters=None):
    """Reference CPU solver on the same SoA arrays and right-hand side, 1 core (the reference is
    single-threaded by construction). Returns (it/s, kind, iterations run, seconds)."""
    from oracle import orc
    fm = orc.FaceMesh(mesh.n_cells, mesh.face_cell, mesh.face_area, mesh.face_dist, mesh.cell_vol,
                      mesh.bface_cell, mesh.bface_area, mesh.bface_dist)
    op = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
    b = op.apply(x_star)
    kind = "reference" if orc.have_ref() else "port"

    def run(iters):
        t = time.perf_counter()
        if kind == "reference":
            r = orc.ref_solve(solver, op, b, num_iterations=iters, abs_tol=0.0, rel_tol=0.0, trace_cap=16)
        else:
            r = orc.solve(solver, op, b, num_iterations=iters, abs_tol=0.0, rel_tol=0.0)
        assert r.iterations == iters
        return time.perf_counter() - t, r

    t1, _ = run(1)                    # init + 1 iteration: calibrates the budget
    per_it = max(t1 / 2.0, 1e-6)
    iters = int(max(2, min(budget_s / per_it, 2000)))
    if max_iters is not None:
        iters = max(2, min(iters, max_iters))
    t_n, res = run(iters)
    rate = (iters - 1) / max(t_n - t1, 1e-9)   # subtract the initialisation (residual + first iteration)
    return rate, kind, iters, t_n, res.hist


def cpu_all_cores_rate(mesh, x_star, solver, budget_s=6.0):
    """The same iteration on every host core: an OpenMP port on the operator's coefficient rows
    (oracle/sb_oracle_omp.c; the reference itself has no threading). Measurement aid only: its 5-iteration residual is
    checked against the single-threaded reference before the time is quoted. Returns a cpu_baseline-shaped dict."""
    import ctypes as C
    from oracle import orc
    path = os.path.join(ROOT, "oracle", "liboracle_omp.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/liboracle_omp.so not built"}
    L = C.CDLL(path)
    L.orc_omp_solve.restype = C.c_double
    L.orc_omp_threads.restype = C.c_int
    fm = orc.FaceMesh(mesh.n_cells, mesh.face_cell, mesh.face_area, mesh.face_dist, mesh.cell_vol,
                      mesh.bface_cell, mesh.bface_area, mesh.bface_dist)
    op = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
    b = op.apply(x_star)
    w, ld, col, a, diag = op.rows_coef()
    n = mesh.n_cells
    ptr = lambda arr: arr.ctypes.data_as(C.c_void_p)  # noqa: E731
    kind = {"cg": 0, "bicgstab": 1}[solver]

    def run(iters):
        x = np.zeros(n)
        secs = C.c_double()
        err = L.orc_omp_solve(kind, C.c_int64(n), int(w), C.c_int64(ld), ptr(col), ptr(a), ptr(diag), ptr(b), ptr(x),
                              C.c_int64(iters), C.byref(secs))
        return err, secs.value

    err5, t5 = run(5)
    run_ref = orc.ref_solve if orc.have_ref() else orc.solve
    want = run_ref(solver, op, b, num_iterations=5, abs_tol=0.0, rel_tol=0.0).abs_err
    if not abs(err5 - want) <= 1e-9 * abs(want):
        return {"unavailable": f"port disagrees with the reference after 5 iterations ({err5!r} vs {want!r})"}
    iters = int(max(10, min(budget_s / max(t5 / 5.0, 1e-6), 5000)))
    _, t = run(iters)
    return {"value": iters / max(t, 1e-9), "unit": "it/s", "cores": int(L.orc_omp_threads()), "kind": "port",
            "sample": f"{iters} {solver} iterations of the same {n}-cell problem ({t:.1f} s), OpenMP port of the reference's "
                      f"statement sequence on the coefficient rows (oracle/sb_oracle_omp.c, gcc -O2 -fopenmp), all host "
                      f"threads; residual after 5 iterations equal to the single-threaded reference's within 1e-9"}


# ---------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # inputs through the host-only build of the mesh sources: this process never maps the CUDA library
    prep = os.path.join(ROOT, "oracle", "libsb_meshprep.so")
    if os.path.exists(prep):
        from stormruler_b200 import capi
        capi.use_host_library(prep)
    mesh, x_star = build_problem(args)
    t0 = time.perf_counter()
    from oracle import orc
    fm = orc.FaceMesh(mesh.n_cells, mesh.face_cell, mesh.face_area, mesh.face_dist, mesh.cell_vol,
                      mesh.bface_cell, mesh.bface_area, mesh.bface_dist)
    op = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
    b = op.apply(x_star)
    kind = "reference" if orc.have_ref() else "port"
    solve = (lambda it: orc.ref_solve(args.solver, op, b, num_iterations=it, abs_tol=0.0, rel_tol=0.0, trace_cap=16)) \
        if kind == "reference" else (lambda it: orc.solve(args.solver, op, b, num_iterations=it, abs_tol=0.0, rel_tol=0.0))
    # warm-up iterations are part of the same solve: time W+K and W iterations and subtract
    tw0 = time.perf_counter(); solve(args.warmup); tw = time.perf_counter() - tw0
    tk0 = time.perf_counter(); r = solve(args.warmup + args.steps); tk = time.perf_counter() - tk0
    dt = max(tk - tw, 1e-9)
    value = args.steps / dt
    line = {
        "impl": "reference", "metric": "krylov_iterations_per_sec", "value": value, "unit": "it/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, mesh),
        "cpu_baseline": {"value": value, "unit": "it/s", "cores": 1, "kind": kind,
                         "sample": f"{args.steps} {args.solver} iterations at full size (after {args.warmup} warm-up "
                                   f"iterations), reference solver headers on a host vector + face-loop operator, "
                                   f"g++ -O2 -ffp-contract=off, single thread (the reference has no threading)"},
        "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "residual_after": r.abs_err, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, mesh):
    return {"workload": f"{args.solver} on 3-D FVM Poisson (Dirichlet mirror ghosts), synthetic jittered "
                        f"{'Kuhn-tetrahedral' if args.cell == 'tet' else 'hexahedral'} box mesh {args.n}^3, "
                        f"RCM-renumbered",
            "cells": int(mesh.n_cells), "interior_faces": int(mesh.n_faces), "boundary_faces": int(mesh.n_bfaces),
            "solver": args.solver, "operator_form": "coef (12 B/entry ELL + diagonal)",
            "l2_policy": "inputs_larger_than_L2 (per-iteration working set >> 126 MB)"}


def history_parity(gpu_hist, ref_hist):
    """Residual history of the GPU path against the reference's own CPU run of the same problem: relative difference
    per iteration (north_star bar 1e-10), its maximum, and the first iteration that exceeds the bar."""
    k = min(len(gpu_hist), len(ref_hist))
    rel = np.abs(np.asarray(gpu_hist[:k]) - np.asarray(ref_hist[:k])) / np.abs(np.asarray(ref_hist[:k]))
    over = np.flatnonzero(rel > 1e-10)
    return {"iterations_compared": int(k - 1), "max_rel_diff": float(rel.max()),
            "first_iteration_over_1e-10": int(over[0]) if over.size else None,
            "rel_diff_at": {str(i): float(rel[i]) for i in (1, 2, 5, 10, 20, 40) if i < k}}


def phase_times(timeline, solver):
    """Mean microseconds per step of an iteration from the persistent kernel's own timeline (globaltimer stamps of
    CTA 0 behind every grid barrier), plus the waits it records."""
    tl = np.asarray(timeline, dtype=np.int64)
    names = {"bicgstab": ["direction", "apply+dot", "half_update", "apply+2dots", "final_update+2dots"],
             "cg": ["apply+dot", "update+dot", "direction"]}[solver]
    nb = len(names)
    if tl.shape[0] < 2:
        return None
    tl = tl[1:]                                                  # the first iteration warms the caches
    steps = np.diff(tl[:, :nb + 1], axis=1)
    out = {"us_per_step": {nm: float(steps[:, k].mean()) * 1e-3 for k, nm in enumerate(names)},
           "us_per_iteration": float((tl[:, nb] - tl[:, 0]).mean()) * 1e-3,
           "us_in_barrier_cta0": {nm: float(tl[:, 6 + k].mean()) * 1e-3 for k, nm in enumerate(names)},
           "us_wait_for_other_ranks": {nm: float(tl[:, 11 + k].mean()) * 1e-3 for k, nm in enumerate(names) if "dot" in nm},
           "us_halo_wait_max": [float(tl[:, 16 + k].mean()) * 1e-3 for k in range(2 if solver == "bicgstab" else 1)],
           "iterations_sampled": int(tl.shape[0])}
    return out


def SCHEDULES(capi):
    return {"auto": capi.SCHEDULE_AUTO, "stepwise": capi.SCHEDULE_STEPWISE, "persistent": capi.SCHEDULE_PERSISTENT,
            "folded": capi.SCHEDULE_FOLDED}


def schedule_name(capi, used):
    return {capi.SCHEDULE_PERSISTENT: "persistent", capi.SCHEDULE_STEPWISE: "stepwise", capi.SCHEDULE_FOLDED: "folded"}[used]


SLOTS = {"bicgstab": ["direction", "apply+dot", "half_update", "apply+2dots", "final_update+2dots"],
         "cg": ["apply+dot", "update+dot", "direction"]}
APPLY_EPILOGUE = {"bicgstab": {"apply+dot": "EpiUY", "apply+2dots": "EpiYYandYX"}, "cg": {"apply+dot": "EpiXY"}}
KERNEL_NAME = {"bicgstab": "sb::krylov_persistent_kernel<BiCgStab, W> (the whole iteration loop, ONE cooperative launch)",
               "cg": "sb::krylov_persistent_kernel<Cg, W> (the whole iteration loop, ONE cooperative launch)"}


def run_own_arm(args):
    import stormruler_b200 as sb
    from stormruler_b200 import capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from stormruler_b200 import multigpu
        return multigpu.bench_main(args, build_problem, workload_config, peaks, ClockSampler)
    mesh, x_star = build_problem(args)
    sampler = ClockSampler(local_rank).start()
    ctx = sb.Context(local_rank)
    t0 = time.time()
    op = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
    log(f"[bench] operator upload: {time.time() - t0:.1f}s, width {op.info.width}, "
        f"{op.info.device_bytes / 1e6:.0f} MB in HBM, {op.info.algorithmic_bytes_per_apply / mesh.n_cells:.1f} B/cell/apply")
    n = mesh.n_cells
    xs = ctx.vector(x_star)
    b = ctx.zeros(n)
    op.mul(b, xs)                                   # b = A x*
    schedule = SCHEDULES(capi)[args.schedule]
    peak, peak_src = peaks()
    alg_apply = op.info.algorithmic_bytes_per_apply

    def measure(solver):
        """W untimed + exactly K timed iterations of `solver` (device time of the iteration loop, CUDA events on the
        solver's stream), the in-kernel timeline of the same schedule, and the per-kernel times of the stepwise one."""
        Solver = sb.BiCgStabSolver if solver == "bicgstab" else sb.CgSolver
        applies_per_it, passes_per_it = (2, 15) if solver == "bicgstab" else (1, 9)

        def solve(iters, **kw):
            s = Solver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=0.0,
                       record=kw.pop("record", False), **kw)
            x = ctx.zeros(n)
            s.solve(x, b, op)
            assert s.iteration == iters, (s.iteration, iters)
            return s, x

        solve(max(args.warmup, 3), use_graph=True, schedule=schedule, tuning=args.tuning)   # warm-up (untimed)
        s, x = solve(args.steps, use_graph=True, schedule=schedule, tuning=args.tuning)     # timed: exactly K iterations
        value = args.steps / (s.iter_ms * 1e-3)
        alg_iter = applies_per_it * alg_apply + passes_per_it * 8 * n
        out = {"solver": solver, "s": s, "x": x, "value": value, "ms_per_step": s.iter_ms / args.steps,
               "alg_iter": alg_iter, "applies_per_it": applies_per_it,
               "schedule": schedule_name(capi, s.schedule_used),
               "iteration_roofline": {"algorithmic_bytes_per_iteration": int(alg_iter),
                                      "achieved_gbs": alg_iter * value / 1e9,
                                      "frac_of_measured_peak": alg_iter * value / 1e9 / peak,
                                      "frac_of_nominal_8TBs": alg_iter * value / 8e12}}
        if s.schedule_used == capi.SCHEDULE_PERSISTENT:
            st, _ = solve(min(args.steps, 64), schedule=schedule, timeline_iters=min(args.steps, 64), tuning=args.tuning)
            out["phases"] = phase_times(st.timeline, solver)
        # per-kernel breakdown of the same K iterations in the stepwise schedule (events around every launch, no graph)
        sp, _ = solve(args.steps, profile=True, schedule=capi.SCHEDULE_FOLDED if s.schedule_used == capi.SCHEDULE_FOLDED else capi.SCHEDULE_STEPWISE,
                      tuning=args.tuning)
        kms = list(sp.kernel_ms)
        fms = list(sp.final_ms)                       # the one-CTA final stage behind a reducing kernel (its own event)
        slots = SLOTS[solver]
        apply_slots = [k for k, nm in enumerate(slots) if nm.startswith("apply")]
        # every apply variant against ITS OWN algorithmic bytes: the <r~,v> epilogue reads a third vector (+ 8 N)
        variants = []
        for k in apply_slots:
            epi = APPLY_EPILOGUE[solver][slots[k]]
            bytes_k = alg_apply + (8 * n if epi == "EpiUY" else 0)
            ms_k = (kms[k] - fms[k]) / args.steps     # the apply kernel alone
            variants.append({"slot": slots[k], "kernel": f"sb::apply_kernel_tma<W, ., ., {epi}>",
                             "algorithmic_bytes_per_launch": int(bytes_k), "avg_launch_ms": ms_k,
                             "final_stage_ms": fms[k] / args.steps,
                             "achieved": bytes_k / (ms_k * 1e-3) / 1e9, "frac": bytes_k / (ms_k * 1e-3) / 1e9 / peak})
        tot_bytes = sum(v["algorithmic_bytes_per_launch"] for v in variants)
        tot_ms = sum(v["avg_launch_ms"] for v in variants)
        out["stepwise"] = {"kernel_ms_per_iteration": {nm: kms[k] / args.steps for k, nm in enumerate(slots)},
                           "final_stage_ms_per_iteration": {nm: fms[k] / args.steps for k, nm in enumerate(slots)},
                           "in_kernel_wait_ms_per_iteration": {nm: sp.wait_ms[k] / args.steps for k, nm in enumerate(slots)},
                           "apply_kernel": {"kernel": "sb::apply_kernel_tma (stepwise schedule, profiled: CUDA events around every "
                                                      "launch; the one-CTA final stage behind it is timed separately)",
                                            "algorithmic_bytes_per_launch": tot_bytes / len(variants),
                                            "avg_launch_ms": tot_ms / len(variants),
                                            "achieved": tot_bytes / (tot_ms * 1e-3) / 1e9,
                                            "frac": tot_bytes / (tot_ms * 1e-3) / 1e9 / peak,
                                            "share_of_step": sum(kms[k] - fms[k] for k in apply_slots) / sum(kms[:len(slots)]),
                                            "launches_timed": len(variants) * args.steps,
                                            "variants": variants}}
        return out

    main = measure(args.solver)
    other = None
    if not args.single_solver:
        other = measure("cg" if args.solver == "bicgstab" else "bicgstab")   # the other target solver, same problem
    s, x = main["s"], main["x"]
    err = np.linalg.norm(x.numpy() - x_star) / np.linalg.norm(x_star)
    # DRAM bytes per launch of the same kernels from the committed `ncu --set full` capture (per apply variant; only
    # valid for the workload it was captured on), averaged over the variants like `achieved`
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "apply_traffic.json")
    if main["schedule"] != "persistent" and os.path.exists(tpath):
        try:
            t = json.load(open(tpath))
            if t.get("workload") == {"cell": args.cell, "axis": args.n}:
                per = [t["variants"][APPLY_EPILOGUE[args.solver][v["slot"]]] for v in main["stepwise"]["apply_kernel"]["variants"]]
                for v, b_ in zip(main["stepwise"]["apply_kernel"]["variants"], per):
                    v["traffic"] = b_
                traffic, traffic_src = sum(per) / len(per), t["source"]
        except Exception:
            traffic = None

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (copies inside the timed region) ----
    import torch
    hx = torch.zeros(n, dtype=torch.float64).pin_memory()
    hb = torch.from_numpy(b.numpy()).pin_memory()
    hxn, hbn = hx.numpy(), hb.numpy()
    sb.solve_host(ctx, op, args.solver, hxn, hbn, num_iterations=3, abs_tol=0.0, rel_tol=0.0, use_graph=True, schedule=schedule,
                  tuning=args.tuning)
    hxn[:] = 0.0
    ctx.sync()
    t = time.perf_counter()
    rep = sb.solve_host(ctx, op, args.solver, hxn, hbn, num_iterations=args.steps, abs_tol=0.0, rel_tol=0.0,
                        use_graph=True, schedule=schedule, tuning=args.tuning)
    e2e_s = time.perf_counter() - t
    assert rep.iterations == args.steps
    e2e_value = args.steps / e2e_s
    clocks = sampler.stop()

    # ---- CPU baseline beside it (bounded sample of the same workload) + parity of the residual histories ----
    cpu = parity = None
    if not args.no_cpu_baseline:
        rate, kind, iters, secs, ref_hist = cpu_reference_rate(mesh, x_star, args.solver, budget_s=args.cpu_budget)
        cpu = {"value": rate, "unit": "it/s", "cores": 1, "kind": kind,
               "sample": f"{iters} {args.solver} iterations of the same {n}-cell problem ({secs:.1f} s), reference "
                         f"solver headers + face-loop operator, g++ -O2 -ffp-contract=off, 1 thread "
                         f"(the reference is single-threaded by construction)"}
        # the same iterations on the GPU, history recorded: the measured path (coefficient rows, tree reductions) and
        # the faithful rows (the face loop's own arithmetic: only the reduction order differs from the reference)
        Solver = sb.BiCgStabSolver if args.solver == "bicgstab" else sb.CgSolver
        parity = {"against": f"{kind}: the reference's solver headers + face loop, sequential sums, first {iters} iterations "
                             f"of the same {n}-cell problem", "bar": "1e-10 relative per iteration (north_star)"}
        for label, form in (("measured_path_coef_rows", sb.FORM_COEF), ("faithful_rows", sb.FORM_FAITHFUL)):
            o = op if form == sb.FORM_COEF else sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=form, dirichlet=True)
            sg = Solver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, record=True)
            xg = ctx.zeros(n)
            sg.solve(xg, b, o)
            parity[label] = history_parity(sg.history, ref_hist)
            del o, xg

    # the other target solver's history against the reference too (a shorter sample: CG's whole history is expected to
    # stay within the bar, BiCGStab's is not -- SURVEY.md F8)
    parity_other = None
    if not args.no_cpu_baseline and other is not None:
        try:
            o_solver = other["solver"]
            _, kind_o, iters_o, _, ref_hist_o = cpu_reference_rate(mesh, x_star, o_solver, budget_s=args.cpu_budget / 4)
            OSolver = sb.BiCgStabSolver if o_solver == "bicgstab" else sb.CgSolver
            sg = OSolver(num_iterations=iters_o, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, record=True)
            xg = ctx.zeros(n)
            sg.solve(xg, b, op)
            parity_other = {"solver": o_solver, "against": f"{kind_o}: first {iters_o} iterations of the same problem, sequential sums",
                            "measured_path_coef_rows": history_parity(sg.history, ref_hist_o)}
            del xg
        except Exception as e:  # noqa: BLE001 -- an extra, never allowed to take the bench line down
            parity_other = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    cpu_all = None
    if not args.no_cpu_baseline:
        try:   # an extra, never allowed to take the bench line down
            cpu_all = cpu_all_cores_rate(mesh, x_star, args.solver)
        except Exception as e:  # noqa: BLE001
            cpu_all = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    persistent = main["schedule"] == "persistent"
    alg_launch = main["alg_iter"] * args.steps if persistent else main["stepwise"]["apply_kernel"]["algorithmic_bytes_per_launch"]
    launch_ms = s.iter_ms if persistent else main["stepwise"]["apply_kernel"]["avg_launch_ms"]
    achieved = alg_launch / (launch_ms * 1e-3) / 1e9
    line = {
        "metric": "krylov_iterations_per_sec", "value": main["value"], "unit": "it/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, mesh), schedule=main["schedule"]),
        "roofline": {"bound": "hbm",
                     "kernel": KERNEL_NAME[args.solver] if persistent else main["stepwise"]["apply_kernel"]["kernel"],
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(alg_launch), "avg_launch_ms": launch_ms,
                     "launches_timed": 1 if persistent else main["stepwise"]["apply_kernel"]["launches_timed"],
                     "variants": None if persistent else main["stepwise"]["apply_kernel"]["variants"],
                     "share_of_step": 1.0 if persistent else main["stepwise"]["apply_kernel"]["share_of_step"],
                     "note": "persistent schedule: one launch = K iterations = K x (applies x (24 N + 12 entries) + passes x 8 N) "
                             "algorithmic bytes, timed by CUDA events around the launch" if persistent else None},
        "iteration_roofline": main["iteration_roofline"],
        "phases": main.get("phases"),
        "stepwise": main["stepwise"],
        "applies_per_sec": main["applies_per_it"] * main["value"],
        "other_solver": None if other is None else {k: other.get(k) for k in ("solver", "value", "ms_per_step", "schedule",
                                                                            "iteration_roofline", "phases", "stepwise")},
        "cpu_baseline": cpu,
        "cpu_baseline_all_cores": cpu_all,
        "parity": parity,
        "parity_other_solver": parity_other,
        "e2e": {"value": e2e_value, "unit": "it/s", "h2d_bytes_per_step": 16 * n / args.steps,
                "d2h_bytes_per_step": 8 * n / args.steps,
                "note": f"one sb_solve_host call = H2D(x0,b) + init + {args.steps} iterations + D2H(x), pinned host buffers"},
        "gpu_launches": int(s.launches), "clocks": clocks,
        "rel_error_vs_exact_after_steps": float(err), "residual_after_steps": float(s.absolute_error),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--solver", default="bicgstab", choices=["cg", "bicgstab"])
    ap.add_argument("--axis", "--n", dest="n", type=int, default=119,
                    help="hexes per axis (119 -> 10.1 M tets); spell it --axis under torchrun, whose parser finds --n ambiguous")
    ap.add_argument("--cell", default="tet", choices=["tet", "hex"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: halo exchange + reductions by in-kernel NVLink peer stores (p2p) or NCCL send/recv + allreduce")
    ap.add_argument("--partition", default="metis", choices=["metis", "slab"], help="N>1: METIS k-way or contiguous RCM slabs")
    ap.add_argument("--schedule", default="auto", choices=["auto", "stepwise", "persistent", "folded"],
                    help="fused-solver schedule of the timed run (auto = stepwise: one kernel per step, graph replay)")
    ap.add_argument("--tuning", type=lambda v: int(v, 0), default=0,
                    help="SB_TUNE_* bits of the stepwise schedule (include/stormb200.h); 0 = the library's defaults")
    ap.add_argument("--single-solver", action="store_true", help="skip the leg of the other target solver (cg <-> bicgstab)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
