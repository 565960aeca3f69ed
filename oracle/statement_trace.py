"""TEST / ANALYSIS INFRASTRUCTURE -- not product code, computes nothing on vectors.

What exactly do the reference's solvers ask of the vector backend per iteration? The product's drop-in TU
(stormruler_b200/host/dropin.cpp: StormRuler's own solver templates on Storm::DeviceVector, unmodified) is linked
against a logging stand-in of the C ABI (oracle/trace/trace_stormb200.c, built by `make -C oracle trace`), run for K
and 2K iterations with tolerances 0, and the difference of the two statement streams is the steady-state cost of K
iterations: operator applies, reductions, and vector passes (V = one read or write of one vector)

  * as written -- every statement its own pass: `y (op)= expr` reads its distinct operands (and y for += -= *= /=)
    and writes y; dot reads its distinct operands; fill writes; copy reads and writes. This is what the generic
    drop-in path executes today, and what SURVEY.md 8a/8d quote as the contract for the solvers without a fused
    schedule;
  * fused -- what a statement-fusing backend would move: consecutive element-wise statements and the reductions
    that follow them form ONE pass over the union of their operands (a vector written earlier in the group is
    forwarded in registers, not re-read); a group ends where a reduction result is needed on the host (i.e. at the
    first statement after a reduction) or at an operator apply; reductions directly behind an apply ride on the apply
    (operands equal to its input or output are free, the kernel has them in registers). Element-wise statements are
    independent per element, so fusion changes no floating-point result.

    python -m oracle.statement_trace [--solvers cg,bicgstab,...] [--json out.json]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "trace", "libstorm_dropin_trace.so")
TRACE = os.path.join(HERE, "_ref", "trace", "libstormb200_trace.so")

# solver -> (iterations per cycle, default inner size argument)
SOLVERS = {"cg": (1, 0), "cgs": (1, 0), "bicgstab": (1, 0), "bicgstabl": (2, 0), "tfqmr": (1, 0), "tfqmr1": (1, 0),
           "idrs": (4, 0), "gmres": (50, 50), "fgmres": (50, 50), "richardson": (1, 0),
           # Storm/B200/GroupedSolvers.hpp: the same algorithms, statements issued as sb_eval_group launches
           "grouped_idrs": (4, 0), "grouped_bicgstabl": (2, 0)}
READS_TARGET = {1, 2, 3, 4}   # += -= *= /= read y as well


class Opts(C.Structure):      # dropin_opts of stormruler_b200/host/dropin.cpp
    _fields_ = [("num_iterations", C.c_int64), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
                ("num_inner_iterations", C.c_int64), ("relaxation_factor", C.c_double),
                ("use_graph", C.c_int32), ("precond", C.c_int32), ("pre_side", C.c_int32),
                ("cheb_degree", C.c_int32), ("cheb_power_iterations", C.c_int32), ("cheb_eig_ratio", C.c_double)]


class Report(C.Structure):
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int64), ("abs_err", C.c_double),
                ("rel_err", C.c_double), ("n_hist", C.c_int64), ("n_trace", C.c_int64), ("n_apply", C.c_int64)]


def available() -> bool:
    return os.path.exists(LIB) and os.path.exists(TRACE)


_libs = None


def _load():
    global _libs
    if _libs is None:
        # RTLD_LOCAL on purpose: the stand-in's sb_* symbols must never leak into the global scope of a process that
        # may also load the real libstormb200.so; the drop-in finds them through its own DT_NEEDED entry
        dr = C.CDLL(LIB)
        tr = C.CDLL(TRACE)          # same library instance (already loaded as the drop-in's dependency)
        tr.sbtrace_log.restype = C.c_char_p
        dr.dropin_solve.restype = C.c_int
        dr.dropin_last_error.restype = C.c_char_p
        _libs = (tr, dr)
    return _libs


def statement_stream(solver: str, iterations: int, precond: int = 0, pre_side: int = 1, grouping: bool = False):
    """Run `solver` for exactly `iterations` iterations on the tracer; returns the log as a list of tuples.
    grouping: Storm::B200::set_statement_grouping(true) -- the generic path queues chain-shaped statements and
    launches them as sb_eval_group with the reduction behind them."""
    tr, dr = _load()
    dr.dropin_set_statement_grouping(int(grouping))  # 0 / 1 / 2
    tr.sbtrace_reset()
    n = 1000
    fake_ctx, fake_op = C.c_void_p(0x1000), C.c_void_p(0x2000)
    x = (C.c_double * 2)()      # two distinct addresses for x and b; never dereferenced by the tracer
    b = (C.c_double * 2)()
    opts = Opts(iterations, 0.0, 0.0, SOLVERS[solver][1], 0.0, 0, precond, pre_side, 0, 0, 0.0)
    rep = Report()
    rc = dr.dropin_solve(solver.encode(), fake_ctx, fake_op, C.cast(x, C.c_void_p), C.cast(b, C.c_void_p),
                         C.c_size_t(n), C.byref(opts), C.byref(rep), None, C.c_int64(0), None, C.c_int64(0))
    dr.dropin_set_statement_grouping(0)
    if rc != 0:
        raise RuntimeError(f"dropin_solve({solver}) on the tracer failed: {dr.dropin_last_error().decode()}")
    assert rep.iterations == iterations, (solver, rep.iterations, iterations)
    out = []
    for line in tr.sbtrace_log().decode().splitlines():
        parts = line.split()
        out.append((parts[0],) + tuple(int(p) for p in parts[1:]))
    return out


def _reads_writes(st):
    """(reads, writes) vector-id sets of one statement."""
    kind = st[0]
    if kind == "eval":
        y, aop, nv = st[1], st[2], st[3]
        reads = set(st[4:4 + nv])
        if aop in READS_TARGET:
            reads.add(y)
        return reads, {y}
    if kind == "fill":
        return set(), {st[1]}
    if kind == "copy":
        return {st[2]}, {st[1]}
    if kind == "dot":
        return {st[1], st[2]}, set()
    if kind == "norm":
        return {st[1]}, set()
    if kind == "upload":
        return set(), set()
    raise ValueError(kind)


def count(stream):
    """Totals of a statement stream: applies, reductions, passes as written, passes fused, kernel launches."""
    applies = reductions = written = fused = 0
    launches_written = launches_fused = 0
    group_reads, group_writes, group_has_red = set(), set(), False
    group_open = False
    apply_ctx = None            # (x, y) of the apply the following reductions may ride on

    def close():
        nonlocal fused, launches_fused, group_reads, group_writes, group_has_red, group_open
        if group_open:
            fused += len(group_reads) + len(group_writes)
            launches_fused += 1
        group_reads, group_writes, group_has_red, group_open = set(), set(), False, False

    for st in stream:
        kind = st[0]
        if kind in ("alloc", "free", "mark", "upload", "download"):
            continue
        if kind == "group":     # group n_reads r.. n_writes w.. n_stmt n_dots: one launch over the union of its operands
            close()
            nr = st[1]
            nw = st[2 + nr]
            n_dots = st[2 + nr + 1 + nw + 1]
            reductions += n_dots
            written += nr + nw
            fused += nr + nw
            launches_written += 1
            launches_fused += 1
            apply_ctx = None
            continue
        if kind == "applydot2":  # apply with <y,y> and <y,x>: both dots' operands are in the apply's registers
            close()
            applies += 1
            reductions += 2
            launches_written += 1
            launches_fused += 1
            apply_ctx = (st[2], st[1])
            continue
        if kind == "applydot":  # applydot y x u: the dot's operands are in the apply's registers, except u when u != x
            close()
            applies += 1
            reductions += 1
            extra = 0 if st[3] == st[2] else 1
            written += extra
            fused += extra
            launches_written += 1
            launches_fused += 1
            apply_ctx = (st[2], st[1])
            continue
        if kind in ("apply", "jacobi"):
            close()
            applies += kind == "apply"
            if kind == "jacobi":
                written += 2
                fused += 2
            launches_written += 1
            launches_fused += 1
            apply_ctx = (st[2], st[1]) if kind == "apply" else None
            continue
        reads, writes = _reads_writes(st)
        is_red = kind in ("dot", "norm")
        reductions += is_red
        written += len(reads) + len(writes)
        launches_written += 1
        if is_red and apply_ctx is not None and not group_open:
            fused += len(reads - set(apply_ctx))          # rides on the apply: its x and y are in registers
            continue
        if not is_red:
            apply_ctx = None
            if group_has_red:       # a reduction result is needed on the host before this statement
                close()
        group_open = True
        group_reads |= reads - group_writes               # written earlier in the group: forwarded
        group_writes |= writes
        group_has_red |= is_red
    close()
    return dict(applies=applies, reductions=reductions, passes_written=written, passes_fused=fused,
                launches_written=launches_written, launches_fused=launches_fused)


def per_iteration(solver: str, cycles: int = 4, grouping: bool = False):
    """Steady-state counts per iteration: (2K iterations) - (K iterations), K a multiple of the solver's cycle."""
    cyc = SOLVERS[solver][0]
    K = cyc * cycles
    a = count(statement_stream(solver, K, grouping=grouping))
    b = count(statement_stream(solver, 2 * K, grouping=grouping))
    return {k: (b[k] - a[k]) / K for k in a}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--solvers", default=",".join(SOLVERS))
    ap.add_argument("--json", default="")
    ap.add_argument("--grouping", nargs="?", const=1, default=0, type=int,
                    help="with Storm::B200::set_statement_grouping(true); --grouping 2: + dependency-aware scheduling")
    args = ap.parse_args()
    if not available():
        sys.exit("tracer not built: make -C oracle trace (needs the StormRuler sources)")
    rows = {}
    print(f"{'solver':12s} {'applies':>8s} {'reductions':>11s} {'V written':>10s} {'V fused':>8s} {'launches':>9s} {'fused':>6s}")
    for s in args.solvers.split(","):
        r = per_iteration(s, grouping=args.grouping)
        rows[s] = r
        print(f"{s:12s} {r['applies']:8.2f} {r['reductions']:11.2f} {r['passes_written']:10.2f} {r['passes_fused']:8.2f} "
              f"{r['launches_written']:9.2f} {r['launches_fused']:6.2f}")
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
