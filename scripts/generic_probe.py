"""Where does a generic-path (reference template on DeviceVector) iteration spend its time? Times the individual
statement kinds of BiCGStab at full size through the same C-ABI calls, each bracketed by a stream sync."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stormruler_b200 as sb  # noqa: E402
from stormruler_b200 import dropin  # noqa: E402
from stormruler_b200.mesh import CELL_TET, Mesh  # noqa: E402

axis = int(sys.argv[1]) if len(sys.argv) > 1 else 119
mesh = Mesh.box(CELL_TET, axis, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)
mesh.renumber_rcm()
n = mesh.n_cells
ctx = sb.Context(0)
op = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
rng = np.random.default_rng(0)
p, r, v, t, x, rt = (ctx.vector(rng.standard_normal(n)) for _ in range(6))
V = sb.expr.v


def timeit(name, f, reps=20):
    for _ in range(3):
        f()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    ctx.sync()
    print(f"{name:40s} {(time.perf_counter() - t0) / reps * 1e6:9.1f} us", flush=True)


timeit("p = r + b*(p - w*v)", lambda: (V(r) + 0.7 * (V(p) - 0.3 * V(v))).assign_to(p))
timeit("v = A p", lambda: op.mul(v, p))
timeit("dot(rt, v)", lambda: ctx.dot(rt, v))
timeit("x += a*p", lambda: (0.5 * V(p)).assign_to(x, sb.ADD_ASSIGN))
timeit("r -= a*v", lambda: (1e-9 * V(v)).assign_to(r, sb.SUB_ASSIGN))
timeit("dot(t, t)", lambda: ctx.dot(t, t))
timeit("norm2(r)", lambda: ctx.norm2(r))
b = ctx.vector(np.sin(0.37 * np.arange(n)))
for name in ("cg", "bicgstab", "cgs", "bicgstab", "idrs", "tfqmr"):
    for iters in (20, 60):
        xx = ctx.zeros(n)
        ctx.sync()
        t0 = time.perf_counter()
        res = dropin.solve(name, op, xx, b, num_iterations=iters, abs_tol=0.0, rel_tol=0.0, trace_cap=64)
        ctx.sync()
        print(f"{name:10s} {iters:4d} iterations: {time.perf_counter() - t0:8.4f} s   abs_err {res.abs_err:.3e}  applies {res.n_apply}", flush=True)
