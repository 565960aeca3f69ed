"""Run-to-run determinism probes at a given box size (default 100 -> 6 M tets)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import stormruler_b200 as sb
from test_gpu_scale import box

n_axis = int(sys.argv[1]) if len(sys.argv) > 1 else 100
forms = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 0]
mesh, fm = box(n_axis)
n = mesh.n_cells
ctx = sb.Context(0)
print(f"n={n} tiles={(n+2047)//2048} V1={os.environ.get('SB_APPLY_V1','0')} PDL={os.environ.get('SB_PDL','0')}", flush=True)
rng = np.random.default_rng(1)
xh = rng.standard_normal(n)
for form in forms:
    gpu = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=form, dirichlet=True)
    x = ctx.vector(xh)
    ys = []
    for rep in range(6):
        y = ctx.zeros(n)
        gpu.mul(y, x)
        ys.append(y.numpy())
    bad = [int((ys[k] != ys[0]).sum()) for k in range(1, 6)]
    print(f"form {form}: plain apply, #differing elements vs run 0: {bad}", flush=True)
    if any(bad):
        k = int(np.argmax(bad)) + 1
        idx = np.flatnonzero(ys[k] != ys[0])
        print("   first differing rows:", idx[:16], " tiles:", np.unique(idx // 2048)[:10], " rows mod 64:", np.unique(idx % 64)[:8])
    # chained applies (input freshly written by the previous kernel)
    outs = []
    for rep in range(4):
        a, b = ctx.zeros(n), ctx.zeros(n)
        gpu.mul(a, x); gpu.mul(b, a); gpu.mul(a, b); gpu.mul(b, a)
        outs.append(b.numpy())
    print(f"form {form}: 4 chained applies, #differing vs run 0: {[int((o != outs[0]).sum()) for o in outs[1:]]}", flush=True)
    d = [ctx.dot(x, x) for _ in range(5)]
    print(f"form {form}: dot repeat equal: {len(set(d)) == 1}")
    bd = ctx.vector(np.sin(0.37 * np.arange(n)))
    for solver, S in (("cg", sb.CgSolver), ("bicgstab", sb.BiCgStabSolver)):
        traces, xs = [], []
        for rep in range(4):
            s = S(num_iterations=12, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, use_graph=(rep % 2 == 1))
            xv = ctx.zeros(n)
            s.solve(xv, bd, gpu)
            traces.append(s.trace.copy()); xs.append(xv.numpy())
        for k in range(1, 4):
            bad = np.flatnonzero(traces[k] != traces[0])
            print(f"form {form} {solver} run {k} vs 0: first differing trace index {int(bad[0]) if len(bad) else None} of {len(traces[0])}; x differs in {int((xs[k] != xs[0]).sum())} elements", flush=True)
