"""Localise a mismatch between the fused BiCGStab and the oracle: first differing reduction scalar."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import stormruler_b200 as sb
from oracle import orc
from test_gpu_scale import box

n_axis = int(sys.argv[1]) if len(sys.argv) > 1 else 58
mesh, fm = box(n_axis)
n = mesh.n_cells
ctx = sb.Context(0)
cpu = orc.FaceOp(fm, prefill=0, dt=-1.0, dirichlet=True)
rows_op = orc.RowsOp(n, *cpu.rows_coef())
rng = np.random.default_rng(1)
a, b2, c3 = (rng.standard_normal(n) for _ in range(3))
A, B, Cc = ctx.vector(a), ctx.vector(b2), ctx.vector(c3)
got = ctx.dot_batch([(A, B), (B, Cc), (A, Cc)])
want = [orc.dot(a, b2, orc.RED_TREE), orc.dot(b2, c3, orc.RED_TREE), orc.dot(a, c3, orc.RED_TREE)]
print("dot_batch3", [g == w for g, w in zip(got, want)])
got = ctx.dot_batch([(A, B), (B, Cc)])
print("dot_batch2", [g == w for g, w in zip(got, want[:2])])
names = {"bicgstab": ["rho0"] + ["<rt,v>", "<t,t>", "<t,r>", "|r|", "<rt,r>"] * 100, "cg": ["g0"] + ["<p,z>", "<r,r>"] * 100}
for form, oracle_op in ((sb.FORM_COEF, rows_op), (sb.FORM_FAITHFUL, cpu)):
    gpu = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=form, dirichlet=True)
    bvec = np.sin(0.37 * np.arange(n))
    bd = ctx.vector(bvec)
    for solver, S in (("cg", sb.CgSolver), ("bicgstab", sb.BiCgStabSolver)):
        want = orc.solve(solver, oracle_op, bvec, num_iterations=6, abs_tol=0.0, rel_tol=0.0, mode=orc.RED_TREE)
        for rep in range(4):
            s = S(num_iterations=6, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, use_graph=(rep % 2 == 1))
            x = ctx.zeros(n)
            s.solve(x, bd, gpu)
            k = min(len(s.trace), len(want.trace))
            bad = np.flatnonzero(s.trace[:k] != want.trace[:k])
            if len(bad):
                i = int(bad[0])
                print(f"form {form} {solver} rep {rep}: first mismatch at trace[{i}] = {names[solver][i]}: gpu {s.trace[i]!r} oracle {want.trace[i]!r} rel {abs(s.trace[i]-want.trace[i])/abs(want.trace[i]):.2e}")
            else:
                print(f"form {form} {solver} rep {rep}: trace identical ({k} entries); x equal: {np.array_equal(x.numpy(), want.x)}")
