"""Time plain sb_apply and the streaming kernels in isolation (wall clock around N launches)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stormruler_b200 as sb
from stormruler_b200.mesh import Mesh, CELL_TET

n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
ctx = sb.Context(0)
m = Mesh.box(CELL_TET, n, shuffle=True); m.renumber_rcm()
op = sb.FvmOperator(ctx, m, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
N = m.n_cells
rng = np.random.default_rng(0)
x = ctx.vector(rng.standard_normal(N)); y = ctx.zeros(N); z = ctx.zeros(N); w = ctx.zeros(N)
alg = op.info.algorithmic_bytes_per_apply

def timeit(f, reps=40):
    for _ in range(5): f()
    ctx.sync(); t = time.perf_counter()
    for _ in range(reps): f()
    ctx.sync(); return (time.perf_counter() - t) / reps

t = timeit(lambda: op.mul(y, x))
print(f"SB_DEBUG={os.environ.get('SB_DEBUG','0')} V1={os.environ.get('SB_APPLY_V1','0')} apply(same x): {t*1e6:7.1f} us  {alg/t/1e9:6.0f} GB/s", flush=True)
def chain():
    op.mul(y, x); op.mul(z, y); op.mul(w, z); op.mul(x, w)
t = timeit(chain, 10) / 4
print(f"   apply(chained y=Ax, z=Ay, ...): {t*1e6:7.1f} us  {alg/t/1e9:6.0f} GB/s", flush=True)
v = sb.expr.v
t = timeit(lambda: (v(x) + 0.5 * v(y)).assign_to(z))
print(f"   z = x + 0.5 y (3V): {t*1e6:7.1f} us  {24*N/t/1e9:6.0f} GB/s")
t = timeit(lambda: (v(x) + 0.5 * (v(y) - 0.25 * v(w))).assign_to(z))
print(f"   z = x + b(y - c w) (4V): {t*1e6:7.1f} us  {32*N/t/1e9:6.0f} GB/s")
t = timeit(lambda: ctx.dot(x, y))
print(f"   dot(x,y) incl. host sync (2V): {t*1e6:7.1f} us  {16*N/t/1e9:6.0f} GB/s")
t = timeit(lambda: z.copy_from(x))
print(f"   cudaMemcpy D2D (2V): {t*1e6:7.1f} us  {16*N/t/1e9:6.0f} GB/s")
