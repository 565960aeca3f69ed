"""Measure the operator-apply kernel under different cell orderings (gather-locality experiment).
Usage: python scripts/order_sweep.py [n_per_axis]   (GPU box)"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stormruler_b200 as sb
from stormruler_b200.mesh import Mesh, CELL_TET, CELL_HEX


def morton_perm(c, bits=10):
    q = np.minimum((c * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    def spread(v):
        v = v & np.uint64(0x3FF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x30000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x300F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x30C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x9249249)
        return v
    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    return np.argsort(key, kind="stable").astype(np.int32)


def time_apply(ctx, mesh, iters=20, label=""):
    op = sb.FvmOperator(ctx, mesh, prefill=0, dt=-1.0, form=sb.FORM_COEF, dirichlet=True)
    n = mesh.n_cells
    rng = np.random.default_rng(0)
    b = ctx.vector(rng.standard_normal(n))
    for solver, slots in ((sb.CgSolver, (0,)), (sb.BiCgStabSolver, (1, 3))):
        s = solver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, record=False)
        x = ctx.zeros(n)
        s.solve(x, b, op)
        s = solver(num_iterations=iters, absolute_error_tolerance=0.0, relative_error_tolerance=0.0, record=False, profile=True)
        x = ctx.zeros(n)
        s.solve(x, b, op)
        alg = op.info.algorithmic_bytes_per_apply
        ms = [s.kernel_ms[k] / iters for k in slots]
        print(f"{label:28s} {solver.__name__:14s} apply ms {['%.4f' % m for m in ms]}  GB/s {[round(alg / m / 1e6) for m in ms]}"
              f"  iter ms {s.iter_ms / iters:.4f}  bw {mesh.bandwidth}", flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
    ctx = sb.Context(0)
    m = Mesh.box(CELL_TET, n, shuffle=False)
    time_apply(ctx, m, label="tet natural (hex-major)")
    m.permute_cells(morton_perm(m.cell_centers()))
    time_apply(ctx, m, label="tet morton")
    m.renumber_rcm()
    time_apply(ctx, m, label="tet rcm(after morton)")
    m = Mesh.box(CELL_TET, n, shuffle=True)
    m.renumber_rcm()
    time_apply(ctx, m, label="tet shuffle+rcm (bench)")
    m = Mesh.box(CELL_TET, n, shuffle=True)
    time_apply(ctx, m, label="tet random order")
    # chain mesh: perfect gather locality, same width 4 and same bytes
    N = m.n_cells
    class Chain: pass
    ch = Chain()
    i = np.arange(N - 2, dtype=np.int32)
    ch.n_cells = N
    ch.face_cell = np.concatenate([np.stack([i, i + 1], 1), np.stack([i, i + 2], 1)]).astype(np.int32)
    F = ch.face_cell.shape[0]
    ch.face_area = np.ones(F); ch.face_dist = np.ones(F); ch.cell_vol = np.ones(N)
    ch.bface_cell = np.zeros(0, np.int32); ch.bface_area = np.zeros(0); ch.bface_dist = np.zeros(0)
    ch.bandwidth = 2
    time_apply(ctx, ch, label="chain (i+-1,i+-2)")


if __name__ == "__main__":
    main()
