#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE ITSELF (run in the build container).

Needs /root/reference (read-only mount) and the oracle/_ref build (`make -C oracle`):
  * oracle/_ref/ref_mesh_tool   -- the reference's own mesh reader / UnstructuredMesh / CellField /
                                   CgSolver, compiled verbatim (oracle/ref_build/ref_mesh.cpp)
  * oracle/_ref/libref_solvers.so -- the reference's nine solver headers compiled verbatim on a host
                                   vector (oracle/ref_build/ref_solvers.cpp)

Outputs (small, committed):
  tests/golden/mesh_<name>.npz       face-list SoA of tests/_data/mesh/<name>.1.* as the reference's
                                     mesh classes see it (entity order, inner/outer, areas, volumes)
  tests/golden/cg_native_<name>.npz  CgSolver on the reference's own CellField: history, final x
  tests/golden/solvers_<name>.npz    all ten solvers through ref_solve: histories, final errors,
                                     iteration counts, reduction traces (head), x of cg/bicgstab
  tests/golden/cahn_hilliard_square_nb.npz  the playground's Cahn-Hilliard time step (Playground.cpp:133-210)
                                     run by the reference's own mesh / CellField / map / CgSolver on
                                     square_nb.1: initial c (glibc rand()), c after each of 2 steps, CG reports
                                     and residual histories (`--only ch` regenerates just this file and
                                     cahn_hilliard_uniformed_square_nb.npz: 3 steps through the reference's
                                     solve_non_uniform, which is what makes its CG converge on this affine operator)
  tests/golden/mesh_step.npz, cg_native_step.npz   the same two for the largest reference mesh, step.1
                                     (`--only step`)
  tests/golden/blas1_kat.npz         known answers of tests/unit/BitternReductions.cpp /
                                     BitternMath.cpp evaluated by the reference templates

Config 1 of BASELINE.json / SURVEY.md 8d: y = x - dt * div grad x, dt = 0.05, b[k] = sin(0.37 k),
x0 = 0, rel tol 1e-10, abs tol 0, <= 500 iterations.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

REF_DATA = "/root/reference/tests/_data/mesh"
OUT = os.path.dirname(os.path.abspath(__file__))
DT, ITERS, RTOL = 0.05, 500, 1e-10
TRACE_HEAD = 64


CH_STEPS = 2


def make_cahn_hilliard(tmp):
    name = "square_nb"
    ch_bin = f"{tmp}/{name}_ch.bin"
    subprocess.run([orc.REF_MESH_TOOL, "ch", f"{REF_DATA}/{name}.1.", str(CH_STEPS), ch_bin], check=True)
    ch = orc.read_ch_dump(ch_bin)
    out = dict(c0=ch["c0"], tau=1.0e-3, Gamma=1.0e-4, sigma=2.0, num_steps=CH_STEPS)
    for k, st in enumerate(ch["steps"]):
        out[f"step{k}_c"] = st["c"]
        out[f"step{k}_hist"] = st["hist"]
        out[f"step{k}_stats"] = np.array([st["converged"], st["iterations"], st["abs_err"], st["rel_err"]])
    np.savez_compressed(f"{OUT}/cahn_hilliard_{name}.npz", **out)
    # the same steps through the reference's solve_non_uniform (the operator is affine): CG converges in ~50 iterations
    subprocess.run([orc.REF_MESH_TOOL, "ch", f"{REF_DATA}/{name}.1.", "3", ch_bin, "uniformed"], check=True)
    ch = orc.read_ch_dump(ch_bin)
    out = dict(c0=ch["c0"], num_steps=3)
    for k, st in enumerate(ch["steps"]):
        out[f"step{k}_c"] = st["c"]
        out[f"step{k}_stats"] = np.array([st["converged"], st["iterations"], st["abs_err"], st["rel_err"]])
    np.savez_compressed(f"{OUT}/cahn_hilliard_uniformed_{name}.npz", **out)


def make_step(tmp):
    """The third config-1 mesh, step.1 (79 672 triangles): the reference mesh classes' export and the reference's own
    CgSolver on its own CellField (500 iterations, not converged: abs 0.23301109656816443). The solution is
    kept at every 8th cell to keep the fixture small; the residual history is complete."""
    name = "step"
    prefix = f"{REF_DATA}/{name}.1."
    mesh_bin, cg_bin = f"{tmp}/{name}.bin", f"{tmp}/{name}_cg.bin"
    subprocess.run([orc.REF_MESH_TOOL, "export", prefix, mesh_bin], check=True)
    subprocess.run([orc.REF_MESH_TOOL, "cg", prefix, str(DT), str(ITERS), str(RTOL), cg_bin], check=True)
    mesh, extra = orc.read_mesh_export(mesh_bin)
    np.savez_compressed(
        f"{OUT}/mesh_{name}.npz", n_cells=mesh.n_cells, face_cell=mesh.face_cell,
        face_area=mesh.face_area, face_dist=mesh.face_dist, cell_vol=mesh.cell_vol,
        bface_cell=mesh.bface_cell, bface_area=mesh.bface_area, bface_dist=mesh.bface_dist,
        bface_label=extra["bface_label"], n_nodes=extra["n_nodes"],
        n_faces_total=extra["n_faces_total"], n_face_labels=extra["n_face_labels"])
    cg = orc.read_cg_dump(cg_bin)
    np.savez_compressed(f"{OUT}/cg_native_{name}.npz", converged=cg["converged"], iterations=cg["iterations"],
                        abs_err=cg["abs_err"], rel_err=cg["rel_err"], hist=cg["hist"], x_every_8th=cg["x"][::8],
                        x_norm=np.linalg.norm(cg["x"]), dt=DT, num_iterations=ITERS, rel_tol=RTOL)


def main():
    assert os.path.isdir(REF_DATA), "reference tree not mounted"
    orc.build()
    tmp = tempfile.mkdtemp()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "step":
        make_step(tmp)
        return
    make_cahn_hilliard(tmp)
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "ch":
        return
    make_step(tmp)
    for name in ("square_nb", "rectangle"):
        prefix = f"{REF_DATA}/{name}.1."
        mesh_bin, cg_bin = f"{tmp}/{name}.bin", f"{tmp}/{name}_cg.bin"
        subprocess.run([orc.REF_MESH_TOOL, "export", prefix, mesh_bin], check=True)
        subprocess.run([orc.REF_MESH_TOOL, "cg", prefix, str(DT), str(ITERS), str(RTOL), cg_bin], check=True)
        mesh, extra = orc.read_mesh_export(mesh_bin)
        np.savez_compressed(
            f"{OUT}/mesh_{name}.npz", n_cells=mesh.n_cells, face_cell=mesh.face_cell,
            face_area=mesh.face_area, face_dist=mesh.face_dist, cell_vol=mesh.cell_vol,
            bface_cell=mesh.bface_cell, bface_area=mesh.bface_area, bface_dist=mesh.bface_dist,
            bface_label=extra["bface_label"], n_nodes=extra["n_nodes"],
            n_faces_total=extra["n_faces_total"], n_face_labels=extra["n_face_labels"])
        cg = orc.read_cg_dump(cg_bin)
        np.savez_compressed(f"{OUT}/cg_native_{name}.npz", converged=cg["converged"],
                            iterations=cg["iterations"], abs_err=cg["abs_err"], rel_err=cg["rel_err"],
                            hist=cg["hist"], x=cg["x"], dt=DT, num_iterations=ITERS, rel_tol=RTOL)
        if name != "square_nb":
            continue
        # all solvers, on the Helmholtz operator (interior faces only: the playground's operator) ...
        op = orc.FaceOp(mesh, prefill=1, dt=-DT)
        b = np.sin(0.37 * np.arange(mesh.n_cells))
        out = {}
        for s in orc.REF_SOLVERS:
            r = orc.ref_solve(s, op, b, num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
            out[f"{s}_hist"] = r.hist
            out[f"{s}_stats"] = np.array([r.converged, r.iterations, r.abs_err, r.rel_err, r.n_apply,
                                          len(r.trace)], dtype=np.float64)
            out[f"{s}_trace_head"] = r.trace[:TRACE_HEAD]
            out[f"{s}_x"] = r.x
        # ... and CG/BiCGStab on the Dirichlet Poisson operator (boundary ghost rows, configs 2/4)
        opd = orc.FaceOp(mesh, prefill=0, dt=-1.0, dirichlet=True)
        for s in ("cg", "bicgstab"):
            r = orc.ref_solve(s, opd, b, num_iterations=ITERS, abs_tol=0.0, rel_tol=RTOL)
            out[f"poisson_{s}_hist"] = r.hist
            out[f"poisson_{s}_stats"] = np.array([r.converged, r.iterations, r.abs_err, r.rel_err,
                                                  r.n_apply, len(r.trace)], dtype=np.float64)
            out[f"poisson_{s}_x"] = r.x
        np.savez_compressed(f"{OUT}/solvers_{name}.npz", **out)
    # BLAS-1 known answers: BitternReductions.cpp:59-76,99-115 ; BitternMath.cpp:136-151
    R = orc.ref()
    m1, m2, m3 = np.array([1.0, -2.0, 3.0, -4.0]), np.array([5.0, -6.0, 7.0, -8.0]), np.array([3.0, 9.0, 2.0, -1.0])
    p = lambda a: a.ctypes.data_as(orc._f64p)  # noqa: E731
    rng = np.random.default_rng(7)
    big_a, big_b = rng.standard_normal(5000), rng.standard_normal(5000)
    rnd = np.zeros(16)
    R.ref_fill_randomly_generic(16, p(rnd))  # first 16 draws of the reference's static engine
    np.savez_compressed(
        f"{OUT}/blas1_kat.npz", mat1=m1, mat2=m2, mat3=m3,
        dot_mat1_mat2=R.ref_dot(4, p(m1), p(m2)),         # CHECK_EQ 70.0
        norm2_mat1=R.ref_norm2(4, p(m1)),                 # CHECK_NEAR 5.47723 (eps 1e-5)
        expr1=np.array([21.0, -152.0, 53.0, -74.0]),      # mat1 + 10*(mat2 - mat3)
        big_a=big_a, big_b=big_b, dot_big=R.ref_dot(5000, p(big_a), p(big_b)),
        norm2_big=R.ref_norm2(5000, p(big_a)), fill_randomly_head=rnd)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f"  {f:32s} {os.path.getsize(os.path.join(OUT, f)) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
