#!/usr/bin/env python
"""gather_locality.py -- how local are the gathers of the operator apply? (host only, no GPU)

    python scripts/gather_locality.py [--axis 119] [--parts 8]

For the bench problem (jittered Kuhn tetrahedra, shuffled then RCM-renumbered) and for the local meshes of its METIS and
slab partitions: per owned row the distance |column - row|, the number of distinct 32-byte sectors of `x` a 64-row slice
(one warp stage of apply_kernel_tma) gathers, and the number of distinct 128-byte lines a 2048-row tile (one CTA) touches.
Written to answer one question of DESIGN.md 6: is the apply of a METIS part slower than the apply of a cube of the same
size because its gathers are less local? (No: 62-64 sectors per slice against 57 for the 1.23 M-cell cube and 63 for the
10.1 M-cell cube, where the kernel runs at its roofline.)"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from stormruler_b200 import capi  # noqa: E402
from stormruler_b200.mesh import CELL_TET, Mesh, Partition  # noqa: E402


def stats(name, face_cell, n_rows):
    a, b = face_cell[:, 0].astype(np.int64), face_cell[:, 1].astype(np.int64)
    rows, cols = np.concatenate([a, b]), np.concatenate([b, a])
    owned = rows < n_rows
    rows, cols = rows[owned], cols[owned]
    d = np.abs(cols - rows)
    sectors = np.unique((rows >> 6) * (1 << 32) + (cols >> 2)).size / (n_rows / 64)
    lines = np.unique((rows >> 11) * (1 << 32) + (cols >> 4)).size / (n_rows / 2048)
    print(f"{name:30s} rows {n_rows:9d}  entries/row {rows.size / n_rows:.2f}  |col-row| median {np.median(d):8.0f} "
          f"p90 {np.percentile(d, 90):8.0f}  32-B sectors per 64-row slice {sectors:5.1f}  128-B lines per 2048-row tile "
          f"{lines:6.1f}", flush=True)


def box(axis):
    m = Mesh.box(CELL_TET, axis, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43)
    m.renumber_rcm()
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--axis", type=int, default=119)
    ap.add_argument("--parts", type=int, default=8)
    args = ap.parse_args()
    per_rank_axis = max(2, round(args.axis / args.parts ** (1.0 / 3.0)))
    small = box(per_rank_axis)
    stats(f"cube {per_rank_axis}, RCM (one rank's size)", np.asarray(small.face_cell), small.n_cells)
    m = box(args.axis)
    stats(f"cube {args.axis}, RCM", np.asarray(m.face_cell), m.n_cells)
    for method, name in ((capi.PART_METIS, "metis"), (capi.PART_SLAB, "slab")):
        t = time.time()
        part = Partition(m, args.parts, method)
        print(f"{name}: {args.parts} parts in {time.time() - t:.1f} s, edge cut {part.info.edge_cut}", flush=True)
        for r in range(args.parts):
            loc = part.local(r)
            stats(f"  {name} rank {r} ({loc.n_nbr} nbrs, halo {loc.n_halo})", np.asarray(loc.face_cell), loc.n_owned)


if __name__ == "__main__":
    main()
