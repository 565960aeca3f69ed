"""Host-side mesh handle (sb_mesh): synthetic box meshes, cell-soup ingestion, RCM renumbering.
Pure host code in libstormb200.so -- works without a GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

CELL_TET, CELL_HEX, CELL_FACELIST = 0, 1, 2


class Mesh:
    def __init__(self, handle):
        self.lib = capi.load()
        self.handle = handle
        self._refresh()

    # -- constructors -----------------------------------------------------------------------------
    @staticmethod
    def box(kind: int, nx: int, ny: int | None = None, nz: int | None = None, jitter: float = 0.2,
            seed_jitter: int = 42, shuffle: bool = True, seed_shuffle: int = 43) -> "Mesh":
        """SURVEY.md 8d config 2/4 geometry: [0,1]^3, nx*ny*nz hexes (x6 Kuhn tets for CELL_TET)."""
        lib = capi.load()
        h = C.c_void_p()
        ny, nz = ny or nx, nz or nx
        capi.check(lib.sb_mesh_generate_box(kind, nx, ny, nz, jitter, seed_jitter, int(shuffle), seed_shuffle,
                                            C.byref(h)))
        return Mesh(h)

    @staticmethod
    def from_cells(kind: int, xyz: np.ndarray, cells: np.ndarray) -> "Mesh":
        lib = capi.load()
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        cells = np.ascontiguousarray(cells, np.int32)
        h = C.c_void_p()
        capi.check(lib.sb_mesh_from_cells(kind, xyz.shape[0], xyz.ctypes.data_as(capi.f64p), cells.shape[0],
                                          cells.ctypes.data_as(capi.i32p), C.byref(h)))
        return Mesh(h)

    @staticmethod
    def from_faces(mesh, centers=None, face_normals=None, bface_normals=None) -> "Mesh":
        """sb_mesh_from_faces: a handle over a face list (`mesh` needs n_cells, face_cell [F,2], face_area,
        face_dist, cell_vol, bface_cell, bface_area, bface_dist -- e.g. an export of the reference's own mesh
        classes, or a PolyMesh), so it can be RCM-renumbered and partitioned like a node-based mesh."""
        lib = capi.load()
        f64 = lambda a: np.ascontiguousarray(a, np.float64)  # noqa: E731
        fc = np.ascontiguousarray(mesh.face_cell, np.int32).reshape(-1)
        fa, fd, cv = f64(mesh.face_area), f64(mesh.face_dist), f64(mesh.cell_vol)
        bc = np.ascontiguousarray(mesh.bface_cell, np.int32)
        ba, bd = f64(mesh.bface_area), f64(mesh.bface_dist)
        soa = capi.MeshSoa(int(mesh.n_cells), int(fa.shape[0]), fc.ctypes.data_as(capi.i32p),
                           fa.ctypes.data_as(capi.f64p), fd.ctypes.data_as(capi.f64p), cv.ctypes.data_as(capi.f64p),
                           int(ba.shape[0]), bc.ctypes.data_as(capi.i32p), ba.ctypes.data_as(capi.f64p),
                           bd.ctypes.data_as(capi.f64p))
        opt = lambda a, cols: None if a is None else f64(a).reshape(-1, cols)  # noqa: E731
        ctr, fn, bn = opt(centers, 3), opt(face_normals, 3), opt(bface_normals, 3)
        ptr = lambda a: None if a is None else a.ctypes.data_as(capi.f64p)  # noqa: E731
        h = C.c_void_p()
        capi.check(lib.sb_mesh_from_faces(C.byref(soa), ptr(ctr), ptr(fn), ptr(bn), C.byref(h)))
        return Mesh(h)

    @staticmethod
    def read_tetgen(path_prefix: str) -> "Mesh":
        """3-D TetGen `<prefix>.node` + `<prefix>.ele` (file grammar of Mallard/IoTetgen.hpp:44-235)."""
        lib = capi.load()
        h = C.c_void_p()
        capi.check(lib.sb_mesh_read_tetgen(str(path_prefix).encode(), C.byref(h)))
        return Mesh(h)

    @staticmethod
    def read_tetgen_2d(path_prefix: str) -> "Mesh":
        """2-D Triangle files `<prefix>.node/.edge/.ele`, as the reference's read_mesh_from_tetgen +
        UnstructuredMesh<2,2> build them (face order, inner/outer, labels, geometry): sb_mesh_read_tetgen_2d."""
        lib = capi.load()
        h = C.c_void_p()
        capi.check(lib.sb_mesh_read_tetgen_2d(str(path_prefix).encode(), C.byref(h)))
        return Mesh(h)

    def bface_labels(self) -> np.ndarray:
        out = np.empty(self.n_bfaces, np.int32)
        capi.check(self.lib.sb_mesh_bface_labels(self.handle, out.ctypes.data_as(capi.i32p)))
        return out

    def __del__(self):
        try:
            if self.handle:
                self.lib.sb_mesh_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- views (zero-copy numpy views of the library-owned SoA) -----------------------------------
    def _refresh(self):
        soa = capi.MeshSoa()
        capi.check(self.lib.sb_mesh_get_soa(self.handle, C.byref(soa)))
        self.soa = soa
        self.n_cells, self.n_faces, self.n_bfaces = int(soa.n_cells), int(soa.n_faces), int(soa.n_bfaces)
        view = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)) if n > 0 else np.zeros(0)  # noqa: E731
        self.face_cell = view(soa.face_cell, 2 * self.n_faces).reshape(-1, 2) if self.n_faces else np.zeros((0, 2), np.int32)
        self.face_area, self.face_dist = view(soa.face_area, self.n_faces), view(soa.face_dist, self.n_faces)
        self.cell_vol = view(soa.cell_vol, self.n_cells)
        self.bface_cell = view(soa.bface_cell, self.n_bfaces) if self.n_bfaces else np.zeros(0, np.int32)
        self.bface_area, self.bface_dist = view(soa.bface_area, self.n_bfaces), view(soa.bface_dist, self.n_bfaces)

    def renumber_rcm(self) -> np.ndarray:
        """Reverse Cuthill-McKee; returns perm with perm[new] = old."""
        perm = np.empty(self.n_cells, np.int32)
        capi.check(self.lib.sb_mesh_renumber_rcm(self.handle, perm.ctypes.data_as(capi.i32p)))
        self._refresh()
        return perm

    def permute_cells(self, perm: np.ndarray):
        perm = np.ascontiguousarray(perm, np.int32)
        capi.check(self.lib.sb_mesh_permute_cells(self.handle, perm.ctypes.data_as(capi.i32p)))
        self._refresh()

    def cell_centers(self) -> np.ndarray:
        out = np.empty((self.n_cells, 3))
        capi.check(self.lib.sb_mesh_cell_centers(self.handle, out.ctypes.data_as(capi.f64p)))
        return out

    def face_normals(self):
        """(interior [F,3] inner -> outer, boundary [B,3] outward) unit normals."""
        fn, bn = np.empty((self.n_faces, 3)), np.empty((self.n_bfaces, 3))
        capi.check(self.lib.sb_mesh_face_normals(self.handle, fn.ctypes.data_as(capi.f64p), bn.ctypes.data_as(capi.f64p)))
        return fn, bn

    def face_flux(self, beta):
        """beta . n per interior / boundary face for a uniform velocity `beta` ((beta_x*n_x + beta_y*n_y) + beta_z*n_z)."""
        fn, bn = self.face_normals()
        bx, by, bz = (float(v) for v in beta)
        return (bx * fn[:, 0] + by * fn[:, 1]) + bz * fn[:, 2], (bx * bn[:, 0] + by * bn[:, 1]) + bz * bn[:, 2]

    @property
    def bandwidth(self) -> int:
        return int(self.lib.sb_mesh_bandwidth(self.handle))

    def write_vtk(self, path: str, fields: dict | None = None):
        """Legacy-VTK dump of the mesh and per-cell scalar fields (host arrays, current cell order), in the file
        grammar of the playground's save_vtk (Playground.cpp:65-109)."""
        fields = fields or {}
        arrs = [np.ascontiguousarray(v, np.float64) for v in fields.values()]
        for a in arrs:
            assert a.shape == (self.n_cells,), "one value per cell"
        names = (C.c_char_p * max(len(arrs), 1))(*[k.encode() for k in fields])
        ptrs = (capi.f64p * max(len(arrs), 1))(*[a.ctypes.data_as(capi.f64p) for a in arrs])
        capi.check(self.lib.sb_mesh_write_vtk(self.handle, str(path).encode(), len(arrs), names, ptrs))


class HexLattice:
    """Uniform hexahedral box [0,1]^3 of nx*ny*nz cells as a face list, generated directly (vectorised numpy, no
    node matching): the integer layout is that of `Mesh.box(CELL_HEX, ..., jitter=0, shuffle=False)` bit for bit
    (lattice cell order; faces created cell by cell in the local-face order of Hexahedron::faces, Shape.hpp:833-837:
    z-, y-, x+, y+, x-, z+; the creating cell is the inner one), the geometry is analytic (areas hy*hz etc., centre
    distances hx etc., mirror-ghost distance = the same spacing). For the sweep points where the node-based
    generator is too slow or too large (1e8 cells and beyond). Duck-types the `mesh` argument of FvmOperator."""

    def __init__(self, nx: int, ny: int | None = None, nz: int | None = None):
        ny, nz = ny or nx, nz or nx
        n = nx * ny * nz
        assert n < 2 ** 31 - 4096
        self.dims = (nx, ny, nz)
        hx, hy, hz = 1.0 / nx, 1.0 / ny, 1.0 / nz
        ids = np.arange(n, dtype=np.int32)
        i, j, k = ids % nx, (ids // nx) % ny, ids // (nx * ny)
        self.n_cells = n
        self.cell_vol = np.full(n, hx * hy * hz)
        # interior faces: cell c creates its x+, y+, z+ faces (local faces 2, 3, 5), in that order
        has = np.stack([i < nx - 1, j < ny - 1, k < nz - 1], axis=1)            # [n, 3]
        step = np.array([1, nx, nx * ny], np.int32)
        nbr = ids[:, None] + step[None, :]
        owner = np.broadcast_to(ids[:, None], has.shape)
        axis = np.broadcast_to(np.arange(3, dtype=np.int8)[None, :], has.shape)
        self.face_cell = np.stack([owner[has], nbr[has]], axis=1)
        self.face_axis = axis[has]
        area = np.array([hy * hz, hx * hz, hx * hy])
        dist = np.array([hx, hy, hz])
        self.face_area, self.face_dist = area[self.face_axis], dist[self.face_axis]
        del has, nbr, owner, axis
        # boundary faces per cell in local-face order z-, y-, x+, y+, x-, z+
        bnd = np.stack([k == 0, j == 0, i == nx - 1, j == ny - 1, i == 0, k == nz - 1], axis=1)   # [n, 6]
        baxis = np.array([2, 1, 0, 1, 0, 2], np.int8)
        bowner = np.broadcast_to(ids[:, None], bnd.shape)
        blf = np.broadcast_to(np.arange(6, dtype=np.int8)[None, :], bnd.shape)
        self.bface_cell = bowner[bnd]
        self.bface_lf = blf[bnd]
        self.bface_area, self.bface_dist = area[baxis[self.bface_lf]], dist[baxis[self.bface_lf]]
        self.n_faces, self.n_bfaces = int(self.face_cell.shape[0]), int(self.bface_cell.shape[0])

    def cell_centers(self) -> np.ndarray:
        nx, ny, nz = self.dims
        ids = np.arange(self.n_cells)
        i, j, k = ids % nx, (ids // nx) % ny, ids // (nx * ny)
        return np.stack([(i + 0.5) / nx, (j + 0.5) / ny, (k + 0.5) / nz], axis=1)

    @property
    def bandwidth(self) -> int:
        nx, ny, nz = self.dims
        return nx * ny if nz > 1 else (nx if ny > 1 else (1 if nx > 1 else 0))


class HexLatticeSlab:
    """One rank's local mesh of a HexLattice split into contiguous slabs of the lattice cell order (the split of
    SB_PART_SLAB: rank r owns cells [n*r//P, n*(r+1)//P)), built DIRECTLY from the lattice arithmetic: no global mesh, no
    partitioner, work and memory proportional to the slab. Every array is what
    `Partition(Mesh.from_faces(HexLattice(...)), P, PART_SLAB).local(rank)` produces, bit for bit (local order
    [interior owned | boundary owned | pad | halo by owner and id], faces in ascending global face index with their
    global numbers, halo send / receive maps) -- tests/test_mesh_host.py compares them. For config 4 at full size
    (49.8 M hexahedra on 8 GPUs: building and partitioning the global mesh costs minutes and 29 GB on rank 0).
    `local` is a LocalArrays (the `local` argument of multigpu.DistOperator), `info` the partition summary."""

    def __init__(self, nx: int, ny: int, nz: int, rank: int, n_parts: int):
        n = nx * ny * nz
        assert n < 2 ** 31 - 4096 and 0 <= rank < n_parts
        self.dims, self.n_global, self.rank, self.n_parts = (nx, ny, nz), n, rank, n_parts
        self.bounds = np.array([n * p // n_parts for p in range(n_parts + 1)], np.int64)
        lo, hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.lo, self.hi = lo, hi
        step = (1, nx, nx * ny)
        sxy = nx * ny
        hx, hy, hz = 1.0 / nx, 1.0 / ny, 1.0 / nz
        area = np.array([hy * hz, hx * hz, hx * hy])
        dist = np.array([hx, hy, hz])
        vol = hx * hy * hz

        # owned cells: interior first, then those with a face to another slab; both in ascending global id
        own = np.arange(lo, hi, dtype=np.int64)
        boundary = np.zeros(hi - lo, bool)
        for a in range(3):
            has_up, has_dn = self._has(own, a, +1), self._has(own, a, -1)
            boundary |= (has_up & (own + step[a] >= hi)) | (has_dn & (own - step[a] < lo))
        n_owned, n_interior = hi - lo, int((~boundary).sum())
        owned = np.concatenate([own[~boundary], own[boundary]])
        l_owned = np.empty(n_owned, np.int64)
        l_owned[~boundary] = np.arange(n_interior)
        l_owned[boundary] = n_interior + np.arange(n_owned - n_interior)
        halo = self._halo_of(lo, hi)                      # ascending id == (owner rank, id) order for slabs
        halo_base = _pad_tile(n_owned)

        def g2l(g):
            g = np.asarray(g, np.int64)
            inside = (g >= lo) & (g < hi)
            out = np.empty(g.shape, np.int64)
            out[inside] = l_owned[g[inside] - lo]
            out[~inside] = halo_base + np.searchsorted(halo, g[~inside])
            return out

        # faces with an owned end, in ascending global face index: created cell by cell (x+, y+, z+) by the lower cell
        creators = np.arange(max(0, lo - sxy), hi, dtype=np.int64)
        present = np.stack([self._has(creators, a, +1) for a in range(3)], axis=1)          # [m, 3]
        nbr = creators[:, None] + np.array(step, np.int64)[None, :]
        c_own = ((creators >= lo) & (creators < hi))[:, None]
        keep = present & (c_own | ((nbr >= lo) & (nbr < hi)))
        pos = np.cumsum(present, axis=1) - present                                           # index among the cell's faces
        fglob = (self._faces_before(creators)[:, None] + pos)[keep]
        owner_b = np.broadcast_to(creators[:, None], keep.shape)[keep]
        nbr_k = nbr[keep]
        axis = np.broadcast_to(np.arange(3)[None, :], keep.shape)[keep]
        face_cell = np.stack([g2l(owner_b), g2l(nbr_k)], axis=1).astype(np.int32)

        # halo exchange maps
        part_of = lambda g: np.searchsorted(self.bounds, g, side="right") - 1  # noqa: E731
        halo_part = part_of(halo)
        nbr_rank = np.unique(halo_part)
        recv_ptr = np.concatenate([[0], np.cumsum([(halo_part == q).sum() for q in nbr_rank])]).astype(np.int64)
        o_in, n_in = (owner_b >= lo) & (owner_b < hi), (nbr_k >= lo) & (nbr_k < hi)
        send, send_ptr, send_dst = [], [0], []
        for q in nbr_rank:
            qlo, qhi = int(self.bounds[q]), int(self.bounds[q + 1])
            mine = np.concatenate([owner_b[o_in & (nbr_k >= qlo) & (nbr_k < qhi)],
                                   nbr_k[n_in & (owner_b >= qlo) & (owner_b < qhi)]])
            cells = np.unique(mine)
            send.append(g2l(cells))
            send_ptr.append(send_ptr[-1] + len(cells))
            q_halo = self._halo_of(qlo, qhi)
            send_dst.append(_pad_tile(qhi - qlo) + int((q_halo < lo).sum()))

        # boundary faces of owned cells, per cell in local-face order z-, y-, x+, y+, x-, z+
        i, j, k = own % nx, (own // nx) % ny, own // sxy
        bnd = np.stack([k == 0, j == 0, i == nx - 1, j == ny - 1, i == 0, k == nz - 1], axis=1)
        baxis = np.array([2, 1, 0, 1, 0, 2])
        bpos = np.cumsum(bnd, axis=1) - bnd
        bglob = (self._bfaces_before(own)[:, None] + bpos)[bnd]
        bcell = np.broadcast_to(own[:, None], bnd.shape)[bnd]
        blf = np.broadcast_to(np.arange(6)[None, :], bnd.shape)[bnd]

        cell_vol = np.ones(halo_base + len(halo))
        cell_vol[:n_owned] = vol
        cell_vol[halo_base:] = vol
        scalars = dict(rank=rank, n_parts=n_parts, n_owned=n_owned, n_interior=n_interior, n_halo=len(halo),
                       halo_base=halo_base, n_cells=halo_base + len(halo), n_faces=len(fglob), n_bfaces=len(bglob),
                       n_nbr=len(nbr_rank))
        arrays = dict(local_to_global=np.concatenate([owned, halo]), face_cell=face_cell.reshape(-1),
                      face_area=area[axis], face_dist=dist[axis], cell_vol=cell_vol, bface_cell=g2l(bcell),
                      bface_area=area[baxis[blf]], bface_dist=dist[baxis[blf]], face_global=fglob, nbr_rank=nbr_rank,
                      send_ptr=np.array(send_ptr), recv_ptr=recv_ptr,
                      send_idx=np.concatenate(send) if send else np.zeros(0), send_dst=np.array(send_dst),
                      bface_global=bglob)
        self.local = LocalArrays(scalars, arrays)
        self.cell_vol_value = vol

    # -- lattice arithmetic --------------------------------------------------------------------------------------
    def _has(self, g, axis: int, sign: int):
        """Does cell g have a lattice neighbour along `axis` in direction `sign`?"""
        nx, ny, nz = self.dims
        c = (g % nx, (g // nx) % ny, g // (nx * ny))[axis]
        return c < (nx, ny, nz)[axis] - 1 if sign > 0 else c > 0

    def _halo_of(self, lo: int, hi: int) -> np.ndarray:
        """Ascending global ids of the cells outside [lo, hi) that share a face with a cell inside."""
        nx, ny, _ = self.dims
        step, sxy = (1, nx, nx * ny), nx * ny
        below = np.arange(max(0, lo - sxy), lo, dtype=np.int64)
        above = np.arange(hi, min(self.n_global, hi + sxy), dtype=np.int64)
        kb, ka = np.zeros(below.shape, bool), np.zeros(above.shape, bool)
        for a in range(3):
            t = below + step[a]
            kb |= self._has(below, a, +1) & (t >= lo) & (t < hi)
            t = above - step[a]
            ka |= self._has(above, a, -1) & (t >= lo) & (t < hi)
        return np.concatenate([below[kb], above[ka]])

    def _faces_before(self, c):
        """Number of interior faces created by the cells with id < c (each creates its x+, y+, z+ faces)."""
        nx, ny, nz = self.dims
        i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
        fx = (c // nx) * (nx - 1) + i
        fy = k * nx * (ny - 1) + j * nx + np.where(j < ny - 1, i, 0)
        fz = np.minimum(c, (nz - 1) * nx * ny)
        return fx + fy + fz

    def _bfaces_before(self, c):
        """Number of boundary faces of the cells with id < c."""
        nx, ny, nz = self.dims
        sxy = nx * ny
        i, j, k = c % nx, (c // nx) % ny, c // sxy
        return (np.minimum(c, sxy) + np.maximum(0, c - (nz - 1) * sxy)            # k == 0, k == nz-1
                + k * nx + np.where(j > 0, nx, i)                                 # j == 0
                + k * nx + np.where(j == ny - 1, i, 0)                            # j == ny-1
                + c // nx + (i > 0)                                               # i == 0
                + c // nx)                                                        # i == nx-1

    # -- what the drivers need besides the local mesh ---------------------------------------------------------------
    def owned_centers(self) -> np.ndarray:
        nx, ny, nz = self.dims
        g = self.local.owned_global.astype(np.int64)
        return np.stack([(g % nx + 0.5) / nx, ((g // nx) % ny + 0.5) / ny, (g // (nx * ny) + 0.5) / nz], axis=1)

    def info(self) -> dict:
        """Partition summary (the keys of sb_part_info) without building any other rank's mesh."""
        P, b = self.n_parts, self.bounds
        owned = np.diff(b)
        halos = np.array([len(self._halo_of(int(b[r]), int(b[r + 1]))) for r in range(P)])
        cap = max(_pad_tile(int(owned[r])) + _pad_tile(max(int(halos[r]), 1)) for r in range(P))
        nx, ny, _ = self.dims
        sxy = nx * ny
        cand = np.unique(np.concatenate([np.arange(max(0, int(B) - sxy), int(B), dtype=np.int64) for B in b[1:-1]]
                                        + [np.zeros(0, np.int64)]))
        part_of = lambda g: np.searchsorted(b, g, side="right") - 1  # noqa: E731
        cut = 0
        for a, st in enumerate((1, nx, sxy)):
            ok = self._has(cand, a, +1)
            cut += int((part_of(cand[ok]) != part_of(cand[ok] + st)).sum())
        return dict(n_cells=self.n_global, edge_cut=cut, min_owned=int(owned.min()), max_owned=int(owned.max()),
                    max_halo=int(halos.max()), vec_capacity=int(cap))


def _pad_tile(n: int) -> int:
    return -(-int(n) // 2048) * 2048   # the kernels' 2048-row CTA tile (include/stormb200.h: sb_local_mesh.halo_base)


class PolyMesh:
    """Synthetic polyhedral mesh for the dual-polyhedra leg of the apply sweep:
    the Voronoi tessellation of a body-centred cubic lattice, i.e. truncated octahedra with 14 faces
    (6 squares to the axis neighbours, 8 hexagons to the diagonal neighbours; F ~ 7 N), optionally
    stretched per axis. N = 2 n^3 cells: corner sites (i, j, k) a and centre sites (i+1/2, j+1/2, k+1/2) a,
    a = 1/n, interleaved so neighbours stay close in memory. Faces follow the reference's conventions:
    created cell by cell in local-face order, the creating (lower) cell is the inner one
    (MeshUnstructured.hpp:509-554); a face whose neighbour site lies outside the lattice is a boundary
    face with the mirror-ghost distance. Duck-types the `mesh` argument of FvmOperator / oracle FaceMesh.
    Host numpy only (a workload generator, not part of the hot path)."""

    #: the 14 neighbour directions in half-lattice units (2 = one lattice step along an axis)
    DIRS = np.array([(-2, 0, 0), (2, 0, 0), (0, -2, 0), (0, 2, 0), (0, 0, -2), (0, 0, 2)] +
                    [(dx, dy, dz) for dz in (-1, 1) for dy in (-1, 1) for dx in (-1, 1)], np.int64)

    def __init__(self, n: int, stretch=(1.0, 1.0, 1.0)):
        n = int(n)
        assert n >= 1 and 2 * n ** 3 < 2 ** 31 - 4096
        self.n, self.stretch = n, tuple(float(v) for v in stretch)
        sx, sy, sz = self.stretch
        a = 1.0 / n
        N = 2 * n ** 3
        ids = np.arange(N, dtype=np.int64)
        s, q = ids & 1, ids >> 1
        i, j, k = q % n, (q // n) % n, q // (n * n)
        # positions in half-lattice units: corner (2i, 2j, 2k), centre (2i+1, 2j+1, 2k+1)
        hx, hy, hz = 2 * i + s, 2 * j + s, 2 * k + s
        self._half = (hx, hy, hz)
        det = sx * sy * sz
        sq_area = (a * a / 8.0) * np.array([sy * sz, sx * sz, sx * sy])            # squares, normal x / y / z
        hex_area = (3.0 * np.sqrt(3.0) * a * a / 16.0) / np.sqrt(3.0) * det * np.sqrt(sx ** -2 + sy ** -2 + sz ** -2)
        nbr = np.empty((N, 14), np.int64)
        area = np.empty(14)
        dist = np.empty(14)
        for d, (dx, dy, dz) in enumerate(self.DIRS):
            px, py, pz = hx + dx, hy + dy, hz + dz
            ps = px & 1                              # parity picks the sub-lattice (all three coordinates agree)
            pi, pj, pk = (px - ps) >> 1, (py - ps) >> 1, (pz - ps) >> 1
            ok = (pi >= 0) & (pi < n) & (pj >= 0) & (pj < n) & (pk >= 0) & (pk < n)
            nbr[:, d] = np.where(ok, 2 * ((pk * n + pj) * n + pi) + ps, -1)
            area[d] = sq_area[d // 2] if d < 6 else hex_area
            dist[d] = 0.5 * a * np.sqrt((sx * dx) ** 2 + (sy * dy) ** 2 + (sz * dz) ** 2)
        owner = np.broadcast_to(ids[:, None], nbr.shape)
        dcol = np.broadcast_to(np.arange(14)[None, :], nbr.shape)
        interior = nbr > owner                       # the lower cell creates the face and is its inner cell
        boundary = nbr < 0
        self.n_cells = N
        self.face_cell = np.stack([owner[interior], nbr[interior]], axis=1).astype(np.int32)
        self.face_dir = dcol[interior].astype(np.int8)
        self.face_area, self.face_dist = area[self.face_dir], dist[self.face_dir]
        self.bface_cell = owner[boundary].astype(np.int32)
        self.bface_dir = dcol[boundary].astype(np.int8)
        self.bface_area, self.bface_dist = area[self.bface_dir], dist[self.bface_dir]
        self.n_faces, self.n_bfaces = int(self.face_cell.shape[0]), int(self.bface_cell.shape[0])
        self.cell_vol = np.full(N, 0.5 * a ** 3 * det)

    @staticmethod
    def bcc(n: int, stretch=(1.0, 1.0, 1.0)) -> "PolyMesh":
        return PolyMesh(n, stretch)

    def cell_centers(self) -> np.ndarray:
        hx, hy, hz = self._half
        h = 0.5 / self.n
        sx, sy, sz = self.stretch
        return np.stack([sx * h * hx, sy * h * hy, sz * h * hz], axis=1).astype(np.float64)

    def face_normals(self):
        """Unit normals (interior inner -> outer, boundary outward): M^-T d normalised, M = diag(stretch)."""
        nrm = self.DIRS / np.asarray(self.stretch)[None, :]
        nrm = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
        return nrm[self.face_dir], nrm[self.bface_dir]

    def face_flux(self, beta):
        fn, bn = self.face_normals()
        bx, by, bz = (float(v) for v in beta)
        return (bx * fn[:, 0] + by * fn[:, 1]) + bz * fn[:, 2], (bx * bn[:, 0] + by * bn[:, 1]) + bz * bn[:, 2]

    @property
    def bandwidth(self) -> int:
        return int(np.abs(self.face_cell[:, 1].astype(np.int64) - self.face_cell[:, 0]).max()) if self.n_faces else 0

    def to_mesh(self) -> Mesh:
        """The same mesh as a library handle (sb_mesh_from_faces) with centres and normals attached: can be
        renumbered (RCM) and partitioned for the multi-GPU path."""
        fn, bn = self.face_normals()
        return Mesh.from_faces(self, self.cell_centers(), fn, bn)


class LocalView:
    """One rank's local mesh (sb_local_mesh) as numpy views; duck-types the `mesh` argument of
    FvmOperator / oracle FaceMesh (n_cells, face_cell, ...)."""

    def __init__(self, part: "Partition", rank: int):
        lm = capi.LocalMesh()
        capi.check(part.lib.sb_part_local(part.handle, rank, C.byref(lm)))
        self.struct, self._part = lm, part
        self.rank, self.n_parts = int(lm.rank), int(lm.n_parts)
        self.n_owned, self.n_interior = int(lm.n_owned), int(lm.n_interior)
        self.n_halo, self.halo_base = int(lm.n_halo), int(lm.halo_base)
        arr = lambda p, n, dt=None: (np.ctypeslib.as_array(p, shape=(n,)) if n > 0 else np.zeros(0, dt))  # noqa: E731
        self.local_to_global = arr(lm.local_to_global, self.n_owned + self.n_halo, np.int32)
        s = lm.soa
        self.n_cells, self.n_faces, self.n_bfaces = int(s.n_cells), int(s.n_faces), int(s.n_bfaces)
        self.face_cell = arr(s.face_cell, 2 * self.n_faces, np.int32).reshape(-1, 2)
        self.face_area, self.face_dist = arr(s.face_area, self.n_faces), arr(s.face_dist, self.n_faces)
        self.cell_vol = arr(s.cell_vol, self.n_cells)
        self.bface_cell = arr(s.bface_cell, self.n_bfaces, np.int32)
        self.bface_area, self.bface_dist = arr(s.bface_area, self.n_bfaces), arr(s.bface_dist, self.n_bfaces)
        self.face_global = arr(lm.face_global, self.n_faces, np.int64)
        self.n_nbr = int(lm.n_nbr)
        self.nbr_rank = arr(lm.nbr_rank, self.n_nbr, np.int32)
        self.send_ptr = arr(lm.send_ptr, self.n_nbr + 1, np.int64)
        self.recv_ptr = arr(lm.recv_ptr, self.n_nbr + 1, np.int64)
        self.send_idx = arr(lm.send_idx, int(self.send_ptr[-1]) if self.n_nbr else 0, np.int32)
        self.send_dst = arr(lm.send_dst, self.n_nbr, np.int64)
        self.bface_global = arr(lm.bface_global, self.n_bfaces, np.int64)

    @property
    def owned_global(self) -> np.ndarray:
        return self.local_to_global[:self.n_owned]

    @property
    def halo_global(self) -> np.ndarray:
        return self.local_to_global[self.n_owned:]


LOCAL_ARRAYS = (("local_to_global", np.int32), ("face_cell", np.int32), ("face_area", np.float64),
                ("face_dist", np.float64), ("cell_vol", np.float64), ("bface_cell", np.int32),
                ("bface_area", np.float64), ("bface_dist", np.float64), ("face_global", np.int64),
                ("nbr_rank", np.int32), ("send_ptr", np.int64), ("recv_ptr", np.int64), ("send_idx", np.int32),
                ("send_dst", np.int64), ("bface_global", np.int64))
LOCAL_SCALARS = ("rank", "n_parts", "n_owned", "n_interior", "n_halo", "halo_base", "n_cells", "n_faces", "n_bfaces",
                 "n_nbr")


class LocalArrays:
    """A rank's local mesh rebuilt from plain arrays (what LocalView exposes), e.g. after it travelled from the rank
    that partitioned the global mesh (multigpu.scatter_mesh). Same attributes as LocalView, and a `struct`
    (sb_local_mesh) pointing into the arrays it keeps alive."""

    def __init__(self, scalars: dict, arrays: dict):
        for k in LOCAL_SCALARS:
            setattr(self, k, int(scalars[k]))
        self._keep = {}
        for k, dt in LOCAL_ARRAYS:
            a = np.ascontiguousarray(arrays[k], dt)
            self._keep[k] = a
            setattr(self, k, a.reshape(-1, 2) if k == "face_cell" else a)
        K = self._keep
        ip, lp, dp = capi.i32p, capi.i64p, capi.f64p
        ptr = lambda k, t: K[k].ctypes.data_as(t)  # noqa: E731
        soa = capi.MeshSoa(self.n_cells, self.n_faces, ptr("face_cell", ip), ptr("face_area", dp), ptr("face_dist", dp),
                           ptr("cell_vol", dp), self.n_bfaces, ptr("bface_cell", ip), ptr("bface_area", dp),
                           ptr("bface_dist", dp))
        self.struct = capi.LocalMesh(self.rank, self.n_parts, self.n_owned, self.n_interior, self.n_halo, self.halo_base,
                                     ptr("local_to_global", ip), soa, ptr("face_global", lp), self.n_nbr,
                                     ptr("nbr_rank", ip), ptr("send_ptr", lp), ptr("send_idx", ip), ptr("recv_ptr", lp),
                                     ptr("send_dst", lp), ptr("bface_global", lp))

    @staticmethod
    def from_view(v) -> "LocalArrays":
        return LocalArrays({k: getattr(v, k) for k in LOCAL_SCALARS},
                           {k: np.array(getattr(v, k), dt, copy=True).reshape(-1) for k, dt in LOCAL_ARRAYS})

    @property
    def owned_global(self) -> np.ndarray:
        return self.local_to_global[:self.n_owned]

    @property
    def halo_global(self) -> np.ndarray:
        return self.local_to_global[self.n_owned:]


class Partition:
    """sb_part: a split of the mesh's cell graph into n_parts (METIS k-way or RCM slabs)."""

    def __init__(self, mesh: Mesh, n_parts: int, method: int = capi.PART_METIS, part=None):
        self.lib, self.mesh = capi.load(), mesh   # the partition borrows the mesh arrays: keep it alive
        h = C.c_void_p()
        if part is not None:
            part = np.ascontiguousarray(part, np.int32)
            capi.check(self.lib.sb_part_from_array(mesh.handle, n_parts, part.ctypes.data_as(capi.i32p), C.byref(h)))
        else:
            capi.check(self.lib.sb_part_create(mesh.handle, n_parts, method, C.byref(h)))
        self.handle, self.n_parts = h, n_parts
        info = capi.PartInfo()
        capi.check(self.lib.sb_part_get_info(h, C.byref(info)))
        self.info = info
        p = capi.i32p()
        capi.check(self.lib.sb_part_get_array(h, C.byref(p)))
        self.part = np.ctypeslib.as_array(p, shape=(mesh.n_cells,))

    def local(self, rank: int) -> LocalView:
        return LocalView(self, rank)

    def __del__(self):
        try:
            if self.handle:
                self.lib.sb_part_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
