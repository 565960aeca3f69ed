"""TEST INFRASTRUCTURE -- numpy restatement of the host-side integer work of the hot path.

Independent of stormruler_b200/csrc/sb_mesh_host.cpp and sb_part_host.cpp (different language,
different algorithms: lexsort / set operations / Python BFS instead of C++ record sorts), written from
the same conventions, which come from the reference (paths under /root/reference):

* local faces of a tetrahedron / hexahedron      source/Storm/Mallard/Shape.hpp:589-592, 833-837
* hexahedron -> 5 tetrahedral pieces             Shape.hpp:845-852; quadrangle -> 2 triangles :395-402
* complex-shape volume / barycentre              Shape.hpp:170-215
* a face is created by the first cell (cell order, then local-face order) that touches it; that cell
  is its inner cell                              source/Storm/Mallard/MeshUnstructured.hpp:509-554
* label-0 (interior) faces first                 MeshUnstructured.hpp:464-500
* permutations: perm[new] = old                  source/Storm/Utils/Permutations.hpp:77-103
* face distance = norm_2(centre_o - centre_i)    source_apps/playground/Playground.cpp:126

and, for what the reference does not have, from the rules stated in
include/stormb200.h ("partitioning") and sb_mesh_host.cpp ("Reverse Cuthill-McKee. Deterministic rules").
Parity status: the conventions are pinned by the reference's own 2-D mesh exports (tests/golden/mesh_*.npz
come from the reference mesh classes); the 3-D generalisation is a restatement ("parity unpinned" against
the reference itself, which is 2-D only at this commit) -- the tests pin the product against
this file bit for bit. Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

TET_FACES = np.array([[0, 2, 1, -1], [0, 1, 3, -1], [1, 2, 3, -1], [2, 0, 3, -1]])
HEX_FACES = np.array([[0, 3, 2, 1], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [0, 4, 7, 3], [4, 5, 6, 7]])
HEX_PIECES = np.array([[0, 3, 1, 4], [3, 2, 1, 6], [4, 5, 6, 1], [4, 6, 7, 3], [4, 3, 1, 6]])
KUHN = np.array([[0, 1, 2, 6], [0, 1, 5, 6], [0, 3, 2, 6], [0, 3, 7, 6], [0, 4, 5, 6], [0, 4, 7, 6]])
TILE = 2048


# ---- std::mt19937_64 + libstdc++ uniform_real_distribution<double> ---------------------------------
class MT19937_64:
    NN, MM = 312, 156
    MASK = (1 << 64) - 1

    def __init__(self, seed: int = 5489):
        mt = [0] * self.NN
        mt[0] = seed & self.MASK
        for i in range(1, self.NN):
            mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & self.MASK
        self.mt, self.idx = mt, self.NN

    def _twist(self):
        mt, NN, MM = self.mt, self.NN, self.MM
        UM, LM, A = 0xFFFFFFFF80000000, 0x7FFFFFFF, 0xB5026F5AA96619E9
        for i in range(NN):
            x = (mt[i] & UM) | (mt[(i + 1) % NN] & LM)
            mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ (A if x & 1 else 0)
        self.idx = 0

    def __call__(self) -> int:
        if self.idx >= self.NN:
            self._twist()
        x = self.mt[self.idx]
        self.idx += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        return x & self.MASK

    def canonical(self) -> float:
        """std::generate_canonical<double, 53> with a 64-bit engine: double(u) / 2^64, clamped below 1."""
        c = float(self()) / 18446744073709551616.0
        return np.nextafter(1.0, 0.0) if c >= 1.0 else c

    def uniform(self, a: float, b: float) -> float:
        return (b - a) * self.canonical() + a


# ---- box meshes ------------------------------------------------------------------------------------
def box_cells(kind: str, nx: int, ny: int, nz: int, jitter=0.2, seed_jitter=42, shuffle=True, seed_shuffle=43):
    """Node coordinates and cell-node table of sb_mesh_generate_box."""
    hx, hy, hz = 1.0 / nx, 1.0 / ny, 1.0 / nz
    eng = MT19937_64(seed_jitter)
    xyz = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))

    def nid(i, j, k):
        return (k * (ny + 1) + j) * (nx + 1) + i
    for k in range(nz + 1):
        for j in range(ny + 1):
            for i in range(nx + 1):
                x, y, z = i * hx, j * hy, k * hz
                if 0 < i < nx and 0 < j < ny and 0 < k < nz and jitter > 0.0:
                    x += eng.uniform(-jitter * hx, jitter * hx)
                    y += eng.uniform(-jitter * hy, jitter * hy)
                    z += eng.uniform(-jitter * hz, jitter * hz)
                xyz[nid(i, j, k)] = (x, y, z)
    cells = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                h = [nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k),
                     nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)]
                if kind == "hex":
                    cells.append(h)
                else:
                    cells.extend([[h[q] for q in t] for t in KUHN])
    cells = np.array(cells, np.int64)
    if shuffle:
        se = MT19937_64(seed_shuffle)
        for i in range(len(cells) - 1, 0, -1):
            j = se() % (i + 1)
            if j != i:
                cells[[i, j]] = cells[[j, i]]
    return xyz, cells


# ---- geometry, operation by operation ------------------------------------------------------------------
def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1)


def _dot3(a, b):
    return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]


def _length(a):
    return np.sqrt(((0.0 + a[:, 0] * a[:, 0]) + a[:, 1] * a[:, 1]) + a[:, 2] * a[:, 2])


def _tri_area(v1, v2, v3):
    return _length(_cross(v2 - v1, v3 - v1)) / 2.0


def _tri_center(v1, v2, v3):
    return ((v1 + v2) + v3) / 3.0


def _tet_volume(v1, v2, v3, v4):
    return np.abs(_dot3(v2 - v1, _cross(v3 - v1, v4 - v1))) / 6.0


def _tet_center(v1, v2, v3, v4):
    return (((v1 + v2) + v3) + v4) / 4.0


def cell_geometry(xyz, cells):
    if cells.shape[1] == 4:
        v = [xyz[cells[:, q]] for q in range(4)]
        return _tet_volume(*v), _tet_center(*v)
    vol = vc = None
    for p, piece in enumerate(HEX_PIECES):
        v = [xyz[cells[:, q]] for q in piece]
        dv = _tet_volume(*v)
        w = dv[:, None] * _tet_center(*v)
        vol, vc = (dv, w) if p == 0 else (vol + dv, vc + w)
    return vol, vc / vol[:, None]


def face_list(xyz, cells):
    """Face-list SoA of a cell soup in the current cell order: dict with face_cell [F,2] (inner, outer),
    face_area, face_dist, cell_vol, cell_ctr, bface_cell, bface_area, bface_dist."""
    n, npc = cells.shape
    lf = TET_FACES if npc == 4 else HEX_FACES
    nfc = len(lf)
    vol, ctr = cell_geometry(xyz, cells)
    # every (cell, local face) side with its sorted node key
    side_cell = np.repeat(np.arange(n), nfc)
    side_lf = np.tile(np.arange(nfc), n)
    nodes = np.where(lf[side_lf] >= 0, cells[side_cell[:, None], np.maximum(lf[side_lf], 0)], np.iinfo(np.int64).max)
    key = np.sort(nodes, axis=1)
    order = np.lexsort((side_cell, key[:, 3], key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    new = np.ones(len(order), bool)
    new[1:] = (ks[1:] != ks[:-1]).any(1)
    start = np.flatnonzero(new)
    count = np.diff(np.append(start, len(order)))
    assert count.max() <= 2, "non-manifold mesh"
    first = order[start]                                    # side of the lower cell id: the creator
    second = np.where(count == 2, order[np.minimum(start + 1, len(order) - 1)], -1)
    creator, creator_lf = side_cell[first], side_lf[first]
    other = np.where(second >= 0, side_cell[np.maximum(second, 0)], -1)
    interior = other >= 0
    # creation order: (creator cell, local face), interior faces first
    rank_key = np.lexsort((creator_lf, creator, ~interior))
    creator, creator_lf, other, interior = creator[rank_key], creator_lf[rank_key], other[rank_key], interior[rank_key]
    fn = np.maximum(lf[creator_lf], 0)
    v1, v2, v3 = xyz[cells[creator, fn[:, 0]]], xyz[cells[creator, fn[:, 1]]], xyz[cells[creator, fn[:, 2]]]
    if npc == 4:
        area, fc = _tri_area(v1, v2, v3), _tri_center(v1, v2, v3)
    else:
        v4 = xyz[cells[creator, fn[:, 3]]]
        a1, a2 = _tri_area(v1, v2, v3), _tri_area(v3, v4, v1)
        area = a1 + a2
        fc = (a1[:, None] * _tri_center(v1, v2, v3) + a2[:, None] * _tri_center(v3, v4, v1)) / area[:, None]
    xi = ctr[creator]
    F = int(interior.sum())
    # unit normals: triangle cross(v2-v1, v3-v1), quadrangle cross of the diagonals; oriented inner -> outer
    # (boundary: along face centre - cell centre)
    nrm = _cross(v2 - v1, v3 - v1) if npc == 4 else _cross(v3 - v1, v4 - v2)
    nrm = nrm / _length(nrm)[:, None]
    direction = np.where(interior[:, None], ctr[np.maximum(other, 0)] - xi, fc - xi)
    nrm = np.where((_dot3(nrm, direction) < 0.0)[:, None], -nrm, nrm)
    return dict(
        face_normal=nrm[:F], bface_normal=nrm[F:],
        n_cells=n, cell_vol=vol, cell_ctr=ctr,
        face_cell=np.stack([creator[:F], other[:F]], 1).astype(np.int32), face_area=area[:F],
        face_dist=_length(ctr[other[:F]] - xi[:F]),
        bface_cell=creator[F:].astype(np.int32), bface_area=area[F:], bface_dist=2.0 * _length(fc[F:] - xi[F:]))


# ---- reverse Cuthill-McKee ---------------------------------------------------------------------------
def rcm(n, face_cell):
    """perm[new] = old. Rules: sb_mesh_host.cpp 'Reverse Cuthill-McKee. Deterministic rules'."""
    adj = [[] for _ in range(n)]
    for a, b in face_cell.tolist():
        adj[a].append(b)
        adj[b].append(a)
    deg = [len(a) for a in adj]
    key = lambda v: (deg[v], v)  # noqa: E731
    by_deg = sorted(range(n), key=key)
    visited = [False] * n
    order, seed_pos = [], 0
    while len(order) < n:
        while visited[by_deg[seed_pos]]:
            seed_pos += 1
        seed = by_deg[seed_pos]
        # last BFS level from the seed (unvisited component only)
        level, seen = [seed], {seed}
        while True:
            nxt = []
            for v in level:
                for w in adj[v]:
                    if not visited[w] and w not in seen:
                        seen.add(w)
                        nxt.append(w)
            if not nxt:
                break
            level = nxt
        root = min(level, key=key)
        first = len(order)
        order.append(root)
        visited[root] = True
        h = first
        while h < len(order):
            nb = [w for w in adj[order[h]] if not visited[w]]
            nb = sorted(set(nb), key=key)
            for w in nb:
                visited[w] = True
            order.extend(nb)
            h += 1
    return np.array(order[::-1], np.int32)


def permute(xyz, cells, perm):
    """New cell soup with cell `new` = old cell perm[new]."""
    return xyz, cells[perm]


def permute_face_list(mesh: dict, perm, face_normal=None, bface_normal=None, cell_ctr=None):
    """Renumbering of a mesh that is given as its face list (include/stormb200.h: sb_mesh_from_faces).
    The local face index of a face in a cell is its ordinal among that cell's faces in the ORIGINAL list
    (interior faces first, then boundary faces); after the renumbering (cell `new` = old cell perm[new]) the
    creating (inner) cell of an interior face is the lower new id, faces are ordered by (creating cell, its local
    face index), interior faces first, and a normal flips when inner and outer swap."""
    n = int(mesh["n_cells"])
    perm = np.asarray(perm, np.int64)
    iperm = np.empty(n, np.int64)
    iperm[perm] = np.arange(n)
    fc = np.asarray(mesh["face_cell"], np.int64).reshape(-1, 2)
    bc = np.asarray(mesh["bface_cell"], np.int64)
    F, B = fc.shape[0], bc.shape[0]
    count = np.zeros(n, np.int64)
    lf = np.zeros((F, 2), np.int64)
    for f in range(F):
        for side in (0, 1):
            lf[f, side] = count[fc[f, side]]
            count[fc[f, side]] += 1
    blf = np.zeros(B, np.int64)
    for q in range(B):
        blf[q] = count[bc[q]]
        count[bc[q]] += 1
    a, b = iperm[fc[:, 0]], iperm[fc[:, 1]]
    swap = b < a
    inner, outer = np.where(swap, b, a), np.where(swap, a, b)
    ilf = np.where(swap, lf[:, 1], lf[:, 0])
    order = np.lexsort((ilf, inner))
    border = np.lexsort((blf, iperm[bc]))
    out = dict(n_cells=n, cell_vol=np.asarray(mesh["cell_vol"])[perm],
               face_cell=np.stack([inner[order], outer[order]], 1).astype(np.int32),
               face_area=np.asarray(mesh["face_area"])[order], face_dist=np.asarray(mesh["face_dist"])[order],
               bface_cell=iperm[bc][border].astype(np.int32), bface_area=np.asarray(mesh["bface_area"])[border],
               bface_dist=np.asarray(mesh["bface_dist"])[border])
    if face_normal is not None:
        sign = np.where(swap, -1.0, 1.0)[:, None]
        out["face_normal"] = (sign * np.asarray(face_normal))[order]
        out["bface_normal"] = np.asarray(bface_normal)[border]
    if cell_ctr is not None:
        out["cell_ctr"] = np.asarray(cell_ctr)[perm]
    return out


# ---- partitions ----------------------------------------------------------------------------------------
def slab_partition(n, n_parts):
    part = np.zeros(n, np.int32)
    for p in range(n_parts):
        part[n * p // n_parts: n * (p + 1) // n_parts] = p
    return part


def pad_up(n):
    return -(-n // TILE) * TILE


def local_maps(mesh: dict, part, rank: int, n_parts: int):
    """Local numbering, halo maps and local face list of one rank (include/stormb200.h: sb_local_mesh)."""
    fc = mesh["face_cell"].astype(np.int64)
    n = mesh["n_cells"]
    pa, pb = part[fc[:, 0]], part[fc[:, 1]]
    cut_a, cut_b = (pa == rank) & (pb != rank), (pb == rank) & (pa != rank)
    boundary = np.zeros(n, bool)
    boundary[fc[cut_a, 0]] = True
    boundary[fc[cut_b, 1]] = True
    halo = np.unique(np.concatenate([fc[cut_a, 1], fc[cut_b, 0]]))
    halo = halo[np.lexsort((halo, part[halo]))]
    owned = np.flatnonzero(part == rank)
    owned = np.concatenate([owned[~boundary[owned]], owned[boundary[owned]]])
    n_owned, n_interior = len(owned), int((~boundary[owned]).sum())
    halo_base = pad_up(n_owned)
    g2l = np.full(n, -1, np.int64)
    g2l[owned] = np.arange(n_owned)
    g2l[halo] = halo_base + np.arange(len(halo))
    nbr = np.unique(part[halo])
    recv_ptr = np.concatenate([[0], np.cumsum([(part[halo] == q).sum() for q in nbr])]).astype(np.int64)
    send, send_ptr = [], [0]
    for q in nbr:
        cells = np.unique(np.concatenate([fc[cut_a & (pb == q), 0], fc[cut_b & (pa == q), 1]]))
        send.append(g2l[cells])
        send_ptr.append(send_ptr[-1] + len(cells))
    # where my block starts in neighbour q's vectors
    send_dst = []
    for q in nbr:
        q_cut_a, q_cut_b = (pa == q) & (pb != q), (pb == q) & (pa != q)
        q_halo = np.unique(np.concatenate([fc[q_cut_a, 1], fc[q_cut_b, 0]]))
        send_dst.append(pad_up(int((part == q).sum())) + int((part[q_halo] < rank).sum()))
    keep = (pa == rank) | (pb == rank)
    bkeep = part[mesh["bface_cell"]] == rank
    cell_vol = np.ones(halo_base + len(halo))
    cell_vol[:n_owned] = mesh["cell_vol"][owned]
    cell_vol[halo_base:] = mesh["cell_vol"][halo]
    return dict(
        n_owned=n_owned, n_interior=n_interior, n_halo=len(halo), halo_base=halo_base,
        local_to_global=np.concatenate([owned, halo]).astype(np.int32),
        nbr_rank=nbr.astype(np.int32), send_ptr=np.array(send_ptr, np.int64), recv_ptr=recv_ptr,
        send_idx=(np.concatenate(send) if send else np.zeros(0)).astype(np.int32),
        send_dst=np.array(send_dst, np.int64),
        face_global=np.flatnonzero(keep).astype(np.int64), face_cell=g2l[fc[keep]].astype(np.int32),
        face_area=mesh["face_area"][keep], face_dist=mesh["face_dist"][keep], cell_vol=cell_vol,
        bface_global=np.flatnonzero(bkeep).astype(np.int64),
        bface_cell=g2l[mesh["bface_cell"][bkeep]].astype(np.int32), bface_area=mesh["bface_area"][bkeep],
        bface_dist=mesh["bface_dist"][bkeep])
