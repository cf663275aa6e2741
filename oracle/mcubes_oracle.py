"""CPU restatement (numpy) of marching cubes for the mesh-extraction step -- TEST INFRASTRUCTURE ONLY.

Reference call site: utils.generate_mesh (reference utils.py:172-182) -> `mcubes.marching_cubes(-1.0*gridvalues, 0)`
followed by `(v - v.min()) / v.max() * (bb_max - bb_min) + bb_min` when there are more than 10 vertices.
PyMCubes is a third-party dependency of the reference (absent from /root/reference, version unpinned, not installed
here), so this restates the published algorithm (Lorensen & Cline 1987; corner / edge numbering of P. Bourke) with
case tables derived by construction (scripts/gen_mc_tables.py).  PARITY UNPINNED against PyMCubes itself: in cubes
with an ambiguous face the triangulation may differ from its table; tests check invariants instead.

Conventions shared with csrc/mcubes.cu: a corner is "set" when its value is < iso; one vertex per crossed grid edge,
numbered by ascending (flat grid vertex index * 3 + axis); triangles by ascending cube index, table order inside."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import gen_mc_tables as _G  # noqa: E402

_EDGE_MASK, _N_TRIS, _TABLE, _MAX_T = _G.build_tables()
_TABLE = np.asarray(_TABLE, dtype=np.int64)
_N_TRIS = np.asarray(_N_TRIS, dtype=np.int64)
_OWNER = np.asarray(_G.EDGE_OWNER, dtype=np.int64)
_CORNERS = np.asarray(_G.CORNERS, dtype=np.int64)


def marching_cubes(u: np.ndarray, iso: float = 0.0):
    """u: (n, n, n) float32.  Returns (vertices (nv,3) float32 in index coordinates, triangles (nt,3) int64)."""
    u = np.asarray(u, dtype=np.float32)
    n = u.shape[0]
    assert u.shape == (n, n, n)
    below = u < np.float32(iso)
    # ---- vertices: crossed grid edges in ascending (vertex, axis) order ----
    flags = np.zeros((n, n, n, 3), dtype=bool)
    flags[:-1, :, :, 0] = below[:-1] != below[1:]
    flags[:, :-1, :, 1] = below[:, :-1] != below[:, 1:]
    flags[:, :, :-1, 2] = below[:, :, :-1] != below[:, :, 1:]
    flat = flags.reshape(-1)
    index = np.cumsum(flat) - flat                      # exclusive scan
    ids = np.nonzero(flat)[0]
    v, axis = ids // 3, ids % 3
    i, j, k = v // (n * n), (v // n) % n, v % n
    ua = u[i, j, k]
    ub = u[i + (axis == 0), j + (axis == 1), k + (axis == 2)]
    t = (np.float32(iso) - ua) / (ub - ua)             # float32, as the kernel
    verts = np.stack([i, j, k], axis=1).astype(np.float32)
    verts[np.arange(len(ids)), axis] += t
    # ---- triangles ----
    case = np.zeros((n - 1, n - 1, n - 1), dtype=np.int64)
    for c, (dx, dy, dz) in enumerate(_CORNERS):
        case |= below[dx:n - 1 + dx, dy:n - 1 + dy, dz:n - 1 + dz].astype(np.int64) << c
    case = case.reshape(-1)
    cells = np.nonzero(_N_TRIS[case])[0]
    ci, cj, ck = cells // ((n - 1) ** 2), (cells // (n - 1)) % (n - 1), cells % (n - 1)
    tris = []
    for t_idx in range(_MAX_T):
        sel = _N_TRIS[case[cells]] > t_idx
        if not sel.any():
            break
        e = _TABLE[case[cells[sel]], 3 * t_idx:3 * t_idx + 3]                       # (m, 3) edge ids
        ov = ((ci[sel, None] + _OWNER[e, 0]) * n + (cj[sel, None] + _OWNER[e, 1])) * n + (ck[sel, None] + _OWNER[e, 2])
        tris.append((cells[sel], np.full(sel.sum(), t_idx), index[ov * 3 + _OWNER[e, 3]]))
    if not tris:
        return verts, np.zeros((0, 3), dtype=np.int64)
    cell_id = np.concatenate([t[0] for t in tris])
    t_id = np.concatenate([t[1] for t in tris])
    tri = np.concatenate([t[2] for t in tris])
    order = np.lexsort((t_id, cell_id))
    return verts, tri[order]


def generate_mesh(gridvalues: np.ndarray, bb_min: float, bb_max: float):
    """reference utils.py:172-182 with the restated marching cubes."""
    vertices, triangles = marching_cubes(-1.0 * np.asarray(gridvalues, dtype=np.float32), 0.0)
    if len(vertices) > 10:
        vertices = (vertices - vertices.min()) / vertices.max()
        vertices = vertices * (bb_max - bb_min) + bb_min
    return vertices, triangles


def mesh_invariants(verts: np.ndarray, tris: np.ndarray):
    """(euler characteristic, boundary edges, non-manifold edges, signed volume, area)"""
    tris = np.asarray(tris, dtype=np.int64)
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    und = np.sort(e, axis=1)
    uniq, counts = np.unique(und, axis=0, return_counts=True)
    p = verts[tris].astype(np.float64)
    cross = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    vol = float(np.einsum("ij,ij->i", p[:, 0], cross).sum() / 6.0)
    area = float(np.linalg.norm(cross, axis=1).sum() / 2.0)
    used = np.unique(tris)
    return (len(used) - len(uniq) + len(tris), int((counts == 1).sum()), int((counts > 2).sum()), vol, area)
