"""Marching TETRAHEDRA in numpy -- TEST INFRASTRUCTURE ONLY: a second, table-free iso-surface extractor used to bound the
marching-cubes step (reference utils.py:172-173, `mcubes.marching_cubes(-grid, 0)`; PyMCubes itself is not available
offline, see oracle/mcubes_oracle.py).

Every cube is split into the six tetrahedra around its main diagonal (Freudenthal / Kuhn split: corner 000, then one
axis step at a time to 111, one tetrahedron per axis order), which is face-consistent between neighbouring cubes, so the
surface is watertight without any case table: a tetrahedron with one corner on the other side yields one triangle, with
two corners a quad.  Nothing here shares code or tables with scripts/gen_mc_tables.py.  The vertices that fall on
axis-parallel grid edges are computed with the same float32 formula as marching cubes, so they must coincide with the
marching-cubes vertex set exactly; Euler characteristic, enclosed volume and area of the two surfaces must agree for
shapes the grid resolves (tests/test_mcubes.py)."""
import itertools

import numpy as np


def marching_tets(u: np.ndarray, iso: float = 0.0):
    """u: (n, n, n) float32.  Returns (vertices (nv, 3) float32 in index coordinates, triangles (nt, 3) int64,
    on_axis (nv,) bool: the vertex lies on an axis-parallel grid edge).  A corner is "set" when its value is < iso;
    triangles are oriented with the normal pointing from the set side to the unset side, as marching cubes does."""
    u = np.asarray(u, dtype=np.float32)
    n = u.shape[0]
    assert u.shape == (n, n, n)
    iso = np.float32(iso)
    ci, cj, ck = np.meshgrid(np.arange(n - 1), np.arange(n - 1), np.arange(n - 1), indexing="ij")
    base = np.stack([ci.ravel(), cj.ravel(), ck.ravel()], axis=1)             # (cells, 3)
    tri_a, tri_b = [], []                                                     # endpoints (flat grid ids) of the 3 crossed edges
    flat = lambda p: (p[..., 0] * n + p[..., 1]) * n + p[..., 2]              # noqa: E731
    for perm in itertools.permutations(range(3)):
        steps = np.zeros((4, 3), dtype=np.int64)
        for s, axis in enumerate(perm):
            steps[s + 1] = steps[s]
            steps[s + 1, axis] += 1
        corners = base[:, None, :] + steps[None, :, :]                        # (cells, 4, 3)
        vals = u[corners[..., 0], corners[..., 1], corners[..., 2]]           # (cells, 4)
        below = vals < iso
        nset = below.sum(axis=1)
        # orientation reference: signed volume of the tetrahedron (p1-p0, p2-p0, p3-p0), the same for all cells
        e = (steps[1:] - steps[0]).astype(np.float64)
        tet_sign = np.sign(np.linalg.det(e))
        for lone_is_set, count in ((True, 1), (False, 3)):                    # one corner alone on its side -> one triangle
            sel = np.nonzero(nset == count)[0]
            if len(sel) == 0:
                continue
            b = below[sel]
            lone = np.argmax(b if lone_is_set else ~b, axis=1)                # index of the lone corner
            others = np.array([[j for j in range(4) if j != l] for l in range(4)])[lone]    # (m, 3) in ascending order
            c = corners[sel]
            pl = c[np.arange(len(sel)), lone]
            po = c[np.arange(len(sel))[:, None], others]                      # (m, 3, 3)
            # (lone, o0, o1, o2) is an even or odd permutation of (0,1,2,3): removing index l from 0..3 costs l swaps
            parity = np.where(lone % 2 == 0, 1.0, -1.0) * tet_sign
            # triangle (lone->o0, lone->o1, lone->o2) has its normal pointing towards the lone corner iff the
            # tetrahedron (lone, o0, o1, o2) is positively oriented; it must point to the UNSET side
            flip = (parity > 0) == lone_is_set
            a = np.repeat(flat(pl)[:, None], 3, axis=1)
            bb = flat(po)
            bb = np.where(flip[:, None], bb[:, ::-1], bb)
            tri_a.append(a)
            tri_b.append(bb)
        sel = np.nonzero(nset == 2)[0]                                         # two and two: a quad, split into two triangles
        if len(sel):
            b = below[sel]
            order = np.argsort(~b, axis=1, kind="stable")                      # set corners first, ascending inside each group
            s0, s1, u0, u1 = order[:, 0], order[:, 1], order[:, 2], order[:, 3]
            c = corners[sel]
            r = np.arange(len(sel))
            P = lambda idx: c[r, idx]                                          # noqa: E731
            # quad around the set pair: (s0-u0, s0-u1, s1-u1, s1-u0); orientation from the permutation parity of
            # (s0, s1, u0, u1) relative to (0, 1, 2, 3)
            perm_idx = np.stack([s0, s1, u0, u1], axis=1)
            inv = np.zeros(len(sel), dtype=np.int64)
            for x in range(4):
                for y in range(x + 1, 4):
                    inv += perm_idx[:, x] > perm_idx[:, y]
            parity = np.where(inv % 2 == 0, 1.0, -1.0) * tet_sign
            qa = np.stack([flat(P(s0)), flat(P(s0)), flat(P(s1)), flat(P(s1))], axis=1)
            qb = np.stack([flat(P(u0)), flat(P(u1)), flat(P(u1)), flat(P(u0))], axis=1)
            # for a positively oriented (s0, s1, u0, u1) the cycle above is seen clockwise from the unset side
            flip = parity > 0
            qa = np.where(flip[:, None], qa[:, ::-1], qa)
            qb = np.where(flip[:, None], qb[:, ::-1], qb)
            tri_a.append(qa[:, [0, 1, 2]]); tri_b.append(qb[:, [0, 1, 2]])
            tri_a.append(qa[:, [0, 2, 3]]); tri_b.append(qb[:, [0, 2, 3]])
    if not tri_a:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int64), np.zeros((0,), bool)
    A = np.concatenate(tri_a)
    B = np.concatenate(tri_b)
    lo, hi = np.minimum(A, B), np.maximum(A, B)
    key = lo * (n ** 3) + hi
    uniq, inverse = np.unique(key.ravel(), return_inverse=True)
    tris = inverse.reshape(-1, 3)
    lo_u, hi_u = uniq // (n ** 3), uniq % (n ** 3)
    pa = np.stack([lo_u // (n * n), (lo_u // n) % n, lo_u % n], axis=1)
    pb = np.stack([hi_u // (n * n), (hi_u // n) % n, hi_u % n], axis=1)
    ua = u[pa[:, 0], pa[:, 1], pa[:, 2]]
    ub = u[pb[:, 0], pb[:, 1], pb[:, 2]]
    t = (iso - ua) / (ub - ua)                                                 # float32, lower flat index first (as marching cubes)
    verts = pa.astype(np.float32) + t[:, None] * (pb - pa).astype(np.float32)
    on_axis = (pb - pa).sum(axis=1) == 1
    return verts, tris.astype(np.int64), on_axis
