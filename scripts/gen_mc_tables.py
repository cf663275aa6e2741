"""Generates the marching-cubes case tables (csrc/mc_tables.h) by construction.

The reference meshes its SDF grids with PyMCubes (`mcubes.marching_cubes(-grid, 0)`, reference utils.py:172-182), a
third-party dependency that is neither vendored in /root/reference nor installed here (version unpinned by the
reference), and the classic 256-entry triangle table it uses is not available offline.  The tables are therefore
derived from the published algorithm (Lorensen & Cline 1987, corner / edge numbering of P. Bourke's "Polygonising a
scalar field"):
  * a cube corner is "set" when its value is below the isovalue; edge e is crossed when its two corners differ;
  * on every cube face the crossed edges are joined by segments; a face with two diagonal set corners (the ambiguous
    configuration) always separates the set corners.  The rule only looks at the face's four corners, so the two
    cubes sharing a face agree and the surface has no cracks;
  * segments are oriented with the set side on their left when the face is seen from outside the cube, chained
    into closed loops and fan-triangulated.  Triangle normals then point towards the set (below-isovalue) side, which
    for the reference's `-sdf` input is the outside of the shape.
Topology can differ from PyMCubes' table in ambiguous cubes only; that parity is unpinned (no PyMCubes here) and
the tests check invariants instead (watertightness, Euler characteristic, orientation, area).

Importable (oracle/mcubes_oracle.py uses build_tables()); run as a script to rewrite csrc/mc_tables.h."""
import os

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# faces as corner cycles, counter-clockwise when seen from OUTSIDE the cube
FACES = [(0, 3, 2, 1),   # z = 0 (outward normal -z)
         (4, 5, 6, 7),   # z = 1
         (0, 1, 5, 4),   # y = 0
         (2, 3, 7, 6),   # y = 1
         (0, 4, 7, 3),   # x = 0
         (1, 2, 6, 5)]   # x = 1
EDGE_OF = {}
for _e, (_a, _b) in enumerate(EDGES):
    EDGE_OF[(_a, _b)] = _e
    EDGE_OF[(_b, _a)] = _e
# edge e of cube (i,j,k) is owned by grid vertex (i,j,k)+EDGE_OWNER[e][:3] along axis EDGE_OWNER[e][3] (0=x,1=y,2=z)
EDGE_OWNER = []
for _a, _b in EDGES:
    _pa, _pb = CORNERS[_a], CORNERS[_b]
    _axis = [i for i in range(3) if _pa[i] != _pb[i]][0]
    _lo = _pa if _pa[_axis] == 0 else _pb
    EDGE_OWNER.append((_lo[0], _lo[1], _lo[2], _axis))


def _face_normal_check():
    """FACES are CCW from outside: (c1-c0) x (c3-c0) points along the outward normal."""
    import numpy as np
    for f in FACES:
        p = [np.array(CORNERS[c], float) for c in f]
        n = np.cross(p[1] - p[0], p[3] - p[0])
        centre = sum(p) / 4 - 0.5
        assert np.dot(n, centre) > 0, f


def case_triangles(case: int):
    """List of triangles (edge-index triples) of one of the 256 corner configurations."""
    inside = [(case >> c) & 1 for c in range(8)]
    nxt = {}                                    # directed segments: crossing edge -> next crossing edge
    for f in FACES:
        s = [inside[c] for c in f]
        # walk the face cycle; a segment starts on the edge where we ENTER the set region (unset -> set) ... see below
        # crossing edges of the face in cycle order
        cross = [(i, EDGE_OF[(f[i], f[(i + 1) % 4])]) for i in range(4) if s[i] != s[(i + 1) % 4]]
        if not cross:
            continue
        # For every maximal run of set corners along the cycle, the segment cutting the run off goes from the edge
        # where the run ENDS (set -> unset) to the edge where it BEGINS (unset -> set): seen from outside with the
        # cycle counter-clockwise, the set run is then on the left of the directed segment.
        if sum(s) == 2 and s[0] == s[2]:        # ambiguous face: two diagonal set corners, each cut off on its own
            runs = [[i] for i in range(4) if s[i]]
        else:
            start = next(i for i in range(4) if s[i] and not s[(i - 1) % 4])
            run = []
            i = start
            while s[i % 4] and len(run) < 4:
                run.append(i % 4)
                i += 1
            runs = [run]
        for run in runs:
            first, last = run[0], run[-1]
            e_begin = EDGE_OF[(f[(first - 1) % 4], f[first])]       # unset -> set
            e_end = EDGE_OF[(f[last], f[(last + 1) % 4])]           # set -> unset
            assert e_end not in nxt
            nxt[e_end] = e_begin
    tris = []
    seen = set()
    for e0 in sorted(nxt):
        if e0 in seen:
            continue
        loop = [e0]
        seen.add(e0)
        e = nxt[e0]
        while e != e0:
            loop.append(e)
            seen.add(e)
            e = nxt[e]
        assert len(loop) >= 3
        for i in range(1, len(loop) - 1):
            tris.append((loop[0], loop[i], loop[i + 1]))
    return tris


def build_tables():
    """(edge_mask[256], n_tris[256], tri_table[256][MAX*3])"""
    _face_normal_check()
    all_tris = [case_triangles(c) for c in range(256)]
    max_t = max(len(t) for t in all_tris)
    edge_mask, n_tris, table = [], [], []
    for tris in all_tris:
        m = 0
        for t in tris:
            for e in t:
                m |= 1 << e
        edge_mask.append(m)
        n_tris.append(len(tris))
        flat = [e for t in tris for e in t]
        table.append(flat + [-1] * (max_t * 3 - len(flat)))
    return edge_mask, n_tris, table, max_t


def write_header(path):
    edge_mask, n_tris, table, max_t = build_tables()
    with open(path, "w") as f:
        f.write("// GENERATED by scripts/gen_mc_tables.py -- do not edit.  Marching-cubes case tables derived by construction\n"
                "// (corner / edge numbering of Bourke's \"Polygonising a scalar field\"; ambiguous faces separate the set corners).\n"
                "#pragma once\n\n")
        f.write(f"#define LIST_MC_MAX_TRIS {max_t}\n\n")
        f.write("__constant__ unsigned char kMcNumTris[256] = {\n  " + ", ".join(map(str, n_tris)) + "};\n\n")
        f.write("// triangle corners as cube-edge indices (0..11), -1 padded\n")
        f.write(f"__constant__ signed char kMcTriTable[256][{max_t * 3}] = {{\n")
        for row in table:
            f.write("  {" + ", ".join(f"{v:2d}" for v in row) + "},\n")
        f.write("};\n\n")
        f.write("// cube edge e -> (di, dj, dk, axis) of the grid vertex / axis that owns it\n")
        f.write("__constant__ unsigned char kMcEdgeOwner[12][4] = {\n")
        for o in EDGE_OWNER:
            f.write("  {" + ", ".join(map(str, o)) + "},\n")
        f.write("};\n")
    return max_t


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "learning-implicitly-from-spatial-transformers-network_b200", "csrc", "mc_tables.h")
    mt = write_header(out)
    print(f"wrote {out} (max {mt} triangles per cube)")
