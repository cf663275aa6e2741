"""Marching cubes: the numpy restatement's invariants on CPU, the CUDA kernels against it on the GPU."""
import numpy as np
import pytest
import torch

from oracle import mcubes_oracle as M


def sphere_sdf(n, r=0.35, c=(0.0, 0.0, 0.0)):
    ax = np.linspace(-0.5, 0.5, n)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    return (np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - r).astype(np.float32)


def bumpy_sdf(n, seed=0):
    ax = np.linspace(-0.5, 0.5, n)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    torus = np.sqrt((np.sqrt(X ** 2 + Y ** 2) - 0.3) ** 2 + Z ** 2) - 0.1
    ball = np.sqrt((X - 0.2) ** 2 + Y ** 2 + (Z - 0.3) ** 2) - 0.12
    noise = 0.01 * np.random.default_rng(seed).standard_normal(X.shape)
    return (np.minimum(torus, ball) + noise).astype(np.float32)


def test_tables_are_consistent():
    import gen_mc_tables as G                      # scripts/ is put on sys.path by the oracle
    mask, ntri, table, max_t = G.build_tables()
    assert max_t == 5 and ntri[0] == 0 and ntri[255] == 0
    assert all(mask[c] == mask[255 - c] for c in range(256))       # complement crosses the same edges
    assert G.case_triangles(1) == [(0, 8, 3)]                      # the classic table's first entry
    for c in range(256):
        crossed = sum(1 for a, b in G.EDGES if ((c >> a) & 1) != ((c >> b) & 1))
        assert bin(mask[c]).count("1") == crossed                   # every crossed edge is used


def test_oracle_sphere_invariants():
    n = 40
    v, t = M.marching_cubes(-sphere_sdf(n), 0.0)
    chi, boundary, nonmanifold, vol, area = M.mesh_invariants(v, t)
    h = 1.0 / (n - 1)
    assert (chi, boundary, nonmanifold) == (2, 0, 0)
    assert vol > 0                                                  # normals point out of the shape for the -sdf input
    assert abs(vol * h ** 3 - 4 / 3 * np.pi * 0.35 ** 3) < 0.01 * 4 / 3 * np.pi * 0.35 ** 3
    assert abs(area * h * h - 4 * np.pi * 0.35 ** 2) < 0.01 * 4 * np.pi * 0.35 ** 2
    # every vertex lies on a grid edge and within half a cell of the analytic surface
    r = np.linalg.norm(v * h - 0.5, axis=1)
    assert np.abs(r - 0.35).max() < 0.5 * h


def test_oracle_is_crack_free_on_ambiguous_cases():
    v, t = M.marching_cubes(-bumpy_sdf(32), 0.0)
    assert len(t) > 1000
    assert M.mesh_invariants(v, t)[1] == 0                          # no boundary edge anywhere inside the grid


def test_oracle_generate_mesh_normalisation():
    v, t = M.generate_mesh(sphere_sdf(24), -0.5, 0.5)
    assert v.min() == -0.5 and len(t) > 0 and v.max() <= 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [("sphere", 33), ("bumpy", 48), ("bumpy", 97), ("empty", 16)])
def test_gpu_marching_cubes_equals_oracle(kind, n):
    from list_b200 import hotpath
    sdf = {"sphere": lambda: sphere_sdf(n), "bumpy": lambda: bumpy_sdf(n, seed=n), "empty": lambda: np.ones((n, n, n), np.float32)}[kind]()
    v_ref, t_ref = M.marching_cubes(-sdf, 0.0)
    v, t = hotpath.marching_cubes(torch.from_numpy(sdf).cuda(), 0.0, negate=True)
    assert v.shape == (len(v_ref), 3) and t.shape == (len(t_ref), 3)
    if len(v_ref):
        assert np.abs(v.cpu().numpy() - v_ref).max() <= 1e-6
        assert np.array_equal(t.cpu().numpy().astype(np.int64), t_ref)


@pytest.mark.gpu
def test_gpu_generate_mesh_matches_reference_recipe(tmp_path):
    from list_b200.network import executors
    sdf = sphere_sdf(64)
    mesh = executors.generate_mesh(sdf, -0.5, 0.5)
    v_ref, t_ref = M.generate_mesh(sdf, -0.5, 0.5)
    assert np.abs(np.asarray(mesh.vertices) - v_ref).max() <= 1e-6 and np.array_equal(np.asarray(mesh.faces), t_ref)
    mesh.export(str(tmp_path / "m.obj"))
    assert (tmp_path / "m.obj").stat().st_size > 1000


def two_balls_and_torus_sdf(n):
    ax = np.linspace(-0.5, 0.5, n)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    torus = np.sqrt((np.sqrt(X ** 2 + Y ** 2) - 0.3) ** 2 + Z ** 2) - 0.1
    b1 = np.sqrt((X - 0.25) ** 2 + (Y + 0.2) ** 2 + (Z - 0.33) ** 2) - 0.11
    b2 = np.sqrt((X + 0.3) ** 2 + (Y - 0.3) ** 2 + (Z + 0.3) ** 2) - 0.13
    return np.minimum(np.minimum(torus, b1), b2).astype(np.float32)


def _compare_with_marching_tets(u, verts, tris, smooth, vol_tol):
    """Bounds a marching-cubes mesh by the table-free marching-tetrahedra surface of the same grid (oracle/mtets_oracle.py):
    identical vertex set on the grid edges, same orientation, volume within vol_tol; for shapes the grid resolves also the
    same Euler characteristic and the same area within 1 %."""
    from oracle import mtets_oracle as T
    v2, t2, on_axis = T.marching_tets(u, 0.0)
    chi2, boundary2, nonmanifold2, vol2, area2 = M.mesh_invariants(v2, t2)
    chi, boundary, nonmanifold, vol, area = M.mesh_invariants(np.asarray(verts), np.asarray(tris))
    assert boundary2 == 0 and boundary == 0
    a = np.asarray(verts, dtype=np.float32)
    b = v2[on_axis]
    assert len(a) == len(b)
    assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])    # bit-identical vertex sets
    assert vol * vol2 > 0 and abs(vol - vol2) <= vol_tol * abs(vol2)
    if smooth:
        assert (nonmanifold, nonmanifold2) == (0, 0) and chi == chi2
        assert abs(area - area2) <= 0.01 * area2
    return chi, vol, vol2


@pytest.mark.parametrize("name,n,chi_expected", [("sphere", 40, 2), ("torus_balls", 48, 4)])
def test_oracle_agrees_with_marching_tetrahedra(name, n, chi_expected):
    """Second, independent extractor (no case table at all): pins what can be pinned of the marching-cubes step without
    PyMCubes -- vertex set, orientation, topology and enclosed volume."""
    u = -(sphere_sdf(n) if name == "sphere" else two_balls_and_torus_sdf(n))
    v, t = M.marching_cubes(u, 0.0)
    chi, vol, vol2 = _compare_with_marching_tets(u, v, t, smooth=True, vol_tol=5e-3)
    assert chi == chi_expected                                      # sphere: 2; torus (0) + two balls (2 + 2)
    # ambiguous faces everywhere (noise at the cell scale): the surfaces differ inside cells, the volume still agrees
    u = -bumpy_sdf(32)
    v, t = M.marching_cubes(u, 0.0)
    _compare_with_marching_tets(u, v, t, smooth=False, vol_tol=1e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["torus_balls", "bumpy"])
def test_gpu_marching_cubes_against_marching_tetrahedra(kind):
    from list_b200 import hotpath
    n = 48 if kind == "torus_balls" else 32
    u = -(two_balls_and_torus_sdf(n) if kind == "torus_balls" else bumpy_sdf(n))
    verts, tris = hotpath.marching_cubes(torch.from_numpy(u).cuda(), 0.0, negate=False)
    torch.cuda.synchronize()
    v_ref, t_ref = M.marching_cubes(u, 0.0)
    # the kernel's vertices equal the oracle's up to the last bit of the interpolation (same order); the bit-exact
    # vertex-set comparison with the tetrahedra then runs on the oracle's coordinates and the kernel's triangles
    assert verts.shape[0] == len(v_ref) and np.allclose(verts.cpu().numpy(), v_ref, atol=1e-5, rtol=0)
    _compare_with_marching_tets(u, v_ref, tris.cpu().numpy(), smooth=(kind == "torus_balls"),
                                vol_tol=5e-3 if kind == "torus_balls" else 1e-2)
