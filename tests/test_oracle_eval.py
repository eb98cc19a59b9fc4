"""CPU: evaluation-path oracle (marching cubes, sampling, chamfer C restatement, metrics)."""
import os

import numpy as np
import torch
from scipy.spatial import cKDTree

from oracle import eval3d as E
from oracle.mc_tables import validate_tri_table, TRI_TABLE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _field(kind, n):
    g = np.linspace(-1.5, 1.5, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    if kind == "sphere":
        return (np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 1.0).astype(np.float32)
    return (np.sqrt((np.sqrt(X ** 2 + Y ** 2) - 0.9) ** 2 + Z ** 2) - 0.35).astype(np.float32)


def _edge_counts(f):
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    u, c = np.unique(e, axis=0, return_counts=True)
    return u, c


def test_case_table_is_consistent_and_watertight():
    assert validate_tri_table()
    assert sum(len(r) // 3 for r in TRI_TABLE) == 820   # classic table: 820 triangles over 256 cases


def test_generated_header_in_sync():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_mc_header
    assert open(os.path.join(ROOT, "zeroshape_b200", "csrc", "mc_tables.h")).read() == gen_mc_header.render()


def test_marching_cubes_known_answers():
    n = 33
    for kind, chi, area in (("sphere", 2, 4 * np.pi), ("torus", 0, 4 * np.pi ** 2 * 0.9 * 0.35)):
        v, f = E.marching_cubes(_field(kind, n), 0.0)
        u, c = _edge_counts(f)
        assert set(c) == {2}                      # closed 2-manifold
        assert len(v) - len(u) + len(f) == chi    # Euler characteristic
        a = E.mesh_area(v / (n - 1) * 3.0 - 1.5, f)
        assert abs(a - area) / area < 0.01
        # every vertex lies on a grid edge: exactly one non-integer coordinate at most
        frac = np.abs(v - np.round(v)) > 1e-12
        assert (frac.sum(axis=1) <= 1).all()


def test_marching_cubes_iso_is_inclusive_and_empty_cases():
    vol = np.zeros((5, 5, 5), np.float32)
    v, f = E.marching_cubes(vol, 0.0)            # all values <= iso -> everything inside -> empty
    assert len(v) == 0 and len(f) == 0
    vol[2, 2, 2] = 1.0                           # one outside voxel -> closed octahedron-like surface
    v, f = E.marching_cubes(vol, 0.5)
    assert len(v) == 6 and len(f) == 8
    np.testing.assert_allclose(np.sort(np.abs(v - 2).sum(axis=1)), 0.5)


def test_reference_vertex_scaling_quirk():
    v = np.array([[0.0, 64.0, 128.0]])
    np.testing.assert_allclose(E.scale_vertices(v, 129, -1.5, 1.5), v / 129 * 3 - 1.5)


def test_surface_sampling_is_area_weighted_and_on_surface():
    n = 33
    v, f = E.marching_cubes(_field("sphere", n), 0.0)
    v = v / (n - 1) * 3.0 - 1.5
    pts = E.sample_surface(v, f, 20000, np.random.RandomState(0))
    r = np.linalg.norm(pts, axis=1)
    assert abs(r.mean() - 1.0) < 0.01 and r.max() < 1.01
    assert np.abs(pts.mean(axis=0)).max() < 0.03     # uniform over the sphere
    assert E.sample_surface(v, f[:0], 7, np.random.RandomState(0)).shape == (7, 3)


def test_chamfer_c_oracle_vs_kdtree_and_ties():
    rs = np.random.RandomState(1)
    a, b = rs.rand(2, 300, 3).astype(np.float32), rs.rand(2, 517, 3).astype(np.float32)
    d1, d2, i1, i2 = E.chamfer_nn(a, b)
    for k in range(2):
        dd, ii = cKDTree(b[k].astype(np.float64)).query(a[k].astype(np.float64))
        assert (ii == i1[k]).all()
        np.testing.assert_allclose(np.sqrt(d1[k]), dd, rtol=1e-5, atol=1e-7)
    # duplicates: lowest index wins
    b2 = np.concatenate([b[:, :5], b[:, :5], b], axis=1)
    _, _, i1b, _ = E.chamfer_nn(b[:, :5], b2)
    assert (i1b == np.arange(5)[None]).all()
    # hand-computable lattice case
    p = np.array([[[0, 0, 0], [1, 0, 0]]], np.float32)
    q = np.array([[[0, 0, 2], [1, 1, 0], [5, 5, 5]]], np.float32)
    d1, d2, i1, i2 = E.chamfer_nn(p, q)
    assert d1.tolist() == [[2.0, 1.0]] and i1.tolist() == [[1, 1]]
    assert d2.tolist() == [[4.0, 1.0, 66.0]] and i2.tolist() == [[0, 1, 1]]


def test_fscore_and_normalize():
    d1 = torch.tensor([[0.004, 0.006, 0.3]])
    d2 = torch.tensor([[0.001, 0.5]])
    f = E.fscore(d1, d2, (0.005, 0.01))
    p, r = 1 / 3, 1 / 2
    np.testing.assert_allclose(f[0, 0].item(), 2 * p * r / (p + r), rtol=1e-6)
    assert E.fscore(torch.ones(1, 3), torch.ones(1, 3), (0.5,)).item() == 0.0       # NaN -> 0
    pc = torch.rand(2, 100, 3) * torch.tensor([2.0, 1.0, 9.0])
    npc = E.normalize_pc(pc)
    ext = npc.max(dim=1)[0] - npc.min(dim=1)[0]
    assert torch.allclose(ext[:, :2].max(dim=1)[0], torch.ones(2), atol=1e-5)        # z extent ignored


def test_rotation_sphere_is_orthonormal():
    R = E.rotation_sphere(3, 2, 2)
    assert R.shape == (12, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(12, 3, 3), atol=1e-6)
