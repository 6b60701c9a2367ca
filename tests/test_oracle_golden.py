"""
Pins the oracle against every known-answer value the reference's own tests hold
for the hot path: tests/test_cvs.py, tests/test_rmsd.py,
tests/test_grids.py of the reference.  CPU only.
"""
from typing import Any, NamedTuple

import numpy as np
import pytest
import torch

from oracle import colvars, grids


class FakeSnapshot(NamedTuple):  # reference tests/test_cvs.py:15-20
    positions: Any
    indices: Any


def test_distance_groups(golden):
    # reference tests/test_cvs.py:23-32
    pos = golden["distance_positions"]
    snap = FakeSnapshot(torch.as_tensor(pos), np.arange(len(pos)))
    cv = colvars.Distance([0, 1])
    assert len(cv.groups) == 0
    assert np.isclose(colvars.build(cv, differentiate=False)(snap).item(), golden["distance_expected"])
    cv = colvars.Distance([[0], [1]])
    assert len(cv.groups) == 2
    assert np.isclose(colvars.build(cv, differentiate=False)(snap).item(), golden["distance_expected"])


@pytest.mark.parametrize("f1, f2, k", [("ci2_1", "ci2_2", 0), ("ci2_1", "ci2_1t", 1), ("ci2_2", "ci2_12", 2)])
def test_rmsd(golden, f1, f2, k):
    # reference tests/test_rmsd.py:9-26
    base, reference = golden[f1], golden[f2]
    cv = colvars.RMSD(np.arange(len(base)), reference)
    value = cv.function(torch.as_tensor(base)).item()
    assert np.allclose(golden["rmsd_expected"][k], value)


def test_grid_indexing(golden):
    # reference tests/test_grids.py:56-92
    pi = np.pi
    grid = grids.Grid((-pi,), (pi,), (64,))
    cheb = grids.Grid((-pi,), (pi,), (64,), kind=grids.Chebyshev)
    per = grids.Grid((-pi,), (pi,), (64,), periodic=True)
    for g, tag in ((grid, "regular"), (cheb, "cheb"), (per, "periodic")):
        get_index = grids.build_indexer(g)
        for x, i in zip(golden[f"grid1d_x_{tag}"], golden[f"grid1d_i_{tag}"]):
            assert get_index(x) == (np.uint32(i),), (tag, x)
    g2 = grids.Grid((-pi, -1.0), (pi, 1.0), (64, 32))
    get_index = grids.build_indexer(g2)
    for x, i in zip(golden["grid2d_x"], golden["grid2d_i"]):
        assert get_index(x) == tuple(np.uint32(v) for v in i)


def test_float_floor_divide_matches_numpy_and_differs_from_naive_floor():
    # SURVEY.md appendix B trap 1: 1.0 // 0.1 == 9, floor(1.0 / 0.1) == 10
    assert grids.jax_floor_divide(1.0, 0.1) == 9.0
    rng = np.random.default_rng(0)
    a = rng.uniform(-50, 50, 200000)
    b = rng.uniform(0.01, 3.0, 200000)
    assert np.array_equal(grids.jax_floor_divide(a, b), a // b)
    # exact bin edges of cfg1's periodic grid (lower=-pi, 50 bins)
    h = 2 * np.pi / 50
    edges = -np.pi + h * np.arange(51)
    assert np.array_equal(grids.jax_floor_divide(edges + np.pi, h), (edges + np.pi) // h)


def test_jacobians_against_finite_differences():
    rng = np.random.default_rng(1)
    pos = rng.normal(size=(12, 3))
    ids = rng.permutation(12)
    snap = FakeSnapshot(torch.as_tensor(pos), ids)
    ref = rng.normal(size=(5, 3))
    cvs = [
        colvars.Component([1, 2, 3], 1),
        colvars.Distance([[0, 1, 2], [5, 6]]),
        colvars.Angle([3, 4, 5]),
        colvars.DihedralAngle([0, 3, 6, 9]),
        colvars.RadiusOfGyration([1, 3, 5, 7]),
        colvars.RMSD([2, 4, 6, 8, 10], ref),
        colvars.CoordinationNumber([[0, 1, 2], [7, 8, 9, 10]], r0=1.3),
        colvars.Displacement([[0, 1], [2, 3]]),
    ]
    xi, J = colvars.build(*cvs)(snap)
    assert xi.shape == (1, 10) and J.shape == (10, 36)
    f = colvars.build(*cvs, differentiate=False)
    eps = 1e-6
    Jfd = np.zeros((10, 36))
    for a in range(12):
        for c in range(3):
            p = pos.copy(); p[a, c] += eps
            m = pos.copy(); m[a, c] -= eps
            up = f(FakeSnapshot(torch.as_tensor(p), ids)).numpy()[0]
            dn = f(FakeSnapshot(torch.as_tensor(m), ids)).numpy()[0]
            Jfd[:, 3 * a + c] = (up - dn) / (2 * eps)
    assert np.allclose(J.numpy(), Jfd, rtol=1e-6, atol=1e-8)


def test_mesh_is_bin_centres():
    g = grids.Grid((0.0, -1.0), (1.0, 1.0), (4, 2))
    mesh = grids.compute_mesh(g)
    assert mesh.shape == (8, 2)
    assert np.allclose(mesh[0], [0.125, -0.5]) and np.allclose(mesh[1], [0.125, 0.5])
    assert np.allclose(mesh[-1], [0.875, 0.5])
