"""
Host-side API mirror, CPU only: the reference's own constructor / copy tests restated on
pysages_b200 (tests/test_grids.py:17-53, tests/test_snapshots.py:8-40) and the error conventions of
SURVEY.md 8(b) (same exception types as the reference).
"""
import math

import numpy as np
import pytest
import torch

import pysages_b200 as ps
from pysages_b200.backends.snapshot import Box, Snapshot, copy
from pysages_b200.grids import Chebyshev, Grid, Periodic, Regular, build_indexer, convert

pi = math.pi


def test_grid_constructor_like_reference():
    lower, upper, shape = (-pi,), (pi,), (64,)
    grid = Grid(lower, upper, shape)
    periodic_grid = Grid(lower, upper, shape, periodic=True)
    grid_tv = Grid[Regular](lower, upper, shape)
    cheb_grid_tv = Grid[Chebyshev](lower, upper, shape)
    periodic_grid_tv = Grid[Periodic](lower, upper, shape)
    assert grid == grid_tv
    assert periodic_grid == periodic_grid_tv
    assert grid != periodic_grid
    assert cheb_grid_tv == convert(grid, Grid[Chebyshev])
    assert not grid.is_periodic
    assert periodic_grid.is_periodic
    with pytest.raises(TypeError):
        Grid(lower, upper, shape, periodic=None)
    with pytest.raises(TypeError):
        Grid(lower, upper, shape, periodic=1)
    with pytest.raises(TypeError):
        Grid[1](lower, upper, shape)
    with pytest.raises(TypeError):
        Grid[bool](lower, upper, shape)
    with pytest.raises(ValueError):
        Grid(lower, upper, shape, invalid_keyword=True)
    with pytest.raises(ValueError):
        Grid[Periodic](lower, upper, shape, periodic=False)
    with pytest.raises(ValueError):
        Grid[Chebyshev](lower, upper, shape, periodic=True)


def test_snapshot_copy_like_reference():
    old = Snapshot(torch.ones((2, 3)), torch.zeros((2, 4)), torch.zeros((2, 3)), torch.zeros(2, dtype=torch.int32),
                   Box(np.eye(3), np.zeros(3)), 0.1)
    new = copy(old)
    assert torch.equal(old.positions, new.positions) and np.array_equal(old.box.H, new.box.H)
    assert old.positions.data_ptr() != new.positions.data_ptr()
    assert old.box.H is not new.box.H
    new_cpu = copy(old, to_cpu=True)
    assert type(new_cpu.positions) is np.ndarray


def test_cv_constructor_errors():
    with pytest.raises(RuntimeError, match="Invalid Cartesian axis"):
        ps.colvars.Component([0, 1], 3)  # colvars/core.py:83-84
    with pytest.raises(RuntimeError, match="Invalid Cartesian axis"):
        ps.colvars.PrincipalMoment([0, 1, 2], -1)
    with pytest.raises(ValueError, match="Exactly 2"):
        ps.colvars.Distance([0, 1, 2])  # colvars/core.py:171-176
    with pytest.raises(ValueError, match="Exactly 4"):
        ps.colvars.DihedralAngle([0, 1, 2])
    with pytest.raises(ValueError, match="Exactly 3"):
        ps.colvars.Angle([[0, 1], [2, 3]])
    with pytest.raises(RuntimeError, match="acylindrity axes"):
        ps.colvars.Acylindricity([0, 1, 2], axes="xx")
    with pytest.raises(ValueError):
        ps.colvars.Distance([0, "a"])
    assert ps.colvars.Distance([[0, 1, 2], [5, 6]]).points()[1].tolist() == [5, 6]
    # mixing bare indices and groups silently drops the bare ones (colvars/core.py:204-207, quirk kept)
    assert [p.tolist() for p in ps.colvars.Angle([0, [1, 2], [3, 4]]).points()] == [[1, 2], [3, 4]]


def test_method_constructor_errors():
    d = ps.colvars.Distance([0, 1])
    with pytest.raises(KeyError, match="kB"):  # metad.py:133-138
        ps.methods.Metadynamics([d], 1.0, [0.1], 10, 20, deltaT=300.0)
    with pytest.raises(ValueError, match="dimensions must match"):  # methods/core.py:435-438
        ps.methods.ABF([d], Grid((0.0, 0.0), (1.0, 1.0), (4, 4)))
    with pytest.raises(ValueError, match="dimensions must match"):
        ps.methods.Metadynamics([d, d], 1.0, [0.1, 0.1], 10, 20, grid=Grid((0.0,), (1.0,), (4,)))
    with pytest.raises(RuntimeError, match="kspring"):  # harmonic_bias.py:92-104
        ps.methods.HarmonicBias([d, d], [1.0, 2.0, 3.0], [0.0, 0.0])
    with pytest.raises(RuntimeError):
        ps.methods.HarmonicBias([d, d], np.ones((3, 3)), [0.0, 0.0])
    with pytest.raises(TypeError):
        ps.methods.HarmonicBias(["not a cv"], 1.0, [0.0])
    hb = ps.methods.HarmonicBias([d, d], 2.0, [0.0, 1.0])
    assert np.array_equal(hb.kspring, 2.0 * np.eye(2))  # scalar -> k x k matrix (harmonic_bias.py:106)
    with pytest.raises(RuntimeError, match="different CVs"):  # umbrella_integration.py:88-114
        ps.methods.UmbrellaIntegration([ps.methods.HarmonicBias([d], 1.0, [0.0]),
                                        ps.methods.HarmonicBias([ps.colvars.Distance([2, 3])], 1.0, [0.0])], 10)
    ui = ps.methods.UmbrellaIntegration([d], 5.0, [[0.1], [0.2], [0.3]], 10)
    assert len(ui.submethods) == 3 and len(ui.histograms) == 3


def test_run_rejects_unknown_backends_and_non_methods():
    class Foreign:
        pass

    with pytest.raises(ValueError, match="Invalid backend"):  # backends/core.py:42-44
        ps.run(ps.methods.Unbiased([ps.colvars.Component([0], 0)]), lambda **kw: Foreign(), 1)
    with pytest.raises(TypeError):
        ps.run(object(), lambda **kw: Foreign(), 1)
    assert set(ps.supported_backends()) >= {"hoomd", "openmm", "lammps"}


def test_restraints_canonicalize_like_reference():
    r = ps.CVRestraints((0.0, 1.0), (2.0, 3.0), (10.0, 0.0), (10.0, 5.0))
    canon = ps.methods.canonicalize(r, [ps.colvars.Distance([0, 1])] * 2)
    # k == 0 -> the bound is moved to +-inf (restraints.py:49-61)
    assert canon.lower[1] == -math.inf and canon.upper[1] == 3.0 and canon.lower[0] == 0.0


def test_indexer_needs_a_gpu_and_says_so():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        build_indexer(Grid((0.0,), (1.0,), (4,)))(np.array([[0.5]]))
