"""C-ABI error behaviour on the GPU box: status codes map to the reference's exception types and a
failed psb_create leaves nothing behind."""
import ctypes

import numpy as np
import pytest
import torch

import pysages_b200 as ps
from parity_utils import device_helpers, device_snapshot
from pysages_b200 import _native as N
from pysages_b200 import synthetic

pytestmark = pytest.mark.gpu


def build(method, snap, layout="hoomd"):
    return method.build(device_snapshot(snap), device_helpers(layout))


def test_create_errors_map_to_reference_exception_types():
    snap = synthetic.make_snapshot(100, layout="hoomd", dtype=np.float32, seed=1)
    with pytest.raises(ValueError, match="out of range"):
        build(ps.methods.HarmonicBias([ps.colvars.Distance([5, 100])], 1.0, [0.0]), snap)
    with pytest.raises(ValueError, match="between 1 and 8 collective variables"):
        build(ps.methods.Unbiased([ps.colvars.Component([i], 0) for i in range(9)]), snap)
    with pytest.raises(ValueError, match="at most 8 CV components"):
        build(ps.methods.Unbiased([ps.colvars.Displacement([2 * i, 2 * i + 1]) for i in range(3)]), snap)
    with pytest.raises(ValueError, match="sigma must be positive"):
        build(ps.methods.Metadynamics([ps.colvars.Distance([0, 1])], 1.0, [0.0], 5, 10), snap)
    with pytest.raises(ValueError, match="stride"):
        build(ps.methods.Metadynamics([ps.colvars.Distance([0, 1])], 1.0, [0.1], 0, 10), snap)
    with pytest.raises(NotImplementedError, match="CoordinationNumber supports"):
        build(ps.methods.HarmonicBias([ps.colvars.CoordinationNumber([[0, 1], [2, 3]], n=5, m=10)], 1.0, [0.0]), snap)
    with pytest.raises(ValueError, match="references"):  # wrong number of reference rows
        ps.colvars.RMSD(list(range(10)), np.zeros((9, 3)))


def test_step_errors():
    snap = synthetic.make_snapshot(100, layout="hoomd", dtype=np.float32, seed=1)
    _, init, update = build(ps.methods.HarmonicBias([ps.colvars.Distance([1, 2])], 1.0, [0.0]), snap)
    native = update.native
    d = N.SnapshotDesc()  # all pointers NULL
    assert native.lib.psb_step(native.handle, ctypes.byref(d), 0, None) == N.ERR_VALUE
    assert b"positions is NULL" in native.lib.psb_last_error()
    with pytest.raises(KeyError):
        native.set_state("hist", np.zeros(4))  # HarmonicBias has no histogram
    assert native.view("heights") is None
    n = ctypes.c_int64()
    assert native.lib.psb_state_info(native.handle, b"no-such-field", ctypes.byref(n), None, None) == N.ERR_KEY
    state = update(device_snapshot(snap), init())
    torch.cuda.synchronize()
    assert int(state.ncalls) == 1


def test_force_net_errors():
    snap = synthetic.make_snapshot(100, layout="hoomd", dtype=np.float32, seed=1)
    _, _, update = build(ps.methods.HarmonicBias([ps.colvars.Distance([1, 2])], 1.0, [0.0]), snap)
    with pytest.raises(ValueError, match="ABF-family"):
        update.native.set_force_net([1, 4, 1], np.zeros(13), [0.0], [1.0], predict_after=0)
    grid = ps.Grid((0.0,), (5.0,), (8,))
    _, _, update = build(ps.methods.ABF([ps.colvars.Distance([1, 2])], grid), snap)
    native = update.native
    with pytest.raises(ValueError, match="input and output width"):
        native.set_force_net([2, 4, 1], np.zeros(17), [0.0], [1.0], predict_after=0)
    with pytest.raises(ValueError, match="layer widths"):
        native.set_force_net([1, 129, 1], np.zeros(388), [0.0], [1.0], predict_after=0)
    with pytest.raises(ValueError, match="dense layers"):
        native.set_force_net([1] * 11, np.zeros(20), [0.0], [1.0], predict_after=0)
    with pytest.raises(ValueError, match="unknown activation"):
        native.set_force_net([1, 4, 1], np.zeros(13), [0.0], [1.0], predict_after=0, activation=7)
    native.set_force_net([1, 4, 1], np.zeros(13), [0.0], [1.0], predict_after=0)  # and a valid one is accepted
    with pytest.raises(ValueError, match="expected 13 network parameters"):
        ps.methods.FUNN([ps.colvars.Distance([1, 2])], grid, (4,), initial_params=np.zeros(12))


def test_ring_cv_errors():
    snap = synthetic.make_snapshot(100, layout="hoomd", dtype=np.float32, seed=1)
    with pytest.raises(ValueError, match="at least 3 atoms"):
        build(ps.methods.Unbiased([ps.colvars.RingAmplitude([1, 2])]), snap)
    d = N.CvDesc()
    keep = []
    good = ps.colvars.RingPhaseAngle([1, 2, 3, 4, 5]).describe(keep)
    good.axis = 7
    handle = ctypes.c_void_p()
    _, _, update = build(ps.methods.Unbiased([ps.colvars.RingPhaseAngle([1, 2, 3, 4, 5])]), snap)
    native = update.native
    mdesc = ps.methods.Unbiased([ps.colvars.RingPhaseAngle([1, 2, 3, 4, 5])]).describe(device_helpers("hoomd"))
    from pysages_b200.backends import snapshot as S
    ldesc = S.describe_layout(device_snapshot(snap), device_helpers("hoomd").layout, unwrap=True, need_momenta=False,
                              dense_outputs=False)
    cvs = (N.CvDesc * 1)(good)
    st = native.lib.psb_create(cvs, 1, ctypes.byref(mdesc), ctypes.byref(ldesc), ctypes.byref(handle))
    assert st == N.ERR_VALUE and b"amplitude (0) or phase (1)" in native.lib.psb_last_error()


def test_force_series_errors():
    snap = synthetic.make_snapshot(100, layout="hoomd", dtype=np.float32, seed=1)
    ex = np.array([[1], [2]], dtype=np.int32)
    _, _, update = build(ps.methods.HarmonicBias([ps.colvars.Distance([1, 2])], 1.0, [0.0]), snap)
    with pytest.raises(ValueError, match="ABF-family"):
        update.native.set_force_series(N.SERIES_MONOMIAL, ex, np.zeros(3), predict_after=0)
    grid = ps.Grid((0.0,), (5.0,), (8,))
    _, _, update = build(ps.methods.ABF([ps.colvars.Distance([1, 2])], grid), snap)
    native = update.native
    with pytest.raises(ValueError, match="unknown kind"):
        native.set_force_series(9, ex, np.zeros(3), predict_after=0)
    with pytest.raises(ValueError, match="out of range"):
        native.set_force_series(N.SERIES_MONOMIAL, np.array([[-1]], dtype=np.int32), np.zeros(2), predict_after=0)
    native.set_force_series(N.SERIES_MONOMIAL, ex, np.zeros(3), predict_after=0)
    native.set_force_series(None, None, None, predict_after=0)  # removed again
    grid3 = ps.Grid((0.0,) * 3, (5.0,) * 3, (4, 4, 4))
    cvs3 = [ps.colvars.Component([i], i) for i in range(3)]
    _, _, update = build(ps.methods.ABF(cvs3, grid3), snap)
    with pytest.raises(ValueError, match="Only 1D and 2D grids"):
        update.native.set_force_series(N.SERIES_MONOMIAL, np.ones((2, 3), dtype=np.int32), np.zeros(3), predict_after=0)
    with pytest.raises(ValueError, match="Only 1D and 2D grids"):  # approxfun/core.py:74-75 at construction
        ps.methods.SpectralABF(cvs3, grid3)


def test_ermsd_errors():
    snap = synthetic.make_snapshot(100, layout="hoomd", dtype=np.float32, seed=1)
    with pytest.raises(ValueError, match="3 \\* n_nucleotides"):
        ps.colvars.ERMSD(list(range(7)), np.zeros((7, 3)))
    with pytest.raises(ValueError, match="references must have shape"):
        ps.colvars.ERMSD(list(range(6)), np.zeros((5, 3)))
    with pytest.raises(ValueError, match="between 2 and 256 nucleotides"):
        build(ps.methods.Unbiased([ps.colvars.ERMSD([1, 2, 3], np.eye(3))]), snap)
