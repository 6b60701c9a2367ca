"""
psb_step_host with few selected atoms: only their rows travel (one pinned block up, the force rows back) -- against
the whole-array staging of the same handle type, on the three host-resident layouts (HOOMD-style host arrays,
OpenMM CPU platform = PLAIN3, LAMMPS without Kokkos).  Same kernels on the same rows: states and forces bitwise equal.
"""
import ctypes

import numpy as np
import pytest
import torch

import pysages_b200 as ps
from parity_utils import device_helpers, device_snapshot
from pysages_b200 import _native as N
from pysages_b200 import synthetic

pytestmark = pytest.mark.gpu

NATOMS = 5000
G1, G2, G3 = [7, 4200, 333], list(range(1000, 1040)), [4999, 0, 2500, 17]


def methods():
    d = [ps.colvars.Distance([G1, G2]), ps.colvars.Angle([G3[0], G3[1], G3[2]])]
    grid = ps.Grid[ps.Regular]((0.0, 0.0), (30.0, np.pi), (16, 16))
    yield "abf", lambda: ps.methods.ABF(d, grid, N=2)
    yield "metad", lambda: ps.methods.Metadynamics([ps.colvars.DihedralAngle(G3)], 1.2, [0.35], 2, 8, None)
    yield "harmonic-rog", lambda: ps.methods.HarmonicBias([ps.colvars.RadiusOfGyration(G2)], 2.0, [50.0])


@pytest.mark.parametrize("layout, dtype", [("hoomd", np.float32), ("hoomd", np.float64), ("openmm-cpu", np.float64), ("lammps", np.float64)])
@pytest.mark.parametrize("name", ["abf", "metad", "harmonic-rog"])
def test_selected_rows_equal_whole_arrays(layout, dtype, name, monkeypatch):
    make = dict(methods())[name]
    snap = synthetic.make_snapshot(NATOMS, layout=layout, dtype=dtype, seed=12)
    traj = synthetic.make_trajectory(snap, frames=5, step=0.05)
    outs = []
    f0 = np.array(snap["arrays"]["forces"], dtype=np.float64, copy=True)
    for compact in ("1", "0"):
        monkeypatch.setenv("PSB_HOST_COMPACT", compact)
        own = dict(snap, arrays={k: np.array(v, copy=True) for k, v in snap["arrays"].items()})  # torch views alias numpy
        host = device_snapshot(own, device="cpu")
        method = make()
        _, initialize, update = method.build(host, device_helpers(layout))
        native = update.native
        assert native.host_buffers
        state = initialize()
        for i in range(4):
            host.positions.copy_(torch.as_tensor(np.ascontiguousarray(traj[i])))
            state = update(host, state, i)
        # the fifth step through the C entry point itself, with its byte counters
        host.positions.copy_(torch.as_tensor(np.ascontiguousarray(traj[4])))
        h2d, d2h = ctypes.c_int64(), ctypes.c_int64()
        N.check(native.lib.psb_step_host(native.handle, ctypes.byref(native._desc(host)), 4, native._stream(),
                                         ctypes.byref(h2d), ctypes.byref(d2h)))
        moved = (h2d.value, d2h.value)
        torch.cuda.synchronize()
        xi = native.view("xi").cpu().numpy().copy()
        extra = {k: native.view(k).cpu().numpy().copy() for k in ("hist", "Fsum", "heights", "centers") if native.view(k) is not None}
        outs.append((xi, extra, host.forces.numpy().copy(), moved))
        del native, update, initialize, method
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    for k in outs[0][1]:
        np.testing.assert_array_equal(outs[0][1][k], outs[1][1][k], err_msg=k)
    np.testing.assert_array_equal(outs[0][2], outs[1][2])
    assert np.abs(outs[0][2].astype(np.float64) - f0).max() > 0  # the bias was applied
    # bytes per step: the rows of <= 47 atoms (256-byte aligned segments) against the arrays of 5000
    assert outs[0][3][0] <= 8192 and outs[0][3][1] <= 47 * 32 and outs[0][3][0] * 20 < outs[1][3][0]
