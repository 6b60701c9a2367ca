"""
psb_batch_step / psb_batch_run_cycle: many independent simulations in one cooperative launch
(methods/umbrella_integration.py:128-203 runs its windows one replica after the other).  Every simulation must come
out BIT-IDENTICAL to stepping its own handle with psb_step: state, forces, histograms.
"""
import ctypes

import numpy as np
import pytest
import torch

import pysages_b200 as ps
from parity_utils import device_helpers, device_snapshot
from pysages_b200 import _native as N
from pysages_b200 import parallel, synthetic
from pysages_b200.backends import mock

pytestmark = pytest.mark.gpu


def build_windows(nwin, natoms, make_method, seed0, nframes=3, dtype=np.float32):
    wins = []
    for w in range(nwin):
        snap = synthetic.make_snapshot(natoms, layout="hoomd", dtype=dtype, seed=seed0 + w)
        traj = synthetic.make_trajectory(snap, frames=nframes, step=0.05)
        frames = [device_snapshot(snap, positions=traj[f]) for f in range(nframes)]
        # all frames of a window share ONE force array, like an engine's buffer
        frames = [fr._replace(forces=frames[0].forces) for fr in frames]
        method = make_method(w)
        _, init, update = method.build(frames[0], device_helpers("hoomd"))
        init()
        wins.append(dict(frames=frames, native=update.native, method=method, f0=frames[0].forces.clone()))
    return wins


def state_of(native, fields):
    torch.cuda.synchronize()
    out = {}
    for name in fields:
        v = native.view(name)
        if v is not None:
            out[name] = v.cpu().numpy().copy()
    return out


CASES = {
    # name: (natoms, make_method, state fields, expected CTAs per window > 1?)
    "harmonic-component-1000": (20_000, lambda w: ps.methods.HarmonicBias(
        [ps.colvars.Component(list(range(1000)), 0)], 50.0, [-3.0 + 0.4 * w]), ("xi", "cv_force", "energy", "ncalls"), True),
    "harmonic-distance-small": (500, lambda w: ps.methods.HarmonicBias(
        [ps.colvars.Distance([list(range(0, 20)), list(range(100, 140))])], 10.0 + w, [2.0]), ("xi", "cv_force", "ncalls"), False),
    "abf-2d": (20_000, lambda w: ps.methods.ABF(
        [ps.colvars.Distance([list(range(0, 600)), list(range(1000, 1700))]), ps.colvars.Component(list(range(3000, 3600)), 1)],
        ps.Grid[ps.Regular]((0.0, -6.0), (20.0, 6.0), (16, 16)), N=2), ("xi", "hist", "Fsum", "force", "Wp", "Wp_", "ncalls"), True),
    "metad-grid-dihedrals": (64, lambda w: ps.methods.Metadynamics(
        [ps.colvars.DihedralAngle([4, 6, 8, 14]), ps.colvars.DihedralAngle([6, 8, 14, 16])], 1.2, [0.35, 0.35], 2, 16,
        deltaT=5000.0, kB=0.0083144626, grid=ps.Grid[ps.Periodic]((-np.pi, -np.pi), (np.pi, np.pi), (20, 20))),
        ("xi", "heights", "centers", "grid_potential", "grid_gradient", "idx", "ncalls"), False),
    "metad-direct": (3000, lambda w: ps.methods.Metadynamics(
        [ps.colvars.Distance([list(range(0, 700)), list(range(1000, 1800))])], 1.0, [0.4], 2, 16),
        ("xi", "heights", "centers", "idx", "cv_force", "energy", "ncalls"), True),
    "rmsd-general-class": (2000, lambda w: ps.methods.HarmonicBias(
        [ps.colvars.RMSD(list(range(0, 800)), np.random.default_rng(3).normal(size=(800, 3)))], 5.0, [1.0]),
        ("xi", "cv_force", "ncalls"), True),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_batch_is_bitwise_equal_to_stepping_each_simulation_alone(case):
    natoms, make_method, fields, multi_cta = CASES[case]
    nwin, nsteps = 5, 5
    dtype = np.float64 if case.startswith(("metad-grid", "rmsd")) else np.float32
    alone = build_windows(nwin, natoms, make_method, seed0=70, dtype=dtype)
    batched = build_windows(nwin, natoms, make_method, seed0=70, dtype=dtype)
    if multi_cta:
        assert alone[0]["native"].launch_info()["grid_blocks"] > 1
    for t in range(nsteps):
        for w in alone:
            w["native"].step(w["frames"][t % 3], t)
    batch = parallel.WindowBatch([w["native"] for w in batched])
    for t in range(nsteps):
        batch.step([w["frames"][t % 3] for w in batched], t)
    assert batch.launches == nsteps  # ONE launch per step for all windows
    for a, b in zip(alone, batched):
        sa, sb = state_of(a["native"], fields), state_of(b["native"], fields)
        assert sorted(sa) == sorted(sb)
        for name in sa:
            np.testing.assert_array_equal(sa[name], sb[name], err_msg=f"{case}: {name}")
        fa, fb = a["frames"][0].forces.cpu().numpy(), b["frames"][0].forces.cpu().numpy()
        np.testing.assert_array_equal(fa, fb)
        assert np.any(fa != a["f0"].cpu().numpy()) or case.startswith("metad")  # the bias was applied
        assert int(sb["ncalls"][0]) == nsteps and b["native"].ncalls == nsteps


def test_batch_run_cycle_graph_replay_matches_plain_steps():
    natoms, make_method, fields, _ = CASES["harmonic-component-1000"]
    nwin, nsteps = 8, 12
    plain = build_windows(nwin, natoms, make_method, seed0=90)
    cyc = build_windows(nwin, natoms, make_method, seed0=90)
    pb = parallel.WindowBatch([w["native"] for w in plain])
    for t in range(nsteps):
        pb.step([w["frames"][t % 3] for w in plain], t)
    cb = parallel.WindowBatch([w["native"] for w in cyc])
    ring = cb.ring([[w["frames"][r] for w in cyc] for r in range(3)])
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        cb.run_cycle(ring, 3, nsteps)
    stream.synchronize()
    assert cb.launches == nsteps
    for a, b in zip(plain, cyc):
        sa, sb = state_of(a["native"], fields), state_of(b["native"], fields)
        for name in sa:
            np.testing.assert_array_equal(sa[name], sb[name], err_msg=name)
        np.testing.assert_array_equal(a["frames"][0].forces.cpu().numpy(), b["frames"][0].forces.cpu().numpy())


def test_batch_refuses_mixed_or_unsupported_members():
    w1 = build_windows(1, 500, CASES["harmonic-distance-small"][1], seed0=1)[0]
    w2 = build_windows(1, 64, CASES["metad-grid-dihedrals"][1], seed0=2, dtype=np.float64)[0]
    with pytest.raises(ValueError, match="one kernel instantiation"):
        parallel.WindowBatch([w1["native"], w2["native"]])
    with pytest.raises(ValueError, match="twice"):
        parallel.WindowBatch([w1["native"], w1["native"]])
    coord = build_windows(1, 2000, lambda w: ps.methods.HarmonicBias(
        [ps.colvars.CoordinationNumber([list(range(0, 100)), list(range(200, 300))], r0=1.0)], 1.0, [10.0]), seed0=3)[0]
    with pytest.raises(NotImplementedError, match="helper launches"):
        parallel.WindowBatch([coord["native"]])


def test_umbrella_integration_lockstep_equals_replica_by_replica_run():
    """pysages_b200.run(UmbrellaIntegration, ..., lockstep=True): all windows in one launch per step; states,
    histogram logs and analyze() identical to the reference-style sequential run."""
    snap = synthetic.make_snapshot(4000, layout="hoomd", dtype=np.float64, seed=9)
    traj = synthetic.make_trajectory(snap, frames=7, step=0.05)
    cv = [ps.colvars.Distance([list(range(0, 300)), list(range(1000, 1400))])]
    centers = [6.0, 6.5, 7.0, 7.5]

    def engine(**kw):
        return mock.MockEngine("hoomd", snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda", trajectory=traj)

    seq = ps.run(ps.methods.UmbrellaIntegration(cv, 50.0, centers, 1), engine, 7)
    lock = ps.run(ps.methods.UmbrellaIntegration(cv, 50.0, centers, 1), engine, 7, lockstep=True)
    torch.cuda.synchronize()
    for a, b in zip(seq.states, lock.states):
        np.testing.assert_array_equal(a.xi.cpu().numpy(), b.xi.cpu().numpy())
        assert int(a.ncalls) == int(b.ncalls) == 7
    for ca, cb in zip(seq.callbacks, lock.callbacks):
        np.testing.assert_array_equal(np.asarray(ca.data), np.asarray(cb.data))
    oa, ob = ps.analyze(seq), ps.analyze(lock)
    np.testing.assert_array_equal(oa["free_energy"], ob["free_energy"])
    for sa, sb in zip(seq.snapshots, lock.snapshots):
        np.testing.assert_array_equal(sa.forces.cpu().numpy(), sb.forces.cpu().numpy())
