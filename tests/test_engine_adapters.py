"""
The real engine adapters (backends/hoomd.py, openmm.py, lammps.py) driven through `pysages_b200.run`
on stand-in engine modules (tests/fake_engines.py): same states and forces as the (oracle-checked)
mock backend, on device arrays and -- OpenMM CPU platform, LAMMPS without Kokkos -- on host arrays.
"""
import numpy as np
import pytest
import torch

import fake_engines
import pysages_b200 as ps
from pysages_b200 import synthetic
from pysages_b200.backends import mock

pytestmark = pytest.mark.gpu

G1, G2 = list(range(0, 30)), list(range(100, 160))
NSTEPS = 6


def method():
    cvs = [ps.colvars.Distance([G1, G2]), ps.colvars.Component([5, 6, 7], 1)]
    grid = ps.Grid[ps.Regular]((0.0, -6.0), (12.0, 6.0), (16, 16))
    return ps.methods.ABF(cvs, grid, N=2)


def reference_run(snap, traj, units=None):
    """The same trajectory through the mock backend (device arrays)."""
    eng = mock.MockEngine(snap["layout"], snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda",
                          trajectory=traj, units=units)
    res = ps.run(method(), lambda **kw: eng, NSTEPS)
    torch.cuda.synchronize()
    return res.states[0], eng.t["forces"].cpu().numpy()


def reference_run2(snap, traj, nsteps, units=None):
    eng = mock.MockEngine(snap["layout"], snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda",
                          trajectory=traj, units=units)
    res = ps.run(method(), lambda **kw: eng, nsteps)
    torch.cuda.synchronize()
    return res.states[0], eng.t["forces"].cpu().numpy()


def check(result, forces, ref_state, ref_forces):
    s = result.states[0]
    torch.cuda.synchronize()
    for name in ("xi", "hist", "Fsum", "force", "Wp", "Wp_"):
        np.testing.assert_array_equal(getattr(s, name).cpu().numpy(), getattr(ref_state, name).cpu().numpy(), err_msg=name)
    assert int(s.ncalls) == NSTEPS
    np.testing.assert_array_equal(forces, ref_forces)


def test_hoomd_adapter():
    snap = synthetic.make_snapshot(400, layout="hoomd", dtype=np.float32, seed=5)
    traj = synthetic.make_trajectory(snap, frames=NSTEPS, step=0.05)
    ref_state, ref_forces = reference_run(snap, traj)
    Simulation = fake_engines.install_hoomd()
    eng = fake_engines.engine(snap, "cuda", traj)
    seen = []
    res = ps.run(method(), lambda **kw: Simulation(eng), NSTEPS, callback=lambda snapshot, state, t: seen.append(t))
    check(res, eng.t["forces"].cpu().numpy(), ref_state, ref_forces)
    assert seen == list(range(NSTEPS))
    from pysages_b200.backends import hoomd as adapter
    assert len(adapter.CONTEXTS_SAMPLERS) == 1
    adapter.detach(next(iter(adapter.CONTEXTS_SAMPLERS)))
    assert not adapter.CONTEXTS_SAMPLERS


def test_lammps_adapter_hands_over_atom_tags(monkeypatch):
    """Kokkos-CUDA LAMMPS, most atoms selected: the fix asks lammps-dlext for atom->tag (slot -> 1-based id) and the
    slot-ordered class streams with it instead of inverting tags_map; same bias as without the accessor."""
    monkeypatch.setenv("PSB_SLOT_MAJOR", "1")
    monkeypatch.setenv("PSB_GRID_BLOCKS", "24")
    n = 6000
    snap = synthetic.make_snapshot(n, layout="lammps", dtype=np.float64, seed=9)
    traj = synthetic.make_trajectory(snap, frames=NSTEPS, step=0.05)

    def bias():
        return ps.methods.HarmonicBias([ps.colvars.Distance([list(range(0, n // 2)), list(range(n // 2, n))])], 50.0, [1.0])

    results = []
    for with_tags in (False, True):
        lammps = fake_engines.install_lammps(kokkos_cuda=True, with_tags=with_tags)
        eng = fake_engines.engine(snap, "cuda", traj, units="real")
        res = ps.run(bias(), lambda **kw: lammps(eng), NSTEPS)
        torch.cuda.synchronize()
        assert (getattr(eng, "tags_calls", 0) >= NSTEPS) == with_tags
        results.append((res.states[0].xi.cpu().numpy().copy(), eng.t["forces"].cpu().numpy().copy()))
    np.testing.assert_allclose(results[1][0], results[0][0], rtol=1e-13)
    assert np.abs(results[1][1] - results[0][1]).max() <= 1e-12 * np.abs(results[0][1]).max()
    assert np.abs(results[0][1] - snap["arrays"]["forces"]).max() > 0


def test_hoomd_adapter_without_a_tags_accessor_inverts_rtags_itself(monkeypatch):
    """A hoomd-dlext build that does not export `tags`: the adapter passes the reference's five arrays only and the
    slot-ordered class inverts rtags in the kernel; same bias as the mock backend."""
    monkeypatch.setenv("PSB_SLOT_MAJOR", "1")
    monkeypatch.setenv("PSB_GRID_BLOCKS", "24")
    n = 6000
    snap = synthetic.make_snapshot(n, layout="hoomd", dtype=np.float32, seed=9)
    traj = synthetic.make_trajectory(snap, frames=NSTEPS, step=0.05)

    def bias():
        return ps.methods.HarmonicBias([ps.colvars.Distance([list(range(0, n // 2)), list(range(n // 2, n))])], 50.0, [1.0])

    eng0 = mock.MockEngine(snap["layout"], snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda", trajectory=traj)
    ref = ps.run(bias(), lambda **kw: eng0, NSTEPS)
    torch.cuda.synchronize()
    Simulation = fake_engines.install_hoomd(with_tags=False)
    eng = fake_engines.engine(snap, "cuda", traj)
    res = ps.run(bias(), lambda **kw: Simulation(eng), NSTEPS)
    torch.cuda.synchronize()
    from pysages_b200.backends import hoomd as adapter
    sampler = next(iter(adapter.CONTEXTS_SAMPLERS.values()))
    assert not sampler._want_tags and not hasattr(eng, "tags_calls")
    np.testing.assert_array_equal(res.states[0].xi.cpu().numpy(), ref.states[0].xi.cpu().numpy())
    np.testing.assert_array_equal(eng.t["forces"].cpu().numpy(), eng0.t["forces"].cpu().numpy())
    adapter.detach(next(iter(adapter.CONTEXTS_SAMPLERS)))


def test_hoomd_adapter_hands_over_the_slot_to_tag_array(monkeypatch):
    """Most of the system selected -> the slot-ordered class: the sampler fetches hoomd-dlext's `tags` (slot -> tag)
    each step and the kernel skips its tag -> slot inversion pass; same bias as with the reference's five arrays."""
    monkeypatch.setenv("PSB_SLOT_MAJOR", "1")
    monkeypatch.setenv("PSB_GRID_BLOCKS", "24")
    n = 6000
    snap = synthetic.make_snapshot(n, layout="hoomd", dtype=np.float32, seed=9)
    traj = synthetic.make_trajectory(snap, frames=NSTEPS, step=0.05)

    def bias():
        return ps.methods.HarmonicBias([ps.colvars.Distance([list(range(0, n // 2)), list(range(n // 2, n))])], 50.0, [1.0])

    eng0 = mock.MockEngine(snap["layout"], snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda", trajectory=traj)
    ref = ps.run(bias(), lambda **kw: eng0, NSTEPS)  # rtags only
    torch.cuda.synchronize()
    Simulation = fake_engines.install_hoomd()
    eng = fake_engines.engine(snap, "cuda", traj)
    res = ps.run(bias(), lambda **kw: Simulation(eng), NSTEPS)
    torch.cuda.synchronize()
    from pysages_b200.backends import hoomd as adapter
    sampler = next(iter(adapter.CONTEXTS_SAMPLERS.values()))
    assert sampler._want_tags and eng.tags_calls >= NSTEPS
    np.testing.assert_allclose(res.states[0].xi.cpu().numpy(), ref.states[0].xi.cpu().numpy(), rtol=1e-13)
    f, f0 = eng.t["forces"].cpu().numpy(), eng0.t["forces"].cpu().numpy()
    assert np.abs(f - f0).max() <= 2e-7 * np.abs(f0).max()
    assert np.abs(f0 - snap["arrays"]["forces"]).max() > 0
    adapter.detach(next(iter(adapter.CONTEXTS_SAMPLERS)))


@pytest.mark.parametrize("layout, device", [("openmm-cuda", "cuda"), ("openmm-cpu", "cpu")])
def test_openmm_adapter(layout, device):
    snap = synthetic.make_snapshot(400, layout=layout, dtype=np.float64, seed=6)
    traj = synthetic.make_trajectory(snap, frames=NSTEPS, step=0.05)
    ref_state, ref_forces = reference_run(snap, traj)
    Simulation = fake_engines.install_openmm()
    eng = fake_engines.engine(snap, device, traj)
    res = ps.run(method(), lambda **kw: Simulation(eng), NSTEPS)
    check(res, eng.t["forces"].cpu().numpy(), ref_state, ref_forces)


@pytest.mark.parametrize("kokkos_cuda", [True, False])
def test_lammps_adapter(kokkos_cuda):
    snap = synthetic.make_snapshot(400, layout="lammps", dtype=np.float64, seed=7)
    traj = synthetic.make_trajectory(snap, frames=NSTEPS, step=0.05)
    ref_state, ref_forces = reference_run(snap, traj, units="real")
    lammps = fake_engines.install_lammps(kokkos_cuda)
    eng = fake_engines.engine(snap, "cuda" if kokkos_cuda else "cpu", traj, units="real")
    ctx = lammps(eng)
    res = ps.run(method(), lambda **kw: ctx, NSTEPS)
    check(res, eng.t["forces"].cpu().numpy(), ref_state, ref_forces)


def test_openmm_cuda_integrator_on_its_own_stream_sees_the_bias():
    """OpenMM integrates on its context's stream: after the callback returns the integrator must find the biased
    forces even when the bias kernel was queued behind other work (openmm.py:160-170 blocks in sync_forces())."""
    snap = synthetic.make_snapshot(400, layout="openmm-cuda", dtype=np.float64, seed=6)
    traj = synthetic.make_trajectory(snap, frames=2, step=0.05)
    _, ref_forces = reference_run2(snap, traj, 2)
    Simulation = fake_engines.install_openmm(own_stream=True)
    sim = Simulation(fake_engines.engine(snap, "cuda", traj))
    ps.run(method(), lambda **kw: sim, 2)
    seen = sim.context.seen_forces
    torch.cuda.synchronize()
    np.testing.assert_array_equal(seen.cpu().numpy(), ref_forces)


def test_lammps_kokkos_integrator_on_its_own_stream_sees_the_bias():
    snap = synthetic.make_snapshot(400, layout="lammps", dtype=np.float64, seed=7)
    traj = synthetic.make_trajectory(snap, frames=2, step=0.05)
    _, ref_forces = reference_run2(snap, traj, 2, units="real")
    lammps = fake_engines.install_lammps(True, own_stream=True)
    ctx = lammps(fake_engines.engine(snap, "cuda", traj, units="real"))
    ps.run(method(), lambda **kw: ctx, 2)
    seen = ctx.seen_forces
    torch.cuda.synchronize()
    np.testing.assert_array_equal(seen.cpu().numpy(), ref_forces)


def test_lammps_bigbig_image_packing_is_rejected():
    lammps = fake_engines.install_lammps(True)
    import lammps as fake

    fake.dlext.kImgBitSize, fake.dlext.kImgBits = 64, 21
    snap = synthetic.make_snapshot(64, layout="lammps", dtype=np.float64, seed=8)
    with pytest.raises(RuntimeError, match="image flags"):
        ps.run(ps.methods.Unbiased([ps.colvars.Component([1], 0)]), lambda **kw: lammps(fake_engines.engine(snap, "cuda")), 1)


def test_sampling_context_enters_and_leaves_the_engine_context():
    """backends/core.py:59-73: `with sampling_context` trampolines to the engine's context (hoomd 2.x: the global
    hoomd.run steps whichever SimulationContext is entered)."""
    snap = synthetic.make_snapshot(64, layout="hoomd", dtype=np.float32, seed=8)
    events = []

    class Engine(mock.MockEngine):
        def __enter__(self):
            events.append("enter")
            return self

        def __exit__(self, *exc):
            events.append("exit")

        def run(self, n, **kw):
            events.append("run")
            super().run(n, **kw)

    Engine.__module__ = mock.__name__
    eng = Engine("hoomd", snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda")
    ps.run(ps.methods.Unbiased([ps.colvars.Component([1], 0)]), lambda **kw: eng, 2)
    assert events == ["enter", "run", "exit"]


def test_states_are_live_views_and_freeze_gives_values():
    """`update` returns views of live device memory (documented, methods.freeze): a kept state follows the run,
    a frozen one does not -- the reference's states are values (methods/core.py:108-116)."""
    snap = synthetic.make_snapshot(400, layout="hoomd", dtype=np.float32, seed=5)
    traj = synthetic.make_trajectory(snap, frames=4, step=0.05)
    kept, frozen = [], []

    def cb(snapshot, state, t):
        kept.append(state)
        frozen.append(ps.freeze(state))

    eng = mock.MockEngine(snap["layout"], snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda", trajectory=traj)
    ps.run(method(), lambda **kw: eng, 4, callback=cb)
    torch.cuda.synchronize()
    xs = [f.xi.cpu().numpy().copy() for f in frozen]
    assert all(np.any(xs[i] != xs[i + 1]) for i in range(3))          # values: one per step
    assert all(np.array_equal(k.xi.cpu().numpy(), xs[-1]) for k in kept)  # views: all show the last step
    assert [int(f.hist.sum()) for f in frozen] == [1, 2, 3, 4]


def test_openmm_variable_step_integrators_are_rejected():
    Simulation = fake_engines.install_openmm()
    import openmm

    snap = synthetic.make_snapshot(64, layout="openmm-cuda", dtype=np.float64, seed=8)
    sim = Simulation(fake_engines.engine(snap, "cuda"))
    sim.context.getIntegrator = lambda: openmm.VariableVerletIntegrator()
    with pytest.raises(ValueError, match="Variable step size"):
        ps.run(ps.methods.Unbiased([ps.colvars.Component([1], 0)]), lambda **kw: sim, 1)


def test_save_load_restart_matches_uninterrupted_run(tmp_path):
    """serialization.py:44-93 + methods/core.py:239-318: 4 steps, save, load, 4 more steps ==
    8 uninterrupted steps (metadynamics: hills, grids, counters all travel through the pickle)."""
    snap = synthetic.make_snapshot(300, layout="hoomd", dtype=np.float64, seed=9)
    traj = synthetic.make_trajectory(snap, frames=8, step=0.05)

    def make_method():
        cvs = [ps.colvars.Distance([G1, G2])]
        grid = ps.Grid[ps.Regular]((0.0,), (12.0,), (32,))
        return ps.methods.Metadynamics(cvs, 1.0, [0.4], 2, 16, deltaT=300.0, kB=0.0083, grid=grid)

    def generator(first=0, **kw):
        return mock.MockEngine("hoomd", snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda",
                               trajectory=traj[first:])

    full = ps.run(make_method(), generator, 8)
    part = ps.run(make_method(), generator, 4)
    ps.save(part, tmp_path / "ckpt.pkl")
    loaded = ps.load(tmp_path / "ckpt.pkl")
    assert isinstance(loaded.states[0].heights, np.ndarray) and int(loaded.states[0].ncalls) == 4
    resumed = ps.run(loaded, generator, 4, context_args=dict(first=4))
    torch.cuda.synchronize()
    a, b = resumed.states[0], full.states[0]
    for name in ("xi", "heights", "centers", "grid_potential", "grid_gradient"):
        np.testing.assert_array_equal(getattr(a, name).cpu().numpy(), getattr(b, name).cpu().numpy(), err_msg=name)
    assert int(a.idx) == int(b.idx) == 3 and int(a.ncalls) == int(b.ncalls) == 8


def test_funn_run_save_load_restart(tmp_path):
    """FUNN through `pysages_b200.run` (funn.py:97-210): 10 steps with train_freq = 2 (refits before calls
    5, 7, 9), checkpoint after 6, resume for 4 == the uninterrupted run: the network travels with the state."""
    from pysages_b200 import ml

    snap = synthetic.make_snapshot(300, layout="hoomd", dtype=np.float64, seed=9)
    traj = synthetic.make_trajectory(snap, frames=10, step=0.05)

    def make_method():
        cvs = [ps.colvars.Distance([G1, G2]), ps.colvars.Component([5, 6, 7], 1)]
        grid = ps.Grid[ps.Regular]((0.0, -6.0), (12.0, 6.0), (8, 8))
        opt = ml.LevenbergMarquardt(reg=ml.L2Regularization(1e-6), max_iters=3)
        return ps.methods.FUNN(cvs, grid, (6, 4), N=2, train_freq=2, optimizer=opt, seed=5)

    def generator(first=0, **kw):
        return mock.MockEngine("hoomd", snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda",
                               trajectory=traj[first:])

    full = ps.run(make_method(), generator, 10)
    part = ps.run(make_method(), generator, 6)
    ps.save(part, tmp_path / "funn.pkl")
    loaded = ps.load(tmp_path / "funn.pkl")
    assert isinstance(loaded.states[0].nn.params, np.ndarray) and int(loaded.states[0].ncalls) == 6
    resumed = ps.run(loaded, generator, 4, context_args=dict(first=6))
    torch.cuda.synchronize()
    a, b = resumed.states[0], full.states[0]
    for name in ("xi", "hist", "Fsum", "F", "Wp", "Wp_"):
        np.testing.assert_array_equal(getattr(a, name).cpu().numpy(), getattr(b, name).cpu().numpy(), err_msg=name)
    for name in ("params", "mean", "std"):
        np.testing.assert_array_equal(getattr(a.nn, name).cpu().numpy(), getattr(b.nn, name).cpu().numpy(), err_msg=name)
    assert int(a.ncalls) == int(b.ncalls) == 10
    assert not np.array_equal(b.nn.params.cpu().numpy(), make_method().model.parameters)  # it was refitted


def _mock(snap, traj, first=0):
    return mock.MockEngine("hoomd", snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda",
                           trajectory=traj[first:])


def test_analyze_metadynamics_metapotential():
    """metad.py:326-383: the deposited bias potential as a callable, against the oracle's sum_of_gaussians."""
    from oracle import methods as om

    snap = synthetic.make_snapshot(300, layout="hoomd", dtype=np.float64, seed=9)
    traj = synthetic.make_trajectory(snap, frames=9, step=0.05)
    cvs = [ps.colvars.Distance([G1, G2]), ps.colvars.DihedralAngle([1, 50, 120, 200])]
    res = ps.run(ps.methods.Metadynamics(cvs, 1.2, [0.4, 0.3], 2, 16), lambda **kw: _mock(snap, traj), 9)
    out = ps.analyze(res)
    st = ps.serialization.numpyfy(res.states[0])
    assert int(st.idx) == 4 and np.array_equal(out["heights"], st.heights)
    rng = np.random.default_rng(0)
    x = np.stack([rng.uniform(5, 9, 7), rng.uniform(-np.pi, np.pi, 7)], axis=1)
    want = [float(om.sum_of_gaussians(torch.as_tensor(xi).reshape(1, 2), torch.as_tensor(st.heights),
                                      torch.as_tensor(st.centers), torch.as_tensor(st.sigmas).reshape(1, 2),
                                      torch.tensor([np.inf, 2 * np.pi], dtype=torch.float64))) for xi in x]
    np.testing.assert_allclose(out["metapotential"](x), want, rtol=1e-12)
    assert np.all(out["metapotential"](st.centers[:4]) > 0)


def test_analyze_abf_and_umbrella_integration():
    snap = synthetic.make_snapshot(300, layout="hoomd", dtype=np.float64, seed=9)
    traj = synthetic.make_trajectory(snap, frames=6, step=0.05)
    res = ps.run(method(), lambda **kw: _mock(snap, traj), 6)  # ABF (abf.py analyze: binned estimates)
    out = ps.analyze(res, topology=(4,), pre_iters=3, iters=5)  # abf.py:261-306: gradient learning (short fit here)
    st = ps.serialization.numpyfy(res.states[0])
    assert out["histogram"].sum() == st.hist.sum() and out["mesh"].shape == (256, 2)
    mf = st.Fsum / np.maximum(st.hist[..., None].astype(np.float64), 1)
    np.testing.assert_array_equal(out["mean_force"], mf.transpose(1, 0, 2))  # transposed along the grid axes (grids.py:161-178)
    np.testing.assert_array_equal(out["histogram"], st.hist.T)
    assert out["free_energy"].shape == (16, 16) and np.isfinite(out["free_energy"]).all() and out["free_energy"].min() == 0.0
    assert out["fes_fn"]([[3.0, 0.0], [6.0, 1.0]]).shape == (2, 1)
    # umbrella_integration.py:220-261: mean force per window and its running integral
    cv = [ps.colvars.Distance([G1, G2])]
    centers = [6.0, 6.5, 7.0]
    ui = ps.methods.UmbrellaIntegration(cv, 50.0, centers, 1)
    res = ps.run(ui, lambda **kw: _mock(snap, traj), 6)
    out = ps.analyze(res)
    means = np.array([cb.get_means() for cb in res.callbacks]).reshape(3)
    mf = -50.0 * (means - np.array(centers))
    np.testing.assert_allclose(out["mean_forces"].reshape(3), mf, rtol=1e-12)
    np.testing.assert_allclose(out["free_energy"].reshape(3), [0.0, mf[0] * 0.5, mf[0] * 0.5 + mf[1] * 0.5], rtol=1e-12)


def test_spectral_abf_run_save_load_restart(tmp_path):
    """SpectralABF through `pysages_b200.run`: 10 steps (fit_threshold 3, fit_freq 2), checkpoint after 6, resume
    == the uninterrupted run; the fitted series travels with the state."""
    snap = synthetic.make_snapshot(300, layout="hoomd", dtype=np.float64, seed=9)
    traj = synthetic.make_trajectory(snap, frames=10, step=0.05)

    def make_method():
        cvs = [ps.colvars.Distance([G1, G2]), ps.colvars.Component([5, 6, 7], 1)]
        return ps.methods.SpectralABF(cvs, ps.Grid((0.0, -6.0), (12.0, 6.0), (8, 8)), N=2, fit_threshold=3, fit_freq=2)

    full = ps.run(make_method(), lambda **kw: _mock(snap, traj), 10)
    part = ps.run(make_method(), lambda **kw: _mock(snap, traj), 6)
    ps.save(part, tmp_path / "sabf.pkl")
    loaded = ps.load(tmp_path / "sabf.pkl")
    assert isinstance(loaded.states[0].fun.coefficients, np.ndarray)
    resumed = ps.run(loaded, lambda first=0, **kw: _mock(snap, traj, first), 4, context_args=dict(first=6))
    torch.cuda.synchronize()
    a, b = resumed.states[0], full.states[0]
    for name in ("xi", "hist", "Fsum", "force", "Wp", "Wp_"):
        np.testing.assert_array_equal(getattr(a, name).cpu().numpy(), getattr(b, name).cpu().numpy(), err_msg=name)
    np.testing.assert_array_equal(a.fun.coefficients.cpu().numpy(), b.fun.coefficients.cpu().numpy())
    assert float(a.fun.scale) == float(b.fun.scale) and int(a.ncalls) == int(b.ncalls) == 10
    assert np.any(b.fun.coefficients.cpu().numpy() != 0.0)


def test_ann_run_save_load_restart_and_analyze(tmp_path):
    """ANN through `pysages_b200.run`: 9 steps with train_freq = 2, checkpoint after 5, resume == uninterrupted;
    analyze() returns the free energy on the grid and a callable built from the saved network (ann.py:261-322)."""
    from pysages_b200 import ml

    snap = synthetic.make_snapshot(300, layout="hoomd", dtype=np.float64, seed=9)
    traj = synthetic.make_trajectory(snap, frames=9, step=0.05)

    def make_method():
        cvs = [ps.colvars.Distance([G1, G2]), ps.colvars.Component([5, 6, 7], 1)]
        opt = ml.LevenbergMarquardt(reg=ml.L2Regularization(1e-6), max_iters=3)
        return ps.methods.ANN(cvs, ps.Grid((0.0, -6.0), (12.0, 6.0), (8, 8)), (5,), 2.5, train_freq=2, optimizer=opt)

    full = ps.run(make_method(), lambda **kw: _mock(snap, traj), 9)
    part = ps.run(make_method(), lambda **kw: _mock(snap, traj), 5)
    ps.save(part, tmp_path / "ann.pkl")
    resumed = ps.run(ps.load(tmp_path / "ann.pkl"), lambda first=0, **kw: _mock(snap, traj, first), 4,
                     context_args=dict(first=5))
    torch.cuda.synchronize()
    a, b = resumed.states[0], full.states[0]
    for name in ("xi", "hist", "phi", "prob"):
        np.testing.assert_array_equal(getattr(a, name).cpu().numpy(), getattr(b, name).cpu().numpy(), err_msg=name)
    np.testing.assert_array_equal(a.nn.params.cpu().numpy(), b.nn.params.cpu().numpy())
    assert int(a.ncalls) == int(b.ncalls) == 9
    out = ps.analyze(full)
    assert out["free_energy"].shape == (8, 8) and out["free_energy"].min() == 0.0
    fes = out["fes_fn"](out["mesh"])
    assert fes.shape == (64, 1) and fes.min() == 0.0 and np.all(np.isfinite(fes))


def test_cff_run_save_load_restart(tmp_path):
    """CFF through `pysages_b200.run`: 9 steps with train_freq = 2, checkpoint after 5, resume == uninterrupted
    (both networks, prob, fe and histp travel with the state)."""
    from pysages_b200 import ml

    snap = synthetic.make_snapshot(300, layout="hoomd", dtype=np.float64, seed=9)
    traj = synthetic.make_trajectory(snap, frames=9, step=0.05)

    def make_method():
        cvs = [ps.colvars.Distance([G1, G2]), ps.colvars.Component([5, 6, 7], 1)]
        reg = ml.L2Regularization(1e-4)
        return ps.methods.CFF(cvs, ps.Grid((0.0, -6.0), (12.0, 6.0), (6, 6)), (4,), 1.0e4, N=2, train_freq=2,
                              optimizer=ml.LevenbergMarquardt(loss=ml.Sobolev1SSE(), reg=reg, max_iters=2),
                              foptimizer=ml.LevenbergMarquardt(reg=reg, max_iters=2))

    full = ps.run(make_method(), lambda **kw: _mock(snap, traj), 9)
    part = ps.run(make_method(), lambda **kw: _mock(snap, traj), 5)
    ps.save(part, tmp_path / "cff.pkl")
    resumed = ps.run(ps.load(tmp_path / "cff.pkl"), lambda first=0, **kw: _mock(snap, traj, first), 4,
                     context_args=dict(first=5))
    torch.cuda.synchronize()
    a, b = resumed.states[0], full.states[0]
    for name in ("xi", "hist", "histp", "prob", "fe", "Fsum", "force", "Wp", "Wp_"):
        np.testing.assert_array_equal(getattr(a, name).cpu().numpy(), getattr(b, name).cpu().numpy(), err_msg=name)
    for net in ("nn", "fnn"):
        np.testing.assert_array_equal(getattr(a, net).params.cpu().numpy(), getattr(b, net).params.cpu().numpy())
    assert int(a.ncalls) == int(b.ncalls) == 9
    out = ps.analyze(full)
    assert out["histogram"].sum() == 9 - 0 and out["mean_force"].shape == (6, 6, 2)
