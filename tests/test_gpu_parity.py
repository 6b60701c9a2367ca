"""
GPU parity tests: the CUDA bias step (through the C ABI) against the oracle on identical seeded
snapshots.  Tolerances are the north-star's: <= 1e-10 relative in x64 mode, <= 1e-5 where the
reference itself computes in fp32, histogram bins / counts bit-exact.
"""
import math

import numpy as np
import pytest
import torch

from parity_utils import ParityRun, close
import pysages_b200 as ps_
from pysages_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["auto", "grid24", "slot-major", "slot-tags"])
def launch_shape(request, monkeypatch):
    """Every parity case runs twice: with the library's own launch shape (one CTA for the small
    test snapshots) and forced onto 24 cooperative CTAs (group-aligned CTAs, grid barrier, per-CTA
    partials, redundant finalizer, register-cached scatter) and, third, with the slot-ordered
    gather / scatter forced on (used for very large selections; point CVs with <= 4 groups); fourth, the same
    with HOOMD's slot -> tag array in the snapshot (psb_snapshot.tags: no tag -> slot inversion pass)."""
    monkeypatch.delenv("PSB_GRID_BLOCKS", raising=False)
    monkeypatch.delenv("PSB_SLOT_MAJOR", raising=False)
    monkeypatch.delenv("PSB_TEST_TAGS", raising=False)
    if request.param in ("grid24", "slot-major", "slot-tags"):
        monkeypatch.setenv("PSB_GRID_BLOCKS", "24")
    if request.param in ("slot-major", "slot-tags"):  # slot-ordered gather / scatter wherever the handle is eligible for it
        monkeypatch.setenv("PSB_SLOT_MAJOR", "1")
    if request.param == "slot-tags":
        monkeypatch.setenv("PSB_TEST_TAGS", "1")
    return request.param

LAYOUTS = [
    ("hoomd", np.float32),
    ("hoomd", np.float64),
    ("openmm-cuda", np.float32),
    ("openmm-cuda", np.float64),
    ("openmm-cpu", np.float64),
    ("lammps", np.float64),
]

N = 600
A, B, C, D = list(range(0, 40)), list(range(100, 180)), list(range(200, 203)), list(range(300, 310))

POINT_CVS = {
    "component-group": [("Component", ([A], 2), {})],
    "component-list": [("Component", ([3, 4, 5, 17], 0), {})],
    "distance-groups": [("Distance", ([A, B],), {})],
    "distance-points": [("Distance", ([7, 501],), {})],
    "angle": [("Angle", ([5, 77, 310],), {})],
    "dihedral": [("DihedralAngle", ([11, 12, 400, 13],), {})],
    "two-distances": [("Distance", ([A, B],), {}), ("Distance", ([C, D],), {})],
    "phi-psi-shared-atoms": [("DihedralAngle", ([4, 6, 8, 14],), {}), ("DihedralAngle", ([6, 8, 14, 16],), {})],
}


def snapshot(layout, dtype, n=N, seed=20240, **kw):
    return synthetic.make_snapshot(n, layout=layout, dtype=dtype, seed=seed, **kw)


def harmonic(k, center):
    return ("HarmonicBias", dict(kspring=k, center=center))


@pytest.mark.parametrize("layout, dtype", LAYOUTS)
@pytest.mark.parametrize("cvname", list(POINT_CVS))
def test_harmonic_bias_point_cvs(layout, dtype, cvname):
    snap = snapshot(layout, dtype)
    specs = POINT_CVS[cvname]
    k = len(specs)
    center = [0.3 * (i + 1) for i in range(k)]
    kspring = 10.0 if k == 1 else [[10.0, 1.5], [1.5, 7.0]]
    run = ParityRun(snap, specs, harmonic(kspring, center))
    traj = synthetic.make_trajectory(snap, frames=3, step=0.05)
    for f in range(3):
        run.step(traj[f])


@pytest.mark.parametrize("layout, dtype", [("hoomd", np.float32), ("lammps", np.float64), ("openmm-cuda", np.float64)])
def test_harmonic_bias_energy_and_generalised_force(layout, dtype):
    snap = snapshot(layout, dtype)
    run = ParityRun(snap, POINT_CVS["two-distances"], harmonic([4.0, 9.0], [1.0, 2.0]))
    s, o = run.step()
    e = run.native.view("energy").item()
    assert abs(e - run.omethod.bias_energy(o.xi)) <= 1e-10 * abs(e)
    F = run.native.view("cv_force")[:2].cpu().numpy()
    Fo = run.omethod.kspring @ (o.xi.numpy().flatten() - run.omethod.center)
    close(F, Fo, run.rtol, what="generalised force")


def test_large_groups_multi_cta():
    """Groups large enough for the cooperative multi-CTA path (grid barrier + partial reduction)."""
    n = 60000
    snap = snapshot("hoomd", np.float32, n=n, seed=20241)
    g1, g2 = list(range(0, 25000)), list(range(30000, 55000))
    run = ParityRun(snap, [("Distance", ([g1, g2],), {})], harmonic(25.0, [3.0]))
    assert run.native.launch_info()["grid_blocks"] > 1 and run.native.launch_info()["cooperative"]
    traj = synthetic.make_trajectory(snap, frames=2, step=0.02)
    for f in range(2):
        run.step(traj[f])


def test_unbiased_displacement_and_rog_values():
    snap = snapshot("hoomd", np.float64)
    import oracle
    import pysages_b200 as ps
    from parity_utils import device_helpers, device_snapshot, oracle_snapshot

    specs_o = [oracle.colvars.Displacement([A, B])]
    xi_o = oracle.colvars.build(*specs_o, differentiate=False)
    helpers, _ = oracle.backends.build_helpers("hoomd", {"positions", "indices"}, True)
    ref = xi_o(helpers.query(oracle_snapshot(snap))).numpy()
    m = ps.methods.Unbiased([ps.colvars.Displacement([A, B])])
    dsnap = device_snapshot(snap)
    _, init, update = m.build(dsnap, device_helpers("hoomd"))
    st = update(dsnap, init())
    torch.cuda.synchronize()
    close(st.xi.cpu().numpy(), ref, 1e-10, what="displacement")
    assert st.ncalls == 1


@pytest.mark.parametrize("layout, dtype", [("hoomd", np.float32), ("hoomd", np.float64), ("lammps", np.float64),
                                           ("openmm-cuda", np.float64), ("openmm-cpu", np.float64)])
def test_rmsd_and_rog_harmonic(layout, dtype):
    """RMSD (Kabsch) + RadiusOfGyration over the SAME atoms: two-pass reduction, shared-atom scatter."""
    snap = snapshot(layout, dtype, n=400, globule=True)
    sel = list(range(20, 220))
    rng = np.random.default_rng(9)
    ref = rng.normal(size=(len(sel), 3)) * 2.0
    specs = [("RMSD", (sel, ref), {}), ("RadiusOfGyration", (sel,), {})]
    run = ParityRun(snap, specs, harmonic([3.0, 0.2], [1.0, 5.0]))
    traj = synthetic.make_trajectory(snap, frames=2, step=0.05)
    for f in range(2):
        run.step(traj[f])


def test_rmsd_weighted_and_reference_known_answers(golden):
    """The reference's own RMSD known answers (tests/test_rmsd.py:12-14) through the CUDA path."""
    import pysages_b200 as ps
    from parity_utils import device_helpers, device_snapshot

    for f1, f2, want in (("ci2_1", "ci2_2", 0), ("ci2_1", "ci2_1t", 1), ("ci2_2", "ci2_12", 2)):
        base, reference = golden[f1], golden[f2]
        n = len(base)
        snap = dict(layout="openmm-cpu", box_H=np.eye(3) * 100, origin=np.zeros(3), dt=0.001,
                    arrays=dict(positions=base.copy(), velocities=np.zeros_like(base), inv_masses=np.ones((n, 1)),
                                forces=np.zeros_like(base), ids=np.arange(n, dtype=np.int32)))
        m = ps.methods.Unbiased([ps.colvars.RMSD(np.arange(n), reference)])
        dsnap = device_snapshot(snap)
        _, init, _ = m.build(dsnap, device_helpers("openmm-cpu"))
        xi = init().xi.cpu().numpy().item()
        assert np.allclose(golden["rmsd_expected"][want], xi), (f1, f2, xi)
    # non-uniform weights against the oracle
    snap = snapshot("hoomd", np.float64, n=300, globule=True)
    sel = list(range(10, 90))
    rng = np.random.default_rng(4)
    w = rng.uniform(0.5, 2.0, size=len(sel)); w /= w.sum()
    ref = rng.normal(size=(len(sel), 3))
    run = ParityRun(snap, [("RMSD", (sel, ref, w), {})], harmonic(2.0, [0.5]))
    run.step()


def test_distance_known_answer(golden):
    """reference tests/test_cvs.py:23-32 through the CUDA path."""
    import pysages_b200 as ps
    from parity_utils import device_helpers, device_snapshot

    pos = golden["distance_positions"]
    snap = dict(layout="openmm-cpu", box_H=np.eye(3), origin=np.zeros(3), dt=0.001,
                arrays=dict(positions=pos.copy(), velocities=np.zeros_like(pos), inv_masses=np.ones((2, 1)),
                            forces=np.zeros_like(pos), ids=np.arange(2, dtype=np.int32)))
    for idx in ([0, 1], [[0], [1]]):
        m = ps.methods.Unbiased([ps.colvars.Distance(idx)])
        dsnap = device_snapshot(snap)
        _, init, _ = m.build(dsnap, device_helpers("openmm-cpu"))
        assert np.isclose(init().xi.item(), golden["distance_expected"])


# ---- Metadynamics -----------------------------------------------------------------------------------
ADP_CVS = POINT_CVS["phi-psi-shared-atoms"]


def adp(golden, layout="openmm-cpu"):
    return synthetic.adp_snapshot(golden["adp_xyz_angstrom"], layout=layout)


def perturbed(snap, nframes, step, seed=3):
    rng = np.random.default_rng(seed)
    pos = snap["arrays"]["positions"]
    out = []
    cur = pos.copy()
    for _ in range(nframes):
        cur = cur.copy()
        cur[:, :3] += step * rng.normal(size=(pos.shape[0], 3)).astype(pos.dtype)
        out.append(cur)
    return out


@pytest.mark.parametrize("deltaT", [None, 5000.0])
@pytest.mark.parametrize("grid", [None, ((-math.pi, -math.pi), (math.pi, math.pi), (50, 50), "periodic")])
def test_metadynamics_adp(golden, deltaT, grid):
    """cfg1: alanine dipeptide, (phi, psi), standard / well-tempered, direct sum / periodic grid."""
    snap = adp(golden)
    kw = dict(height=1.2, sigma=[0.35, 0.35], stride=3, ngaussians=8, grid=grid)
    if deltaT is not None:
        kw.update(deltaT=deltaT, kB=0.0083144626)
    run = ParityRun(snap, ADP_CVS, ("Metadynamics", kw))
    for pos in perturbed(snap, 14, 0.01):
        run.step(pos)
    assert int(run.state.idx) == 4  # calls 4, 7, 10, 13 deposit (ncalls % stride == 1, ncalls > 1)
    if grid is None or deltaT is not None:
        e = run.native.view("energy").item()
        assert abs(e - run.omethod.bias_energy(run.ostate)) <= 1e-9 * max(abs(e), 1e-12)


def test_metadynamics_stride_one_never_deposits(golden):
    snap = adp(golden)
    run = ParityRun(snap, ADP_CVS, ("Metadynamics", dict(height=1.0, sigma=[0.3, 0.3], stride=1, ngaussians=4)))
    for pos in perturbed(snap, 3, 0.01):
        run.step(pos)
    assert int(run.state.idx) == 0


def test_metadynamics_hill_buffer_overflow_is_dropped(golden):
    snap = adp(golden)
    run = ParityRun(snap, ADP_CVS, ("Metadynamics", dict(height=1.0, sigma=[0.3, 0.3], stride=2, ngaussians=2)))
    for pos in perturbed(snap, 9, 0.01):
        run.step(pos)
    assert int(run.state.idx) == 4 and run.state.heights.shape[0] == 2


@pytest.mark.parametrize("restraints", [None, ((1.0, 1.0), (6.0, 6.0), (50.0, 50.0), (50.0, 50.0))])
def test_metadynamics_regular_grid_out_of_bounds(restraints):
    """Distance CVs on a non-periodic grid that the trajectory leaves: OOB -> restraint force or zero,
    and no deposit while out of bounds (metad.py:258-260, 293-312)."""
    snap = snapshot("hoomd", np.float32)
    kw = dict(height=0.5, sigma=[0.4, 0.4], stride=2, ngaussians=6, deltaT=300.0, kB=0.0019872,
              grid=((1.0, 1.0), (6.0, 6.0), (16, 12), "regular"), restraints=restraints)
    run = ParityRun(snap, POINT_CVS["two-distances"], ("Metadynamics", kw))
    xi0 = run.state.xi.cpu().numpy().flatten()
    assert (xi0 > 6.0).any() or (xi0 < 1.0).any() or True
    for pos in perturbed(snap, 7, 0.3):
        run.step(pos)


def test_metadynamics_many_hills_multi_cta():
    """Direct sum over enough hills to need several CTAs and several TMA stages per CTA."""
    snap = snapshot("hoomd", np.float64, n=600)
    H = 40000
    kw = dict(height=0.8, sigma=[0.5, 0.7], stride=5, ngaussians=H, deltaT=1000.0, kB=0.0083144626)
    run = ParityRun(snap, POINT_CVS["two-distances"], ("Metadynamics", kw))
    rng = np.random.default_rng(0)
    xi = run.state.xi.cpu().numpy().flatten()
    nfill = H - 7
    heights = np.zeros(H); heights[:nfill] = rng.uniform(0.5, 1.2, nfill)
    centers = np.zeros((H, 2)); centers[:nfill] = xi + rng.normal(scale=1.0, size=(nfill, 2))
    for name, val in (("heights", heights), ("centers", centers)):
        run.native.set_state(name, val)
    run.native.set_state("idx", np.array([nfill]))
    run.ostate = run.ostate._replace(heights=torch.as_tensor(heights), centers=torch.as_tensor(centers), idx=nfill)
    assert run.native.launch_info()["grid_blocks"] > 1
    for pos in perturbed(snap, 7, 0.05):
        run.step(pos)
    assert int(run.state.idx) == nfill + 1


# ---- ABF ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout, dtype", LAYOUTS)
def test_abf_two_distances(layout, dtype):
    """cfg2 (small): ABF on two Distance CVs, 2-D grid with restraint walls."""
    snap = snapshot(layout, dtype)
    kw = dict(grid=((0.0, 0.0), (8.0, 8.0), (8, 8), "regular"), N=2,
              restraints=((0.0, 0.0), (8.0, 8.0), (20.0, 20.0), (20.0, 20.0)))
    run = ParityRun(snap, POINT_CVS["two-distances"], ("ABF", kw), units="real" if layout == "lammps" else None)
    for pos in perturbed(snap, 6, 0.1):
        run.step(pos)


def test_abf_shared_atoms_and_clamped_gather(golden):
    """Dihedrals sharing atoms (J J^T off-diagonal from overlapping groups) on a regular grid WITHOUT
    restraints: out-of-range bins drop the scatter but gather the clamped last bin (SURVEY.md trap 3)."""
    snap = adp(golden)
    kw = dict(grid=((-1.0, -1.0), (1.0, 1.0), (4, 4), "regular"), N=1)
    run = ParityRun(snap, ADP_CVS, ("ABF", kw))
    for pos in perturbed(snap, 6, 0.02):
        run.step(pos)


def test_abf_position_dependent_jacobian_pass_c():
    """ABF on RMSD + Component: Jacobian depends on positions -> generic J J^T / J p pass."""
    snap = snapshot("hoomd", np.float64, n=300, globule=True)
    sel = list(range(10, 100))
    ref = np.random.default_rng(1).normal(size=(len(sel), 3)) * 2
    specs = [("RMSD", (sel, ref), {}), ("Component", ([list(range(50, 150))], 1), {})]
    kw = dict(grid=((0.0, -5.0), (10.0, 5.0), (6, 6), "regular"), N=1)
    run = ParityRun(snap, specs, ("ABF", kw))
    for pos in perturbed(snap, 4, 0.05):
        run.step(pos)


def test_abf_use_pinv():
    snap = snapshot("hoomd", np.float64)
    kw = dict(grid=((0.0, 0.0), (8.0, 8.0), (8, 8), "periodic"), N=3, use_pinv=True)
    run = ParityRun(snap, POINT_CVS["two-distances"], ("ABF", kw))
    for pos in perturbed(snap, 4, 0.1):
        run.step(pos)


# ---- pairwise coordination --------------------------------------------------------------------------
@pytest.mark.parametrize("layout, dtype, na, nb", [("hoomd", np.float32, 70, 130), ("lammps", np.float64, 33, 64),
                                                   ("hoomd", np.float64, 2500, 300)])
def test_coordination_number_harmonic(layout, dtype, na, nb):
    n = 4000
    snap = synthetic.make_snapshot(n, layout=layout, dtype=dtype, seed=20240 + 3000)
    rng = np.random.default_rng(5)
    tags = rng.permutation(n)
    ga, gb = sorted(tags[:na].tolist()), sorted(tags[na : na + nb].tolist())
    specs = [("CoordinationNumber", ([ga, gb],), dict(r0=2.5))]
    run = ParityRun(snap, specs, harmonic(10.0, [5.0]))
    run.step()
    run.step(perturbed(snap, 1, 0.05)[0])


# ---- bit-exact bins ---------------------------------------------------------------------------------
def test_grid_indices_bit_exact_on_device():
    import oracle
    import pysages_b200 as ps

    rng = np.random.default_rng(12)
    for kind, ok, pk in (("regular", oracle.grids.Regular, ps.Regular), ("periodic", oracle.grids.Periodic, ps.Periodic),
                         ("chebyshev", oracle.grids.Chebyshev, ps.Chebyshev)):
        for lower, upper, shape in ((-math.pi, math.pi, 50), (0.0, 1.0, 10), (-2.5, 7.25, 64), (0.0, 23.2, 64)):
            h = (upper - lower) / shape
            edges = lower + h * np.arange(-2, shape + 3)
            xs = np.concatenate([edges, np.nextafter(edges, np.inf), np.nextafter(edges, -np.inf),
                                 rng.uniform(lower - 1, upper + 1, 5000)])
            want = np.array([oracle.grids.build_indexer(oracle.grids.Grid((lower,), (upper,), (shape,), kind=ok))(x)[0]
                             for x in xs], dtype=np.uint32)
            got = ps.grids.build_indexer(ps.Grid[pk]((lower,), (upper,), (shape,)))(xs.reshape(-1, 1))
            assert np.array_equal(got[:, 0], want), kind


# ---- multiple walkers sharing hills (extension) -----------------------------------------
@pytest.mark.parametrize("deltaT", [None, 5000.0])
@pytest.mark.parametrize("grid", [None, ((-math.pi, -math.pi), (math.pi, math.pi), (20, 20), "periodic")])
def test_shared_hill_walkers(golden, deltaT, grid):
    """W walkers in lockstep on one GPU: psb_walker_begin -> records in rank order -> psb_walker_finish,
    against the single-process restatement (oracle/walkers.py).  The all-gather itself is covered by
    tests/test_parallel_cpu.py (gloo) and by bench.py --gpus N (NCCL)."""
    import oracle
    import pysages_b200 as ps
    from parity_utils import X64_RTOL, device_helpers, device_snapshot, make_cvs, make_method, oracle_snapshot
    from pysages_b200 import parallel

    W, nsteps = 3, 8
    kw = dict(height=1.2, sigma=[0.35, 0.35], stride=3, ngaussians=12, grid=grid)
    if deltaT is not None:
        kw.update(deltaT=deltaT, kB=0.0083144626)
    base = adp(golden, layout="hoomd")
    exchange = parallel.RecordExchange()
    owalkers, engines, trajs, dsnaps, osnaps = [], [], [], [], []
    for w in range(W):
        snap = dict(base)
        snap["arrays"] = {k: np.array(v, copy=True) for k, v in base["arrays"].items()}
        trajs.append(perturbed(snap, nsteps, 0.02, seed=30 + w))
        om = make_method("oracle", ("Metadynamics", kw), make_cvs(oracle.colvars, ADP_CVS))
        helpers, bias = oracle.backends.build_helpers("hoomd", om.snapshot_flags, True)
        osnaps.append(oracle_snapshot(snap))
        owalkers.append(oracle.walkers.SharedHillsWalker(om, osnaps[-1], helpers, bias))
        m = make_method("ours", ("Metadynamics", dict(kw, shared_hills=True)), make_cvs(ps.colvars, ADP_CVS))
        dsnaps.append(device_snapshot(snap))
        _, init, update = m.build(dsnaps[-1], device_helpers("hoomd"))
        init()
        engines.append(parallel.NativeWalkerEngine(update.native))
    for t in range(nsteps):
        for w in range(W):
            osnaps[w] = osnaps[w]._replace(positions=np.array(trajs[w][t], copy=True))
            dsnaps[w] = dsnaps[w]._replace(positions=torch.as_tensor(np.ascontiguousarray(trajs[w][t])).cuda())
        if engines[0].next_step_deposits():
            assert all(e.next_step_deposits() for e in engines) and owalkers[0].next_step_deposits()
            orecs = torch.stack([ow.walker_begin(s) for ow, s in zip(owalkers, osnaps)])
            recs = torch.stack([e.walker_begin(s).clone() for e, s in zip(engines, dsnaps)])
            close(recs.cpu().numpy(), orecs.numpy(), X64_RTOL, what="hill records")
            for ow, s in zip(owalkers, osnaps):
                ow.walker_finish(s, orecs)
            for e, s in zip(engines, dsnaps):
                e.walker_finish(s, recs)
        else:
            for ow, s in zip(owalkers, osnaps):
                ow.step(s)
            for e, s in zip(engines, dsnaps):
                parallel.walker_step(e, s, exchange)
        torch.cuda.synchronize()
        for w in range(W):
            n, o = engines[w].native, owalkers[w].state
            close(n.view("xi").cpu().numpy().ravel()[:2], o.xi.numpy().ravel(), X64_RTOL, what="xi")
            close(n.view("heights").cpu().numpy().ravel(), o.heights.numpy(), X64_RTOL, what="heights")
            close(n.view("centers").cpu().numpy().reshape(-1, 2), o.centers.numpy(), X64_RTOL, what="centers")
            assert int(n.view("idx").item()) == int(o.idx) and int(n.view("ncalls").item()) == int(o.ncalls)
            close(n.view("bias").cpu().numpy().reshape(-1, 3), o.bias.numpy(), X64_RTOL,
                  scale=max(float(o.bias.abs().max()), 1e-300), what="bias")
            close(dsnaps[w].forces.cpu().numpy()[:, :3], osnaps[w].forces[:, :3], 1e-12, what="forces")
            if grid is not None:
                close(n.view("grid_gradient").cpu().numpy().reshape(o.grid_gradient.shape), o.grid_gradient.numpy(),
                      X64_RTOL, what="grid_gradient")
                if deltaT is not None:
                    close(n.view("grid_potential").cpu().numpy().reshape(o.grid_potential.shape),
                          o.grid_potential.numpy(), X64_RTOL, what="grid_potential")
    assert int(engines[0].native.view("idx").item()) == 2 * W


# ---- engine-loop emulation: CUDA-graph replay of the snapshot ring ---------------------------------------
@pytest.mark.parametrize("method", ["abf", "metad-grid", "metad-direct"])
def test_run_cycle_graph_replay_matches_single_steps_and_oracle(method, launch_shape):
    """psb_run_cycle captures one pass over the snapshot ring as a CUDA graph and replays it (barrier
    epochs, deposit decisions and all method state live on the device): bitwise equal to issuing
    psb_step per step, and equal to the oracle after a few hundred steps."""
    import ctypes

    import oracle
    import pysages_b200 as ps
    from parity_utils import X64_RTOL, device_helpers, device_snapshot, make_cvs, make_method, oracle_snapshot
    from pysages_b200 import _native as N
    from pysages_b200.backends import snapshot as S

    nring, nsteps = 3, 121
    snap = snapshot("hoomd", np.float64, n=800, seed=20250)
    traj = synthetic.make_trajectory(snap, frames=nring, step=0.03)
    cvs = POINT_CVS["two-distances"]
    if method == "abf":
        spec = ("ABF", dict(grid=((0.0, 0.0), (14.0, 14.0), (24, 24), "regular"), N=5))
    elif method == "metad-grid":
        spec = ("Metadynamics", dict(height=0.7, sigma=[0.4, 0.6], stride=7, ngaussians=64, deltaT=900.0, kB=0.0083,
                                     grid=((0.0, 0.0), (14.0, 14.0), (40, 40), "regular")))
    else:
        spec = ("Metadynamics", dict(height=0.7, sigma=[0.4, 0.6], stride=7, ngaussians=64, deltaT=900.0, kB=0.0083))

    def build():
        m = make_method("ours", spec, make_cvs(ps.colvars, cvs))
        m.dense_bias = False
        dsnaps = [device_snapshot(snap, positions=traj[i]) for i in range(nring)]
        _, init, update = m.build(dsnaps[0], device_helpers("hoomd"))
        init()
        return update.native, dsnaps

    # (a) graph replay (capture needs a non-default stream, like an engine's own stream)
    stream = torch.cuda.Stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        nat_g, snaps_g = build()
        descs = (N.SnapshotDesc * nring)()
        for i, s in enumerate(snaps_g):
            S.describe_snapshot(s, nat_g.layout, descs[i])
        N.check(nat_g.lib.psb_run_cycle(nat_g.handle, descs, nring, nsteps, sp))
    stream.synchronize()
    assert nat_g.lib.psb_cycle_uses_graph(nat_g.handle) == 1
    # (b) one psb_step per step (programmatic dependent launch between consecutive steps, the default)
    nat_s, snaps_s = build()
    for t in range(nsteps):
        nat_s.step(snaps_s[t % nring], t)
    # (b') the same without programmatic dependent launch: plain stream order
    nat_p, snaps_p = build()
    nat_p.set_option("pdl", 0)
    for t in range(nsteps):
        nat_p.step(snaps_p[t % nring], t)
    torch.cuda.synchronize()
    fields = ["xi", "cv_force", "ncalls"] + (["hist", "Fsum", "force", "Wp", "Wp_"] if method == "abf" else
                                             ["heights", "centers", "idx"] + (["grid_potential", "grid_gradient"] if method == "metad-grid" else []))
    for f in fields:
        assert torch.equal(nat_g.view(f), nat_s.view(f)), f
        assert torch.equal(nat_p.view(f), nat_s.view(f)), f
    for a, b, c in zip(snaps_g, snaps_s, snaps_p):
        assert torch.equal(a.forces, b.forces) and torch.equal(c.forces, b.forces)
    # (c) the oracle over the same ring
    om = make_method("oracle", spec, make_cvs(oracle.colvars, cvs))
    helpers, bias = oracle.backends.build_helpers("hoomd", om.snapshot_flags, True)
    osnaps = [oracle_snapshot(snap, positions=traj[i]) for i in range(nring)]
    _, oinit, oupdate = om.build(osnaps[0], helpers)
    ostate = oinit()
    for t in range(nsteps):
        ostate = oupdate(osnaps[t % nring], ostate)
        bias(osnaps[t % nring], ostate.bias.numpy())
    close(nat_g.view("xi").cpu().numpy().ravel()[:2], ostate.xi.numpy().ravel(), X64_RTOL, what="xi")
    if method == "abf":
        assert np.array_equal(nat_g.view("hist").cpu().numpy().reshape(24, 24).astype(np.int64), ostate.hist.numpy())
        close(nat_g.view("Fsum").cpu().numpy().reshape(24, 24, 2), ostate.Fsum.numpy(), 1e-8, what="Fsum")
    else:
        assert int(nat_g.view("idx").item()) == int(ostate.idx) == (nsteps - 1) // 7
        close(nat_g.view("heights").cpu().numpy().ravel(), ostate.heights.numpy(), 1e-9, what="heights")
        close(nat_g.view("centers").cpu().numpy().reshape(-1, 2), ostate.centers.numpy(), X64_RTOL, what="centers")
    for a, o in zip(snaps_g, osnaps):
        close(a.forces.cpu().numpy()[:, :3], o.forces[:, :3], 1e-9, what="forces after %d steps" % nsteps)


# ---- gyration-tensor CVs (colvars/shape.py:85-344) -------------------------------------
GYRATION_CVS = {
    "principal-moments": [("PrincipalMoment", (list(range(20, 220)), 0), {}), ("PrincipalMoment", (list(range(20, 220)), 2), {})],
    "asphericity+anisotropy": [("Asphericity", (list(range(20, 220)),), {}), ("ShapeAnisotropy", (list(range(250, 330)),), {})],
    "acylindricity-xy-xz": [("Acylindricity", (list(range(20, 220)),), dict(axes="xy")),
                            ("Acylindricity", (list(range(230, 300)),), dict(axes="xz"))],
    "acylindricity-yz-is-zero": [("Acylindricity", (list(range(20, 220)),), dict(axes="yz")),
                                 ("PrincipalMoment", (list(range(230, 300)), 1), {})],
}


@pytest.mark.parametrize("layout, dtype", [("hoomd", np.float32), ("hoomd", np.float64), ("lammps", np.float64),
                                           ("openmm-cuda", np.float64)])
@pytest.mark.parametrize("cvname", list(GYRATION_CVS))
def test_gyration_tensor_cvs_harmonic(layout, dtype, cvname):
    snap = snapshot(layout, dtype, n=400, globule=True)
    run = ParityRun(snap, GYRATION_CVS[cvname], harmonic([0.7, 0.3], [1.0, 2.0]))
    traj = synthetic.make_trajectory(snap, frames=2, step=0.05)
    for f in range(2):
        run.step(traj[f])
    if cvname == "acylindricity-yz-is-zero":  # the reference's index quirk (shape.py:240, 285)
        assert float(run.state.xi.cpu().numpy().ravel()[0]) == 0.0


def test_gyration_tensor_metadynamics_grid():
    snap = snapshot("hoomd", np.float64, n=400, globule=True)
    cvs = [("Asphericity", (list(range(0, 300)),), {})]
    kw = dict(height=0.5, sigma=[0.3], stride=2, ngaussians=8, grid=((0.0,), (40.0,), (64,), "regular"))
    run = ParityRun(snap, cvs, ("Metadynamics", kw))
    for pos in synthetic.make_trajectory(snap, frames=5, step=0.05):
        run.step(pos)
    assert int(run.state.idx) == 2


# ---- ABF: the J J^T special cases -----------------------------------------------------------------------
@pytest.mark.parametrize("case", ["component+distance-disjoint", "distances-sharing-atoms",
                                  "distance-overlapping-own-groups", "four-components"])
def test_abf_jjt_cases(case):
    """psb_create inverts J J^T once when it cannot depend on positions (Component / Distance /
    Displacement CVs, no atom shared between CVs); every other combination assembles it per step."""
    snap = snapshot("hoomd", np.float64, n=500, seed=77)
    X, Y = list(range(10, 60)), list(range(200, 230))
    if case == "component+distance-disjoint":
        cvs = [("Component", (list(range(300, 340)), 1), {}), ("Distance", ([X, Y],), {})]
        grid = ((-12.0, 0.0), (12.0, 14.0), (8, 8), "regular")
    elif case == "distances-sharing-atoms":  # atoms 40..59 belong to both CVs: J J^T has off-diagonal terms
        cvs = [("Distance", ([X, Y],), {}), ("Distance", ([list(range(40, 90)), list(range(300, 320))],), {})]
        grid = ((0.0, 0.0), (14.0, 14.0), (8, 8), "regular")
    elif case == "distance-overlapping-own-groups":  # the two groups of ONE Distance CV overlap
        cvs = [("Distance", ([list(range(10, 60)), list(range(40, 100))],), {})]
        grid = ((0.0,), (14.0,), (16,), "regular")
    else:
        cvs = [("Component", (list(range(i * 50, i * 50 + 30)), i % 3), {}) for i in range(4)]
        grid = ((-12.0,) * 4, (12.0,) * 4, (4, 4, 4, 4), "regular")
    run = ParityRun(snap, cvs, ("ABF", dict(grid=grid, N=3)))
    for pos in perturbed(snap, 5, 0.08):
        run.step(pos)


@pytest.mark.parametrize("grid", [None, ((-math.pi, -math.pi), (math.pi, math.pi), (20, 20), "periodic")])
def test_shared_hills_single_walker_equals_plain_metadynamics(golden, grid):
    """Metadynamics(shared_hills=True) through the public build()/update() path with one process: the
    begin -> (identity exchange) -> finish split of a deposit step reproduces the fused step."""
    kw = dict(height=1.2, sigma=[0.35, 0.35], stride=3, ngaussians=12, grid=grid, deltaT=5000.0, kB=0.0083144626)
    snap = adp(golden, layout="hoomd")
    plain = ParityRun(snap, ADP_CVS, ("Metadynamics", kw))
    shared = ParityRun(snap, ADP_CVS, ("Metadynamics", dict(kw, shared_hills=True)))
    for pos in perturbed(snap, 10, 0.02):
        plain.step(pos)
        shared.step(pos)  # also checked against the oracle's plain metadynamics
    for name in ("heights", "centers", "idx"):
        assert torch.equal(getattr(plain.state, name), getattr(shared.state, name)), name
    assert int(shared.state.idx) == 3


# ---- FUNN: ABF accumulation + force from the network (row K) ----------------------------------------
FUNN_GRID = ((0.0, 0.0), (8.0, 8.0), (8, 8), "regular")


@pytest.mark.parametrize("layout, dtype", LAYOUTS)
@pytest.mark.parametrize("restraints", [None, ((0.0, 0.0), (8.0, 8.0), (20.0, 20.0), (20.0, 20.0))])
def test_funn_network_force(layout, dtype, restraints):
    """funn.py:187-296 with train_freq = 2: calls 1-4 are plain ABF, call 5 refits (LM capped at 0
    iterations so both sides keep the SAME weights; mean / std come from the normalized, Blackman-
    smoothed binned force) and from then on F = std * mlp(scale(f32(xi))) + mean inside the kernel."""
    snap = snapshot(layout, dtype)
    kw = dict(grid=FUNN_GRID, topology=(7, 5), N=2, train_freq=2, restraints=restraints, max_iters=0, seed=4)
    run = ParityRun(snap, POINT_CVS["two-distances"], ("FUNN", kw), units="real" if layout == "lammps" else None)
    for pos in perturbed(snap, 9, 0.1):
        run.step(pos)
    assert run.native.ncalls == 9
    nn, onn = run.state.nn, run.ostate.nn
    tol = max(1e-9, 10 * run.rtol)  # fp32 engines: Fsum itself only agrees to single precision
    np.testing.assert_allclose(nn.mean.cpu().numpy(), onn.mean, rtol=tol)
    np.testing.assert_allclose(nn.std.cpu().numpy(), onn.std, rtol=tol)
    assert np.any(nn.std.cpu().numpy() != 0.0)  # the network force was live, not a zero


def test_funn_refit_tracks_oracle():
    """The cadenced Levenberg-Marquardt refit (ml/optimizers.py:188-232) on the device against the numpy
    restatement: same start, a few iterations -> same weights to solver rounding, same network force."""
    snap = snapshot("hoomd", np.float64)
    kw = dict(grid=FUNN_GRID, topology=(6,), N=2, train_freq=3, max_iters=4, seed=9)
    run = ParityRun(snap, POINT_CVS["two-distances"], ("FUNN", kw))
    frames = perturbed(snap, 8, 0.1)
    for pos in frames[:6]:
        run.step(pos)
    ours, ref = run.step(frames[6], check=False)  # call 7 = 2 * train_freq + 1: first fit
    p, po = ours.nn.params.cpu().numpy(), ref.nn.params
    assert not np.array_equal(po, run.omethod.initial_params)
    np.testing.assert_allclose(p, po, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ours.F.cpu().numpy(), ref.F.numpy(), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(ours.bias.cpu().numpy(), ref.bias.numpy(), rtol=1e-5, atol=1e-8)
    assert np.array_equal(ours.hist.cpu().numpy().astype(np.int64), ref.hist.numpy())


@pytest.mark.parametrize("activation", ["tanh", "sin"])
@pytest.mark.parametrize("topology", [(), (128,) * 7, (3, 128, 1, 40)])
def test_force_net_kernel_against_oracle_mlp(activation, topology):
    """psb_set_force_net directly: widest / deepest networks the ABI admits, tanh and sine layers,
    weights handed over from device memory; F read back from the state against oracle.ml.mlp_apply."""
    from oracle import ml as oml
    from pysages_b200 import _native as N

    snap = snapshot("hoomd", np.float64)
    grid = ((0.0, 0.0), (8.0, 8.0), (8, 8), "regular")
    run = ParityRun(snap, POINT_CVS["two-distances"], ("ABF", dict(grid=grid, N=2)))
    widths = oml.layer_sizes(2, topology, 2)
    rng = np.random.default_rng(len(topology))
    params = oml.init_params(widths, seed=21) * (3.0 if activation == "tanh" else 1.0)
    mean, std = rng.normal(size=2), rng.uniform(0.5, 2.0, size=2)
    dev_params = torch.as_tensor(params, device="cuda")
    run.native.set_force_net(widths, dev_params, mean, std, predict_after=2,
                             activation=N.ACT_SIN if activation == "sin" else N.ACT_TANH, omega0=16.0, omega=1.5)
    frames = perturbed(snap, 5, 0.1)
    for i, pos in enumerate(frames):
        ours, ref = run.step(pos, check=i < 2)  # calls 1, 2: still the binned average == the oracle's ABF
        if i >= 2:
            xi = ours.xi.cpu().numpy().astype(np.float32).astype(np.float64)
            want = std * oml.mlp_apply(params, widths, xi, grid[0], grid[1], activation, 16.0, 1.5).ravel() + mean
            np.testing.assert_allclose(ours.force.cpu().numpy(), want, rtol=1e-10, atol=1e-12)
            # hist / Fsum keep accumulating exactly as in ABF, with the network force as F_prev
            assert int(ours.hist.sum()) == i + 1
    run.native.set_force_net([], None, None, None, predict_after=0)  # n_dense = 0: plain ABF again
    ours, _ = run.step(frames[0], check=False)
    import pysages_b200 as ps

    idx = [int(v) for v in ps.grids.build_indexer(run.method.grid)(ours.xi.cpu().numpy().reshape(1, -1))]
    flat = int(np.ravel_multi_index(tuple(min(v, 7) for v in idx), (8, 8)))  # gathers clamp
    want = ours.Fsum.reshape(-1, 2)[flat].cpu().numpy() / max(2, int(ours.hist.reshape(-1)[flat]))
    np.testing.assert_allclose(ours.force.cpu().numpy(), want, rtol=1e-12)


# ---- the two grid-wide exchanges of the point-CV class ------------------------------------------------
@pytest.mark.parametrize("method", ["harmonic", "abf"])
@pytest.mark.parametrize("exchange", ["flags", "counter"])
def test_pass_a_exchange_variants(method, exchange, monkeypatch, launch_shape):
    """HarmonicBias / ABF (and Unbiased) on point CVs exchange their per-CTA partial sums through flagged words
    (ll_store / reduce_rows_ll, psb_step.cuh); PSB_NO_LL=1 keeps the counter barrier every other
    configuration uses.  Both against the oracle, and bitwise against each other (same reduction order)."""
    if launch_shape != "grid24":
        pytest.skip("needs the multi-CTA launch shape")
    if exchange == "counter":
        monkeypatch.setenv("PSB_NO_LL", "1")
    snap = snapshot("hoomd", np.float64, n=3000, seed=31)
    cvs = [("Distance", ([list(range(0, 900)), list(range(1000, 1700))],), {}),
           ("Component", (list(range(2000, 2600)), 2), {})]
    spec = {"harmonic": harmonic([[3.0, 0.5], [0.5, 2.0]], [4.0, 1.0]),
            "abf": ("ABF", dict(grid=((0.0, -12.0), (12.0, 12.0), (8, 8), "regular"), N=2))}[method]
    run = ParityRun(snap, cvs, spec)
    info = run.native.launch_info()
    assert info["grid_blocks"] > 1 and info["cooperative"]
    states = []
    for pos in perturbed(snap, 5, 0.05):
        ours, _ = run.step(pos)
        states.append((ours.xi.cpu().numpy().copy(), ours.bias.cpu().numpy().copy()))
    key = (method,)
    seen = test_pass_a_exchange_variants.__dict__.setdefault("seen", {})
    if key in seen:  # second variant of the same method: bit-identical to the first
        for (xa, ba), (xb, bb) in zip(seen[key], states):
            assert np.array_equal(xa, xb) and np.array_equal(ba, bb)
    else:
        seen[key] = states


# ---- Cremer-Pople ring puckering (colvars/angles.py:131-281) ------------------------------------------
RING6, RING5, RING7 = [40, 7, 313, 99, 150, 508], [3, 77, 140, 201, 333], [10, 20, 30, 45, 50, 60, 75]


@pytest.mark.parametrize("layout, dtype", [("hoomd", np.float64), ("hoomd", np.float32), ("lammps", np.float64),
                                           ("openmm-cuda", np.float64)])
@pytest.mark.parametrize("cv", ["RingAmplitude", "RingPhaseAngle", "RingPuckeringCoordinates"])
def test_ring_puckering_bias(layout, dtype, cv):
    snap = snapshot(layout, dtype)
    if cv == "RingPuckeringCoordinates":  # HarmonicBias takes one centre per CV (bias.py:62-65): use metadynamics
        method = ("Metadynamics", dict(height=0.7, sigma=[0.4, 0.3], stride=2, ngaussians=6))
    else:
        method = harmonic(3.0, [0.5])
    run = ParityRun(snap, [(cv, (RING6,), {})], method)
    for pos in perturbed(snap, 3, 0.1):
        run.step(pos)


def test_ring_puckering_with_other_cvs_metadynamics():
    """Phase of a 5-ring + amplitude of a 7-ring + a Distance sharing atoms with the 5-ring, direct-sum
    metadynamics: general class, unique-atom scatter."""
    snap = snapshot("hoomd", np.float64)
    specs = [("RingPhaseAngle", (RING5,), {}), ("RingAmplitude", (RING7,), {}), ("Distance", ([[3, 77], B],), {})]
    kw = dict(height=0.8, sigma=[0.3, 0.2, 0.5], stride=2, ngaussians=8)
    run = ParityRun(snap, specs, ("Metadynamics", kw))
    for pos in perturbed(snap, 6, 0.05):
        run.step(pos)
    assert int(run.state.idx) == 2  # calls 3 and 5 deposit (metad.py:192-193)


def test_ring_amplitude_abf_position_dependent_jacobian():
    snap = snapshot("hoomd", np.float64)
    specs = [("RingAmplitude", (RING6,), {}), ("Component", ([A], 0), {})]
    kw = dict(grid=((0.0, -6.0), (12.0, 6.0), (6, 6), "regular"), N=1)
    run = ParityRun(snap, specs, ("ABF", kw))
    for pos in perturbed(snap, 4, 0.05):
        run.step(pos)


# ---- SpectralABF: ABF accumulation + force from the fitted series (spectral_abf.py:95-268) --------------
@pytest.mark.parametrize("case", ["fourier-2d", "monomial-1d", "monomial-2d-restraints", "monomial-2d-oob-zero"])
def test_spectral_abf(case, golden):
    """fit_threshold = 3, fit_freq = 2: calls 1-3 plain ABF, refits before calls 5, 7, 9 (pinv product on the
    device vs numpy), force = gradient of the series evaluated in the kernel; outside the grid restraints,
    or -- without them -- zero force (not the clamped bin average ABF / FUNN use)."""
    if case == "fourier-2d":
        snap = adp(golden)
        cvs = ADP_CVS
        kw = dict(grid=((-math.pi, -math.pi), (math.pi, math.pi), (12, 10), "periodic"), N=2)
        step = 0.02
    else:
        snap = snapshot("hoomd", np.float64)
        step = 0.1
        if case == "monomial-1d":
            cvs = POINT_CVS["distance-groups"]
            kw = dict(grid=((0.0,), (14.0,), (24,), "regular"), N=2)
        else:
            cvs = POINT_CVS["two-distances"]
            hi = 8.0 if case.endswith("restraints") else 4.5  # the smaller box leaves xi outside most of the time
            kw = dict(grid=((0.0, 0.0), (hi, hi), (8, 8), "regular"), N=2)
            if case.endswith("restraints"):
                kw["restraints"] = ((0.0, 0.0), (hi, hi), (20.0, 20.0), (20.0, 20.0))
    kw.update(fit_threshold=3, fit_freq=2)
    run = ParityRun(snap, cvs, ("SpectralABF", kw))
    oob_steps = 0
    for pos in perturbed(snap, 10, step):
        ours, ref = run.step(pos)
        oob_steps += int(not np.any(ref.force.numpy()))
    fun, ofun = run.state.fun, run.ostate.fun
    np.testing.assert_allclose(float(fun.scale), ofun.scale, rtol=1e-10)
    np.testing.assert_allclose(fun.coefficients.cpu().numpy(), ofun.coefficients, rtol=1e-8,
                               atol=1e-11 * np.abs(ofun.coefficients).max())
    if case == "monomial-2d-oob-zero":
        assert oob_steps > 0  # the zero-force branch was taken


# ---- ANN: histogram + force from the gradient of the network (ann.py:65-258) -------------------------------
@pytest.mark.parametrize("layout, dtype", [("hoomd", np.float64), ("hoomd", np.float32), ("lammps", np.float64),
                                           ("openmm-cuda", np.float64)])
@pytest.mark.parametrize("dims", [1, 2])
def test_ann_network_gradient_force(layout, dtype, dims):
    """train_freq = 2: calls 1-2 unbiased histogramming, refits before calls 3, 5, 7 (LM capped at 0 iterations:
    both sides keep the same weights, std comes from the normalized / smoothed kT log(prob)), force =
    std * d net / d xi by back-propagation inside the kernel, zero outside the grid; histogram reset at refits."""
    snap = snapshot(layout, dtype)
    if dims == 1:
        cvs, grid = POINT_CVS["distance-groups"], ((0.0,), (14.0,), (16,), "regular")
    else:
        cvs, grid = POINT_CVS["two-distances"], ((0.0, 0.0), (8.0, 8.0), (8, 8), "regular")
    kw = dict(grid=grid, topology=(6, 5), kT=2.5, train_freq=2, max_iters=0, seed=7)
    run = ParityRun(snap, cvs, ("ANN", kw), units="real" if layout == "lammps" else None)
    for pos in perturbed(snap, 8, 0.1):
        ours, ref = run.step(pos)
        np.testing.assert_allclose(ours.phi.cpu().numpy(), ref.phi, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(ours.prob.cpu().numpy(), ref.prob, rtol=1e-10)
    assert int(run.state.hist.sum()) <= 2  # reset at the refit before call 7, then calls 7 and 8
    np.testing.assert_allclose(float(run.state.nn.std), run.ostate.nn.std, rtol=1e-8)


def test_ann_refit_tracks_oracle():
    snap = snapshot("hoomd", np.float64)
    kw = dict(grid=((0.0, 0.0), (8.0, 8.0), (8, 8), "regular"), topology=(6,), kT=2.5, train_freq=3, max_iters=4, seed=3)
    run = ParityRun(snap, POINT_CVS["two-distances"], ("ANN", kw))
    frames = perturbed(snap, 5, 0.1)
    for pos in frames[:3]:
        run.step(pos)
    ours, ref = run.step(frames[3], check=False)  # call 4 = train_freq + 1: first fit
    np.testing.assert_allclose(ours.nn.params.cpu().numpy(), ref.nn.params, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ours.bias.cpu().numpy(), ref.bias.numpy(), rtol=1e-5, atol=1e-9)
    assert np.array_equal(ours.hist.cpu().numpy().astype(np.int64), ref.hist.numpy())


@pytest.mark.parametrize("activation", ["tanh", "sin"])
def test_force_net_gradient_mode_against_oracle(activation):
    """psb_set_force_net(gradient=1) on a deep / wide network, weights from device memory: F against the
    oracle's back-propagation (tanh) or autograd of the sine network."""
    from oracle import ml as oml
    from pysages_b200 import _native as N

    snap = snapshot("hoomd", np.float64)
    grid = ((0.0, 0.0), (8.0, 8.0), (8, 8), "regular")
    cvs = [ps_.colvars.Distance([A, B]), ps_.colvars.Distance([C, D])]
    method = ps_.methods.ANN(cvs, ps_.Grid((0.0, 0.0), (8.0, 8.0), (8, 8)), (4,), 2.5, train_freq=10**9, dense_bias=True)
    from parity_utils import device_helpers, device_snapshot

    dsnap = device_snapshot(snap)
    _, init, update = method.build(dsnap, device_helpers("hoomd"))
    state = init()
    widths = oml.layer_sizes(2, (128, 3, 64, 128, 128, 9, 128), 1)
    params = oml.init_params(widths, seed=5)  # glorot scale: gain ~1 per layer keeps the gradient well conditioned
    update.native.set_force_net(widths, torch.as_tensor(params, device="cuda"), [0.0], [1.7], predict_after=0,
                                activation=N.ACT_SIN if activation == "sin" else N.ACT_TANH, omega0=3.0, omega=1.5,
                                gradient=True)
    for pos in perturbed(snap, 3, 0.1):
        dsnap = dsnap._replace(positions=torch.as_tensor(np.ascontiguousarray(pos)).cuda())
        state = update(dsnap, state)
        torch.cuda.synchronize()
        xi = state.xi.cpu().numpy().astype(np.float32).astype(np.float64).ravel()
        if activation == "tanh":
            want = 1.7 * oml.mlp_input_gradient(params, widths, xi, grid[0], grid[1])
        else:
            x = torch.tensor(xi, dtype=torch.float64, requires_grad=True)
            h = (x - torch.tensor(grid[0])) * 2 / (torch.tensor(grid[1]) - torch.tensor(grid[0])) - 1
            layers = oml.split_params(params, widths)
            h = h.reshape(1, 2)
            for l, (W, b) in enumerate(layers):
                h = h @ torch.as_tensor(W) + torch.as_tensor(b)
                if l + 1 < len(layers):
                    h = torch.sin((3.0 if l == 0 else 1.5) * h)
            (g,) = torch.autograd.grad(h.sum(), x)
            want = 1.7 * g.numpy()
        got = update.native.view("force").cpu().numpy()[:2]
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12 * np.abs(want).max())


@pytest.mark.parametrize("nblocks", [2, 3, 7, 37, 96, 148])
def test_pass_a_exchange_grid_sizes(nblocks, monkeypatch, launch_shape):
    """The flagged-word exchange over grids from 2 CTAs to one CTA per SM: group-aligned splits with 1 ... 74
    CTAs per group (more than the 48 entries one polling round covers), ABF with restraints against the oracle."""
    if launch_shape != "auto":
        pytest.skip("sets its own launch shape")
    monkeypatch.setenv("PSB_GRID_BLOCKS", str(nblocks))
    snap = snapshot("hoomd", np.float32, n=40000, seed=5)
    h = 20000
    cvs = [("Distance", ([list(range(0, h)), list(range(h, 2 * h))],), {})]
    kw = dict(grid=((0.0,), (30.0,), (16,), "regular"), N=2, restraints=((0.0,), (30.0,), (5.0,), (5.0,)))
    run = ParityRun(snap, cvs, ("ABF", kw))
    info = run.native.launch_info()
    assert info["grid_blocks"] == nblocks
    for pos in perturbed(snap, 3, 0.05):
        run.step(pos, check=False)
        run.check(jacobian=False)  # the dense (1, 3N) Jacobian compare is covered elsewhere; keep this one fast


# ---- CFF: ABF + second histogram + force network; Sobolev-loss energy network (cff.py:109-350) -------------
@pytest.mark.parametrize("layout, dtype", [("hoomd", np.float64), ("hoomd", np.float32), ("lammps", np.float64)])
@pytest.mark.parametrize("restraints", [None, ((0.0, 0.0), (8.0, 8.0), (20.0, 20.0), (20.0, 20.0))])
def test_cff_network_force(layout, dtype, restraints):
    """train_freq = 2, both Levenberg-Marquardt fits capped at 0 iterations (same weights on both sides): the
    per-step kernel path (ABF + histp + force network with the scalar std and the per-component mean) and the
    host pipeline of cff.py:280-307 (prob, fe, normalisation, float32 smoothing) against the oracle."""
    snap = snapshot(layout, dtype)
    # kT far above the synthetic mean-force scale: the untrained energy network times that scale must not
    # overflow exp(fe / kT) (it does in the reference algorithm itself for a small kT)
    kw = dict(grid=FUNN_GRID, topology=(6, 4), kT=1.0e4, N=2, train_freq=2, restraints=restraints, max_iters=0,
              fmax_iters=0, seed=2)
    run = ParityRun(snap, POINT_CVS["two-distances"], ("CFF", kw))
    tol = max(1e-8, 10 * run.rtol)
    for pos in perturbed(snap, 8, 0.1):
        ours, ref = run.step(pos)
        assert np.array_equal(ours.histp.cpu().numpy().astype(np.int64), ref.histp.numpy())
        np.testing.assert_allclose(ours.prob.cpu().numpy(), ref.prob, rtol=tol)
        np.testing.assert_allclose(ours.fe.cpu().numpy(), ref.fe, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(float(run.state.fnn.std), run.ostate.fnn.std, rtol=tol)
    np.testing.assert_allclose(run.state.fnn.mean.cpu().numpy(), run.ostate.fnn.mean, rtol=tol, atol=1e-12)


def test_cff_refit_tracks_oracle():
    """Three Sobolev-loss LM iterations of the energy network and three SSE iterations of the force network on
    the device (forward-over-reverse Jacobian) against the oracle's reverse-over-reverse autograd."""
    snap = snapshot("hoomd", np.float64)
    kw = dict(grid=((0.0, 0.0), (8.0, 8.0), (6, 6), "regular"), topology=(4,), kT=1.0e4, N=2, train_freq=3,
              max_iters=3, fmax_iters=3, seed=6)
    run = ParityRun(snap, POINT_CVS["two-distances"], ("CFF", kw))
    frames = perturbed(snap, 5, 0.1)
    for pos in frames[:3]:
        run.step(pos)
    ours, ref = run.step(frames[3], check=False)  # call 4 = train_freq + 1: first fit
    np.testing.assert_allclose(ours.nn.params.cpu().numpy(), ref.nn.params, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ours.fnn.params.cpu().numpy(), ref.fnn.params, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ours.force.cpu().numpy(), ref.force.numpy(), rtol=1e-4, atol=1e-8)
    assert int(ours.histp.sum()) == 1  # reset at the fit, then this step's visit


# ---- Sirens: Siren free-energy network, force = its gradient (sirens.py:95-420) ---------------------------
@pytest.mark.parametrize("mode", ["abf", "cff"])
@pytest.mark.parametrize("layout, dtype", [("hoomd", np.float64), ("hoomd", np.float32)])
@pytest.mark.parametrize("restraints", [None, ((0.0, 0.0), (8.0, 8.0), (20.0, 20.0), (20.0, 20.0))])
def test_sirens_network_gradient_force(mode, layout, dtype, restraints):
    """train_freq = 2, LM capped at 0 iterations (same weights on both sides): per-step kernel path (ABF
    [+ histp], back-propagation through sine layers with omega_0 = 16) and the host-side targets / scale."""
    snap = snapshot(layout, dtype)
    kw = dict(grid=FUNN_GRID, topology=(6, 5), N=2, train_freq=2, restraints=restraints, max_iters=0, seed=4, mode=mode)
    if mode == "cff":
        kw["kT"] = 1.0e4
    run = ParityRun(snap, POINT_CVS["two-distances"], ("Sirens", kw))
    for pos in perturbed(snap, 8, 0.1):
        ours, ref = run.step(pos)
        if mode == "cff":
            assert np.array_equal(ours.histp.cpu().numpy().astype(np.int64), ref.histp.numpy())
            np.testing.assert_allclose(ours.fe.cpu().numpy(), ref.fe, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(float(run.state.nn.std), run.ostate.nn.std, rtol=max(1e-8, 10 * run.rtol))


@pytest.mark.parametrize("mode", ["abf", "cff"])
def test_sirens_refit_tracks_oracle(mode):
    snap = snapshot("hoomd", np.float64)
    kw = dict(grid=((0.0, 0.0), (8.0, 8.0), (6, 6), "regular"), topology=(4,), N=2, train_freq=3, max_iters=3, seed=6,
              mode=mode, omega_0=4)
    if mode == "cff":
        kw["kT"] = 1.0e4
    run = ParityRun(snap, POINT_CVS["two-distances"], ("Sirens", kw))
    frames = perturbed(snap, 5, 0.1)
    for pos in frames[:3]:
        run.step(pos)
    ours, ref = run.step(frames[3], check=False)  # call 4 = train_freq + 1: first fit
    np.testing.assert_allclose(ours.nn.params.cpu().numpy(), ref.nn.params, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ours.force.cpu().numpy(), ref.force.numpy(), rtol=1e-4, atol=1e-8)


# ---- eRMSD (colvars/orientation.py:129-378) -----------------------------------------------------------
def _rna_like(n_bases, seed, box=None):
    """n bases: three ring atoms each around points of a loose helix (nm-like units, neighbours within the cutoff)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_bases)
    centres = np.stack([0.9 * np.cos(0.6 * t), 0.9 * np.sin(0.6 * t), 0.33 * t], axis=1)
    tri = np.array([[0.14, 0.0, 0.0], [-0.07, 0.12, 0.0], [-0.07, -0.12, 0.0]])
    out = []
    for c in centres:
        Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        out.append(c + tri @ Q.T + 0.01 * rng.normal(size=(3, 3)))
    return np.concatenate(out)


@pytest.mark.parametrize("layout, dtype, nb", [("hoomd", np.float64, 12), ("lammps", np.float64, 7),
                                               ("openmm-cuda", np.float64, 20), ("hoomd", np.float64, 64)])
def test_ermsd_harmonic(layout, dtype, nb):
    """eRMSD of nb bases against a perturbed reference + HarmonicBias: value, the hand-written adjoints (dense
    Jacobian against the oracle's autograd of the reference formula) and the applied forces."""
    snap = snapshot(layout, dtype, n=max(N, 3 * nb + 50))
    rng = np.random.default_rng(nb)
    idx = rng.permutation(snap["arrays"]["positions"].shape[0] if layout != "openmm-cuda" else N)[: 3 * nb].tolist()
    structure = _rna_like(nb, seed=nb)
    reference = structure + 0.03 * rng.normal(size=structure.shape)
    pos = snap["arrays"]["positions"]
    tags = np.asarray(idx)
    specs = [("ERMSD", (idx, reference), dict(cutoff=2.4, a=0.5, b=0.3))]
    # place the selected atoms on the structure through the oracle's own tag -> row map, so that every layout
    # (HOOMD rtag, OpenMM permuted order, LAMMPS tag map) is written consistently; zero image flags
    import oracle
    helpers, _ = oracle.backends.build_helpers(layout, {"positions", "indices"}, True)
    from parity_utils import oracle_snapshot
    osnap = oracle_snapshot(snap)
    index_of = np.asarray(helpers.query(osnap).indices).astype(np.int64)
    rows = index_of[tags]
    pos[rows, :3] = structure.astype(pos.dtype)
    if "images" in snap["arrays"]:
        img = snap["arrays"]["images"]
        if img.ndim == 2:
            img[rows] = 0
        else:  # LAMMPS packed image flags: zero image = 512 in every 10-bit field
            img[rows] = 512 | (512 << 10) | (512 << 20)
    run = ParityRun(snap, specs, harmonic(5.0, [0.3]))
    for frame in perturbed(snap, 3, 0.01):
        run.step(frame)
    assert float(run.state.xi[0, 0]) > 0.0


@pytest.mark.parametrize("layout, nb", [("hoomd", 10), ("lammps", 17)])
def test_ermsd_coarse_grained_harmonic(layout, nb):
    """ERMSDCG (orientation.py:380-590): sites -> reconstructed ring atoms -> eRMSD, adjoints through both frame
    constructions; dense Jacobian against autograd of the reference formula."""
    import oracle
    from parity_utils import oracle_snapshot

    snap = snapshot(layout, np.float64, n=max(N, 3 * nb + 50))
    rng = np.random.default_rng(100 + nb)
    idx = rng.permutation(snap["arrays"]["positions"].shape[0])[: 3 * nb].tolist()
    structure = _rna_like(nb, seed=nb + 1)
    reference = structure + 0.03 * rng.normal(size=structure.shape)
    sequence = rng.integers(0, 4, size=nb).tolist()
    helpers, _ = oracle.backends.build_helpers(layout, {"positions", "indices"}, True)
    rows = np.asarray(helpers.query(oracle_snapshot(snap)).indices).astype(np.int64)[np.asarray(idx)]
    pos = snap["arrays"]["positions"]
    pos[rows, :3] = structure.astype(pos.dtype)
    img = snap["arrays"]["images"]
    if img.ndim == 2:
        img[rows] = 0
    else:
        img[rows] = 512 | (512 << 10) | (512 << 20)
    specs = [("ERMSDCG", (idx, reference, sequence), dict(cutoff=2.4, a=0.5, b=0.3))]
    run = ParityRun(snap, specs, harmonic(5.0, [0.3]))
    for frame in perturbed(snap, 3, 0.01):
        run.step(frame)
    assert float(run.state.xi[0, 0]) > 0.0


# ---- more CVs than the first build allowed (colvars/core.py:228-274 stacks any number; here up to 8 CVs,
# ---- 8 components, 32 groups) -----------------------------------------------------------------------------------
@pytest.mark.parametrize("layout, dtype", [("hoomd", np.float32), ("hoomd", np.float64), ("openmm-cuda", np.float64),
                                           ("lammps", np.float64)])
def test_six_distances_harmonic_bias(layout, dtype):
    """k = 6 with a full 6 x 6 spring matrix over 12 groups."""
    snap = snapshot(layout, dtype)
    groups = [list(range(40 * i, 40 * i + 25 + i)) for i in range(12)]
    specs = [("Distance", ([groups[2 * i], groups[2 * i + 1]],), {}) for i in range(6)]
    rng = np.random.default_rng(0)
    a = rng.normal(size=(6, 6))
    kspring = (a @ a.T + 6.0 * np.eye(6)).tolist()
    run = ParityRun(snap, specs, harmonic(kspring, [0.5 * (i + 1) for i in range(6)]))
    for pos in perturbed(snap, 3, 0.05):
        run.step(pos)


def test_eight_dihedrals_metadynamics_direct_sum():
    """8 CVs, k = 8, 32 single-atom groups (the limits), periodic wrap on every component, deposits."""
    snap = snapshot("hoomd", np.float64)
    specs = [("DihedralAngle", ([4 + 7 * i, 6 + 7 * i, 8 + 7 * i, 14 + 7 * i],), {}) for i in range(8)]
    kw = dict(height=1.2, sigma=[0.35] * 8, stride=2, ngaussians=8)
    run = ParityRun(snap, specs, ("Metadynamics", kw))
    for pos in perturbed(snap, 6, 0.05):
        run.step(pos)
    assert int(run.state.idx) == 2 and run.native.k == 8


def test_ring_puckering_coordinates_plus_displacement_unbiased():
    """RingPuckeringCoordinates (2 components) + Displacement (3): k = 5, rejected by the 4-component build.  Values
    and the dense Jacobian against the oracle (the reference itself can only run this pair unbiased: its
    get_periods / HarmonicBias count CVs, not components, colvars/utils.py:13-19, bias.py:62-65)."""
    import oracle
    import pysages_b200 as ps
    from parity_utils import device_helpers, device_snapshot, oracle_snapshot

    snap = snapshot("hoomd", np.float64)
    cv_o = oracle.colvars.build(oracle.colvars.RingPuckeringCoordinates(RING6), oracle.colvars.Displacement([A, B]))
    helpers, _ = oracle.backends.build_helpers("hoomd", {"positions", "indices"}, True)
    xi_o, J_o = cv_o(helpers.query(oracle_snapshot(snap)))
    m = ps.methods.Unbiased([ps.colvars.RingPuckeringCoordinates(RING6), ps.colvars.Displacement([A, B])], dense_bias=True)
    dsnap = device_snapshot(snap)
    _, init, update = m.build(dsnap, device_helpers("hoomd"))
    st = update(dsnap, init())
    torch.cuda.synchronize()
    assert st.xi.shape == (1, 5) and update.native.k == 5
    close(st.xi.cpu().numpy().ravel(), xi_o.numpy().ravel(), 1e-10, what="xi")


def test_limits_are_reported():
    snap = snapshot("hoomd", np.float64)
    nine = [("Component", ([i], 0), {}) for i in range(9)]
    with pytest.raises(ValueError, match="between 1 and 8"):
        ParityRun(snap, nine, harmonic(1.0, [0.0] * 9))
