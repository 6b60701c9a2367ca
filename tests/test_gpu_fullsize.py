"""
Parity at BASELINE.json's FULL sizes (the other GPU tests use sizes the dense oracle finishes in
milliseconds): the oracle itself where it is still affordable (cfg2 at 1e5 atoms, the 1e6-atom
series incl. the slot-ordered kernels, cfg4 at 5e4 atoms), and independent numpy known answers plus
conservation laws where the dense formulation is not (cfg3: 1e8 pairs).
"""
import math

import numpy as np
import pytest
import torch

from parity_utils import ParityRun, close
from pysages_b200 import synthetic

pytestmark = pytest.mark.gpu


def groups4(n):
    q = n // 4
    return [list(range(i * q, (i + 1) * q)) for i in range(4)]


def test_cfg2_abf_100k_against_oracle():
    n = 100_000
    snap = synthetic.make_snapshot(n, layout="hoomd", dtype=np.float32, seed=20240 + 2000)
    g = groups4(n)
    L = snap["L"]
    cvs = [("Distance", ([g[0], g[1]],), {}), ("Distance", ([g[2], g[3]],), {})]
    spec = ("ABF", dict(grid=((0.0, 0.0), (L / 2, L / 2), (64, 64), "regular"), N=500,
                        restraints=((0.0, 0.0), (L / 2, L / 2), (10.0, 10.0), (10.0, 10.0))))
    run = ParityRun(snap, cvs, spec)
    assert run.native.launch_info()["grid_blocks"] > 1
    traj = synthetic.make_trajectory(snap, frames=3, step=0.02)
    for f in range(3):
        s, o = run.step(traj[f])
    # conservation: a distance between two centroids pushes the two groups with opposite total force
    bias = s.bias.cpu().numpy()
    assert np.abs(bias.sum(axis=0)).max() <= 1e-9 * max(np.abs(bias).max(), 1e-300) * n
    assert int(s.hist.sum().item()) == 3


@pytest.mark.parametrize("method", ["harmonic", "abf"])
def test_million_atoms_slot_ordered_kernels_against_oracle(method, monkeypatch):
    """S = N = 1e6: the library picks the slot-ordered gather / scatter by itself."""
    monkeypatch.delenv("PSB_SLOT_MAJOR", raising=False)
    monkeypatch.delenv("PSB_GRID_BLOCKS", raising=False)
    n = 1_000_000
    snap = synthetic.make_snapshot(n, layout="hoomd", dtype=np.float32, seed=20240 + 7)
    L = snap["L"]
    if method == "harmonic":
        cvs = [("Distance", ([list(range(0, n // 2)), list(range(n // 2, n))],), {})]
        spec = ("HarmonicBias", dict(kspring=10.0, center=[1.0]))
    else:
        g = groups4(n)
        cvs = [("Distance", ([g[0], g[1]],), {}), ("Distance", ([g[2], g[3]],), {})]
        spec = ("ABF", dict(grid=((0.0, 0.0), (L / 2, L / 2), (64, 64), "regular"), N=500))
    run = ParityRun(snap, cvs, spec)
    run.step()
    # dense outputs (parity harness) select the general kernels; the production handle is slot-ordered:
    import bench

    dev = torch.device("cuda:0")
    frames, L2, _ = bench.device_hoomd_frames(n, 1, dev, seed=5)
    snaps = bench.snapshots_of(frames, L2, 0.005)
    cv2, m2 = bench.workload_specs("harmonic" if method == "harmonic" else "abf2d", n, L2)
    monkeypatch.setenv("PSB_SLOT_MAJOR", "0")
    _, tagmajor = bench.build_ours(cv2, m2, snaps[0])
    monkeypatch.delenv("PSB_SLOT_MAJOR")
    rtag_only = bench.snapshots_of(frames, L2, 0.005, with_tags=False)  # the reference's five arrays, nothing else
    _, slotmajor = bench.build_ours(cv2, m2, rtag_only[0])  # planned without the tag array: two CTAs per SM
    _, slottags = bench.build_ours(cv2, m2, snaps[0])       # planned for it: one CTA per SM, gathers overlap
    assert slottags.wants_tags
    f0 = frames[0]["forces"].clone()
    outs = []
    for native, sn in ((tagmajor, rtag_only[0]), (slotmajor, rtag_only[0]), (slottags, snaps[0])):
        frames[0]["forces"].copy_(f0)
        native.step(sn)
        native.step(sn)
        torch.cuda.synchronize()
        outs.append((native.view("xi").cpu().numpy().copy(), native.view("cv_force").cpu().numpy().copy(),
                     frames[0]["forces"].cpu().numpy().copy()))
    for i, what in ((1, "slot-"), (2, "slot- (engine's slot -> tag array)")):
        close(outs[i][0], outs[0][0], 1e-12, what=f"xi {what} vs tag-ordered")
        close(outs[i][1], outs[0][1], 1e-9, what=f"generalized force {what} vs tag-ordered")
        assert np.abs(outs[i][2] - outs[0][2]).max() <= 2e-7 * np.abs(outs[0][2]).max()
    close(outs[2][0], outs[1][0], 1e-13, what="xi with vs without the tags array")
    assert np.abs(outs[0][2] - f0.cpu().numpy()).max() > 0  # the bias was applied


def test_cfg3_coordination_1e8_pairs_known_answer():
    """10k x 10k CoordinationNumber in a 1e6-atom system: value against chunked numpy, Newton's third
    law (the two groups receive opposite total forces) and the harmonic force magnitude."""
    import pysages_b200 as ps
    from parity_utils import device_helpers, device_snapshot

    n = 1_000_000
    snap = synthetic.make_snapshot(n, layout="hoomd", dtype=np.float32, seed=20240 + 3000)
    rng = np.random.Generator(np.random.PCG64(20240 + 3000))
    tags = rng.permutation(n)[:20_000]
    ga, gb = np.sort(tags[:10_000]), np.sort(tags[10_000:])
    r0, kspring, center = 1.0, 10.0, 100.0
    method = ps.methods.HarmonicBias([ps.colvars.CoordinationNumber([ga.tolist(), gb.tolist()], r0=r0)], kspring, [center])
    dsnap = device_snapshot(snap)
    f0 = dsnap.forces.clone()
    _, init, update = method.build(dsnap, device_helpers("hoomd"))
    state = update(dsnap, init())
    torch.cuda.synchronize()
    # numpy: unwrap, then sum 1 / (1 + (r / r0)^6) over all pairs, 500 rows at a time
    a = snap["arrays"]
    L = snap["L"]
    slot = a["rtags"]
    pos = a["positions"][:, :3].astype(np.float64) + L * a["images"].astype(np.float64)
    xa, xb = pos[slot[ga]], pos[slot[gb]]
    total = 0.0
    for i in range(0, len(xa), 500):
        d2 = ((xa[i : i + 500, None, :] - xb[None, :, :]) ** 2).sum(-1)
        total += (1.0 / (1.0 + (d2 / r0**2) ** 3)).sum()
    xi = float(state.xi.cpu().numpy().ravel()[0])
    assert abs(xi - total) <= 1e-10 * abs(total), (xi, total)
    df = (dsnap.forces - f0).cpu().numpy()[:, :3].astype(np.float64)
    fa, fb = df[slot[ga]].sum(0), df[slot[gb]].sum(0)
    scale = np.abs(df).max() * 1e4
    assert np.abs(fa + fb).max() <= 1e-5 * scale  # fp32 force array: one rounding per atom
    untouched = np.ones(n, bool); untouched[slot[ga]] = False; untouched[slot[gb]] = False
    assert not df[untouched].any()
    # every per-atom gradient of the 1e8-pair sum against chunked numpy (SURVEY.md appendix A:
    # d xi / d r_i = sum_j s'(r_ij) (r_i - r_j) / r_ij, s'(r) = -6 x^5 / (r0 (1 + x^6)^2), x = r / r0; opposite sign for j)
    gA = np.zeros((len(xa), 3)); gB = np.zeros((len(xb), 3))
    for i in range(0, len(xa), 500):
        d = xa[i : i + 500, None, :] - xb[None, :, :]
        r2 = (d**2).sum(-1)
        x6 = (r2 / r0**2) ** 3
        coef = -6.0 * x6 / (r2 * (1.0 + x6) ** 2)  # s'(r) / r
        t = coef[:, :, None] * d
        gA[i : i + 500] = t.sum(1); gB -= t.sum(0)
    F_np = kspring * (total - center)
    want = np.zeros((n, 3)); want[slot[ga]] = -F_np * gA; want[slot[gb]] = -F_np * gB
    f0n = f0.cpu().numpy()[:, :3].astype(np.float64)
    # the force array is fp32: one rounding of (f0 + bias) per component
    tol = 1.2e-7 * np.maximum(np.abs(f0n + want), np.abs(f0n))
    assert np.all(np.abs(df - want) <= tol + 1e-9 * np.abs(want).max()), np.abs(df - want).max()
    F = float(update.native.view("cv_force").cpu().numpy().ravel()[0])
    assert abs(F - kspring * (xi - center)) <= 1e-12 * abs(F)


def test_cfg4_rmsd_rog_50k_direct_hills_against_oracle_and_numpy():
    n, H = 50_000, 20_000
    snap = synthetic.make_snapshot(n, layout="hoomd", dtype=np.float64, seed=20240 + 4000, globule=True)
    snap["arrays"]["images"][:] = 0
    rng = np.random.default_rng(4)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    pos_by_tag = snap["arrays"]["positions"][snap["arrays"]["rtags"]][:, :3]
    ref = pos_by_tag @ q.T + rng.normal(scale=0.1, size=(n, 3))
    allatoms = list(range(n))
    cvs = [("RMSD", (allatoms, ref), {}), ("RadiusOfGyration", (allatoms,), {})]
    kw = dict(height=1.0, sigma=[0.05, 5.0], stride=3, ngaussians=H)
    run = ParityRun(snap, cvs, ("Metadynamics", kw))
    xi0 = run.state.xi.cpu().numpy().ravel()
    # known answers at full size, independent of the oracle: numpy Kabsch + <r.r>
    P = pos_by_tag - pos_by_tag.mean(0); Q = ref - ref.mean(0)
    V, S, Wt = np.linalg.svd(P.T @ Q)
    d = np.sign(np.linalg.det(V) * np.linalg.det(Wt)); V[:, -1] *= d
    rmsd = math.sqrt((((P @ (V @ Wt)) - Q) ** 2).sum() / n)
    assert abs(xi0[0] - rmsd) <= 1e-10 * rmsd
    assert abs(xi0[1] - (pos_by_tag**2).sum() / n) <= 1e-12 * xi0[1]
    # pre-filled hills on both sides, then steps through a deposit
    heights = np.zeros(H); heights[: H - 5] = rng.uniform(0.5, 1.2, H - 5)
    centers = np.zeros((H, 2)); centers[: H - 5] = xi0 + rng.normal(size=(H - 5, 2)) * np.array([0.05, 5.0])
    run.native.set_state("heights", heights); run.native.set_state("centers", centers)
    run.native.set_state("idx", np.array([H - 5], dtype=np.int64))
    run.ostate = run.ostate._replace(heights=torch.tensor(heights), centers=torch.tensor(centers), idx=H - 5)
    traj = synthetic.make_trajectory(snap, frames=4, step=0.01)
    for f in range(4):
        run.step(traj[f])
    assert int(run.state.idx) == H - 4


@pytest.mark.parametrize("layout", ["hoomd", "openmm-cuda"])
def test_radius_of_gyration_slot_ordered_against_general_class(layout, monkeypatch):
    """RadiusOfGyration over all atoms + a Distance CV + grid metadynamics: the slot-ordered class (sum of r.r in
    slot order, force rog_scale * r_i from a second streaming read of the positions) against the general class
    (tag-ordered gather), which the parity suite checks against the oracle.  OpenMM: the tag -> group table path."""
    import bench

    monkeypatch.delenv("PSB_GRID_BLOCKS", raising=False)
    n = 300_000
    dev = torch.device("cuda:0")
    if layout == "hoomd":
        frames, L, _ = bench.device_hoomd_frames(n, 1, dev, seed=11)
        snaps = bench.snapshots_of(frames, L, 0.005)
        forces = frames[0]["forces"]
    else:
        snaps, L = bench.device_layout_snapshots(layout, np.float32, n, 1, dev, seed=12)
        forces = snaps[0].forces
    half = n // 2
    cvs = [("RadiusOfGyration", (list(range(0, half)),), {}),
           ("Distance", ([list(range(half, half + n // 4)), list(range(half + n // 4, n))],), {})]
    # (the image flags put unwrapped coordinates up to 1.5 L from the origin: <r.r> reaches a few L^2)
    md = dict(height=1.0e6, sigma=[L * L / 10.0, 0.5], stride=2, ngaussians=16,
              grid=((0.0, 0.0), (4.0 * L * L, L), (64, 32), "regular"))
    monkeypatch.setenv("PSB_SLOT_MAJOR", "0")
    _, general = bench.build_ours(cvs, ("Metadynamics", md), snaps[0], layout=layout)
    monkeypatch.setenv("PSB_SLOT_MAJOR", "1")
    _, slot = bench.build_ours(cvs, ("Metadynamics", md), snaps[0], layout=layout)
    assert slot.launch_info()["grid_blocks"] > 1
    f0 = forces.clone()
    positions = snaps[0].positions
    p0 = positions.clone()
    outs = []
    for native in (general, slot):
        forces.copy_(f0)
        positions.copy_(p0)
        for _ in range(3):  # call 3 deposits a hill (stride 2) at the current xi
            native.step(snaps[0])
        positions[:, 0] += 0.05  # move off the hill centre: the bias force is non-zero now
        native.step(snaps[0])
        torch.cuda.synchronize()
        outs.append((native.view("xi").cpu().numpy().copy(), native.view("cv_force").cpu().numpy().copy(),
                     forces.cpu().numpy().astype(np.float64).copy()))
    close(outs[1][0], outs[0][0], 1e-12, what="xi slot- vs tag-ordered")
    close(outs[1][1], outs[0][1], 1e-9, what="generalized force slot- vs tag-ordered")
    scale = np.abs(outs[0][2] - f0.cpu().numpy().astype(np.float64)).max()
    assert scale > 0  # the bias was applied
    assert np.abs(outs[1][2] - outs[0][2]).max() <= max(2e-7 * np.abs(outs[0][2]).max(), 4.0 if layout != "hoomd" else 0.0)


def test_cfg4_bench_generator_1e5_hills_through_a_deposit_against_numpy():
    """BASELINE config 4 at its stated size, on bench.py's own generator and production handle (no dense
    outputs): RMSD + RoG over a 5e4-atom globule, direct sum over 1e5 TMA-staged hills, through a deposit.
    Known answers from numpy: Kabsch by SVD (orientation.py:63-95), <r.r> (shape.py:52-57), the Gaussian sum and
    its gradient (metad.py:318-323), -J^T F with the closed-form Jacobians."""
    import bench

    dev = torch.device("cuda:0")
    n, H = 50_000, 100_000
    frames, L, cvs4, md, prefill = bench.cfg4_case(dev, natoms=n, H=H)
    assert md["ngaussians"] == H
    snaps = bench.snapshots_of(frames, L, 0.005)
    _, native = bench.build_ours(cvs4, ("Metadynamics", dict(md)), snaps[0])
    prefill(native)
    native.set_state("ncalls", np.array([md["stride"]], dtype=np.int64))  # the next call deposits (metad.py:192-193)
    heights = native.view("heights").cpu().numpy().copy()
    centers = native.view("centers", (-1, 2)).cpu().numpy().copy()
    idx0 = int(native.view("idx").item())
    assert idx0 == H - 8 and np.count_nonzero(heights[:H]) == H - 8
    sigma = np.asarray(md["sigma"])
    ref_by_tag = np.asarray(cvs4[0][1][1], dtype=np.float64)

    def numpy_step(frame, heights, centers, nh):
        a = {k: v.cpu().numpy() for k, v in frame.items()}
        slot = a["rtags"].astype(np.int64)
        r = a["positions"][:, :3].astype(np.float64) + L * a["images"].astype(np.float64)
        x = r[slot]  # tag order
        P, Q = x - x.mean(0), ref_by_tag - ref_by_tag.mean(0)
        V_, _, Wt = np.linalg.svd(P.T @ Q)
        V_[:, -1] *= np.sign(np.linalg.det(V_) * np.linalg.det(Wt))
        U = V_ @ Wt
        rmsd = math.sqrt(((P @ U - Q) ** 2).sum() / n)
        rog = (x**2).sum() / n
        xi = np.array([rmsd, rog])
        d = xi - centers[:nh]
        e = heights[:nh] * np.exp(-0.5 * ((d / sigma) ** 2).sum(-1))
        F = (-e[:, None] * d / sigma**2).sum(0)
        terms = np.abs(e[:, None] * d / sigma**2).sum(0)  # conditioning of the signed sum
        J0 = (P - Q @ U.T) / (n * rmsd)
        J1 = 2.0 * x / n
        bias_by_tag = -(F[0] * J0 + F[1] * J1)
        bias = np.zeros((n, 3))
        bias[slot] = bias_by_tag
        return xi, e.sum(), F, terms, bias

    for step in range(2):
        frame = frames[step]
        f0 = frame["forces"].clone()
        xi, V, F, terms, bias = numpy_step(frame, heights, centers, min(idx0 + step, H))
        native.step(snaps[step])
        torch.cuda.synchronize()
        close(native.view("xi").cpu().numpy().ravel()[:2], xi, 1e-10, what="xi")
        got_F = native.view("cv_force").cpu().numpy().ravel()[:2]
        assert np.all(np.abs(got_F - F) <= 1e-10 * np.maximum(terms, np.abs(F))), (got_F, F, terms)
        energy = float(native.view("energy").item())
        deposited = step == 0
        want_E = V + (md["height"] if deposited else 0.0)
        assert abs(energy - want_E) <= 1e-10 * abs(want_E)
        df = (frame["forces"] - f0).cpu().numpy()[:, :3].astype(np.float64)
        fscale = np.abs(frame["forces"].cpu().numpy()).max()
        assert np.abs(df - bias).max() <= 2e-7 * max(fscale, np.abs(bias).max())
        if deposited:  # metad.py:262-271: plain height, centre = xi, stored at idx
            heights[idx0], centers[idx0] = md["height"], xi
            assert int(native.view("idx").item()) == idx0 + 1
            close(native.view("heights").cpu().numpy()[idx0], heights[idx0], 1e-15, what="deposited height")
            close(native.view("centers", (-1, 2)).cpu().numpy()[idx0], centers[idx0], 1e-10, what="deposited centre")
    assert int(native.view("ncalls").item()) == md["stride"] + 2


def test_slot_ordered_handle_takes_snapshots_with_and_without_the_tag_array(monkeypatch):
    """A handle planned for HOOMD's slot -> tag array also steps snapshots that lack it (the reference's five arrays):
    the two slot-ordered instantiations alternate on one handle -- same grid, same flag stamps -- and the trajectory of
    ABF states equals that of a handle that never saw the array."""
    import bench

    monkeypatch.setenv("PSB_SLOT_MAJOR", "1")
    monkeypatch.setenv("PSB_GRID_BLOCKS", "24")
    n = 40_000
    dev = torch.device("cuda:0")
    frames, L, _ = bench.device_hoomd_frames(n, 6, dev, seed=21)
    with_tags = bench.snapshots_of(frames, L, 0.005, with_tags=True)
    without = bench.snapshots_of(frames, L, 0.005, with_tags=False)
    cvs, m = bench.workload_specs("abf2d", n, L)
    f0 = [f["forces"].clone() for f in frames]
    outs = []
    for pattern in ("mixed", "never"):
        for f, f_init in zip(frames, f0):
            f["forces"].copy_(f_init)
        _, native = bench.build_ours(cvs, m, with_tags[0] if pattern == "mixed" else without[0])
        assert native.launch_info()["grid_blocks"] == 24 and native.wants_tags
        for i in range(6):
            native.step(with_tags[i] if (pattern == "mixed" and i % 2 == 0) else without[i])
        torch.cuda.synchronize()
        outs.append((native.view("xi").cpu().numpy().copy(), native.view("hist").cpu().numpy().copy(),
                     native.view("Fsum").cpu().numpy().copy(), np.stack([f["forces"].cpu().numpy() for f in frames])))
    close(outs[0][0], outs[1][0], 1e-13, what="xi")
    assert np.array_equal(outs[0][1], outs[1][1])
    close(outs[0][2], outs[1][2], 1e-10, what="Fsum")
    assert np.abs(outs[0][3] - outs[1][3]).max() <= 2e-7 * np.abs(outs[1][3]).max()


def test_dual_grid_handle_switches_between_replay_and_single_steps(monkeypatch):
    """Handles planned for the tag array run one CTA per SM under psb_run_cycle (graph replay: consecutive steps
    overlap) and two when stepped one launch at a time; switching back and forth restarts every grid-dependent counter
    and leaves the ABF state exactly as a plain sequence of single steps does."""
    import ctypes

    import bench
    from pysages_b200 import _native as N

    monkeypatch.delenv("PSB_SLOT_MAJOR", raising=False)
    monkeypatch.delenv("PSB_GRID_BLOCKS", raising=False)
    n = 400_000
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    frames, L, _ = bench.device_hoomd_frames(n, 4, dev, seed=33)
    snaps = bench.snapshots_of(frames, L, 0.005)
    cvs, m = bench.workload_specs("abf2d", n, L)
    f0 = [f["forces"].clone() for f in frames]
    results = []
    for mode in ("switching", "single"):
        for f, f_init in zip(frames, f0):
            f["forces"].copy_(f_init)
        _, native = bench.build_ours(cvs, m, snaps[0])
        descs = bench.snapdesc_array(native, snaps)
        grids = []
        with torch.cuda.stream(stream):
            if mode == "switching":
                for part in range(3):
                    N.check(native.lib.psb_run_cycle(native.handle, descs, 4, 8, ctypes.c_void_p(stream.cuda_stream)))
                    grids.append(native.launch_info()["grid_blocks"])
                    for i in range(4):
                        native.step(snaps[i])
                    grids.append(native.launch_info()["grid_blocks"])
            else:
                for part in range(3):
                    for i in range(12):
                        native.step(snaps[i % 4])
        stream.synchronize()
        if mode == "switching":
            assert grids[0] * 2 == grids[1] and grids[0::2] == [grids[0]] * 3 and grids[1::2] == [grids[1]] * 3, grids
        results.append((native.view("hist").cpu().numpy().copy(), native.view("Fsum").cpu().numpy().copy(),
                        int(native.view("ncalls").item()), np.stack([f["forces"].cpu().numpy() for f in frames])))
    assert results[0][2] == results[1][2] == 36
    assert np.array_equal(results[0][0], results[1][0])
    close(results[0][1], results[1][1], 1e-10, what="Fsum")
    assert np.abs(results[0][3] - results[1][3]).max() <= 2e-7 * np.abs(results[1][3]).max()


@pytest.mark.parametrize("method", ["harmonic", "abf2d"])
def test_openmm_slot_ordered_class_at_full_grid_against_tag_ordered(method, monkeypatch):
    """OpenMM-CUDA layout, 3e5 atoms, the library's own launch shape (one CTA per SM, two-level flagged exchange,
    gather before the dependency wait): same CVs, generalized forces and fixed-point force updates as the tag-ordered
    class, over several steps of graph replay and single launches."""
    import ctypes

    import bench
    from pysages_b200 import _native as N

    monkeypatch.delenv("PSB_GRID_BLOCKS", raising=False)
    n = 300_000
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    snaps, L = bench.device_layout_snapshots("openmm-cuda", np.float32, n, 4, dev, seed=44)
    cvs, m = bench.workload_specs(method, n, L)
    f0 = [s.forces.clone() for s in snaps]
    outs = []
    for slot in ("0", "1"):
        monkeypatch.setenv("PSB_SLOT_MAJOR", slot)
        for s, f in zip(snaps, f0):
            s.forces.copy_(f)
        _, native = bench.build_ours(cvs, m, snaps[0], layout="openmm-cuda")
        descs = bench.snapdesc_array(native, snaps)
        with torch.cuda.stream(stream):
            N.check(native.lib.psb_run_cycle(native.handle, descs, 4, 8, ctypes.c_void_p(stream.cuda_stream)))
            for i in range(4):
                native.step(snaps[i])
        stream.synchronize()
        outs.append((native.view("xi").cpu().numpy().copy(), native.view("cv_force").cpu().numpy().copy(),
                     np.stack([s.forces.cpu().numpy() for s in snaps]), native.launch_info()["grid_blocks"]))
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert outs[1][3] == 2 * sms  # slot-ordered, last stepped one launch at a time: two CTAs per SM (one under replay)
    close(outs[1][0], outs[0][0], 1e-12, what="xi slot- vs tag-ordered")
    close(outs[1][1], outs[0][1], 1e-9, what="generalized force slot- vs tag-ordered")
    moved = np.abs(outs[0][2].astype(np.float64) - np.stack([f.cpu().numpy() for f in f0]).astype(np.float64)).max()
    assert moved > 0
    # int64 fixed point (2^-32): the two classes add the same twelve biases per atom up to a few units in the last place
    assert np.abs(outs[1][2].astype(np.float64) - outs[0][2].astype(np.float64)).max() <= max(8.0, 1e-9 * moved)
