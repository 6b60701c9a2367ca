"""
CPU tests (gloo, world size 2) of the multi-GPU host logic: hill-record exchange in rank order and
window partitioning.  The stepping engine is the oracle (no GPU here); on the GPU box the same
`walker_step` drives `NativeWalkerEngine` (tests/test_gpu_parity.py::test_shared_hill_walkers).
"""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from parity_utils import make_cvs, make_method, oracle_snapshot
from pysages_b200 import parallel, synthetic

CVS = [("DihedralAngle", ([4, 6, 8, 14],), {}), ("DihedralAngle", ([6, 8, 14, 16],), {})]
GRID = ((-math.pi, -math.pi), (math.pi, math.pi), (12, 12), "periodic")
NSTEPS = 9


def _method(grid):
    kw = dict(height=1.2, sigma=[0.35, 0.35], stride=3, ngaussians=16, deltaT=5000.0, kB=0.0083144626, grid=grid)
    return make_method("oracle", ("Metadynamics", kw), make_cvs(oracle.colvars, CVS))


def _walker(rank, grid):
    snap = synthetic.make_snapshot(40, layout="hoomd", dtype=np.float64, seed=100 + rank)
    traj = synthetic.make_trajectory(snap, frames=NSTEPS, step=0.05, seed=7 + rank)
    method = _method(grid)
    helpers, bias = oracle.backends.build_helpers("hoomd", method.snapshot_flags, True)
    osnap = oracle_snapshot(snap)
    return oracle.walkers.SharedHillsWalker(method, osnap, helpers, bias), osnap, traj


def _run_lockstep(world, grid):
    """Single-process restatement: all walkers in one loop, records stacked in rank order."""
    ws = [_walker(r, grid) for r in range(world)]
    for t in range(NSTEPS):
        snaps = [osnap._replace(positions=np.array(traj[t], copy=True)) for (_, osnap, traj) in ws]
        if ws[0][0].next_step_deposits():
            recs = torch.stack([w.walker_begin(s) for (w, _, _), s in zip(ws, snaps)])
            for (w, _, _), s in zip(ws, snaps):
                w.walker_finish(s, recs)
        else:
            for (w, _, _), s in zip(ws, snaps):
                w.step(s)
    return [w.state for (w, _, _) in ws]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, grid, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        exchange = parallel.RecordExchange()
        assert (exchange.world_size, exchange.rank) == (world, rank)
        w, osnap, traj = _walker(rank, grid)
        for t in range(NSTEPS):
            parallel.walker_step(w, osnap._replace(positions=np.array(traj[t], copy=True)), exchange)
        s = w.state
        res = dict(xi=s.xi.numpy(), bias=s.bias.numpy(), heights=s.heights.numpy(), centers=s.centers.numpy(), idx=int(s.idx),
                   gp=None if s.grid_potential is None else s.grid_potential.numpy(),
                   gg=None if s.grid_gradient is None else s.grid_gradient.numpy())
        # window partition + final gather of the per-window means (umbrella integration)
        mine = parallel.partition_windows(7, world, rank)
        means = parallel.gather_window_means({i: [float(i), float(rank)] for i in mine}, 7)
        res["means"] = means.numpy()
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid", [None, GRID], ids=["direct", "grid"])
def test_shared_hills_over_gloo_match_single_process_restatement(grid):
    world = 2
    ref = _run_lockstep(world, grid)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), grid, out), nprocs=world, join=True)
    for r in range(world):
        got, want = out[r], ref[r]
        assert got["idx"] == int(want.idx) == 2 * world  # deposits at steps 4 and 7, one hill per walker each
        np.testing.assert_array_equal(got["heights"], want.heights.numpy())
        np.testing.assert_array_equal(got["centers"], want.centers.numpy())
        np.testing.assert_array_equal(got["xi"], want.xi.numpy())
        np.testing.assert_array_equal(got["bias"], want.bias.numpy())
        if grid is not None:
            np.testing.assert_array_equal(got["gp"], want.grid_potential.numpy())
            np.testing.assert_array_equal(got["gg"], want.grid_gradient.numpy())
    # every rank holds the same hills (rank order) although each walker saw its own trajectory
    np.testing.assert_array_equal(out[0]["heights"], out[1]["heights"])
    np.testing.assert_array_equal(out[0]["centers"], out[1]["centers"])
    assert not np.array_equal(out[0]["xi"], out[1]["xi"])
    want_means = np.array([[i, i % world] for i in range(7)], dtype=np.float64)
    for r in range(world):
        np.testing.assert_array_equal(out[r]["means"], want_means)


def test_partition_windows():
    for n, w in ((32, 8), (32, 4), (32, 2), (5, 8), (0, 3)):
        parts = [parallel.partition_windows(n, w, r) for r in range(w)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        parallel.partition_windows(4, 2, 2)


def test_record_exchange_single_process_is_identity():
    ex = parallel.RecordExchange()
    rec = torch.arange(4, dtype=torch.float64)
    assert ex(rec).shape == (1, 4) and torch.equal(ex(rec)[0], rec)
