"""2+ GPUs under torchrun: `pysages_b200.run` with Metadynamics(shared_hills=True), one walker per rank,
hill records all-gathered over NCCL; checks every rank against the single-process lockstep restatement
(oracle/walkers.py) -- run on the GPU box:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/walkers_nccl_check.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))

import oracle
import pysages_b200 as ps
from parity_utils import make_cvs, make_method, oracle_snapshot
from pysages_b200 import synthetic
from pysages_b200.backends import mock

CVS = [("DihedralAngle", ([4, 6, 8, 14],), {}), ("DihedralAngle", ([6, 8, 14, 16],), {})]
NSTEPS = 11
for grid in (None, ((-math.pi, -math.pi), (math.pi, math.pi), (16, 16), "periodic")):
    kw = dict(height=1.2, sigma=[0.35, 0.35], stride=3, ngaussians=32, deltaT=5000.0, kB=0.0083144626, grid=grid)

    def walker_inputs(r):
        snap = synthetic.make_snapshot(64, layout="hoomd", dtype=np.float64, seed=300 + r)
        return snap, synthetic.make_trajectory(snap, frames=NSTEPS, step=0.05, seed=17 + r)

    snap, traj = walker_inputs(rank)
    method = make_method("ours", ("Metadynamics", dict(kw, shared_hills=True)), make_cvs(ps.colvars, CVS))
    method.dense_bias = False
    eng = mock.MockEngine("hoomd", snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda", trajectory=traj)
    res = ps.run(method, lambda **kw_: eng, NSTEPS)
    torch.cuda.synchronize()
    st = res.states[0]
    # the single-process restatement of all walkers
    ws = []
    for r in range(world):
        s, t = walker_inputs(r)
        om = make_method("oracle", ("Metadynamics", kw), make_cvs(oracle.colvars, CVS))
        helpers, bias = oracle.backends.build_helpers("hoomd", om.snapshot_flags, True)
        osnap = oracle_snapshot(s)
        ws.append((oracle.walkers.SharedHillsWalker(om, osnap, helpers, bias), osnap, t))
    for step in range(NSTEPS):
        snaps = [o._replace(positions=np.array(t[step], copy=True)) for (_, o, t) in ws]
        if ws[0][0].next_step_deposits():
            recs = torch.stack([w.walker_begin(s) for (w, _, _), s in zip(ws, snaps)])
            for (w, _, _), s in zip(ws, snaps):
                w.walker_finish(s, recs)
        else:
            for (w, _, _), s in zip(ws, snaps):
                w.step(s)
    want = ws[rank][0].state
    ok = (np.allclose(st.heights.cpu().numpy().ravel(), want.heights.numpy(), rtol=1e-10, atol=0)
          and np.allclose(st.centers.cpu().numpy(), want.centers.numpy(), rtol=1e-10, atol=1e-14)
          and int(st.idx) == int(want.idx) == 3 * world
          and np.allclose(st.xi.cpu().numpy(), want.xi.numpy(), rtol=1e-10)
          and np.allclose(eng.t["forces"].cpu().numpy()[:, :3], ws[rank][1].forces[:, :3], rtol=1e-9, atol=1e-12))
    if grid is not None:
        ok = ok and np.allclose(st.grid_gradient.cpu().numpy(), want.grid_gradient.numpy(), rtol=1e-10, atol=1e-14)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"walkers over NCCL, world={world}, grid={'yes' if grid else 'no'}: {'OK' if flag.item() else 'MISMATCH'}")
    assert flag.item() == 1
dist.destroy_process_group()
