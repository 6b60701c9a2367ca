"""e2e through psb_step_host for a small selection in a large host-resident system: selected rows vs whole arrays."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pysages_b200 import _native as N
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
n = 100_000
frames, L, _ = bench.device_hoomd_frames(n, 1, dev, seed=3)
snaps = bench.snapshots_of(frames, L, 0.005, with_tags=False)
# cfg1-like: two dihedrals of a solute in 1e5 atoms (8 selected atoms), well-tempered metadynamics on a grid
pi = float(np.pi)
cvs = [("DihedralAngle", ([4, 6, 8, 14],), {}), ("DihedralAngle", ([6, 8, 14, 16],), {})]
md = dict(height=1.2, sigma=[0.35, 0.35], stride=500, ngaussians=64, deltaT=5000.0, kB=0.0083144626, grid=((-pi, -pi), (pi, pi), (50, 50), "periodic"))
abf = bench.workload_specs("abf2d", n, L)
small_abf = ([("Distance", ([list(range(0, 50)), list(range(50, 100))],), {}), ("Distance", ([list(range(100, 150)), list(range(150, 200))],), {})], abf[1])
for label, (c, m) in (("metad, 2 dihedrals (8 atoms)", (cvs, ("Metadynamics", md))), ("ABF, 2 distances (200 atoms)", small_abf)):
    for compact in ("1", "0"):
        os.environ["PSB_HOST_COMPACT"] = compact
        method, native = bench.build_ours(c, m, snaps[0])
        rate, h2d, d2h = bench.time_e2e(native, frames[0], L, 0.005, 300, 5, stream)
        print(f"{label}, N=1e5, selected rows={compact}: {rate:.0f} steps/s = {1e6 / rate:.1f} us/step, {h2d} B up, {d2h} B down", flush=True)
        del native, method
