"""Experiment (build with PSB_NVCC_EXTRA=-DPSB_PROBE_FIN): finer stamps inside the ABF finalizer of the headline workload."""
import os, sys
os.environ.setdefault("PSB_DEBUG_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
natoms = 100_000
frames, L, _ = bench.device_hoomd_frames(natoms, 8, dev, seed=1)
snaps = bench.snapshots_of(frames, L, 0.005)
cvs, m = bench.workload_specs("abf2d", natoms, L)
method, native = bench.build_ours(cvs, m, snaps[0])
descs = bench.snapdesc_array(native, snaps)
ms, _ = bench.time_cycle(native, descs, 200, 20, stream)
torch.cuda.synchronize()
tt = native.view("debug_timing", (2, -1, 16)).cpu().numpy().astype(np.float64)
t = tt[0]
c = t[:, 8:16] - t[:, 8:9]
names = ["pre-reduce", "reduce done", "cv values done", "pair_sync(1) passed", "J p done", "W p done + helper joined", "F done", "finish_force done"]
print("us/step", ms * 1e3 / 200)
for i, n in enumerate(names):
    print(f"  {n:28s} median {np.median(c[:, i]):7.0f}  (CTA0 {c[0, i]:7.0f})  delta {np.median(c[:, i] - (c[:, i - 1] if i else 0)):6.0f}")
h = (t[:, 7] - t[:, 8]); h2 = (t[:, 5] - t[:, 8])
print(f"  helper warp: reaches pair_sync(1) at median {np.median(h):7.0f} cycles after pre-reduce, finishes bin + gathers at {np.median(h2):7.0f}")
