"""OpenMM-CUDA layout, slot-ordered class (ids is slot -> atom): grid of one vs two CTAs per SM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
for wl, natoms in [(c.split(":")[0], int(c.split(":")[1])) for c in sys.argv[1:]]:
  for with_tags in (False, True):
    snaps, L = bench.device_layout_snapshots("lammps", np.float64, natoms, 4, dev, seed=5, with_tags=with_tags)
    cvs, m = bench.workload_specs(wl, natoms, L)
    method, native = bench.build_ours(cvs, m, snaps[0], layout="lammps")
    descs = bench.snapdesc_array(native, snaps)
    ms, _ = bench.time_cycle(native, descs, 300, 20, stream, plain=bool(os.environ.get("PSB_NO_GRAPH")))
    info = native.launch_info()
    if os.environ.get("PSB_DEBUG_TIMING"):
        torch.cuda.synchronize()
        tt = native.view("debug_timing", (2, -1, 16)).cpu().numpy().astype(np.float64)
        b2 = 1 if tt[0, :, 0].min() < tt[1, :, 0].min() else 0
        t = tt[b2]; t0 = t[:, 0].min()
        names = {1: "passA", 2: "barrier", 3: "finalize", 5: "scatter"}
        print("   " + " | ".join(f"{names[i]} {((t[:, i] - t0) / 1e3).min():.1f}/{np.median((t[:, i] - t0) / 1e3):.1f}/{((t[:, i] - t0) / 1e3).max():.1f}" for i in (1, 2, 3, 5)))
    print(f"lammps {wl} tags={with_tags} N={natoms} grid {info['grid_blocks']}: {ms * 1e3 / 300:.2f} us/step = {info['algorithmic_bytes_per_step'] / (ms / 300 * 1e-3) / 1e9:.0f} GB/s", flush=True)
    del native, method, snaps
