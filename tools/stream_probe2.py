"""Experiment: phase stamps of the streaming kernel at 1e6 atoms under env knobs."""
import os, sys
os.environ.setdefault("PSB_DEBUG_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
for wl, natoms in (("harmonic", 1_000_000), ("abf2d", 1_000_000)):
    frames, L, _ = bench.device_hoomd_frames(natoms, 8, dev, seed=1)
    snaps = bench.snapshots_of(frames, L, 0.005)
    cvs, m = bench.workload_specs(wl, natoms, L)
    method, native = bench.build_ours(cvs, m, snaps[0])
    descs = bench.snapdesc_array(native, snaps)
    ms, _ = bench.time_cycle(native, descs, 200, 20, stream)
    torch.cuda.synchronize()
    tt = native.view("debug_timing", (2, -1, 16)).cpu().numpy().astype(np.float64)
    a, b2 = (0, 1) if tt[0, :, 0].min() < tt[1, :, 0].min() else (1, 0)
    t = tt[b2]; t0 = t[:, 0].min()
    names = {7: "pass0 done", 1: "passA done", 2: "barrier passed", 3: "finalize done", 5: "scatter end"}
    print(os.environ.get("TAG", ""), wl, "us/step %.2f" % (ms * 1e3 / 200), "grid", t.shape[0], " | ".join(
        f"{names[i]} {((t[:, i] - t0) / 1e3).min():.1f}/{np.median((t[:, i] - t0) / 1e3):.1f}/{((t[:, i] - t0) / 1e3).max():.1f}" for i in (7, 1, 2, 3, 5)))
    c = (t[:, 9:16] - t[:, 8:9]) / 1965.0  # SM cycles -> us at 1965 MHz
    cn = ["chunk0 landed", "chunk0 done", "chunk1 landed", "chunk1 done", "chunk2 landed", "chunk2 done", "all chunks done"]
    print("    warp 0 of each CTA, us after its gather start: " + " | ".join(f"{n} {c[:, i].min():.2f}/{np.median(c[:, i]):.2f}/{c[:, i].max():.2f}" for i, n in enumerate(cn)))
