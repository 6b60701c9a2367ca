"""1e6-row slot-ordered kernels with / without HOOMD's slot -> tag array: us per step (graph replay) and phase stamps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
cases = [(c.split(":")[0], int(c.split(":")[1])) for c in sys.argv[1:]] or [("harmonic", 1_000_000), ("abf2d", 1_000_000)]
for wl, natoms in cases:
    frames, L, _ = bench.device_hoomd_frames(natoms, 8 if natoms > 200_000 else 64, dev, seed=1)
    for with_tags in (False, True):
        snaps = bench.snapshots_of(frames, L, 0.005, with_tags)
        cvs, m = bench.workload_specs(wl, natoms, L)
        method, native = bench.build_ours(cvs, m, snaps[0])
        descs = bench.snapdesc_array(native, snaps)
        ms, _ = bench.time_cycle(native, descs, 400, 20, stream, plain=bool(os.environ.get("PSB_NO_GRAPH")))
        torch.cuda.synchronize()
        info = native.launch_info()
        line = f"{wl} tags={with_tags} wants={native.wants_tags} grid {info['grid_blocks']}: {ms * 1e3 / 400:.2f} us/step = {info['algorithmic_bytes_per_step'] / (ms / 400 * 1e-3) / 1e9:.0f} GB/s"
        if os.environ.get("PSB_DEBUG_TIMING"):
            tt = native.view("debug_timing", (2, -1, 16)).cpu().numpy().astype(np.float64)
            b2 = 1 if tt[0, :, 0].min() < tt[1, :, 0].min() else 0
            t = tt[b2]; t0 = t[:, 0].min()
            names = {7: "pass0", 1: "passA", 2: "barrier", 3: "finalize", 5: "scatter"}
            line += " | " + " | ".join(f"{names[i]} {((t[:, i] - t0) / 1e3).min():.1f}/{np.median((t[:, i] - t0) / 1e3):.1f}/{((t[:, i] - t0) / 1e3).max():.1f}" for i in ((1, 2, 3, 5) if with_tags else (7, 1, 2, 3, 5)))
            c = t[:, 8:16]
            line += "\n      finalizer clocks (us after slot 8, median over CTAs; slots 9..15): " + " ".join(f"{np.median((c[:, i] - c[:, 0]) / 1965.0):.2f}" for i in range(1, 8))
        print(line, flush=True)
        del native, method, snaps
