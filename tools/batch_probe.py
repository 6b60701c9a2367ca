"""32 umbrella windows on one GPU: one stream per window (round 1) vs one batched launch per step."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pysages_b200 import _native as N, parallel

dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
nwin, natoms, steps = int(os.environ.get("NWIN", 32)), 100_000, 2000
natives, rows, keep = [], [], []
for w in range(nwin):
    frames, L, _ = bench.device_hoomd_frames(natoms, 4, dev, seed=20240 + 5500 + w)
    snaps = bench.snapshots_of(frames, L, 0.005)
    cvs, m = bench.workload_specs("window", natoms, L)
    mm = (m[0], dict(m[1], center=[-L / 4 + (L / 2) * w / max(nwin - 1, 1)]))
    natives.append(bench.build_ours(cvs, mm, snaps[0])[1])
    keep.append((frames, snaps))
print("grid per window", natives[0].launch_info()["grid_blocks"])
batch = parallel.WindowBatch(natives)
ring = batch.ring([[keep[w][1][r] for w in range(nwin)] for r in range(4)])
with torch.cuda.stream(stream):
    batch.run_cycle(ring, 4, 40)
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    batch.run_cycle(ring, 4, steps)
    e1.record(stream)
    stream.synchronize()
ms = e0.elapsed_time(e1)
print(f"batched: {nwin} windows x {steps} steps in {ms:.2f} ms -> {nwin * steps / ms * 1e3:,.0f} window-steps/s, {ms * 1e3 / steps:.2f} us per batched step")
os.environ["PSB_NO_GRAPH"] = "1"
