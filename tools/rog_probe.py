import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
natoms = 1_000_000
frames, L, _ = bench.device_hoomd_frames(natoms, 8, dev, seed=20240 + natoms % 971)
rog = [("RadiusOfGyration", (list(range(natoms)),), {})]
md = dict(height=1.0, sigma=[L * L / 10.0], stride=500, ngaussians=64, grid=((0.0,), (4.0 * L * L,), (256,), "regular"))
for with_tags in (True, False):
    snaps = bench.snapshots_of(frames, L, 0.005, with_tags)
    method, native = bench.build_ours(rog, ("Metadynamics", md), snaps[0])
    ms, _ = bench.time_cycle(native, bench.snapdesc_array(native, snaps), 300, 20, stream)
    print("RoG+metad 1e6 tags", with_tags, "grid", native.launch_info()["grid_blocks"], f"{ms*1e3/300:.2f} us/step", flush=True)
    del native, method
