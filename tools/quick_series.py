"""Short timing of a few workloads (us/step, CUDA-graph replay): python tools/quick_series.py [name:natoms ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
cases = [c.split(":") for c in sys.argv[1:]] or [("harmonic", "1000000"), ("abf2d", "1000000"), ("abf2d", "100000"), ("harmonic", "100000")]
for wl, n in cases:
    n = int(n)
    nfr = 64 if n <= 100_000 else 8
    frames, L, _ = bench.device_hoomd_frames(n, nfr, dev, seed=3)
    snaps = bench.snapshots_of(frames, L, 0.005)
    cvs, m = bench.workload_specs(wl, n, L)
    method, native = bench.build_ours(cvs, m, snaps[0])
    descs = bench.snapdesc_array(native, snaps)
    steps = 2000 if n <= 100_000 else 400
    best = min(bench.time_cycle(native, descs, steps, 20, stream)[0] for _ in range(3))
    info = native.launch_info()
    us = best * 1e3 / steps
    print(f"{wl:9s} N={n:8d} grid={info['grid_blocks']:4d} {us:8.2f} us/step  {info['algorithmic_bytes_per_step']/us/1e3:8.1f} GB/s algorithmic")
    del frames, snaps, native, method; torch.cuda.empty_cache()
