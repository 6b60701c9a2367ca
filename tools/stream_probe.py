"""Experiment: clock stamps inside the streaming gather (build with PSB_NVCC_EXTRA=-DPSB_PROBE_STREAM) + arithmetic peaks."""
import os, sys, ctypes
os.environ.setdefault("PSB_DEBUG_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pysages_b200 import _native as N

lib = N.load()
out = (ctypes.c_double * 5)()
N.check(lib.psb_fma_peak(out))
sms = torch.cuda.get_device_properties(0).multi_processor_count
for name, v in zip(("DFMA", "FFMA", "F2F f32->f64 (pair up/down)", "I2F int->f64 (pair)", "DADD"), out):
    print(f"{name:30s} {v/1e12:8.3f} Tops/s  = {v/sms/1.965e9:6.1f} per clk per SM at 1965 MHz")
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
    t = tt[b2]
    t0 = t[:, 0].min()
    names = {0: "start", 7: "pass0 done", 8: "gather start", 9: "first chunk landed", 10: "chunks done", 11: "block reduce done",
             1: "passA done", 2: "barrier passed", 3: "finalize done", 5: "scatter end", 6: "kernel end"}
    print(wl, natoms, "us/step", ms * 1e3 / 200, "grid", t.shape[0], native.launch_info())
    for i in (0, 7, 8, 9, 10, 11, 1, 2, 3, 5, 6):
        col = (t[:, i] - t0) / 1e3
        print(f"   {names[i]:20s} min {col.min():7.2f} median {np.median(col):7.2f} max {col.max():7.2f}")
    del frames, snaps, native, method
    torch.cuda.empty_cache()
