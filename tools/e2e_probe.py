"""psb_step_host on the headline workload: torch-pinned vs pageable (registered by the library) host arrays."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pysages_b200 import _native as N

dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
natoms = 100_000
frames, L, _ = bench.device_hoomd_frames(natoms, 2, dev, seed=3)
snaps = bench.snapshots_of(frames, L, 0.005)
cvs, m = bench.workload_specs("abf2d", natoms, L)
for kind in ("torch pin_memory()", "numpy (pageable) + library cudaHostRegister"):
    method, native = bench.build_ours(cvs, m, snaps[0])
    if kind.startswith("torch"):
        host = {k: v.cpu().pin_memory() for k, v in frames[0].items()}
        ptr = {k: v.data_ptr() for k, v in host.items()}
    else:
        host = {k: np.ascontiguousarray(v.cpu().numpy()).copy() for k, v in frames[0].items()}
        ptr = {k: v.ctypes.data for k, v in host.items()}
    d = N.SnapshotDesc()
    d.positions, d.vel_mass, d.forces, d.ids, d.images = ptr["positions"], ptr["vel_mass"], ptr["forces"], ptr["rtags"], ptr["images"]
    for i in range(3):
        d.box_diag[i] = L
    d.dt = 0.005
    h2d, d2h = ctypes.c_int64(), ctypes.c_int64()
    sp = ctypes.c_void_p(stream.cuda_stream)
    for _ in range(5):
        N.check(native.lib.psb_step_host(native.handle, ctypes.byref(d), 0, sp, ctypes.byref(h2d), ctypes.byref(d2h)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    n = 300
    for i in range(n):
        N.check(native.lib.psb_step_host(native.handle, ctypes.byref(d), i, sp, ctypes.byref(h2d), ctypes.byref(d2h)))
    e1.record(stream); stream.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"{kind:48s} {us:7.1f} us/step  {1e6/us:7.0f} steps/s  {(h2d.value + d2h.value) / us / 1e3:5.1f} GB/s effective")
    N.check(native.lib.psb_host_release(native.handle))
    del native, method
