"""Per-step GPU time (CUDA events around each psb_step) for a workload: finds steps that are slower than others."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pysages_b200 import _native as N
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4grid"
frames, L, cvs, md, prefill = bench.cfg4_case(dev)
if wl == "cfg4grid":
    md = dict(md, grid=((0.0, 0.0), (40.0, 4000.0), (256, 256), "regular"))
snaps = bench.snapshots_of(frames, L, 0.005)
method, native = bench.build_ours(cvs, ("Metadynamics", md), snaps[0])
if wl == "cfg4":
    prefill(native)
descs = bench.snapdesc_array(native, snaps)
sp = ctypes.c_void_p(stream.cuda_stream)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
with torch.cuda.stream(stream):
    for i in range(10):
        N.check(native.lib.psb_step(native.handle, ctypes.byref(descs[i % 16]), i, sp))
    ev[0].record(stream)
    for i in range(40):
        N.check(native.lib.psb_step(native.handle, ctypes.byref(descs[i % 16]), i, sp))
        ev[i + 1].record(stream)
stream.synchronize()
print(wl, "per-step us:", " ".join(f"{ev[i].elapsed_time(ev[i+1])*1e3:.0f}" for i in range(40)))
print("xi:", native.view("xi").cpu().numpy().ravel()[:2])
