"""Launch-latency floor on the GPU box: us per back-to-back launch of an empty kernel."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysages_b200 import _native as N

lib = N.load()
stream = torch.cuda.Stream()
sp = ctypes.c_void_p(stream.cuda_stream)
def run(grid, coop, big, bar, n=4000):
    N.check(lib.psb_launch_floor(grid, coop, big, bar, 200, sp)); stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); N.check(lib.psb_launch_floor(grid, coop, big, bar, n, sp)); e1.record(stream); stream.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
for grid in (1, 148, 296):
    for coop, big, bar in ((0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1)):
        print(f"grid={grid:4d} cooperative={coop} big_params={big} grid_barrier={bar}: {run(grid, coop, big, bar):6.2f} us/launch")
out = (ctypes.c_double * 16)()
N.check(lib.psb_latency_probe(out))
names = ["DFMA", "DADD", "FFMA", "IMAD", "LDS", "sqrt(f64)", "1/x (f64)", "rsqrt(f64)", "floor+DFMA", "fmod(f64)", "atan2(f64)", "exp(f64)"]
print("dependent-chain latency, cycles/op:", ", ".join(f"{n}={out[i]:.1f}" for i, n in enumerate(names)))
