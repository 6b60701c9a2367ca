"""50 000 pipelined steps (CUDA graph + PDL) of the headline workload: counters and sums stay consistent."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pysages_b200 import _native as N
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
n = int(os.environ.get("PSB_LONG_ATOMS", 100_000))  # 300000: the slot-ordered class with the tag array (one CTA per SM, two-level exchange)
frames, L, _ = bench.device_hoomd_frames(n, 64, dev, seed=4)
snaps = bench.snapshots_of(frames, L, 0.005)
cvs, m = bench.workload_specs("abf2d", n, L)
method, native = bench.build_ours(cvs, m, snaps[0])
descs = bench.snapdesc_array(native, snaps)
f0 = [f["forces"].clone() for f in frames]
steps = int(os.environ.get("PSB_LONG_STEPS", 50_000))
steps -= steps % 64
N.check(native.lib.psb_run_cycle(native.handle, descs, 64, steps, ctypes.c_void_p(stream.cuda_stream)))
stream.synchronize()
hist = native.view("hist").cpu().numpy(); Fsum = native.view("Fsum").cpu().numpy()
print("ncalls", int(native.view("ncalls").item()), "hist sum", int(hist.sum()), "Fsum finite", bool(np.isfinite(Fsum).all()),
      "xi", native.view("xi").cpu().numpy().ravel()[:2])
assert int(native.view("ncalls").item()) == steps and int(hist.sum()) == steps and np.isfinite(Fsum).all()
# every frame received steps/64 bias applications; compare one frame against the same work without graph / PDL
method2, native2 = bench.build_ours(cvs, m, snaps[0])
native2.set_option("pdl", 0)
for fr, f in zip(frames, f0):
    fr["forces"].copy_(f)
for i in range(640):
    native2.step(snaps[i % 64], i)
torch.cuda.synchronize()
h2 = native2.view("hist").cpu().numpy()
method3, native3 = bench.build_ours(cvs, m, snaps[0])
for fr, f in zip(frames, f0):
    fr["forces"].copy_(f)
N.check(native3.lib.psb_run_cycle(native3.handle, descs, 64, 640, ctypes.c_void_p(stream.cuda_stream)))
stream.synchronize()
assert np.array_equal(native3.view("hist").cpu().numpy(), h2)
g2, g3 = native2.launch_info()["grid_blocks"], native3.launch_info()["grid_blocks"]
if g2 == g3:
    assert torch.equal(native3.view("Fsum"), native2.view("Fsum"))
    print("640 steps: graph+PDL == plain steps without PDL (hist, Fsum bitwise)")
else:  # dual launch shape (slot-ordered class with the tag array): other CTA count, other order of the partial sums
    a, b = native3.view("Fsum").cpu().numpy(), native2.view("Fsum").cpu().numpy()
    assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max()
    print(f"640 steps: graph+PDL ({g3} CTAs) == plain steps without PDL ({g2} CTAs): hist bitwise, Fsum to 1e-10")
