import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda:0")
natoms = 100_000
L = float((natoms / 0.8) ** (1.0 / 3.0))
cv, m = bench.workload_specs("abf2d", natoms, L)
print(bench.time_api_hook(cv, m, natoms, 5000, dev, seed=1))
import cProfile, pstats
# where does the host time go?
import time, ctypes
from pysages_b200 import _native as N
import pysages_b200 as ps
from pysages_b200 import synthetic
from pysages_b200.backends import mock
from torch.utils.dlpack import to_dlpack
snap = synthetic.make_snapshot(natoms, layout="hoomd", dtype=__import__("numpy").float32, seed=1)
eng = mock.MockEngine("hoomd", snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda")
kw = dict(m[1]); lower, upper, shape, kind = kw.pop("grid"); kw["restraints"] = ps.CVRestraints(*kw["restraints"])
method = ps.methods.ABF([getattr(ps.colvars, n)(*a, **k) for (n, a, k) in cv], ps.Grid[ps.Regular](lower, upper, shape), **kw)
sampler = mock.Sampler(eng, method)
caps = tuple(to_dlpack(eng.t[k]) for k in ("positions", "vel_mass", "rtags", "images", "forces"))
sampler.on_hoomd_half_step(*caps, 0)
nat = sampler._native
def burst(fn, n=256):
    best = 1e30
    for _ in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(n): fn(i)
        best = min(best, (time.perf_counter() - t0) * 1e6 / n)
    return best
print("full hook            %.2f us" % burst(lambda i: sampler.on_hoomd_half_step(*caps, i)))
print("step_current         %.2f us" % burst(lambda i: sampler.step_current(i)))
print("native.step_desc     %.2f us" % burst(lambda i: nat.step_desc(sampler._desc, i)))
br = ctypes.byref(sampler._desc); f = nat.lib.psb_step; h = nat.handle; rs = torch._C._cuda_getCurrentRawStream
print("raw ctypes psb_step  %.2f us" % burst(lambda i: f(h, br, i, rs(0))))
from pysages_b200.backends import _common as C
print("5 capsule reads      %.2f us" % burst(lambda i: [C.capsule_data(c) for c in caps]))
nat.set_option("pdl", 0)
print("raw psb_step, no PDL attribute  %.2f us" % burst(lambda i: f(h, br, i, rs(0))))
