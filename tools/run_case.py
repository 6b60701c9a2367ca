"""Run one named workload for a few plain-launch steps (profiling driver): python tools/run_case.py <case> [steps]
cases: cfg3 (pair kernel), harmonic1m / abf1m (slot-ordered kernels), cfg4 / cfg4grid, abf100k"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pysages_b200 import _native as N

case = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
extra = None
if case == "cfg3":
    n = 1_000_000
    frames, L, _ = bench.device_hoomd_frames(n, 2, dev, seed=20240 + 3000)
    gen = np.random.Generator(np.random.PCG64(20240 + 3000))
    tags = gen.permutation(n)[:20_000]
    specs = ([("CoordinationNumber", ([sorted(tags[:10_000].tolist()), sorted(tags[10_000:].tolist())],), dict(r0=1.0))],
             ("HarmonicBias", dict(kspring=10.0, center=[100.0])))
elif case in ("harmonic1m", "abf1m"):
    n = 1_000_000
    frames, L, _ = bench.device_hoomd_frames(n, 4, dev, seed=3)
    specs = bench.workload_specs("harmonic" if case == "harmonic1m" else "abf2d", n, L)
elif case in ("cfg4", "cfg4grid"):
    frames, L, cvs4, md, prefill = bench.cfg4_case(dev)
    if case == "cfg4grid":
        md = dict(md, grid=((0.0, 0.0), (40.0, 4000.0), (256, 256), "regular"))
    else:
        extra = prefill
    specs = (cvs4, ("Metadynamics", md))
elif case in ("openmm100k", "lammps100k"):
    layout, dtype = ("openmm-cuda", np.float32) if case == "openmm100k" else ("lammps", np.float64)
    snaps, L = bench.device_layout_snapshots(layout, dtype, 100_000, 16, dev, seed=20240 + 2000)
    specs = bench.workload_specs("abf2d", 100_000, L)
    frames = None
else:
    n = 100_000
    frames, L, _ = bench.device_hoomd_frames(n, 8, dev, seed=3)
    specs = bench.workload_specs("abf2d", n, L)
if frames is not None:
    snaps = bench.snapshots_of(frames, L, 0.005)
layout_name = {"openmm100k": "openmm-cuda", "lammps100k": "lammps"}.get(case, "hoomd")
method, native = bench.build_ours(specs[0], specs[1], snaps[0], layout=layout_name)
if extra:
    extra(native)
descs = bench.snapdesc_array(native, snaps)
ms, launches = bench.time_cycle(native, descs, steps, 3, stream, plain=bool(os.environ.get("PSB_NO_GRAPH")))
print(case, "us/step", ms * 1e3 / steps, "launches/step", launches / steps, native.launch_info())
