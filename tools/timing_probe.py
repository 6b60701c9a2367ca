"""Per-phase timestamps of the fused step kernel (PSB_DEBUG_TIMING=1), for latency work.
   python tools/timing_probe.py            (on the GPU box)"""
import os, sys
os.environ.setdefault("PSB_DEBUG_TIMING", "1")  # "graph": keep CUDA-graph replay + PDL while stamping
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench

dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
cases = [tuple([c.split(":")[0], int(c.split(":")[1])]) for c in sys.argv[1:]] or [("abf2d", 100_000), ("harmonic", 10_000), ("harmonic", 100_000), ("harmonic", 1_000_000), ("abf2d", 1_000_000), ("walkers", 100_000), ("window", 100_000)]
for wl, natoms in cases:
    if wl.startswith("cfg1"):
        frames, L, _ = bench.device_hoomd_frames(natoms, 8, dev, seed=1, dtype=torch.float64)
        snaps = bench.snapshots_of(frames, L, 0.002)
        pi = float(np.pi)
        cvs = [("DihedralAngle", ([4, 6, 8, 14],), {}), ("DihedralAngle", ([6, 8, 14, 16],), {})]
        md = dict(height=1.2, sigma=[0.35, 0.35], stride=500, ngaussians=64, deltaT=5000.0, kB=0.0083144626)
        if wl == "cfg1grid":
            md["grid"] = ((-pi, -pi), (pi, pi), (50, 50), "periodic")
        method, native = bench.build_ours(cvs, ("Metadynamics", md), snaps[0])
    elif wl.startswith("cfg4"):
        frames, L, cvs, md, prefill = bench.cfg4_case(dev)
        if wl == "cfg4grid":
            md = dict(md, grid=((0.0, 0.0), (40.0, 4000.0), (256, 256), "regular"))
        m = ("Metadynamics", md)
        snaps = bench.snapshots_of(frames, L, 0.005)
        method, native = bench.build_ours(cvs, m, snaps[0])
        if wl == "cfg4":
            prefill(native)
    else:
        frames, L, _ = bench.device_hoomd_frames(natoms, 8, dev, seed=1)
        snaps = bench.snapshots_of(frames, L, 0.005)
        cvs, m = bench.workload_specs(wl, natoms, L)
        method, native = bench.build_ours(cvs, m, snaps[0])
    descs = bench.snapdesc_array(native, snaps)
    ms, _ = bench.time_cycle(native, descs, 200, 20, stream)
    torch.cuda.synchronize()
    tt = native.view("debug_timing", (2, -1, 16)).cpu().numpy().astype(np.float64)
    # the two halves hold the last two launches; order them by start time
    a, b2 = (0, 1) if tt[0, :, 0].min() < tt[1, :, 0].min() else (1, 0)
    prev, t = tt[a], tt[b2]
    g = t.shape[0]
    t0 = t[:, 0].min()
    rel = lambda x: (x - t0) / 1e3
    print(f"{wl} N={natoms} grid={g} us/step={ms*1e3/200:.2f}   gap (prev kernel last stamp -> this kernel first stamp) = {(t0 - prev[:, 6].max())/1e3:.2f} us;"
          f" period by stamps = {(t0 - prev[:, 0].min())/1e3:.2f} us")
    names = ["start", "passA done", "barrier passed", "finalize done", "scatter start", "scatter end", "kernel end", "pass0 done (slot-major)"]
    for i, nme in enumerate(names):
        if i == 7 and t[:, 7].max() < t0:
            continue
        col = rel(t[:, i])
        print(f"  {nme:15s}: min {col.min():6.2f} median {np.median(col):6.2f} max {col.max():6.2f}")
    pe = prev[:, 6].max()
    print("  relative to the END of the previous kernel (max over CTAs): " + ", ".join(
        f"{nme}={np.median(t[:, i] - pe) / 1e3:.2f}" for i, nme in enumerate(names[:7])))
    c = t[0, 8:16]
    cn = ["pre-reduce", "reduce_rows", "finalize_cv", "JJt/Jp", "solve", "grid_bin", "hist gather+F", "finish_force(end post_cv)"]
    base = c[0]
    if os.environ.get("PSB_PROBE_PASSA"):
        t0c = t[0, 8]
        print("  CTA0 pass-A clocks (cycles since model copy done): segment+prologue=%d, loads+accumulate=%d, block reduce=%d" % (c[1] - t0c, c[2] - t0c, c[3] - t0c))
    elif wl == "abf2d":
        print("  CTA0 finalize clocks (cycles since pre-reduce):", ", ".join(f"{n}={int(v - base)}" for n, v in zip(cn, c)))
    else:
        print("  CTA0 finalize clocks:", f"reduce_rows={int(c[1]-base)} finalize_cv={int(c[2]-base)} post_cv={int(c[7]-base)}")
    del frames, snaps, native, method
    torch.cuda.empty_cache()
