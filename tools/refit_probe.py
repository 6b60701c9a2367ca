"""Time the device Levenberg-Marquardt refit (psb_lm_fit) against the torch loop on FUNN-sized problems."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pysages_b200 import ml

def run(shape, topology, iters):
    dims = len(shape)
    rng = np.random.default_rng(0)
    lower, upper = -np.pi * np.ones(dims), np.pi * np.ones(dims)
    axes = [np.linspace(lower[d], upper[d], s) for d, s in enumerate(shape)]
    x = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, dims)
    y = np.stack([np.sin(x[:, d]) * (1 + 0.3 * np.cos(x[:, (d + 1) % dims])) + 0.05 * rng.standard_normal(len(x)) for d in range(dims)], -1)
    model = ml.MLP(dims, dims, topology, lower, upper, seed=0)
    dev = torch.device("cuda")
    xt, yt, p0 = (torch.as_tensor(a, device=dev) for a in (x, y, model.parameters))
    opt = ml.LevenbergMarquardt(reg=ml.L2Regularization(1e-6), max_iters=iters)
    out = {}
    for name, fit in (("native", ml.build_fitting_function(model, opt)), ("torch", ml.build_fitting_function(model, opt, force_host=True))):
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            p = fit(p0, xt, yt)
            t_enq = time.perf_counter() - t0
            torch.cuda.synchronize(); t1 = time.perf_counter() - t0
        e = model.apply(p, xt) - yt
        out[name] = (t_enq * 1e3, t1 * 1e3, float((e * e).sum() / 2))
        if name == "native":
            out["ctl"] = fit.workspace.result()
    print(f"grid {shape} topology {topology} max_iters {iters} P={model.parameters.size}: "
          f"native enqueue {out['native'][0]:.2f} ms total {out['native'][1]:.2f} ms (sse {out['native'][2]:.4f}; iters/cost/mu/loops {out['ctl']}) | "
          f"torch {out['torch'][1]:.1f} ms (sse {out['torch'][2]:.4f})", flush=True)

run((64, 64), (14,), 500)      # funn.py docs example scale
run((64, 64), (14, 14), 500)
run((32, 32), (8, 8), 50)      # the harness' ANN / FUNN parity cases
run((128,), (8, 8), 500)
run((50, 50), (30, 20), 100)
