"""One short device LM fit (for ncu launch lists): python tools/refit_one.py H1,H2 GRID ITERS"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from pysages_b200 import ml
topology = tuple(int(t) for t in sys.argv[1].split(","))
g, iters = int(sys.argv[2]), int(sys.argv[3])
lower, upper = -np.pi * np.ones(2), np.pi * np.ones(2)
ax = np.linspace(-np.pi, np.pi, g)
x = np.stack(np.meshgrid(ax, ax, indexing="ij"), -1).reshape(-1, 2)
y = np.stack([np.sin(x[:, d]) * (1 + 0.3 * np.cos(x[:, 1 - d])) for d in range(2)], -1)
model = ml.MLP(2, 2, topology, lower, upper, seed=0)
dev = torch.device("cuda")
xt, yt, p0 = (torch.as_tensor(a, device=dev) for a in (x, y, model.parameters))
fit = ml.build_fitting_function(model, ml.LevenbergMarquardt(reg=ml.L2Regularization(1e-6), max_iters=iters))
p = fit(p0, xt, yt)
torch.cuda.synchronize()
print(model.parameters.size, fit.workspace.result())
