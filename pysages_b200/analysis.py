"""
Free energies from the results of biased runs (pysages/methods/analysis.py:44-175 and the per-method `analyze`
functions: abf.py:261-306, funn.py:300-345, cff.py:353-430, sirens.py:419-500, spectral_abf.py:271-350).

* ABF / FUNN -- "gradient learning": a multilayer perceptron A(xi) (grid dimensions -> 1) whose INPUT-GRADIENT is
  fitted to the binned mean force, in two Levenberg-Marquardt stages (analysis.py:90-146): 250 iterations on the
  Blackman-smoothed, max-normalised forces, then up to 1000 on the raw ones starting from there; free energy
  = max A - A on the mesh, and `fes_fn` interpolates it anywhere in the grid's domain.
* CFF / Sirens: the network the run itself trained (`state.nn`) is the free energy; CFF also reports `state.fe`.
* SpectralABF: the fitted Fourier / monomial series evaluated as a VALUE (approxfun/core.py:149-166, 282-308).

Cadenced, off the per-step path: everything runs on the device the states live on with torch fp64 (`ml.py`).
Arrays mapped to the grid are returned transposed along the grid axes, like the reference (`grid_transposer`,
grids.py:161-178: "for convenience when plotting").
"""

import numpy as np
import torch

from . import ml
from .grids import compute_mesh

F64 = torch.float64


def grid_transposer(grid):
    """grids.py:161-178: reshape to (*grid.shape, m), reverse the grid axes, squeeze."""
    d = len(grid.shape)
    shape = tuple(int(s) for s in grid.shape)
    n = int(np.prod(shape))

    def transpose(array):
        a = np.asarray(array)
        m = a.size // n
        return a.reshape(*shape, m).transpose(*reversed(range(d)), d).squeeze()

    return transpose


def average_forces(hist, Fsum):
    hist = np.asarray(hist, dtype=np.float64)
    return np.asarray(Fsum) / np.maximum(hist.reshape(*np.asarray(Fsum).shape[:-1], 1), 1)


def _device_of(state):
    return state.hist.device if isinstance(state.hist, torch.Tensor) else torch.device("cpu")


def _as_tensor(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=F64)
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=device)


def gradient_learning_fes(grid, hist, Fsum, topology, device=None, pre_iters=250, iters=1000, seed=0):
    """analysis.py:90-146 for one state -> (fes_fn, NNData)."""
    device = torch.device(device) if device is not None else (Fsum.device if isinstance(Fsum, torch.Tensor) else torch.device("cpu"))
    dims = int(grid.shape.size)
    shape = tuple(int(s) for s in grid.shape)
    inputs = _as_tensor(compute_mesh(grid), device)
    model = ml.MLP(dims, 1, tuple(topology), grid.lower, grid.upper, seed=seed)
    reg = ml.L2Regularization(1e-4)
    pre_fit = ml.build_fitting_function(model, ml.LevenbergMarquardt(loss=ml.GradientsSSE(), reg=reg, max_iters=pre_iters))
    fit = ml.build_fitting_function(model, ml.LevenbergMarquardt(loss=ml.GradientsSSE(), reg=reg, max_iters=iters))
    kernel = _as_tensor(ml.blackman_kernel(dims, 7), device).to(torch.float32)
    boundary = "wrap" if grid.is_periodic else "edge"

    h = _as_tensor(hist, device).reshape(*shape, 1)
    F = _as_tensor(Fsum, device).reshape(*shape, dims) / torch.clamp(h, min=1.0)
    s = F.abs().max()
    s = torch.where(s == 0, torch.ones_like(s), s)
    F = F / s  # analysis.py:131-133: scale the mean forces

    def smooth(data):
        # every force component smoothed over the grid axes in float32, as funn.py:226 / ann.py:189 do.  (The
        # reference's analysis.py:111-117 vmaps the same convolution over the LEADING grid axis of the data instead,
        # which for a 2-D grid convolves (component, axis-1) slabs with the 2-D kernel; the stage only seeds the fit
        # of the raw data below, and the run is parity-unpinned anyway -- JAX's initial weights cannot be reproduced.)
        return ml.smooth(data.to(torch.float32), kernel, boundary).to(F64)

    p0 = torch.as_tensor(model.parameters, dtype=F64, device=device)
    p1 = pre_fit(p0, inputs, smooth(F).reshape(-1, dims))
    p2 = fit(p1, inputs, F.reshape(-1, dims))
    nn = ml.NNData(p2, torch.zeros((), dtype=F64, device=device), s)
    return network_fes_fn(model, nn), nn


def network_fes_fn(model, nn):
    """x -> max A - A with A = std * model(x) + mean (analysis.py:139-144, cff.py:392-398, sirens.py:467-473)."""
    params = nn.params if isinstance(nn.params, torch.Tensor) else torch.as_tensor(np.asarray(nn.params), dtype=F64)
    params = params.to(F64)

    def fes_fn(x):
        x = _as_tensor(np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x, dtype=np.float64).reshape(-1, model.widths[0]), params.device)
        std = _as_tensor(nn.std, params.device)
        mean = _as_tensor(nn.mean, params.device)
        A = std * model.apply(params, x) + mean
        return (A.max() - A).cpu().numpy()

    return fes_fn


def series_value_evaluator(model):
    """approxfun/core.py:149-166, 282-308: value of the fitted series at x (unscaled CV coordinates)."""
    ns = np.asarray(model.exponents, dtype=np.float64)  # (K, d)
    grid = model.grid
    lower, size = np.asarray(grid.lower, dtype=np.float64), np.asarray(grid.size, dtype=np.float64)

    def evaluate(fun, x):
        x = np.asarray(x, dtype=np.float64).reshape(-1, ns.shape[1])
        u = (x - lower) * 2 / size - 1  # approxfun/core.py:99-104
        if grid.is_periodic:
            z = np.exp(1j * np.pi * (u[:, None, :] * ns[None]).sum(-1))  # (P, K)
            V = np.hstack((z.real, z.imag))
        else:
            V = np.prod(u[:, None, :] ** ns[None], axis=-1)
        cs = np.asarray(fun.coefficients.cpu() if hasattr(fun.coefficients, "cpu") else fun.coefficients, dtype=np.float64)
        scale = float(np.asarray(fun.scale.cpu() if hasattr(fun.scale, "cpu") else fun.scale))
        c0 = float(np.asarray(getattr(fun, "c0", 0.0)))
        return (scale * (V @ cs + c0)).reshape(x.shape[0], -1)

    return evaluate


def first_or_all(seq):
    return seq[0] if len(seq) == 1 else seq
