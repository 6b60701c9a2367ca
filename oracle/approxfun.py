"""
Oracle (test infrastructure): Fourier / monomial series fits of pysages/approxfun/core.py, in numpy.

  approxfun/core.py:73-96     collect_exponents
  approxfun/core.py:99-136    scale, unit_mesh (Chebyshev- or uniformly distributed)
  approxfun/core.py:170-201   vandergrad_builder (rows: point-major, then CV dimension)
  approxfun/core.py:204-211   pinv of the gradient Vandermonde matrix ([imag, real] blocks when periodic)
  approxfun/core.py:232-258   build_fitter (SpectralGradientFit)
  approxfun/core.py:311-336   build_grad_evaluator

jnp.linalg.pinv's default cut-off (10 * max(M, N) * eps) is passed explicitly to numpy.
Parity unpinned: no reference test holds known answers for these functions.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""

from itertools import product

import numpy as np

from . import grids


def collect_exponents(grid):
    if grid.shape.size > 2:
        raise ValueError("Only 1D and 2D grids are supported")
    if grid.is_periodic:
        r = grid.shape.sum() / 4
    else:
        r = np.square(grid.shape).sum() / 16
    n = int(np.floor(np.sqrt(r))) + 1
    r = float(r)
    if grid.shape.size == 1:
        return np.arange(1, n).reshape(-1, 1)
    m = 1 - n if grid.is_periodic else 1
    pairs = product(range(1, n), range(m, n))
    return np.vstack([
        [(0, i) for i in range(1, n)],
        [(i, 0) for i in range(1, n)],
        [t for t in pairs if t[1] != 0 and t[0] ** 2 + t[1] ** 2 <= r],
    ])


def scale(x, grid):
    return (x - grid.lower) * 2 / grid.size - 1


def unit_mesh(grid):
    axes = []
    for n in grid.shape:
        k = np.arange(int(n), dtype=np.float64)
        if grid.kind is grids.Chebyshev:
            axes.append(-np.cos((k + 0.5) * np.pi / n))
        else:
            axes.append(-1 + (2 * k + 1) / n)
    coords = np.meshgrid(*axes, indexing="ij")
    return np.stack([c.reshape(-1) for c in coords], axis=-1)


def vandergrad(grid, ns, xs):
    """(points * dims, K) complex (periodic) or real matrix."""
    s = 2 * (np.pi if grid.is_periodic else 1) / grid.size
    ns = np.asarray(ns)
    rows = []
    for x in np.asarray(xs, dtype=np.float64).reshape(-1, grid.shape.size):
        if grid.is_periodic:
            z = np.exp(-1j * np.pi * ns * x)
            e = s * ns * z
            other = z
        else:
            z = x ** np.maximum(ns - 1, 0)
            e = s * ns * z
            other = x ** ns
        if grid.shape.size > 1:
            e = e * np.fliplr(other)
        rows.append(e.T)
    return np.concatenate(rows, axis=0).reshape(-1, ns.shape[0])


def gradient_pinv(grid, ns=None, mesh=None):
    ns = collect_exponents(grid) if ns is None else ns
    mesh = unit_mesh(grid) if mesh is None else mesh
    A = vandergrad(grid, ns, mesh)
    if grid.is_periodic:
        A = np.hstack((np.imag(A), np.real(A)))
    return np.linalg.pinv(A, rcond=10 * max(A.shape) * np.finfo(np.float64).eps)


class Fun:
    def __init__(self, scale, coefficients, c0=0.0):
        self.scale, self.coefficients, self.c0 = scale, coefficients, c0


def fit_gradient(grid, pinv, dy):
    axes = tuple(range(grid.shape.size))
    std = dy.std(axis=axes).max()
    std = 1.0 if std == 0 else std
    dy = (dy - dy.mean(axis=axes)) / std if grid.is_periodic else dy / std
    return Fun(std, pinv @ dy.flatten(), 0.0)


def get_gradient(grid, ns, fun, x):
    x = scale(np.asarray(x, dtype=np.float64).reshape(-1, grid.shape.size), grid)
    V = vandergrad(grid, ns, x)
    if grid.is_periodic:
        V = np.hstack((np.imag(V), np.real(V)))
    return (fun.scale * (V @ fun.coefficients)).reshape(x.shape)
