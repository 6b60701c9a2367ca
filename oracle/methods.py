"""
Oracle (test infrastructure): sampling-method state machines.

Restates, in torch-fp64 (CPU):
  pysages/methods/harmonic_bias.py:78-132   HarmonicBias
  pysages/methods/metad.py:156-323          Metadynamics (standard / well-tempered, direct / grid)
  pysages/methods/abf.py:142-258            ABF
  pysages/methods/funn.py:97-296            FUNN (per-step accumulation, network force, cadenced refit via oracle/ml.py)
  pysages/methods/restraints.py:49-73       CVRestraints
  pysages/methods/core.py:463-474           generalize
  pysages/utils/core.py:113-136             gaussian, linear_solver
Every `update(snapshot, state)` is a pure function like the reference's jitted one.
Out-of-bounds indexing follows JAX's defaults: gathers clamp, scatters drop.
"""

from typing import Any, NamedTuple

import numpy as np
import scipy.linalg
import torch

from . import colvars, grids

F64 = torch.float64


# --------------------------------------------------------------------------- #
# restraints.py
# --------------------------------------------------------------------------- #


class CVRestraints(NamedTuple):  # restraints.py:14-37
    lower: Any
    upper: Any
    kl: Any
    ku: Any


def canonicalize(restraints, cvs):
    # restraints.py:49-61
    if restraints is None:
        return None
    lower, upper, kl, ku = restraints
    kl = np.asarray(kl, dtype=np.float64).flatten()
    ku = np.asarray(ku, dtype=np.float64).flatten()
    lower = np.where(kl == 0, -np.inf, np.asarray(lower, dtype=np.float64).flatten())
    upper = np.where(ku == 0, np.inf, np.asarray(upper, dtype=np.float64).flatten())
    return CVRestraints(lower, upper, kl, ku)


def apply_restraints(lo, hi, kl, kh, xi):
    # restraints.py:68-73
    lo, hi, kl, kh = (torch.as_tensor(v, dtype=F64) for v in (lo, hi, kl, kh))
    zero = torch.zeros_like(xi)
    return torch.where(xi < lo, kl * (xi - lo), torch.where(xi > hi, kh * (xi - hi), zero))


# --------------------------------------------------------------------------- #
# JAX indexing semantics
# --------------------------------------------------------------------------- #


def _gather(arr, I):
    """arr[I] with every index clamped into range (JAX default gather mode)."""
    idx = tuple(min(int(i), arr.shape[d] - 1) for d, i in enumerate(I))
    return arr[idx]


def _in_bounds(arr, I):
    return all(int(i) < arr.shape[d] for d, i in enumerate(I))


def _scatter_add(arr, I, val):
    """arr.at[I].add(val): dropped when any index is out of range (JAX default scatter mode)."""
    out = arr.clone()
    if _in_bounds(arr, I):
        idx = tuple(int(i) for i in I)
        out[idx] = out[idx] + val
    return out


def _scatter_set(arr, i, val):
    out = arr.clone()
    if int(i) < arr.shape[0]:
        out[int(i)] = val
    return out


# --------------------------------------------------------------------------- #
# base  (methods/core.py:81-153, 463-474)
# --------------------------------------------------------------------------- #


class SamplingMethod:
    snapshot_flags = set()

    def __init__(self, cvs, **kwargs):
        self.cvs = cvs
        self.cv = colvars.build(*cvs, differentiate=kwargs.get("cv_grad", True))
        self.requires_box_unwrapping = any(cv.requires_box_unwrapping for cv in cvs)
        self.kwargs = kwargs


def generalize(concrete_update, helpers):
    # core.py:463-474
    def update(snapshot, state):
        return concrete_update(state, helpers.query(snapshot))

    return update


def _natoms(snapshot):
    return np.shape(snapshot.positions)[0]


# --------------------------------------------------------------------------- #
# HarmonicBias
# --------------------------------------------------------------------------- #


class HarmonicBiasState(NamedTuple):  # harmonic_bias.py:23-39
    xi: Any
    bias: Any
    ncalls: int = 0


class HarmonicBias(SamplingMethod):
    snapshot_flags = {"positions", "indices"}  # bias.py:32

    def __init__(self, cvs, kspring, center, **kwargs):
        super().__init__(cvs, **kwargs)
        N = len(cvs)
        center = np.asarray(center, dtype=np.float64)
        if center.shape == ():
            center = center.reshape(1)
        if len(center.shape) != 1 or center.shape[0] != N:  # bias.py:62-65
            raise RuntimeError(f"Invalid center shape expected {N} got {center.shape}.")
        self.center = center
        kspring = np.asarray(kspring, dtype=np.float64)  # harmonic_bias.py:88-107
        shape = kspring.shape
        if len(shape) > 2:
            raise RuntimeError(f"Wrong kspring shape {shape} (expected scalar, 1D or 2D)")
        if len(shape) == 2:
            if shape != (N, N):
                raise RuntimeError(f"2D kspring with wrong shape, expected ({N}, {N}), got {shape}")
            if not np.allclose(kspring, kspring.T):
                raise RuntimeError("Spring matrix is not symmetric")
            self.kspring = kspring
        else:
            if kspring.size not in (N, 1):
                raise RuntimeError(f"Wrong kspring size, expected 1 or {N}, got {kspring.size}.")
            self.kspring = np.identity(N) * kspring

    def build(self, snapshot, helpers):
        # harmonic_bias.py:113-132
        cv = self.cv
        center = torch.as_tensor(self.center)
        kspring = torch.as_tensor(self.kspring)
        natoms = _natoms(snapshot)

        def initialize():
            xi, _ = cv(helpers.query(snapshot))
            bias = torch.zeros((natoms, helpers.dimensionality()), dtype=F64)
            return HarmonicBiasState(xi, bias)

        def update(state, data):
            xi, Jxi = cv(data)
            forces = kspring @ (xi - center).flatten()  # no periodic wrap (quirk kept)
            bias = -Jxi.T.to(forces.dtype) @ forces.flatten()
            bias = bias.reshape(state.bias.shape)
            return HarmonicBiasState(xi, bias, state.ncalls + 1)

        return snapshot, initialize, generalize(update, helpers)

    def bias_energy(self, xi):
        """Not computed by the reference; defined as 1/2 (xi-c)^T K (xi-c)."""
        d = (torch.as_tensor(xi, dtype=F64).flatten() - torch.as_tensor(self.center)).flatten()
        return float(0.5 * d @ (torch.as_tensor(self.kspring) @ d))


# --------------------------------------------------------------------------- #
# Metadynamics
# --------------------------------------------------------------------------- #


class MetadynamicsState(NamedTuple):  # metad.py:26-69
    xi: Any
    bias: Any
    heights: Any
    centers: Any
    sigmas: Any
    grid_potential: Any
    grid_gradient: Any
    idx: int
    ncalls: int = 0


def gaussian(a, sigma, x):
    # utils/core.py:113-117
    return a * torch.exp(-torch.sum(((x / sigma) ** 2).reshape(x.shape[0], -1), dim=1) / 2)


def sum_of_gaussians(xi, heights, centers, sigmas, periods):
    # metad.py:318-323
    delta_x = colvars.wrap(xi - centers, periods)
    return gaussian(heights, sigmas, delta_x).sum()


class Metadynamics(SamplingMethod):
    snapshot_flags = {"positions", "indices"}

    def __init__(self, cvs, height, sigma, stride, ngaussians, deltaT=None, **kwargs):
        if deltaT is not None and "kB" not in kwargs:  # metad.py:133-138
            raise KeyError("For well-tempered Metadynamics a keyword argument `kB` must be provided.")
        super().__init__(cvs, **kwargs)
        self.grid = kwargs.get("grid", None)
        self.restraints = canonicalize(kwargs.get("restraints", None), cvs)
        if self.grid is not None and len(cvs) != self.grid.shape.size:  # core.py:435-438
            raise ValueError("Grid and Collective Variable dimensions must match.")
        self.height = height
        self.sigma = sigma
        self.stride = stride
        self.ngaussians = ngaussians
        self.deltaT = deltaT
        self.kB = kwargs.get("kB", None)

    def build(self, snapshot, helpers):
        # metad.py:156-205
        method = self
        cv = self.cv
        stride, ngaussians = self.stride, self.ngaussians
        natoms = _natoms(snapshot)
        periods = colvars.get_periods(self.cvs)
        height_0, deltaT, grid, kB = self.height, self.deltaT, self.grid, self.kB
        restraints = self.restraints

        if grid is not None:
            grid_mesh = torch.as_tensor(grids.compute_mesh(grid))
            get_grid_index = grids.build_indexer(grid)
            gshape = tuple(int(s) for s in grid.shape)

        def out_of_bounds(I):
            return any(int(i) == s for i, s in zip(I, gshape))

        def initialize():  # metad.py:166-185
            xi, _ = cv(helpers.query(snapshot))
            bias = torch.zeros((natoms, helpers.dimensionality()), dtype=F64)
            heights = torch.zeros(ngaussians, dtype=F64)
            centers = torch.zeros((ngaussians, xi.numel()), dtype=F64)
            sigmas = torch.as_tensor(np.array(method.sigma, dtype=np.float64, ndmin=2))
            if grid is None:
                gp = gg = None
            else:
                gp = torch.zeros(gshape, dtype=F64) if deltaT else None
                gg = torch.zeros((*gshape, len(gshape)), dtype=F64)
            return MetadynamicsState(xi, bias, heights, centers, sigmas, gp, gg, 0)

        def next_height(xi, heights, centers, sigmas, gp, I):  # metad.py:219-229
            if deltaT is None:
                return torch.as_tensor(float(height_0), dtype=F64)
            if grid is None:
                V = sum_of_gaussians(xi.to(F64), heights, centers, sigmas, periods)
            else:
                V = _gather(gp, I)
            return height_0 * torch.exp(-V / (deltaT * kB))

        def update_grids(gp, gg, height, xi, sigmas):  # metad.py:251-256
            mesh = grid_mesh.clone().requires_grad_(True)
            # sum over mesh points of the (independent) per-point Gaussians: its gradient
            # w.r.t. the mesh is exactly vmap(grad(current_gaussian))(grid_mesh)
            delta = colvars.wrap(mesh - xi.detach().reshape(1, -1).to(F64), periods)
            vals = gaussian(height, sigmas, delta)
            (grads,) = torch.autograd.grad(vals.sum(), mesh)
            vals = vals.detach()
            new_gp = gp + vals.reshape(gp.shape) if deltaT is not None else gp
            new_gg = gg + grads.reshape(gg.shape)
            return new_gp, new_gg

        def update(state, data):  # metad.py:187-203
            xi, Jxi = cv(data)
            ncalls = state.ncalls + 1
            in_deposition_step = (ncalls > 1) and (ncalls % stride == 1)

            heights, centers, sigmas = state.heights, state.centers, state.sigmas
            gp, gg, idx = state.grid_potential, state.grid_gradient, state.idx
            I_xi = get_grid_index(xi.detach().numpy()) if grid is not None else None
            should = in_deposition_step and (grid is None or not out_of_bounds(I_xi))
            if should:  # metad.py:262-271
                h = next_height(xi, heights, centers, sigmas, gp, I_xi)
                heights = _scatter_set(heights, idx, h)
                centers = _scatter_set(centers, idx, xi.flatten().to(F64))
                if grid is not None:
                    gp, gg = update_grids(gp, gg, h, xi, sigmas)
                idx = idx + 1

            # metad.py:282-314
            if grid is None:
                x = xi.detach().clone().to(F64).requires_grad_(True)
                V = sum_of_gaussians(x, heights, centers, sigmas, periods)
                (generalized_force,) = torch.autograd.grad(V, x)
            elif out_of_bounds(I_xi):
                if restraints:
                    lo, hi, kl, kh = restraints
                    generalized_force = apply_restraints(
                        lo, hi, kl, kh, xi.reshape(len(gshape)).to(F64)
                    )
                else:
                    generalized_force = torch.zeros(len(gshape), dtype=F64)
            else:
                generalized_force = _gather(gg, I_xi)

            bias = -Jxi.T.to(F64) @ generalized_force.flatten()
            bias = bias.reshape(state.bias.shape)
            return MetadynamicsState(xi, bias, heights, centers, sigmas, gp, gg, idx, ncalls)

        return snapshot, initialize, generalize(update, helpers)

    def bias_energy(self, state):
        """V(xi) of the state (direct sum, or the binned grid potential when available)."""
        periods = colvars.get_periods(self.cvs)
        if self.grid is None or state.grid_potential is None:
            return float(
                sum_of_gaussians(
                    state.xi.to(F64), state.heights, state.centers, state.sigmas, periods
                )
            )
        I = grids.build_indexer(self.grid)(state.xi.detach().numpy())
        return float(_gather(state.grid_potential, I))


# --------------------------------------------------------------------------- #
# ABF (+ FUNN accumulation)
# --------------------------------------------------------------------------- #


class ABFState(NamedTuple):  # abf.py:32-72
    xi: Any
    bias: Any
    hist: Any
    Fsum: Any
    force: Any
    Wp: Any
    Wp_: Any
    ncalls: int = 0


def linear_solver(use_pinv):
    # utils/core.py:120-136 + utils/compat.py:105-111
    if use_pinv:

        def tsolve(A, B):
            return torch.linalg.pinv(A.T) @ B

    else:

        def tsolve(A, B):
            a = (A @ A.T).numpy()
            b = (A @ B).numpy()
            return torch.as_tensor(scipy.linalg.solve(a, b, assume_a="pos"))

    return tsolve


class ABF(SamplingMethod):
    snapshot_flags = {"positions", "indices", "momenta"}  # abf.py:114

    def __init__(self, cvs, grid, **kwargs):
        if len(cvs) != grid.shape.size:
            raise ValueError("Grid and Collective Variable dimensions must match.")
        super().__init__(cvs, **kwargs)
        self.grid = grid
        self.restraints = canonicalize(kwargs.get("restraints", None), cvs)
        self.N = np.asarray(kwargs.get("N", 500))
        self.use_pinv = kwargs.get("use_pinv", False)

    def build(self, snapshot, helpers):
        # abf.py:142-227
        cv, grid = self.cv, self.grid
        dt = snapshot.dt
        dims = grid.shape.size
        gshape = tuple(int(s) for s in grid.shape)
        natoms = _natoms(snapshot)
        tsolve = linear_solver(self.use_pinv)
        get_grid_index = grids.build_indexer(grid)
        N = int(self.N)
        restraints = self.restraints
        query, dimensionality, to_force_units = helpers

        def estimate_force(xi, I_xi, Fsum, hist):  # abf.py:230-258
            ob = any(int(i) == s for i, s in zip(I_xi, gshape))
            if restraints is not None and ob:
                lo, hi, kl, kh = restraints
                return apply_restraints(lo, hi, kl, kh, xi.reshape(dims).to(F64))
            return _gather(Fsum, I_xi) / max(N, int(_gather(hist, I_xi)))

        def initialize():  # abf.py:173-190
            xi, _ = cv(query(snapshot))
            bias = torch.zeros((natoms, dimensionality()), dtype=F64)
            hist = torch.zeros(gshape, dtype=torch.int64)  # uint32 in the reference
            Fsum = torch.zeros((*gshape, dims), dtype=F64)
            force = torch.zeros(dims, dtype=F64)
            return ABFState(xi, bias, hist, Fsum, force, torch.zeros(dims, dtype=F64),
                            torch.zeros(dims, dtype=F64))

        def update(state, data):  # abf.py:192-225
            xi, Jxi = cv(data)
            p = torch.as_tensor(data.momenta)
            Jd = Jxi.to(F64) if (Jxi.dtype == F64 or p.dtype == F64) else Jxi
            Wp = tsolve(Jd, p.to(Jd.dtype)).to(F64)
            dWp_dt = (1.5 * Wp - 2.0 * state.Wp + 0.5 * state.Wp_) / dt
            I_xi = get_grid_index(xi.detach().numpy())
            hist = _scatter_add(state.hist, I_xi, 1)
            Fsum = _scatter_add(state.Fsum, I_xi, to_force_units(dWp_dt) + state.force)
            force = estimate_force(xi, I_xi, Fsum, hist).reshape(dims)
            bias = (-Jxi.T.to(F64) @ force).reshape(state.bias.shape)
            return ABFState(xi, bias, hist, Fsum, force, Wp, state.Wp, state.ncalls + 1)

        return snapshot, initialize, generalize(update, helpers)


# --------------------------------------------------------------------------- #
# FUNN (funn.py:97-296)
# --------------------------------------------------------------------------- #


class FUNNState(NamedTuple):  # funn.py:41-86
    xi: Any
    bias: Any
    hist: Any
    Fsum: Any
    F: Any
    Wp: Any
    Wp_: Any
    nn: Any
    ncalls: int = 0


class FUNN(ABF):
    """funn.py:97-210: ABF accumulation every step; after 2 * train_freq calls the force is
    std * mlp(scale(f32(xi))) + mean, the network being refitted every train_freq calls to the
    Blackman-smoothed, normalized binned mean force."""

    def __init__(self, cvs, grid, topology, **kwargs):
        super().__init__(cvs, grid, **kwargs)
        from . import ml

        self.topology = tuple(topology)
        self.train_freq = kwargs.get("train_freq", 5000)
        dims = grid.shape.size
        self.widths = ml.layer_sizes(dims, self.topology, dims)
        self.initial_params = kwargs.get("initial_params", None)
        if self.initial_params is None:
            self.initial_params = ml.init_params(self.widths, kwargs.get("seed", 0))
        self.reg = kwargs.get("reg", 1e-6)  # funn.py:153
        self.max_iters = kwargs.get("max_iters", 500)

    def build(self, snapshot, helpers):
        from . import ml

        cv, grid = self.cv, self.grid
        dt = snapshot.dt
        dims = grid.shape.size
        gshape = tuple(int(s) for s in grid.shape)
        natoms = _natoms(snapshot)
        tsolve = linear_solver(self.use_pinv)
        get_grid_index = grids.build_indexer(grid)
        N = int(self.N)
        restraints = self.restraints
        train_freq = self.train_freq
        widths = self.widths
        lower, upper = np.asarray(grid.lower, dtype=np.float64), np.asarray(grid.upper, dtype=np.float64)
        inputs = grids.compute_mesh(grid)
        kernel = ml.blackman_kernel(dims, 7)
        padding = "wrap" if grid.is_periodic else "edge"
        query, dimensionality, to_force_units = helpers

        def learn(state):  # funn.py:213-249
            hist = state.hist.numpy().astype(np.float64)[..., None]
            y = state.Fsum.numpy() / np.maximum(hist, 1)
            axes = tuple(range(y.ndim - 1))
            y, mean, std = ml.normalize(y, axes=axes)
            reference = ml.smooth(y, kernel, padding)
            params = ml.lm_fit(state.nn.params, widths, inputs, reference, lower, upper,
                               reg=self.reg, max_iters=self.max_iters)
            return ml.NNData(params, mean, std / reference.std(axis=axes))

        def estimate_force(xi, I_xi, Fsum, hist, nn, pred):  # funn.py:252-296
            ob = any(int(i) == s for i, s in zip(I_xi, gshape))
            if restraints is not None and ob:
                lo, hi, kl, kh = restraints
                return apply_restraints(lo, hi, kl, kh, xi.reshape(dims).to(F64))
            if pred:
                x = xi.detach().numpy().astype(np.float32).astype(np.float64)
                out = ml.mlp_apply(nn.params, widths, x, lower, upper).reshape(dims)
                return torch.as_tensor(nn.std * out + nn.mean)
            return _gather(Fsum, I_xi) / max(N, int(_gather(hist, I_xi)))

        def initialize():  # funn.py:175-185
            xi, _ = cv(query(snapshot))
            bias = torch.zeros((natoms, dimensionality()), dtype=F64)
            hist = torch.zeros(gshape, dtype=torch.int64)
            Fsum = torch.zeros((*gshape, dims), dtype=F64)
            z = torch.zeros(dims, dtype=F64)
            nn = ml.NNData(np.asarray(self.initial_params, dtype=np.float64), np.zeros(dims), np.zeros(dims))
            return FUNNState(xi, bias, hist, Fsum, z, z.clone(), z.clone(), nn)

        def update(state, data):  # funn.py:187-210
            ncalls = state.ncalls + 1
            in_training_regime = ncalls > 2 * train_freq
            in_training_step = in_training_regime and (ncalls % train_freq == 1)
            nn = learn(state) if in_training_step else state.nn
            xi, Jxi = cv(data)
            p = torch.as_tensor(data.momenta)
            Jd = Jxi.to(F64) if (Jxi.dtype == F64 or p.dtype == F64) else Jxi
            Wp = tsolve(Jd, p.to(Jd.dtype)).to(F64)
            dWp_dt = (1.5 * Wp - 2.0 * state.Wp + 0.5 * state.Wp_) / dt
            I_xi = get_grid_index(xi.detach().numpy())
            hist = _scatter_add(state.hist, I_xi, 1)
            Fsum = _scatter_add(state.Fsum, I_xi, to_force_units(dWp_dt) + state.F)
            F = estimate_force(xi, I_xi, Fsum, hist, nn, in_training_regime).reshape(dims)
            bias = (-Jxi.T.to(F64) @ F).reshape(state.bias.shape)
            return FUNNState(xi, bias, hist, Fsum, F, Wp, state.Wp, nn, ncalls)

        return snapshot, initialize, generalize(update, helpers)


# --------------------------------------------------------------------------- #
# SpectralABF (spectral_abf.py:95-260)
# --------------------------------------------------------------------------- #


class SpectralABFState(NamedTuple):  # spectral_abf.py:36-84
    xi: Any
    bias: Any
    hist: Any
    Fsum: Any
    force: Any
    Wp: Any
    Wp_: Any
    fun: Any
    ncalls: int = 0


class SpectralABF(ABF):
    """spectral_abf.py:95-260: ABF accumulation; past `fit_threshold` calls the force is the gradient of
    a Fourier (periodic grid) or monomial (grid converted to Chebyshev) series refitted every `fit_freq`
    calls by one pseudo-inverse product.  Outside the grid: restraints, or ZERO force without them."""

    def __init__(self, cvs, grid, **kwargs):
        super().__init__(cvs, grid, **kwargs)
        from . import approxfun

        self.fit_freq = kwargs.get("fit_freq", 100)
        self.fit_threshold = kwargs.get("fit_threshold", 500)
        if not self.grid.is_periodic:  # spectral_abf.py:140: convert(grid, Grid[Chebyshev])
            self.grid = grids.Grid(self.grid.lower, self.grid.upper, self.grid.shape, kind=grids.Chebyshev)
        self.exponents = approxfun.collect_exponents(self.grid)
        self.pinv = approxfun.gradient_pinv(self.grid, self.exponents)

    def build(self, snapshot, helpers):
        from . import approxfun

        cv, grid = self.cv, self.grid
        dt = snapshot.dt
        dims = grid.shape.size
        gshape = tuple(int(s) for s in grid.shape)
        natoms = _natoms(snapshot)
        tsolve = linear_solver(self.use_pinv)
        get_grid_index = grids.build_indexer(grid)
        N = int(self.N)
        restraints = self.restraints
        fit_freq, fit_threshold = self.fit_freq, self.fit_threshold
        query, dimensionality, to_force_units = helpers

        def fit(dy):
            return approxfun.fit_gradient(grid, self.pinv, dy)

        def estimate_force(xi, I_xi, Fsum, hist, fun, pred):  # spectral_abf.py:236-268
            ob = any(int(i) == s for i, s in zip(I_xi, gshape))
            if ob:
                if restraints is None:
                    return torch.zeros(dims, dtype=F64)
                lo, hi, kl, kh = restraints
                return apply_restraints(lo, hi, kl, kh, xi.reshape(dims).to(F64))
            if pred:
                g = approxfun.get_gradient(grid, self.exponents, fun, xi.detach().numpy().astype(np.float64))
                return torch.as_tensor(g.reshape(dims))
            return _gather(Fsum, I_xi) / max(int(_gather(hist, I_xi)), N)

        def initialize():  # spectral_abf.py:169-179
            xi, _ = cv(query(snapshot))
            bias = torch.zeros((natoms, dimensionality()), dtype=F64)
            hist = torch.zeros(gshape, dtype=torch.int64)
            Fsum = torch.zeros((*gshape, dims), dtype=F64)
            z = torch.zeros(dims, dtype=F64)
            return SpectralABFState(xi, bias, hist, Fsum, z, z.clone(), z.clone(), fit(Fsum.numpy()))

        def update(state, data):  # spectral_abf.py:181-205
            ncalls = state.ncalls + 1
            in_fitting_regime = ncalls > fit_threshold
            in_fitting_step = in_fitting_regime and (ncalls % fit_freq == 1)
            fun = state.fun
            if in_fitting_step:  # spectral_abf.py:216-219
                hist = state.hist.numpy().astype(np.float64).reshape(*gshape, 1)
                fun = fit(state.Fsum.numpy() / np.maximum(hist, 1))
            xi, Jxi = cv(data)
            p = torch.as_tensor(data.momenta)
            Jd = Jxi.to(F64) if (Jxi.dtype == F64 or p.dtype == F64) else Jxi
            Wp = tsolve(Jd, p.to(Jd.dtype)).to(F64)
            dWp_dt = (1.5 * Wp - 2.0 * state.Wp + 0.5 * state.Wp_) / dt
            I_xi = get_grid_index(xi.detach().numpy())
            hist = _scatter_add(state.hist, I_xi, 1)
            Fsum = _scatter_add(state.Fsum, I_xi, to_force_units(dWp_dt) + state.force)
            force = estimate_force(xi, I_xi, Fsum, hist, fun, in_fitting_regime).reshape(dims)
            bias = (-Jxi.T.to(F64) @ force).reshape(state.bias.shape)
            return SpectralABFState(xi, bias, hist, Fsum, force, Wp, state.Wp, fun, ncalls)

        return snapshot, initialize, generalize(update, helpers)


# --------------------------------------------------------------------------- #
# ANN (ann.py:65-258)
# --------------------------------------------------------------------------- #


class ANNState(NamedTuple):  # ann.py:38-62
    xi: Any
    bias: Any
    hist: Any
    phi: Any
    prob: Any
    nn: Any
    ncalls: int = 0


class ANN(SamplingMethod):
    """ann.py:65-258: histogram of visits; every train_freq calls a network is fitted to the (smoothed,
    normalized) free energy estimate kT log(prob); the bias force is std * d net / d xi inside the grid once
    ncalls > train_freq, zero otherwise."""

    snapshot_flags = {"positions", "indices"}

    def __init__(self, cvs, grid, topology, kT, **kwargs):
        if len(cvs) != grid.shape.size:
            raise ValueError("Grid and Collective Variable dimensions must match.")
        super().__init__(cvs, **kwargs)
        from . import ml

        self.grid = grid
        self.topology = tuple(topology)
        self.kT = kT
        self.train_freq = kwargs.get("train_freq", 5000)
        dims = grid.shape.size
        self.widths = ml.layer_sizes(dims, self.topology, 1)
        self.initial_params = kwargs.get("initial_params", None)
        if self.initial_params is None:
            self.initial_params = ml.init_params(self.widths, kwargs.get("seed", 0))
        self.reg = kwargs.get("reg", 1e-6)
        self.max_iters = kwargs.get("max_iters", 500)

    def build(self, snapshot, helpers):
        from . import ml

        cv, grid, kT = self.cv, self.grid, self.kT
        dims = grid.shape.size
        gshape = tuple(int(s) for s in grid.shape)
        shape = gshape if dims > 1 else (*gshape, 1)
        natoms = _natoms(snapshot)
        get_grid_index = grids.build_indexer(grid)
        train_freq, widths = self.train_freq, self.widths
        lower, upper = np.asarray(grid.lower, dtype=np.float64), np.asarray(grid.upper, dtype=np.float64)
        inputs = grids.compute_mesh(grid)
        kernel = ml.blackman_kernel(dims, 7)
        padding = "wrap" if grid.is_periodic else "edge"
        query, dimensionality, _ = helpers

        def smooth(y):  # ann.py:189: conv if dims > 1 else vmap(conv)(y.T).T
            if dims > 1:
                return ml.convolve(y, kernel, padding)
            return ml.convolve(y[:, 0], kernel, padding)[:, None]

        def learn(state):  # ann.py:194-211
            hist = state.hist.numpy().astype(np.float64).reshape(shape)
            prob = state.prob + hist * np.exp(state.phi / kT)
            phi = kT * np.log(prob)
            y, mean, std = ml.normalize(phi)
            reference = smooth(y)
            params = ml.lm_fit(state.nn.params, widths, inputs, reference, lower, upper, reg=self.reg,
                               max_iters=self.max_iters)
            nn = ml.NNData(params, mean, std / reference.std())
            phi = nn.std * ml.mlp_apply(params, widths, inputs, lower, upper).reshape(phi.shape)
            phi = phi - phi.min()
            return torch.zeros_like(state.hist), prob, phi, nn

        def initialize():  # ann.py:143-151
            xi, _ = cv(query(snapshot))
            bias = torch.zeros((natoms, dimensionality()), dtype=F64)
            hist = torch.zeros(gshape, dtype=torch.int64)
            nn = ml.NNData(np.asarray(self.initial_params, dtype=np.float64), 0.0, 1.0)
            return ANNState(xi, bias, hist, np.zeros(shape), np.ones(shape), nn)

        def update(state, data):  # ann.py:153-167
            ncalls = state.ncalls + 1
            in_training_regime = ncalls > train_freq
            in_training_step = in_training_regime and (ncalls % train_freq == 1)
            hist, prob, phi, nn = learn(state) if in_training_step else (state.hist, state.prob, state.phi, state.nn)
            xi, Jxi = cv(data)
            I_xi = get_grid_index(xi.detach().numpy())
            hist = _scatter_add(hist, I_xi, 1)
            in_bounds = all(int(i) != s for i, s in zip(I_xi, gshape))
            if in_training_regime and in_bounds:  # ann.py:243-256
                x = xi.detach().numpy().astype(np.float32).astype(np.float64)
                F = torch.as_tensor(nn.std * ml.mlp_input_gradient(nn.params, widths, x, lower, upper))
            else:
                F = torch.zeros(dims, dtype=F64)
            bias = (-Jxi.T.to(F64) @ F).reshape(state.bias.shape)
            return ANNState(xi, bias, hist, phi, prob, nn, ncalls)

        return snapshot, initialize, generalize(update, helpers)


# --------------------------------------------------------------------------- #
# CFF (cff.py:109-350)
# --------------------------------------------------------------------------- #


class CFFState(NamedTuple):  # cff.py:36-98
    xi: Any
    bias: Any
    hist: Any
    histp: Any
    prob: Any
    fe: Any
    Fsum: Any
    force: Any
    Wp: Any
    Wp_: Any
    nn: Any
    fnn: Any
    ncalls: int = 0


class CFF(ABF):
    """cff.py:109-350: ABF accumulation plus a second histogram; every train_freq calls an energy network is
    fitted (Sobolev loss) to the frequency-based free energy and the binned mean force, a force network to the
    mean force; past train_freq calls the bias force is fnn.std * fmodel(f32(xi)) + fnn.mean."""

    def __init__(self, cvs, grid, topology, kT, **kwargs):
        super().__init__(cvs, grid, **kwargs)
        from . import ml

        self.topology = tuple(topology)
        self.kT = kT
        self.train_freq = kwargs.get("train_freq", 5000)
        dims = grid.shape.size
        self.widths = ml.layer_sizes(dims, self.topology, 1)
        self.fwidths = ml.layer_sizes(dims, self.topology, dims)
        seed = kwargs.get("seed", 0)
        self.initial_params = ml.init_params(self.widths, seed)
        self.initial_fparams = ml.init_params(self.fwidths, seed)
        self.reg = kwargs.get("reg", 1e-4)  # cff.py:163
        self.max_iters = kwargs.get("max_iters", 1000)  # cff.py:164
        self.fmax_iters = kwargs.get("fmax_iters", 500)  # cff.py:165 (LevenbergMarquardt default)

    def build(self, snapshot, helpers):
        from . import ml

        cv, grid, kT = self.cv, self.grid, self.kT
        dt = snapshot.dt
        dims = grid.shape.size
        gshape = tuple(int(s) for s in grid.shape)
        shape = gshape if dims > 1 else (*gshape, 1)
        natoms = _natoms(snapshot)
        tsolve = linear_solver(self.use_pinv)
        get_grid_index = grids.build_indexer(grid)
        N = int(self.N)
        restraints = self.restraints
        train_freq = self.train_freq
        lower, upper = np.asarray(grid.lower, dtype=np.float64), np.asarray(grid.upper, dtype=np.float64)
        inputs = grids.compute_mesh(grid).astype(np.float32).astype(np.float64)  # cff.py:263: f32(compute_mesh)
        kernel = ml.blackman_kernel(dims, 7).astype(np.float32).astype(np.float64)
        padding = "wrap" if grid.is_periodic else "edge"
        f32 = lambda a: np.asarray(a).astype(np.float32).astype(np.float64)
        query, dimensionality, to_force_units = helpers

        def vsmooth(y):  # vmap(conv)(y.T).T
            return np.stack([ml.convolve(y[..., c], kernel, padding) for c in range(y.shape[-1])], axis=-1)

        def smooth(y):
            return ml.convolve(y, kernel, padding) if dims > 1 else vsmooth(y)

        def learn(state):  # cff.py:295-307
            histp = state.histp.numpy().astype(np.float64).reshape(shape)
            prob = state.prob + histp * np.exp(state.fe / kT)
            fe = kT * np.log(np.maximum(1, prob))
            hist = state.hist.numpy().astype(np.float64).reshape(*gshape, 1)
            force = state.Fsum.numpy() / np.maximum(1, hist)
            # preprocess, cff.py:280-286
            axes = tuple(range(force.ndim - 1))
            dy, dy_mean, dy_std = ml.normalize(force, axes=axes)
            s = max(fe.std(), dy_std.max())
            y = smooth(f32(fe / s))
            dy = vsmooth(f32(dy * dy_std / s))
            params = ml.lm_fit_sobolev(state.nn.params, self.widths, inputs, y, dy, lower, upper, reg=self.reg,
                                       max_iters=self.max_iters)
            fparams = ml.lm_fit(state.fnn.params, self.fwidths, inputs, dy, lower, upper, reg=self.reg,
                                max_iters=self.fmax_iters)
            nn, fnn = ml.NNData(params, state.nn.mean, s), ml.NNData(fparams, dy_mean, s)
            fe = nn.std * ml.mlp_apply(params, self.widths, inputs, lower, upper).reshape(fe.shape)
            return torch.zeros_like(state.histp), prob, fe - fe.min(), nn, fnn

        def estimate_force(xi, I_xi, Fsum, hist, fnn, pred):  # cff.py:315-350
            ob = any(int(i) == s for i, s in zip(I_xi, gshape))
            if restraints is not None and ob:
                lo, hi, kl, kh = restraints
                return apply_restraints(lo, hi, kl, kh, xi.reshape(dims).to(F64))
            if pred:
                x = xi.detach().numpy().astype(np.float32).astype(np.float64)
                out = ml.mlp_apply(fnn.params, self.fwidths, x, lower, upper).reshape(dims)
                return torch.as_tensor(fnn.std * out + fnn.mean)
            return _gather(Fsum, I_xi) / max(N, int(_gather(hist, I_xi)))

        def initialize():  # cff.py:197-217
            xi, _ = cv(query(snapshot))
            bias = torch.zeros((natoms, dimensionality()), dtype=F64)
            hist = torch.zeros(gshape, dtype=torch.int64)
            histp = torch.zeros(gshape, dtype=torch.int64)
            Fsum = torch.zeros((*gshape, dims), dtype=F64)
            z = torch.zeros(dims, dtype=F64)
            nn = ml.NNData(np.asarray(self.initial_params, dtype=np.float64), 0.0, 1.0)
            fnn = ml.NNData(np.asarray(self.initial_fparams, dtype=np.float64), np.zeros(dims), 1.0)
            return CFFState(xi, bias, hist, histp, np.zeros(shape), np.zeros(shape), Fsum, z, z.clone(), z.clone(),
                            nn, fnn)

        def update(state, data):  # cff.py:223-244
            ncalls = state.ncalls + 1
            in_training_regime = ncalls > train_freq
            in_training_step = in_training_regime and (ncalls % train_freq == 1)
            if in_training_step:
                histp, prob, fe, nn, fnn = learn(state)
            else:
                histp, prob, fe, nn, fnn = state.histp, state.prob, state.fe, state.nn, state.fnn
            xi, Jxi = cv(data)
            p = torch.as_tensor(data.momenta)
            Jd = Jxi.to(F64) if (Jxi.dtype == F64 or p.dtype == F64) else Jxi
            Wp = tsolve(Jd, p.to(Jd.dtype)).to(F64)
            dWp_dt = (1.5 * Wp - 2.0 * state.Wp + 0.5 * state.Wp_) / dt
            I_xi = get_grid_index(xi.detach().numpy())
            hist = _scatter_add(state.hist, I_xi, 1)
            Fsum = _scatter_add(state.Fsum, I_xi, to_force_units(dWp_dt) + state.force)
            histp = _scatter_add(histp, I_xi, 1)
            force = estimate_force(xi, I_xi, Fsum, hist, fnn, in_training_regime).reshape(dims)
            bias = (-Jxi.T.to(F64) @ force).reshape(state.bias.shape)
            return CFFState(xi, bias, hist, histp, prob, fe, Fsum, force, Wp, state.Wp, nn, fnn, ncalls)

        return snapshot, initialize, generalize(update, helpers)


# --------------------------------------------------------------------------- #
# Sirens (sirens.py:95-420)
# --------------------------------------------------------------------------- #


class SirensState(NamedTuple):  # sirens.py:40-92
    xi: Any
    bias: Any
    hist: Any
    histp: Any
    prob: Any
    fe: Any
    Fsum: Any
    force: Any
    Wp: Any
    Wp_: Any
    nn: Any
    ncalls: int = 0


class Sirens(ABF):
    """sirens.py:95-420: ABF accumulation (+ histp / prob / fe in mode "cff"); every train_freq calls a Siren
    network of the free energy is fitted to the binned mean force (mode "abf": gradients loss) or to force and
    frequency-based free energy (mode "cff": Sobolev loss); past train_freq calls the bias force is
    nn.std * d net / d xi at f32(xi)."""

    def __init__(self, cvs, grid, topology, **kwargs):
        mode = kwargs.get("mode", "abf")
        kT = kwargs.get("kT", None)
        if mode not in ("abf", "cff"):
            raise ValueError(f"Invalid mode {mode}. Possible values are 'abf' and 'cff'")
        if mode == "cff" and kT is None:
            raise KeyError("When running in 'cff' mode, a keyword argument `kT` in the same "
                           "units as the backend internal energy units must be provided.")
        super().__init__(cvs, grid, **kwargs)
        from . import ml

        self.mode, self.kT = mode, kT
        self.topology = tuple(topology)
        self.train_freq = kwargs.get("train_freq", 5000)
        self.omega_0, self.omega = kwargs.get("omega_0", 16), kwargs.get("omega", 1)  # ml/models.py:95
        self.widths = ml.layer_sizes(grid.shape.size, self.topology, 1)
        self.initial_params = ml.init_siren_params(self.widths, self.omega, kwargs.get("seed", 0))
        self.reg = kwargs.get("reg", 1e-4)
        self.max_iters = kwargs.get("max_iters", 250)

    def build(self, snapshot, helpers):
        from . import ml

        cv, grid, kT, mode = self.cv, self.grid, self.kT, self.mode
        dt = snapshot.dt
        dims = grid.shape.size
        gshape = tuple(int(s) for s in grid.shape)
        shape = gshape if dims > 1 else (*gshape, 1)
        natoms = _natoms(snapshot)
        tsolve = linear_solver(self.use_pinv)
        get_grid_index = grids.build_indexer(grid)
        N = int(self.N)
        restraints = self.restraints
        train_freq, widths = self.train_freq, self.widths
        lower, upper = np.asarray(grid.lower, dtype=np.float64), np.asarray(grid.upper, dtype=np.float64)
        f32 = lambda a: np.asarray(a).astype(np.float32).astype(np.float64)
        inputs = f32(grids.compute_mesh(grid))
        kernel = f32(ml.blackman_kernel(dims, 5))  # sirens.py:291
        padding = "wrap" if grid.is_periodic else "edge"
        net = dict(activation="sin", omega0=float(self.omega_0), omega=float(self.omega))
        query, dimensionality, to_force_units = helpers

        def vsmooth(y):
            return np.stack([ml.convolve(y[..., c], kernel, padding) for c in range(y.shape[-1])], axis=-1)

        def smooth(y):
            return ml.convolve(y, kernel, padding) if dims > 1 else vsmooth(y)

        def normalize_gradients(data):  # sirens.py:399-417
            axes = tuple(range(data.ndim - 1))
            if grid.is_periodic:
                return data - data.mean(axis=axes), data.std(axis=axes).max()
            return data, data.std(axis=axes).max()

        def learn(state):  # sirens.py:340-347
            hist = state.hist.numpy().astype(np.float64).reshape(*gshape, 1)
            force = state.Fsum.numpy() / np.maximum(1, hist)
            histp, prob, fe = state.histp, state.prob, state.fe
            if mode == "cff":
                prob = state.prob + state.histp.numpy().astype(np.float64).reshape(shape) * np.exp(state.fe / kT)
                fe = kT * np.log(np.maximum(1, prob))
                histp = torch.zeros_like(state.histp)
                dy, dy_std = normalize_gradients(force)
                s = max(fe.std(), dy_std)
                params = ml.lm_fit_autograd(state.nn.params, widths, inputs, (smooth(f32(fe / s)), dy / s), "sobolev",
                                            lower, upper, reg=self.reg, max_iters=self.max_iters, **net)
            else:
                dy, s = normalize_gradients(force)
                params = ml.lm_fit_autograd(state.nn.params, widths, inputs, dy / s, "gradients", lower, upper,
                                            reg=self.reg, max_iters=self.max_iters, **net)
            nn = ml.NNData(params, state.nn.mean, s)
            if mode == "cff":
                fe = nn.std * ml.mlp_apply(params, widths, inputs, lower, upper, "sin", net["omega0"], net["omega"]
                                           ).reshape(fe.shape)
                fe = fe - fe.min()
            return histp, prob, fe, nn

        def estimate_force(xi, I_xi, Fsum, hist, nn, pred):  # sirens.py:350-396
            ob = any(int(i) == s for i, s in zip(I_xi, gshape))
            if restraints is not None and ob:
                lo, hi, kl, kh = restraints
                return apply_restraints(lo, hi, kl, kh, xi.reshape(dims).to(F64))
            if pred:
                x = xi.detach().numpy().astype(np.float32).astype(np.float64)
                return torch.as_tensor(nn.std * ml.net_input_gradient(nn.params, widths, x, lower, upper, **net))
            return _gather(Fsum, I_xi) / max(N, int(_gather(hist, I_xi)))

        def initialize():  # sirens.py:229-251
            xi, _ = cv(query(snapshot))
            bias = torch.zeros((natoms, dimensionality()), dtype=F64)
            hist = torch.zeros(gshape, dtype=torch.int64)
            Fsum = torch.zeros((*gshape, dims), dtype=F64)
            z = torch.zeros(dims, dtype=F64)
            nn = ml.NNData(np.asarray(self.initial_params, dtype=np.float64), 0.0, 1.0)
            if mode == "cff":
                histp, prob, fe = torch.zeros(gshape, dtype=torch.int64), np.zeros(shape), np.zeros(shape)
            else:
                histp = prob = fe = None
            return SirensState(xi, bias, hist, histp, prob, fe, Fsum, z, z.clone(), z.clone(), nn)

        def update(state, data):  # sirens.py:253-275
            ncalls = state.ncalls + 1
            in_training_regime = ncalls > train_freq
            in_training_step = in_training_regime and (ncalls % train_freq == 1)
            histp, prob, fe, nn = learn(state) if in_training_step else (state.histp, state.prob, state.fe, state.nn)
            xi, Jxi = cv(data)
            p = torch.as_tensor(data.momenta)
            Jd = Jxi.to(F64) if (Jxi.dtype == F64 or p.dtype == F64) else Jxi
            Wp = tsolve(Jd, p.to(Jd.dtype)).to(F64)
            dWp_dt = (1.5 * Wp - 2.0 * state.Wp + 0.5 * state.Wp_) / dt
            I_xi = get_grid_index(xi.detach().numpy())
            hist = _scatter_add(state.hist, I_xi, 1)
            Fsum = _scatter_add(state.Fsum, I_xi, to_force_units(dWp_dt) + state.force)
            if mode == "cff":
                histp = _scatter_add(histp, I_xi, 1)
            force = estimate_force(xi, I_xi, Fsum, hist, nn, in_training_regime).reshape(dims)
            bias = (-Jxi.T.to(F64) @ force).reshape(state.bias.shape)
            return SirensState(xi, bias, hist, histp, prob, fe, Fsum, force, Wp, state.Wp, nn, ncalls)

        return snapshot, initialize, generalize(update, helpers)
