"""
Cadenced neural-network work of the ABF family (FUNN): the fit that runs every `train_freq` steps.

The per-step part -- evaluating the network at xi inside the fused step kernel -- lives in
csrc/psb_step.cuh (`force_net_eval`); this module produces the parameters it consumes
(psb_set_force_net, include/pysages_b200.h).  It mirrors, with torch fp64 tensors that stay on the
method's device,

  pysages/ml/models.py:44-82          MLP(indim, outdim, topology, transform=scale)
  pysages/ml/utils.py:43-64,121-131   flat parameter layout, Blackman kernel
  pysages/ml/training.py:17-56        normalize, convolve, build_fitting_function
  pysages/ml/objectives.py            SSE loss (float32 residuals) + L2 regularisation
  pysages/ml/optimizers.py:40-51,121-131,188-232   LevenbergMarquardt

The Jacobian of the network with respect to its parameters is written out in closed form (tanh
layers) instead of being traced.  Initial weights follow stax.Dense's distributions (glorot-normal
W, N(0, 1e-2) b) from numpy's PCG64: the reference's threefry stream is not reproducible without JAX.
"""

from dataclasses import dataclass, field
from typing import NamedTuple

import numpy as np
import torch

F64 = torch.float64


class NNData(NamedTuple):  # ml/training.py:11-14
    params: torch.Tensor
    mean: torch.Tensor
    std: torch.Tensor


class L2Regularization(NamedTuple):  # ml/objectives.py
    coeff: float


class LevenbergMarquardtParams(NamedTuple):  # ml/optimizers.py:40-51
    mu_0: float = 1e-1
    mu_c: float = 10.0
    mu_min: float = 1e-8
    mu_max: float = 1e8
    rho_c: float = 1e-1
    rho_min: float = 1e-4


class SSE:  # ml/objectives.py: sum of squared errors
    pass


class Sobolev1SSE:  # ml/objectives.py: squared errors of the values AND of the input-gradients
    pass


class GradientsSSE:  # ml/objectives.py: squared errors of the input-gradients only
    pass


@dataclass
class LevenbergMarquardt:  # ml/optimizers.py:121-131
    params: LevenbergMarquardtParams = LevenbergMarquardtParams()
    loss: object = field(default_factory=SSE)
    reg: L2Regularization = L2Regularization(0.0)
    max_iters: int = 500


@dataclass
class MLP:
    """ml/models.py:44-82 with the `scale(grid)` input transform of funn.py:150-152."""

    indim: int
    outdim: int
    topology: tuple
    lower: np.ndarray
    upper: np.ndarray
    seed: int = 0
    parameters: np.ndarray = field(default=None, repr=False)
    activation: str = "tanh"  # "sin": Siren layers sin(omega * (x W + b)) (ml/models.py:84-152)
    omega_0: float = 1.0
    omega: float = 1.0

    def __post_init__(self):
        self.topology = tuple(int(t) for t in self.topology)
        self.widths = [int(self.indim), *self.topology, int(self.outdim)]
        self.lower = np.asarray(self.lower, dtype=np.float64)
        self.upper = np.asarray(self.upper, dtype=np.float64)
        if self.parameters is None:
            rng = np.random.Generator(np.random.PCG64(self.seed))
            chunks = []
            for l, (a, b) in enumerate(zip(self.widths[:-1], self.widths[1:])):
                if self.activation == "sin":  # ml/models.py:113-147 + ml/utils.py:67-100 (uniform_scaling, fan_in)
                    ws = 1.0 / a if l == 0 else np.sqrt(6.0 / self.omega**2 / a)
                    chunks.append((rng.uniform(-1.0, 1.0, (a, b)) * ws).ravel())
                    chunks.append(rng.uniform(-1.0, 1.0, b) * np.sqrt(1.0 / a))
                else:
                    chunks.append((rng.standard_normal((a, b)) * np.sqrt(2.0 / (a + b))).ravel())
                    chunks.append(rng.standard_normal(b) * 1e-2)
            self.parameters = np.concatenate(chunks)
        self.parameters = np.asarray(self.parameters, dtype=np.float64)
        expected = sum(a * b + b for a, b in zip(self.widths[:-1], self.widths[1:]))
        if self.parameters.size != expected:
            raise ValueError(f"expected {expected} network parameters, got {self.parameters.size}")

    def _layers(self, ps):
        out, o = [], 0
        for a, b in zip(self.widths[:-1], self.widths[1:]):
            W = ps[o:o + a * b].view(a, b)
            o += a * b
            out.append((W, ps[o:o + b]))
            o += b
        return out

    def _scaled(self, x):
        lo = torch.as_tensor(self.lower, dtype=F64, device=x.device)
        hi = torch.as_tensor(self.upper, dtype=F64, device=x.device)
        return (x - lo) * 2 / (hi - lo) - 1  # approxfun/core.py:99-104

    def apply(self, ps, x, with_jacobian=False):
        """x (n, indim) -> (n, outdim) [, d out / d ps as (n * outdim, nparams)]."""
        h = self._scaled(x.to(F64).reshape(-1, self.widths[0]))
        layers = self._layers(ps)
        if self.activation == "sin":
            for l, (W, b) in enumerate(layers):
                h = h @ W + b
                if l + 1 < len(layers):
                    h = torch.sin((self.omega_0 if l == 0 else self.omega) * h)
            if not with_jacobian:
                return h
            J = torch.func.jacfwd(lambda q: self.apply(q, x).reshape(-1))(ps)  # no closed form kept for sine layers
            return h, J
        acts = [h]
        for l, (W, b) in enumerate(layers):
            h = h @ W + b
            if l + 1 < len(layers):
                h = torch.tanh(h)
            acts.append(h)
        if not with_jacobian:
            return h
        n, outdim = h.shape
        J = torch.empty((n, outdim, ps.numel()), dtype=F64, device=ps.device)
        delta = torch.eye(outdim, dtype=F64, device=ps.device).expand(n, outdim, outdim)
        off = ps.numel()
        for l in range(len(layers) - 1, -1, -1):
            W, b = layers[l]
            a_in = acts[l]
            off -= b.numel()
            J[:, :, off:off + b.numel()] = delta
            off -= W.numel()
            J[:, :, off:off + W.numel()] = (a_in[:, None, :, None] * delta[:, :, None, :]).reshape(n, outdim, -1)
            if l > 0:
                delta = (delta @ W.T) * (1.0 - a_in * a_in)[:, None, :]
        return h, J.reshape(n * outdim, ps.numel())


def normalize(data, axes=None):
    # ml/training.py:17-20 (population standard deviation, like numpy's default); axes=None: all elements
    if axes is None:
        mean, std = data.mean(), data.std(unbiased=False)
    else:
        mean, std = data.mean(dim=axes), data.std(dim=axes, unbiased=False)
    return (data - mean) / std, mean, std


def blackman_kernel(dims, M):
    # ml/utils.py:121-131
    n = M - 2
    inds = np.stack(np.meshgrid(*(np.arange(1 - n, n, 2) for _ in range(dims))), axis=-1)
    r = np.linalg.norm(inds.reshape(-1, dims).astype(np.float64), axis=1) / 2
    x = 2 * np.pi * r / (M - 1)
    kernel = 0.42 + 0.5 * np.cos(x) + 0.08 * np.cos(2 * x)
    return (kernel / kernel.sum()).reshape(*(n for _ in range(dims)))


def convolve(data, kernel, boundary="edge"):
    """ml/training.py:23-39 for a kernel with as many axes as `data`: pad by (s - 1) // 2 per side
    ("edge" or "wrap"), then a "valid" convolution, as shifted adds on the data's device."""
    kernel = torch.as_tensor(kernel, dtype=F64, device=data.device)
    padded = data
    for ax, s in enumerate(kernel.shape):
        p, n = (s - 1) // 2, data.shape[ax]
        idx = torch.arange(-p, n + p, device=data.device)
        idx = idx % n if boundary == "wrap" else idx.clamp(0, n - 1)
        padded = padded.index_select(ax, idx)
    out = torch.zeros_like(data)
    flipped = torch.flip(kernel, dims=tuple(range(kernel.ndim)))  # convolution, not correlation
    for off in np.ndindex(*kernel.shape):
        window = tuple(slice(o, o + n) for o, n in zip(off, data.shape))
        out += float(flipped[off]) * padded[window]
    return out


def smooth(y, kernel, boundary):
    """funn.py:226 `vmap(conv)(y.T).T`: each component smoothed over the grid axes."""
    return torch.stack([convolve(y[..., c], kernel, boundary) for c in range(y.shape[-1])], dim=-1)


class _LMWorkspace:
    """Owns a psb_lm_workspace (device buffers + the captured iteration graph of csrc/psb_refit.cu)."""

    def __init__(self):
        import ctypes

        self.handle = ctypes.c_void_p()
        self.keep = None
        self.lib = None

    def result(self):
        """(accepted iterations, cost, mu, loop passes) of the last native fit; synchronises the current stream."""
        import ctypes

        it, loops, cost, mu = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double(), ctypes.c_float()
        from . import _native as N

        dev = self.keep[2].device
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        N.check(self.lib.psb_lm_result(self.handle, ctypes.byref(it), ctypes.byref(cost), ctypes.byref(mu), ctypes.byref(loops), stream))
        return it.value, cost.value, mu.value, loops.value

    def __del__(self):
        if self.lib is not None and self.handle:
            try:
                self.lib.psb_lm_destroy(self.handle)
            except Exception:  # interpreter shutdown
                pass


def build_fitting_function(model, optimizer, force_host=False):
    """ml/training.py:42-56 + ml/optimizers.py:188-232 -> fit(params, x, y) -> params.  Plain-loss fits of tanh
    networks on a CUDA device run in hand-written kernels (psb_lm_fit); sine layers, the gradient / Sobolev losses,
    CPU tensors and `force_host=True` take the torch implementation below (same loop, same float32 conventions)."""
    _, c, mu_min, mu_max, rho_c, rho_min = optimizer.params
    r = float(optimizer.reg.coeff)
    f32 = np.float32

    sobolev = isinstance(optimizer.loss, Sobolev1SSE)
    gradients_only = isinstance(optimizer.loss, GradientsSSE)

    def input_gradients(p, x):
        """d (sum of outputs) / d x for every point: (n, indim) -- functional, so that forward-mode
        differentiation with respect to p can run through it."""
        point = lambda q, xi: model.apply(q, xi[None]).sum()
        return torch.func.vmap(torch.func.grad(point, argnums=1), in_dims=(None, 0))(p, x)

    def error(p, x, y, with_jacobian=False):
        """objectives.py build_error_function: float32 residuals; Sobolev: (values, input-gradients) stacked."""
        x = x.to(F64)
        if gradients_only:  # objectives.py: build_error_function(model, GradientsLoss)
            e = (input_gradients(p, x).reshape(y.shape) - y).to(torch.float32).flatten().to(F64)
            if not with_jacobian:
                return e
            return e, torch.func.jacfwd(lambda q: input_gradients(q, x).reshape(-1))(p)
        ref = y
        if sobolev:
            ref, dref = y
        if with_jacobian:
            pred, J = model.apply(p, x, True)
        else:
            pred, J = model.apply(p, x), None
        e = (pred.reshape(ref.shape) - ref).to(torch.float32).flatten().to(F64)
        if sobolev:
            ge = (input_gradients(p, x).reshape(dref.shape) - dref).to(torch.float32).flatten().to(F64)
            e = torch.cat([e, ge])
            if with_jacobian:  # forward-over-reverse: one tangent per parameter
                gJ = torch.func.jacfwd(lambda q: input_gradients(q, x).reshape(-1))(p)
                J = torch.cat([J, gJ], dim=0)
        return (e, J) if with_jacobian else e

    def cost(e, p):
        return float((e @ e + r * (p @ p)) / 2)

    native_ok = (not sobolev and not gradients_only and model.activation != "sin" and not force_host
                 and sum(a * b + b for a, b in zip(model.widths[:-1], model.widths[1:])) <= 1024
                 and len(model.widths) - 1 <= 8 and max(model.widths) <= 128)
    workspace = _LMWorkspace() if native_ok else None

    def native_fit(params, x, y):
        """psb_lm_fit (csrc/psb_refit.cu): the same loop with hand-written kernels and device-resident control
        state -- stream-ordered, no host synchronisation."""
        import ctypes

        from . import _native as N

        lib = N.load()
        dev = params.device
        p = params.detach().to(F64).contiguous().clone()
        xx = x.detach().to(device=dev, dtype=F64).reshape(-1, model.widths[0]).contiguous()
        yy = y.detach().to(device=dev, dtype=F64).reshape(xx.shape[0], model.widths[-1]).contiguous()
        q = N.LmProblem()
        q.n_dense = len(model.widths) - 1
        for i, w in enumerate(model.widths):
            q.widths[i] = int(w)
        q.activation = N.ACT_TANH
        q.max_iters = int(optimizer.max_iters)
        q.n = int(xx.shape[0])
        q.x, q.y, q.params = xx.data_ptr(), yy.data_ptr(), p.data_ptr()
        lo, hi = np.asarray(model.lower, dtype=np.float64).ravel(), np.asarray(model.upper, dtype=np.float64).ravel()
        for d in range(model.widths[0]):
            q.lower[d], q.upper[d] = float(lo[d]), float(hi[d])
        q.reg = float(r)
        q.mu_0, q.mu_c, q.mu_min, q.mu_max, q.rho_c, q.rho_min = (float(v) for v in optimizer.params)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            N.check(lib.psb_lm_fit(ctypes.byref(q), ctypes.byref(workspace.handle), stream))
        workspace.keep = (xx, yy, p)  # inputs of the stream-ordered kernels stay alive until the next fit
        workspace.lib = lib
        return p

    def fit(params, x, y, history=None):
        if native_ok and history is None and params.is_cuda:
            return native_fit(params, x, y)
        p_ = params.to(F64)
        e_ = error(p_, x, y)
        C_, mu = np.inf, f32(optimizer.params.mu_0)
        iters, improved = 0, True
        while improved and iters < optimizer.max_iters and mu < mu_max:
            _, J = error(p_, x, y, True)
            H = J.T @ J
            H.diagonal().add_(float(f32(r) + mu))
            Je = J.T @ e_ + r * p_
            L, info = torch.linalg.cholesky_ex(H)
            if int(info) == 0:
                dp = torch.cholesky_solve(Je[:, None], L)[:, 0]
            else:  # not positive definite: the reference's solve yields NaN -> rejected step
                dp = torch.full_like(Je, float("nan"))
            p = p_ - dp
            e = error(p, x, y)
            C = cost(e, p)
            rho = (C_ - C) / float(dp @ (float(mu) * dp + Je))
            if rho > rho_c:
                mu = max(f32(mu / f32(c)), f32(mu_min))
            bad = bool(rho < rho_min) or bool(torch.isnan(p).any())
            if bad:
                mu = min(f32(f32(c) * mu), f32(mu_max))
                p, e, C = p_, e_, C_
            improved = bool(C_ > C) or bad
            iters += 0 if bad else 1
            p_, e_, C_ = p, e, C
            if history is not None:
                history.append((p_.clone(), float(C_), float(mu)))
        return p_

    fit.workspace = workspace
    fit.native = native_ok
    return fit


def Siren(indim, outdim, topology, lower, upper, omega_0=16, omega=1, seed=0, parameters=None):
    """ml/models.py:84-152: sinusoidal representation network (first layer frequency omega_0)."""
    return MLP(indim, outdim, topology, lower, upper, seed=seed, parameters=parameters, activation="sin",
               omega_0=float(omega_0), omega=float(omega))
