"""
Oracle (test infrastructure): the neural-network pieces of the ABF family, restated in numpy fp64.

  pysages/ml/models.py:44-82          MLP = Flatten, elementwise(transform), (Dense, Tanh)*, Dense
  pysages/ml/utils.py:43-64,121-131   flat parameter vector (W (in,out) row-major then b, per dense layer),
                                      Blackman smoothing kernel
  pysages/ml/training.py:17-56        normalize, convolve (pad + "valid" convolution), fitting loop
  pysages/ml/objectives.py            SSE + L2 regularisation: errors cast to float32, cost, damped Hessian, J^T e
  pysages/ml/optimizers.py:188-232    Levenberg-Marquardt (mu kept in float32)
  pysages/approxfun/core.py:99-104    scale(x, grid)

Parity unpinned: the reference draws its initial weights from JAX's threefry PRNG through stax
(glorot-normal W, N(0, 1e-2) b), which cannot be reproduced without JAX; `init_params` uses the same
distributions from numpy's PCG64.  No reference test holds golden values for any of this.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""

import numpy as np
import scipy.linalg
import scipy.signal

F32 = np.float32


def layer_sizes(indim, topology, outdim):
    return [int(indim), *[int(t) for t in topology], int(outdim)]


def number_of_params(widths):
    return sum(a * b + b for a, b in zip(widths[:-1], widths[1:]))


def init_params(widths, seed=0):
    """stax.Dense defaults: W ~ glorot normal, b ~ N(0, 1e-2) (different PRNG than the reference)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for a, b in zip(widths[:-1], widths[1:]):
        out.append((rng.standard_normal((a, b)) * np.sqrt(2.0 / (a + b))).ravel())
        out.append(rng.standard_normal(b) * 1e-2)
    return np.concatenate(out)


def split_params(ps, widths):
    """ml/utils.py:56-64 pack: [(W, b), ...] views of the flat vector."""
    layers, o = [], 0
    for a, b in zip(widths[:-1], widths[1:]):
        W = ps[o:o + a * b].reshape(a, b)
        o += a * b
        bias = ps[o:o + b]
        o += b
        layers.append((W, bias))
    return layers


def scale(x, lower, upper):
    # approxfun/core.py:99-104
    return (x - lower) * 2 / (upper - lower) - 1


def mlp_apply(ps, widths, x, lower=None, upper=None, activation="tanh", omega0=1.0, omega=1.0):
    """ml/models.py:44-82 (and :84-152 for sin).  x: (n, indim); the transform is `scale` when bounds are given."""
    h = np.asarray(x, dtype=np.float64).reshape(-1, widths[0])
    if lower is not None:
        h = scale(h, np.asarray(lower, dtype=np.float64), np.asarray(upper, dtype=np.float64))
    layers = split_params(np.asarray(ps, dtype=np.float64), widths)
    for l, (W, b) in enumerate(layers):
        h = h @ W + b
        if l + 1 < len(layers):
            h = np.tanh(h) if activation == "tanh" else np.sin((omega0 if l == 0 else omega) * h)
    return h


def mlp_jacobian(ps, widths, x, lower=None, upper=None):
    """d mlp(x)[n, o] / d ps -> (n * outdim, nparams), tanh activations (what jax.jacobian(error) yields)."""
    h = np.asarray(x, dtype=np.float64).reshape(-1, widths[0])
    if lower is not None:
        h = scale(h, np.asarray(lower, dtype=np.float64), np.asarray(upper, dtype=np.float64))
    layers = split_params(np.asarray(ps, dtype=np.float64), widths)
    acts = [h]
    for l, (W, b) in enumerate(layers):
        h = h @ W + b
        if l + 1 < len(layers):
            h = np.tanh(h)
        acts.append(h)
    n, outdim = acts[-1].shape
    J = np.zeros((n, outdim, ps.size))
    for o in range(outdim):
        delta = np.zeros((n, outdim))
        delta[:, o] = 1.0
        off = ps.size
        for l in range(len(layers) - 1, -1, -1):
            W, b = layers[l]
            a_in = acts[l]
            off -= b.size
            J[:, o, off:off + b.size] = delta
            off -= W.size
            J[:, o, off:off + W.size] = (a_in[:, :, None] * delta[:, None, :]).reshape(n, -1)
            if l > 0:
                delta = (delta @ W.T) * (1.0 - a_in * a_in)
    return J.reshape(n * outdim, ps.size)


# ---- training data ----------------------------------------------------------------------------

def normalize(data, axes=None):
    # ml/training.py:17-20
    mean = data.mean(axis=axes)
    std = data.std(axis=axes)
    return (data - mean) / std, mean, std


def blackman(M, n):
    # ml/utils.py:121-123
    x = 2 * np.pi * n / (M - 1)
    return 0.42 + 0.5 * np.cos(x) + 0.08 * np.cos(2 * x)


def blackman_kernel(dims, M):
    # ml/utils.py:126-131
    n = M - 2
    inds = np.stack(np.meshgrid(*(np.arange(1 - n, n, 2) for _ in range(dims))), axis=-1)
    kernel = blackman(M, np.linalg.norm(inds.reshape(-1, dims).astype(np.float64), axis=1) / 2)
    return (kernel / kernel.sum()).reshape(*(n for _ in range(dims)))


def convolve(data, kernel, boundary="edge"):
    # ml/training.py:23-39
    if kernel.ndim == 1:
        padding = (kernel.size - 1) // 2
    else:
        padding = [((s - 1) // 2, (s - 1) // 2) for s in kernel.shape]
    return scipy.signal.convolve(np.pad(data, padding, mode=boundary), kernel, mode="valid")


def smooth(y, kernel, boundary):
    """funn.py:226: vmap(conv)(y.T).T -- every output component is smoothed over the grid axes."""
    return np.stack([convolve(y[..., c], kernel, boundary) for c in range(y.shape[-1])], axis=-1)


# ---- Levenberg-Marquardt ------------------------------------------------------------------------

class LMParams:
    # ml/optimizers.py:40-51
    mu_0 = 1e-1
    mu_c = 10.0
    mu_min = 1e-8
    mu_max = 1e8
    rho_c = 1e-1
    rho_min = 1e-4


def lm_fit(ps, widths, x, y, lower, upper, reg=1e-6, max_iters=500, params=LMParams, history=None):
    """ml/optimizers.py:188-232 + ml/training.py:42-56 with SSE loss and L2 regularisation
    (funn.py:153: LevenbergMarquardt(reg=L2Regularization(1e-6)))."""
    y = np.asarray(y, dtype=np.float64)

    def error(p):  # objectives.py build_error_function: float32 residuals
        pred = mlp_apply(p, widths, x, lower, upper).reshape(y.shape)
        return (pred - y).astype(F32).ravel().astype(np.float64)

    def cost(e, p):
        return (np.sum(e * e) + reg * np.sum(p * p)) / 2

    c = F32(params.mu_c)
    p_, e_, C_, mu = np.asarray(ps, dtype=np.float64), error(ps), np.inf, F32(params.mu_0)
    iters, improved = 0, True
    while improved and iters < max_iters and mu < params.mu_max:
        mu = F32(mu)
        J = mlp_jacobian(p_, widths, x, lower, upper)
        H = J.T @ J
        H[np.diag_indices_from(H)] += float(F32(reg) + mu)  # python float + f32 scalar stays f32 in JAX
        Je = J.T @ e_ + reg * p_
        dp = scipy.linalg.solve(H, Je, assume_a="pos")
        p = p_ - dp
        e = error(p)
        C = cost(e, p)
        rho = (C_ - C) / (dp @ (float(mu) * dp + Je))
        if rho > params.rho_c:
            mu = max(F32(mu / c), F32(params.mu_min))
        bad = bool(rho < params.rho_min) or bool(np.any(np.isnan(p)))
        if bad:
            mu = min(F32(c * mu), F32(params.mu_max))
            p, e, C = p_, e_, C_
        improved = bool(C_ > C) or bad
        iters += 0 if bad else 1
        p_, e_, C_ = p, e, C
        if history is not None:
            history.append((p_.copy(), float(C_), float(mu)))
    return p_


class NNData:
    # ml/training.py:11-14
    def __init__(self, params, mean, std):
        self.params, self.mean, self.std = params, mean, std


def mlp_input_gradient(ps, widths, x, lower, upper):
    """jax.grad(lambda p, x: model.apply(p, x).sum(), argnums=1) of ann.py:240 for ONE point x (tanh layers),
    chain rule through `scale` included."""
    u = scale(np.asarray(x, dtype=np.float64).reshape(1, widths[0]), np.asarray(lower, dtype=np.float64),
              np.asarray(upper, dtype=np.float64))
    layers = split_params(np.asarray(ps, dtype=np.float64), widths)
    h, acts = u, []
    for l, (W, b) in enumerate(layers):
        h = h @ W + b
        if l + 1 < len(layers):
            h = np.tanh(h)
            acts.append(h)
    delta = np.ones((1, widths[-1]))
    for l in range(len(layers) - 1, -1, -1):
        delta = delta @ layers[l][0].T
        if l > 0:
            delta = delta * (1.0 - acts[l - 1] ** 2)
    return (delta * 2 / (np.asarray(upper, dtype=np.float64) - np.asarray(lower, dtype=np.float64))).reshape(-1)


# ---- Sobolev-loss Levenberg-Marquardt (CFF, cff.py:164 `LevenbergMarquardt(loss=Sobolev1SSE(), ...)`) ------

def _torch_mlp(p, widths, x, lower, upper, activation="tanh", omega0=1.0, omega=1.0):
    import torch

    h = (x - torch.as_tensor(lower)) * 2 / (torch.as_tensor(upper) - torch.as_tensor(lower)) - 1
    o = 0
    n = len(widths) - 1
    for l, (a, b) in enumerate(zip(widths[:-1], widths[1:])):
        W = p[o:o + a * b].reshape(a, b)
        o += a * b
        h = h @ W + p[o:o + b]
        o += b
        if l + 1 < n:
            h = torch.tanh(h) if activation == "tanh" else torch.sin((omega0 if l == 0 else omega) * h)
    return h


def init_siren_params(widths, omega=1.0, seed=0):
    """ml/models.py:113-147 + ml/utils.py:67-100: U(-1, 1) * scale; first layer 1 / fan_in, others
    sqrt(6 / omega^2 / fan_in), biases sqrt(1 / fan_in) (different PRNG than the reference)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for l, (a, b) in enumerate(zip(widths[:-1], widths[1:])):
        ws = 1.0 / a if l == 0 else np.sqrt(6.0 / omega**2 / a)
        out.append((rng.uniform(-1.0, 1.0, (a, b)) * ws).ravel())
        out.append(rng.uniform(-1.0, 1.0, b) * np.sqrt(1.0 / a))
    return np.concatenate(out)


def lm_fit_autograd(ps, widths, x, targets, loss, lower, upper, reg=1e-4, max_iters=250, activation="tanh",
                    omega0=1.0, omega=1.0, params=LMParams, history=None):
    """ml/optimizers.py:188-232 for the three losses of ml/objectives.py -- "sse" (values), "gradients"
    (input-gradients, sirens.py:162 mode "abf") and "sobolev" (both) -- and either activation, every
    derivative by torch autograd (reverse over reverse)."""
    import torch

    xt = torch.as_tensor(np.asarray(x, dtype=np.float64))
    if loss == "sobolev":
        yt, dyt = (torch.as_tensor(np.asarray(t, dtype=np.float64)) for t in targets)
    elif loss == "gradients":
        yt, dyt = None, torch.as_tensor(np.asarray(targets, dtype=np.float64))
    else:
        yt, dyt = torch.as_tensor(np.asarray(targets, dtype=np.float64)), None
    net = lambda p, xx: _torch_mlp(p, widths, xx, lower, upper, activation, omega0, omega)

    def grads(p, create_graph):
        xx = xt.clone().requires_grad_(True)
        (g,) = torch.autograd.grad(net(p, xx).sum(), xx, create_graph=create_graph)
        return g

    def residuals(p, create_graph=False):
        parts = []
        if yt is not None:
            parts.append((net(p, xt).reshape(yt.shape) - yt).flatten())
        if dyt is not None:
            parts.append((grads(p, create_graph).reshape(dyt.shape) - dyt).flatten())
        return torch.cat(parts)

    def error(p_np):  # float32 residuals (objectives.py)
        return residuals(torch.as_tensor(p_np)).detach().to(torch.float32).to(torch.float64).numpy()

    def jacobian(p_np):
        p = torch.as_tensor(p_np).clone().requires_grad_(True)
        return torch.autograd.functional.jacobian(lambda q: residuals(q, True), p).numpy()

    def cost(e, p):
        return (np.sum(e * e) + reg * np.sum(p * p)) / 2

    c = F32(params.mu_c)
    p_, C_, mu = np.asarray(ps, dtype=np.float64), np.inf, F32(params.mu_0)
    e_ = error(p_)
    iters, improved = 0, True
    while improved and iters < max_iters and mu < params.mu_max:
        mu = F32(mu)
        J = jacobian(p_)
        H = J.T @ J
        H[np.diag_indices_from(H)] += float(F32(reg) + mu)
        Je = J.T @ e_ + reg * p_
        dp = scipy.linalg.solve(H, Je, assume_a="pos")
        p = p_ - dp
        e = error(p)
        C = cost(e, p)
        rho = (C_ - C) / (dp @ (float(mu) * dp + Je))
        if rho > params.rho_c:
            mu = max(F32(mu / c), F32(params.mu_min))
        bad = bool(rho < params.rho_min) or bool(np.any(np.isnan(p)))
        if bad:
            mu = min(F32(c * mu), F32(params.mu_max))
            p, e, C = p_, e_, C_
        improved = bool(C_ > C) or bad
        iters += 0 if bad else 1
        p_, e_, C_ = p, e, C
        if history is not None:
            history.append((p_.copy(), float(C_), float(mu)))
    return p_


def net_input_gradient(ps, widths, x, lower, upper, activation="tanh", omega0=1.0, omega=1.0):
    """grad(lambda p, x: model.apply(p, x).sum(), argnums=-1) at ONE point (sirens.py:367,375-379)."""
    import torch

    xx = torch.tensor(np.asarray(x, dtype=np.float64).reshape(1, widths[0]), requires_grad=True)
    out = _torch_mlp(torch.as_tensor(np.asarray(ps, dtype=np.float64)), widths, xx, lower, upper, activation, omega0, omega)
    (g,) = torch.autograd.grad(out.sum(), xx)
    return g.numpy().reshape(-1)


def lm_fit_sobolev(ps, widths, x, y, dy, lower, upper, reg=1e-4, max_iters=1000, params=LMParams, history=None):
    """ml/optimizers.py:188-232 with ml/objectives.py's Sobolev1SSE pieces: errors (e, ge) of the values and of
    the input-gradients (both float32), cost (|e|^2 + |ge|^2 + r |p|^2) / 2, damped Hessian J^T J + gJ^T gJ +
    (r + mu) I, J^T e + gJ^T ge + r p.  Derivatives by torch autograd (reverse over reverse)."""
    import torch

    xt = torch.as_tensor(np.asarray(x, dtype=np.float64))
    yt, dyt = torch.as_tensor(np.asarray(y, dtype=np.float64)), torch.as_tensor(np.asarray(dy, dtype=np.float64))

    def grads(p, create_graph):
        xx = xt.clone().requires_grad_(True)
        out = _torch_mlp(p, widths, xx, lower, upper).sum()
        (g,) = torch.autograd.grad(out, xx, create_graph=create_graph)
        return g

    def error(p_np):
        p = torch.as_tensor(p_np)
        e = (_torch_mlp(p, widths, xt, lower, upper).reshape(yt.shape) - yt).to(torch.float32).flatten().to(torch.float64)
        ge = (grads(p, False).reshape(dyt.shape) - dyt).to(torch.float32).flatten().to(torch.float64)
        return torch.cat([e, ge]).numpy()

    def jacobians(p_np):
        J = mlp_jacobian(p_np, widths, x, lower, upper)
        p = torch.as_tensor(p_np).clone().requires_grad_(True)
        gJ = torch.autograd.functional.jacobian(lambda q: grads(q, True).reshape(-1), p)
        return np.vstack([J, gJ.numpy()])

    def cost(e, p):
        return (np.sum(e * e) + reg * np.sum(p * p)) / 2

    c = F32(params.mu_c)
    p_, C_, mu = np.asarray(ps, dtype=np.float64), np.inf, F32(params.mu_0)
    e_ = error(p_)
    iters, improved = 0, True
    while improved and iters < max_iters and mu < params.mu_max:
        mu = F32(mu)
        J = jacobians(p_)
        H = J.T @ J
        H[np.diag_indices_from(H)] += float(F32(reg) + mu)
        Je = J.T @ e_ + reg * p_
        dp = scipy.linalg.solve(H, Je, assume_a="pos")
        p = p_ - dp
        e = error(p)
        C = cost(e, p)
        rho = (C_ - C) / (dp @ (float(mu) * dp + Je))
        if rho > params.rho_c:
            mu = max(F32(mu / c), F32(params.mu_min))
        bad = bool(rho < params.rho_min) or bool(np.any(np.isnan(p)))
        if bad:
            mu = min(F32(c * mu), F32(params.mu_max))
            p, e, C = p_, e_, C_
        improved = bool(C_ > C) or bad
        iters += 0 if bad else 1
        p_, e_, C_ = p, e, C
        if history is not None:
            history.append((p_.copy(), float(C_), float(mu)))
    return p_
