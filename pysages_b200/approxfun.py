"""
Series fits of SpectralABF (pysages/approxfun/core.py): set-up on the host, refits on the device.

The pseudo-inverse of the gradient Vandermonde matrix depends only on the grid: it is built once with
numpy when the method is constructed (approxfun/core.py:73-96 collect_exponents, :115-136 unit_mesh,
:170-201 vandergrad_builder, :204-211 pinv) and uploaded.  A refit (core.py:232-258 build_fitter) is then
one normalisation and one matrix-vector product on the method's device, and its result -- [scale,
coefficients] in one device tensor -- goes to the step kernel through psb_set_force_series without
touching the host.  The kernel evaluates the gradient of the series at xi (core.py:311-336; csrc/psb_step.cuh
force_series_eval).
"""

from itertools import product

import numpy as np
import torch

from .grids import Chebyshev, Grid  # noqa: F401


def collect_exponents(grid):
    """approxfun/core.py:73-96"""
    if grid.shape.size > 2:
        raise ValueError("Only 1D and 2D grids are supported")
    r = grid.shape.sum() / 4 if grid.is_periodic else np.square(grid.shape).sum() / 16
    n = int(np.floor(np.sqrt(r))) + 1
    r = float(r)
    if grid.shape.size == 1:
        return np.arange(1, n, dtype=np.int32).reshape(-1, 1)
    m = 1 - n if grid.is_periodic else 1
    mixed = [t for t in product(range(1, n), range(m, n)) if t[1] != 0 and t[0] ** 2 + t[1] ** 2 <= r]
    return np.asarray([(0, i) for i in range(1, n)] + [(i, 0) for i in range(1, n)] + mixed, dtype=np.int32)


def unit_mesh(grid):
    """approxfun/core.py:115-136: bin centres on [-1, 1]^n, Chebyshev-distributed for Chebyshev grids."""
    axes = []
    for n in grid.shape:
        k = np.arange(int(n), dtype=np.float64)
        axes.append(-np.cos((k + 0.5) * np.pi / n) if type(grid).type_parameter is Chebyshev else -1 + (2 * k + 1) / n)
    return np.stack([c.reshape(-1) for c in np.meshgrid(*axes, indexing="ij")], axis=-1)


def gradient_vandermonde(grid, exponents, xs):
    """approxfun/core.py:170-201 for all points at once: rows ordered (point, dimension), columns = terms;
    periodic grids: the [imag, real] blocks of core.py:207-209."""
    ns = np.asarray(exponents, dtype=np.float64)  # (K, d)
    xs = np.asarray(xs, dtype=np.float64).reshape(-1, 1, ns.shape[1])  # (P, 1, d)
    if grid.is_periodic:
        s = 2 * np.pi / grid.size
        phase = np.exp(-1j * np.pi * (ns[None] * xs).sum(-1))  # (P, K)
        A = (s * ns)[None] * phase[..., None]  # (P, K, d): s_d n_kd exp(-i pi n_k.x)
    else:
        s = 2 / grid.size
        ni = np.asarray(exponents)
        lowered = xs ** np.maximum(ni - 1, 0)[None]  # x_d^(n_kd - 1)
        full = xs ** ni[None]
        others = np.stack([np.prod(np.delete(full, d, axis=-1), axis=-1) for d in range(ni.shape[1])], axis=-1)
        A = (s * ns)[None] * lowered * others
    A = np.transpose(A, (0, 2, 1)).reshape(-1, ns.shape[0])  # (P * d, K)
    return np.hstack((A.imag, A.real)) if grid.is_periodic else A


class SpectralGradientFit:
    """approxfun/core.py:30-61: grid, unit mesh, exponents and the pseudo-inverse (jnp.linalg.pinv's
    default cut-off 10 * max(M, N) * eps)."""

    def __init__(self, grid):
        self.grid = grid
        self.exponents = collect_exponents(grid)
        self.mesh = unit_mesh(grid)
        A = gradient_vandermonde(grid, self.exponents, self.mesh)
        self.pinv = np.linalg.pinv(A, rcond=10 * max(A.shape) * np.finfo(np.float64).eps)

    @property
    def is_periodic(self):
        return self.grid.is_periodic


def build_fitter(model, device, force_host=False):
    """approxfun/core.py:232-258 on `device`: dy (*grid.shape, dims) -> tensor [scale, coefficients...].
    On a CUDA device `fit.from_state(hist, Fsum)` runs the whole refit (mean force, normalisation, projection) in two
    hand-written kernels (psb_series_fit, csrc/psb_refit.cu), stream-ordered; `fit(dy)` is the torch form of the same."""
    device = torch.device(device)
    pinv = torch.as_tensor(model.pinv, dtype=torch.float64, device=device).contiguous()
    axes = tuple(range(model.grid.shape.size))
    periodic = model.is_periodic

    def fit(dy):
        dy = dy.to(torch.float64)
        std = dy.std(dim=axes, unbiased=False).max()
        std = torch.where(std == 0, torch.ones_like(std), std)
        dy = (dy - dy.mean(dim=axes)) / std if periodic else dy / std
        return torch.cat([std.reshape(1), pinv @ dy.reshape(-1)])

    def from_state_host(hist, Fsum):
        h = hist.to(torch.float64).reshape(*Fsum.shape[:-1], 1)
        return fit(Fsum / torch.clamp(h, min=1.0))

    fit.native = device.type == "cuda" and not force_host
    if not fit.native:
        fit.from_state = from_state_host
        return fit

    import ctypes

    from . import _native as N

    lib = N.load()
    keep = {}

    def from_state(hist, Fsum):
        k = int(Fsum.shape[-1])
        npoints = int(Fsum.numel() // k)
        if hist.dtype not in (torch.int32, torch.uint32) or Fsum.dtype != torch.float64 or not Fsum.is_contiguous() \
                or not hist.is_contiguous() or pinv.shape[1] != npoints * k:
            return from_state_host(hist, Fsum)
        out = torch.empty(1 + pinv.shape[0], dtype=torch.float64, device=device)
        work = keep.get("work")
        if work is None or work.numel() < npoints * k:
            work = keep["work"] = torch.empty(npoints * k, dtype=torch.float64, device=device)
        d = N.SeriesFitDesc()
        d.npoints, d.k, d.nterms, d.periodic = npoints, k, int(pinv.shape[0]), int(bool(periodic))
        d.hist, d.Fsum, d.pinv, d.out, d.work = hist.data_ptr(), Fsum.data_ptr(), pinv.data_ptr(), out.data_ptr(), work.data_ptr()
        with torch.cuda.device(device):
            N.check(lib.psb_series_fit(ctypes.byref(d), ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
        return out

    fit.from_state = from_state
    return fit
