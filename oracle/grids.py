"""
Oracle (test infrastructure): grids, bin indexers and bin-centre meshes.

Restates pysages/grids.py:30-158 and pysages/approxfun/core.py:99-147 in numpy.
Bin indices are the one place where parity must be BIT-exact, so JAX's float
floor-division / remainder are restated step by step (jax/_src/numpy/ufuncs.py
`_float_divmod` and `remainder`, jax 0.4.x -- the reference pins jax>=0.3.5,
Dockerfile jax==0.4.34) instead of relying on numpy's `//`, which rounds the
quotient slightly differently (floor + 0.5 test instead of round-half-away).
"""

import numpy as np


class Regular:
    pass


class Periodic:
    pass


class Chebyshev:
    pass


class Grid:
    """pysages/grids.py:30-73 (plum parametric type -> explicit `kind` attribute)"""

    def __init__(self, lower, upper, shape, periodic=False, kind=None):
        if kind is None:
            kind = Periodic if periodic else Regular
        shape = np.asarray(shape, dtype=np.int64)
        n = shape.size
        self.kind = kind
        self.lower = np.asarray(lower, dtype=np.float64).reshape(n)
        self.upper = np.asarray(upper, dtype=np.float64).reshape(n)
        self.shape = shape.reshape(n)
        self.size = self.upper - self.lower

    @property
    def is_periodic(self):
        return self.kind is Periodic


def _round_half_away(x):
    # lax.round default rounding method is AWAY_FROM_ZERO
    return np.copysign(np.floor(np.abs(x) + 0.5), x)


def jax_floor_divide(x1, x2):
    """jnp.floor_divide for floats == _float_divmod(x1, x2)[0]"""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mod = np.fmod(x1, x2)  # lax.rem: C fmod, sign of the dividend
        div = (x1 - mod) / x2
        ind = (mod != 0) & (np.sign(x2) != np.sign(mod))
        div = np.where(ind, div - 1.0, div)
        return _round_half_away(div)


def jax_remainder(x1, x2):
    """jnp.remainder for floats"""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        trunc_mod = np.fmod(x1, x2)
        do_plus = ((trunc_mod < 0) != (x2 < 0)) & (trunc_mod != 0)
        return np.where(do_plus, trunc_mod + x2, trunc_mod)


def build_indexer(grid):
    """pysages/grids.py:109-158 -> get_index(x) -> tuple of np.uint32"""
    shape_f = grid.shape.astype(np.float64)

    if grid.kind is Regular:

        def get_index(x):  # grids.py:117-121
            h = grid.size / shape_f
            idx = jax_floor_divide(np.asarray(x, dtype=np.float64).flatten() - grid.lower, h)
            idx = np.where((idx < 0) | (idx > shape_f), shape_f, idx)
            idx = np.where(np.isnan(idx), shape_f, idx)  # uint32(nan): keep it out of bounds
            return tuple(np.uint32(i) for i in idx)

    elif grid.kind is Periodic:

        def get_index(x):  # grids.py:134-138
            h = grid.size / shape_f
            idx = jax_floor_divide(np.asarray(x, dtype=np.float64).flatten() - grid.lower, h)
            idx = jax_remainder(idx, shape_f)
            return tuple(np.uint32(i) for i in idx)

    elif grid.kind is Chebyshev:

        def get_index(x):  # grids.py:152-156
            with np.errstate(invalid="ignore"):
                y = 2 * (grid.lower - np.asarray(x, dtype=np.float64).flatten()) / grid.size + 1
                idx = jax_floor_divide(shape_f * np.arccos(y), np.pi)
            idx = np.where(np.isnan(idx), shape_f, idx)
            return tuple(np.uint32(i) for i in idx)

    else:
        raise TypeError("unknown grid kind")

    return get_index


def compute_mesh(grid):
    """approxfun/core.py:107-147 -> (prod(shape), ndim) bin centres, row-major ('ij')"""
    axes = []
    for n in grid.shape:
        k = np.arange(int(n), dtype=np.float64)
        if grid.kind is Chebyshev:
            axes.append(-np.cos((k + 0.5) * np.pi / n))
        else:
            axes.append(-1 + (2 * k + 1) / n)
    coords = np.meshgrid(*axes, indexing="ij")
    unit = np.stack([c.reshape(-1) for c in coords], axis=-1)
    return (unit + 1) / 2 * grid.size + grid.lower
