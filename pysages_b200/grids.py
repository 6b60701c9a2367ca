"""
Grids over CV space (pysages/grids.py:13-106): host-side description only.

`Grid(lower, upper, shape, periodic=...)`, `Grid[Periodic](...)`, `Grid[Chebyshev](...)` behave
like the reference's plum-parametric class (same TypeError / ValueError rules, tested by the
reference's tests/test_grids.py:17-54).  Binning itself (`build_indexer`, grids.py:109-158) runs on
the device, bit-exactly, inside the step kernel; `build_indexer` here returns a function that
evaluates it through the C ABI (psb_grid_index) for callers that want indices on the host.
"""

import numpy as np

from . import _native as N


class GridType:
    pass


class Periodic(GridType):
    pass


class Regular(GridType):
    pass


class Chebyshev(GridType):
    pass


_KIND = {Regular: N.GRID_REGULAR, Periodic: N.GRID_PERIODIC, Chebyshev: N.GRID_CHEBYSHEV}


class _GridMeta(type):
    _cache = {}

    def __getitem__(cls, T):
        base = cls.__dict__.get("_base", cls)
        if not (isinstance(T, type) and issubclass(T, GridType)):
            raise TypeError("Type parameter must be a subclass of GridType.")
        key = (base, T)
        if key not in _GridMeta._cache:
            sub = _GridMeta(f"Grid[{T.__name__}]", (base,), {"type_parameter": T, "_base": base})
            _GridMeta._cache[key] = sub
        return _GridMeta._cache[key]

    def __call__(cls, lower, upper, shape, **kwargs):
        if "type_parameter" not in cls.__dict__:
            # grids.py:38-40 __infer_type_parameter__
            periodic = kwargs.get("periodic", False)
            if len(kwargs) > 1 or (len(kwargs) == 1 and "periodic" not in kwargs):
                raise ValueError("Invalid keyword argument")
            if type(periodic) is not bool:
                raise TypeError("`periodic` must be a bool.")
            cls = cls[Periodic if periodic else Regular]
        return super(_GridMeta, cls).__call__(lower, upper, shape, **kwargs)


class Grid(metaclass=_GridMeta):
    """pysages/grids.py:30-73"""

    def __init__(self, lower, upper, shape, **kwargs):
        self.__check_init_invariants__(**kwargs)
        shape = np.asarray(shape, dtype=np.int64)
        n = shape.size
        if n > N.MAX_GRID_DIMS:
            raise ValueError(f"grids of up to {N.MAX_GRID_DIMS} dimensions are supported")
        self.lower = np.asarray(lower, dtype=np.float64).reshape(n)
        self.upper = np.asarray(upper, dtype=np.float64).reshape(n)
        self.shape = shape.reshape(n)
        self.size = self.upper - self.lower

    def __check_init_invariants__(self, **kwargs):
        # grids.py:51-63
        T = type(self).type_parameter
        if len(kwargs) > 1 or (len(kwargs) == 1 and "periodic" not in kwargs):
            raise ValueError("Invalid keyword argument")
        periodic = kwargs.get("periodic", T is Periodic)
        if type(periodic) is not bool:
            raise TypeError("`periodic` must be a bool.")
        if (not periodic and T is Periodic) or (periodic and T in (Regular, Chebyshev)):
            raise ValueError("Incompatible type parameter and keyword argument")

    def __reduce__(self):
        # the parametrised classes Grid[T] are created on demand and cannot be found by name
        return (build_grid, (type(self).type_parameter, self.lower, self.upper, self.shape))

    def __eq__(self, other):
        return (
            isinstance(other, Grid)
            and type(self).type_parameter is type(other).type_parameter
            and np.array_equal(self.lower, other.lower)
            and np.array_equal(self.upper, other.upper)
            and np.array_equal(self.shape, other.shape)
        )

    __hash__ = None

    def __repr__(self):
        T = type(self).type_parameter
        P = "" if T is Regular else f"[{T.__name__}]"
        return f"Grid{P} ({' x '.join(map(str, self.shape))})"

    @property
    def is_periodic(self):
        return type(self).type_parameter is Periodic

    def describe(self):
        """-> psb_grid_desc"""
        d = N.GridDesc()
        d.kind = _KIND[type(self).type_parameter]
        d.ndim = int(self.shape.size)
        for i in range(d.ndim):
            d.lower[i] = float(self.lower[i])
            d.upper[i] = float(self.upper[i])
            d.shape[i] = int(self.shape[i])
        return d


def build_grid(T, lower, upper, shape):
    # grids.py:76-78
    return Grid[T](lower, upper, shape)


def convert(grid, T):
    # grids.py:86-90
    if not (isinstance(T, type) and issubclass(T, Grid)):
        raise TypeError(f"Cannot convert Grid to a {repr(T)}")
    return T(grid.lower, grid.upper, grid.shape)


def get_info(grid):
    # grids.py:93-101
    if grid is None:
        return (None,)
    return (
        type(grid).type_parameter,
        tuple(float(x) for x in grid.lower),
        tuple(float(x) for x in grid.upper),
        tuple(int(x) for x in grid.shape),
    )


def build_indexer(grid):
    """grids.py:109-158 -- bins computed ON THE DEVICE by the same code the step kernel uses."""
    desc = grid.describe()

    def get_index(x):
        lib = N.load()
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1, desc.ndim))
        out = np.zeros(x.shape, dtype=np.uint32)
        N.check(
            lib.psb_grid_index(
                desc,
                x.ctypes.data_as(N.POINTER(N.c_double)),
                x.shape[0],
                out.ctypes.data_as(N.POINTER(N.c_uint32)),
            )
        )
        if out.shape[0] == 1:
            return tuple(np.uint32(i) for i in out[0])
        return out

    return get_index


def compute_mesh(grid):
    """approxfun/core.py:107-147: bin centres, row-major ('ij'); host numpy (set-up time only)."""
    axes = []
    for n in grid.shape:
        k = np.arange(int(n), dtype=np.float64)
        if type(grid).type_parameter is Chebyshev:
            axes.append(-np.cos((k + 0.5) * np.pi / n))
        else:
            axes.append(-1 + (2 * k + 1) / n)
    coords = np.meshgrid(*axes, indexing="ij")
    unit = np.stack([c.reshape(-1) for c in coords], axis=-1)
    return (unit + 1) / 2 * grid.size + grid.lower
