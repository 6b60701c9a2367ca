"""
Sampling methods (pysages/methods/*.py) backed by the fused CUDA bias step.

Same classes, constructor arguments, validation errors and `build(snapshot, helpers) ->
(snapshot, initialize, update)` contract as the reference (methods/core.py:81-153):

    HarmonicBias        methods/harmonic_bias.py:42-132
    Metadynamics        methods/metad.py:72-205 (standard / well-tempered, direct sum / grid)
    ABF                 methods/abf.py:75-227
    Unbiased            methods/unbiased.py:36-75
    UmbrellaIntegration methods/umbrella_integration.py:37-125
    CVRestraints        methods/restraints.py:14-73

`update(snapshot, state)` enqueues ONE fused launch on the current CUDA stream: gather of the
selected atoms, CVs + closed-form Jacobians, the method's bias and the in-place scatter into the
engine's force array.  It never synchronises with the host; state fields are zero-copy views of
device memory owned by the native handle.
"""

import ctypes
import math
from typing import Any, NamedTuple

import numpy as np
import torch

from . import _native as N
from . import colvars as _colvars
from .backends import snapshot as _snap
from .grids import Grid


# --------------------------------------------------------------------------------------------
# restraints (methods/restraints.py)
# --------------------------------------------------------------------------------------------
class CVRestraints(NamedTuple):
    """Harmonic restraint parameters (restraints.py:14-37)."""

    lower: Any
    upper: Any
    kl: Any
    ku: Any


def canonicalize(restraints, cvs):
    """restraints.py:49-61: infinite bounds wherever the spring constant is zero."""
    if restraints is None:
        return None
    lower, upper, kl, ku = restraints
    kl = np.asarray(kl, dtype=np.float64).flatten()
    ku = np.asarray(ku, dtype=np.float64).flatten()
    lower = np.where(kl == 0, -np.inf, np.asarray(lower, dtype=np.float64).flatten())
    upper = np.where(ku == 0, np.inf, np.asarray(upper, dtype=np.float64).flatten())
    ncomp = sum(cv.ncomponents for cv in cvs)
    if not (len(lower) == len(upper) == len(kl) == len(ku) == ncomp):
        raise ValueError("The number of restraints must equal the number of CVs.")
    return CVRestraints(lower, upper, kl, ku)


# --------------------------------------------------------------------------------------------
# device state views
# --------------------------------------------------------------------------------------------
class _DeviceBuffer:
    """Exposes handle-owned device memory through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape),
            "typestr": typestr,
            "data": (int(ptr), False),
            "version": 2,
        }
        self._owner = owner


_TYPESTR = {0: "<f8", 1: "<i8", 2: "<i4"}  # u32 histograms are viewed as int32 (torch has no uint32 ops)


class HarmonicBiasState(NamedTuple):  # harmonic_bias.py:23-39
    xi: Any
    bias: Any
    ncalls: int = 0


class UnbiasedState(NamedTuple):  # unbiased.py:17-33
    xi: Any
    bias: Any
    ncalls: int = 0


class MetadynamicsState(NamedTuple):  # metad.py:26-69
    xi: Any
    bias: Any
    heights: Any
    centers: Any
    sigmas: Any
    grid_potential: Any
    grid_gradient: Any
    idx: Any
    ncalls: int = 0


class ABFState(NamedTuple):  # abf.py:32-72
    xi: Any
    bias: Any
    hist: Any
    Fsum: Any
    force: Any
    Wp: Any
    Wp_: Any
    ncalls: int = 0


def freeze(state):
    """A persistent VALUE of a state.  The tuples `update` returns are views of the handle's live device memory
    (zero copy, no per-step allocation): every state of a run aliases the same buffers, so one kept across steps
    shows the newest values -- unlike the reference, whose `update` is pure-functional and returns fresh arrays
    (methods/core.py:108-116).  Callbacks that keep history (`Result.states` of intermediate steps, custom loggers)
    take `freeze(state)`: tensors are cloned on the device (stream-ordered, no host sync), nested tuples recursed."""

    def clone(x):
        if isinstance(x, torch.Tensor):
            return x.detach().clone()
        if isinstance(x, tuple) and hasattr(x, "_fields"):
            return type(x)(*(clone(v) for v in x))
        if isinstance(x, (tuple, list)):
            return type(x)(clone(v) for v in x)
        return x

    return clone(state)


class NativeStep:
    """Owns one psb_handle: the device-resident method state + index tables of one simulation."""

    def __init__(self, method, snapshot, helpers, dense_outputs=False):
        if not torch.cuda.is_available():
            raise RuntimeError(
                "pysages_b200 runs the bias step on a CUDA device only (no CPU fallback): no CUDA device found"
            )
        # Engine arrays on the host (OpenMM CPU / Reference platforms, LAMMPS without Kokkos-CUDA):
        # every step stages them through the device (psb_step_host) -- still no CPU compute path.
        self.host_buffers = not snapshot.positions.is_cuda
        self.lib = N.load()
        self.method = method
        self.layout = helpers.layout
        self.device = (torch.device("cuda", torch.cuda.current_device()) if self.host_buffers
                       else snapshot.positions.device)
        keep = []
        descs = []
        for cv in method.cvs:  # a multicomponent ring CV lowers to two kernel-level CVs
            d = cv.describe(keep)
            descs.extend(d if isinstance(d, list) else [d])
        cvs = (N.CvDesc * len(descs))(*descs)
        k_total = sum(cv.ncomponents for cv in method.cvs)
        if not 1 <= len(descs) <= N.MAX_CVS:  # the library's own messages; raised here before the descriptors overflow
            raise ValueError(f"between 1 and {N.MAX_CVS} collective variables are supported (got {len(descs)})")
        if k_total > N.MAX_K:
            raise ValueError(f"at most {N.MAX_K} CV components are supported")
        mdesc = method.describe(helpers)
        ldesc = _snap.describe_layout(
            snapshot,
            self.layout,
            unwrap=method.requires_box_unwrapping,
            need_momenta="momenta" in method.snapshot_flags,
            dense_outputs=dense_outputs,
        )
        self.k = sum(cv.ncomponents for cv in method.cvs)
        self.natoms = int(ldesc.natoms)
        self.dense_outputs = bool(dense_outputs)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self.lib.psb_create(cvs, len(descs), ctypes.byref(mdesc), ctypes.byref(ldesc),
                                        ctypes.byref(handle)))
        self.handle = handle
        self._snapdesc = N.SnapshotDesc()
        self._psb_step = self.lib.psb_step
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)  # int handle of the current stream
        self._device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._host_keep = {}
        self._views = {}
        self.ncalls = 0

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self.lib.psb_destroy(h)  # also unregisters the host arrays; only then may they be freed
            except Exception:  # interpreter shutdown
                pass
        self._host_keep = {}

    # ---- stepping ------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _desc(self, snapshot):
        if self.host_buffers:
            # psb_step_host page-locks these arrays in place and leaves them registered until the handle dies: keep
            # them alive that long (a freed-but-still-registered range would poison later allocations at its address)
            def hold(x):
                if isinstance(x, (tuple, list)):
                    for y in x:
                        hold(y)
                elif isinstance(x, dict):
                    for y in x.values():
                        hold(y)
                elif isinstance(x, torch.Tensor):
                    self._host_keep.setdefault(x.data_ptr(), x)

            hold((snapshot.positions, snapshot.vel_mass, snapshot.forces, snapshot.ids, snapshot.extras))
        return _snap.describe_snapshot(snapshot, self.layout, self._snapdesc)

    def evaluate(self, snapshot):
        fn = self.lib.psb_evaluate_cvs_host if self.host_buffers else self.lib.psb_evaluate_cvs
        N.check(fn(self.handle, ctypes.byref(self._desc(snapshot)), self._stream()))

    def step(self, snapshot, timestep=0):
        if self.host_buffers:  # H2D of the arrays the method needs -> step -> D2H forces (synchronous)
            N.check(self.lib.psb_step_host(self.handle, ctypes.byref(self._desc(snapshot)), int(timestep),
                                           self._stream(), None, None))
        else:
            N.check(self.lib.psb_step(self.handle, ctypes.byref(self._desc(snapshot)), int(timestep), self._stream()))
        self.ncalls += 1

    def step_desc(self, desc, timestep=0):
        """psb_step on a psb_snapshot the caller keeps filled (engine adapters: device pointers read straight
        out of the plugin's DLPack capsules, backends/_common.py).  The per-step Python path: the raw stream handle
        comes from torch's C accessor, nothing is allocated."""
        if self._raw_stream is not None:
            st = self._psb_step(self.handle, ctypes.byref(desc), timestep, self._raw_stream(self._device_index))
        else:
            st = self._psb_step(self.handle, ctypes.byref(desc), int(timestep), self._stream())
        if st:
            N.check(st)
        self.ncalls += 1

    def release_host_buffers(self):
        """The engine is about to free / reallocate the host arrays psb_step_host page-locked in place."""
        N.check(self.lib.psb_host_release(self.handle))

    def next_step_deposits(self):
        return bool(self.lib.psb_next_step_deposits(self.handle))

    def walker_begin(self, snapshot, record, timestep=0):
        N.check(self.lib.psb_walker_begin(self.handle, ctypes.byref(self._desc(snapshot)), int(timestep),
                                          ctypes.c_void_p(record.data_ptr()), self._stream()))

    def walker_finish(self, snapshot, records, nwalkers):
        N.check(self.lib.psb_walker_finish(self.handle, ctypes.byref(self._desc(snapshot)),
                                           ctypes.c_void_p(records.data_ptr()), int(nwalkers), self._stream()))
        self.ncalls += 1

    @property
    def record_doubles(self):
        return int(self.lib.psb_walker_record_doubles(self.handle))

    def set_option(self, name, value):
        N.check(self.lib.psb_set_option(self.handle, name.encode(), int(value)))

    def set_force_net(self, widths, params, mean, std, predict_after, activation=N.ACT_TANH, omega0=1.0, omega=1.0,
                      gradient=False):
        """psb_set_force_net: `params` is a flat fp64 torch tensor (device or host) or numpy array."""
        d = N.ForceNetDesc()
        d.n_dense = len(widths) - 1 if widths else 0
        for i, w in enumerate(list(widths or ())[: N.MAX_DENSE + 1]):  # longer lists fail on n_dense in the library
            d.widths[i] = int(w)
        d.activation, d.omega0, d.omega = int(activation), float(omega0), float(omega)
        d.predict_after = int(predict_after)
        d.gradient = int(bool(gradient))
        keep = None
        if d.n_dense:
            if isinstance(params, torch.Tensor):
                keep = params.detach().to(torch.float64).contiguous()
                d.params_on_device = int(keep.is_cuda)
                d.params = keep.data_ptr()
            else:
                keep = np.ascontiguousarray(params, dtype=np.float64)
                d.params = keep.ctypes.data
            host = lambda v: np.asarray(v.detach().cpu() if isinstance(v, torch.Tensor) else v, dtype=np.float64).ravel()
            mean, std = host(mean), host(std)  # (python floats must not pass through torch's float32 default)
            for i in range(self.k):  # scalars (ANN: one std for every component) are broadcast
                d.mean[i], d.std[i] = float(mean[i % mean.size]), float(std[i % std.size])
        N.check(self.lib.psb_set_force_net(self.handle, ctypes.byref(d), self._stream()))
        self._force_net_params = keep  # a device source must outlive the stream-ordered copy

    def set_force_series(self, kind, exponents, values, predict_after, oob_zero=True):
        """psb_set_force_series: `values` = [scale, coefficients...], a torch tensor (device or host) or array;
        kind None removes the series."""
        d = N.ForceSeriesDesc()
        keep = None
        if kind is not None:
            ex = np.ascontiguousarray(exponents, dtype=np.int32)
            d.kind, d.n_terms = int(kind), int(ex.shape[0])
            d.oob_zero, d.predict_after = int(bool(oob_zero)), int(predict_after)
            d.exponents = ex.ctypes.data
            if isinstance(values, torch.Tensor):
                vals = values.detach().to(torch.float64).contiguous()
                d.values_on_device = int(vals.is_cuda)
                d.values = vals.data_ptr()
            else:
                vals = np.ascontiguousarray(values, dtype=np.float64)
                d.values = vals.ctypes.data
            keep = (ex, vals)
        N.check(self.lib.psb_set_force_series(self.handle, ctypes.byref(d), self._stream()))
        self._force_series = keep  # a device source must outlive the stream-ordered copy

    @property
    def wants_tags(self):
        """True when snapshots that carry the engine's slot -> tag array step faster (psb_wants_tags)."""
        return bool(self.lib.psb_wants_tags(self.handle))

    def launch_info(self):
        g, b, c, nbytes = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
        N.check(self.lib.psb_launch_info(self.handle, ctypes.byref(g), ctypes.byref(b), ctypes.byref(c),
                                         ctypes.byref(nbytes)))
        return dict(grid_blocks=g.value, block_threads=b.value, cooperative=bool(c.value),
                    algorithmic_bytes_per_step=nbytes.value, launches=int(self.lib.psb_launch_count(self.handle)),
                    cuda_graph=bool(self.lib.psb_cycle_uses_graph(self.handle)))

    # ---- state ---------------------------------------------------------------------------
    def view(self, field, shape=None):
        """Zero-copy torch view of a state field (None when the method has no such field)."""
        if field in self._views:
            return self._views[field]
        nbytes, code, ptr = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_void_p()
        st = self.lib.psb_state_info(self.handle, field.encode(), ctypes.byref(nbytes), ctypes.byref(code),
                                     ctypes.byref(ptr))
        if st == N.ERR_KEY:
            self._views[field] = None
            return None
        N.check(st)
        itemsize = 4 if code.value == 2 else 8
        n = nbytes.value // itemsize
        buf = _DeviceBuffer(ptr.value, (n,), _TYPESTR[code.value], self)
        t = torch.as_tensor(buf, device=self.device)
        if shape is not None:
            t = t.reshape(shape)
        self._views[field] = t
        return t

    def set_state(self, field, value):
        t = self.view(field)
        if t is None:
            raise KeyError(field)
        src = torch.as_tensor(np.asarray(value), device=self.device).to(t.dtype).reshape(t.shape)
        t.copy_(src)
        if field == "ncalls":
            n = np.asarray(int(np.asarray(value).reshape(-1)[0]), dtype=np.int64)
            N.check(self.lib.psb_set_state(self.handle, b"ncalls", n.ctypes.data_as(ctypes.c_void_p), 8, self._stream()))
            self.ncalls = int(n)


# --------------------------------------------------------------------------------------------
# base classes (methods/core.py:81-153)
# --------------------------------------------------------------------------------------------
class SamplingMethod:
    """methods/core.py:81-116"""

    snapshot_flags = set()
    _kind = N.METHOD_UNBIASED

    def __init__(self, cvs, **kwargs):
        cvs = list(cvs)
        if not cvs:
            raise ValueError("at least one collective variable is required")
        for cv in cvs:
            if not isinstance(cv, _colvars.CollectiveVariable):
                raise TypeError(f"{cv!r} is not a CollectiveVariable")
        self.cvs = cvs
        self.requires_box_unwrapping = any(cv.requires_box_unwrapping for cv in cvs)
        self.kwargs = kwargs
        self.dense_bias = bool(kwargs.get("dense_bias", False))

    @property
    def ncomponents(self):
        return sum(cv.ncomponents for cv in self.cvs)

    # -- lowering to psb_method_desc --------------------------------------------------------
    def _base_desc(self, helpers):
        d = N.MethodDesc()
        d.kind = self._kind
        for i in range(N.MAX_K):
            d.periods[i] = math.inf
            d.sigma[i] = 1.0
            d.r_lower[i], d.r_upper[i] = -math.inf, math.inf
        d.force_unit_factor = float(helpers.layout.force_unit_factor)
        d.grid.kind = N.GRID_NONE
        return d

    def describe(self, helpers):
        return self._base_desc(helpers)

    def _set_grid(self, d):
        grid = getattr(self, "grid", None)
        if grid is not None:
            d.grid = grid.describe()
        restraints = getattr(self, "restraints", None)
        if restraints is not None:
            d.has_restraints = 1
            for i in range(len(restraints.lower)):
                d.r_lower[i] = float(restraints.lower[i])
                d.r_upper[i] = float(restraints.upper[i])
                d.r_kl[i] = float(restraints.kl[i])
                d.r_ku[i] = float(restraints.ku[i])

    # -- build contract ---------------------------------------------------------------------
    _state_type = UnbiasedState

    def _make_state(self, native):
        k = native.k
        bias = native.view("bias", (native.natoms, 3)) if native.dense_outputs else None
        return self._state_type(native.view("xi", (1, k)), bias, native.ncalls)

    def build(self, snapshot, helpers, *args, **kwargs):
        """methods/core.py:108-116 -> (snapshot, initialize, update)"""
        native = NativeStep(self, snapshot, helpers, dense_outputs=self.dense_bias)
        self.native = native  # most recent handle, for introspection (launch_info, tests)

        def initialize():
            native.evaluate(snapshot)  # xi of the initial snapshot, e.g. harmonic_bias.py:119-122
            return self._make_state(native)

        def update(snapshot, state, timestep=0):
            native.step(snapshot, timestep)
            return self._make_state(native)

        update.native = native
        update.fused = True  # nothing but the native step: adapters may call native.step_desc directly
        return snapshot, initialize, update


class GriddedSamplingMethod(SamplingMethod):
    """methods/core.py:119-140"""

    def __init__(self, cvs, grid, **kwargs):
        check_dims(cvs, grid)
        super().__init__(cvs, **kwargs)
        self.grid = grid
        self.restraints = canonicalize(kwargs.get("restraints", None), self.cvs)


def check_dims(cvs, grid):
    # methods/core.py:435-443
    if grid is not None and len(list(cvs)) != grid.shape.size:
        raise ValueError("Grid and Collective Variable dimensions must match.")


# --------------------------------------------------------------------------------------------
# concrete methods
# --------------------------------------------------------------------------------------------
class Unbiased(SamplingMethod):
    """methods/unbiased.py:36-75: evaluates the CVs every step, applies no bias."""

    snapshot_flags = {"positions", "indices"}
    _kind = N.METHOD_UNBIASED


class Bias(SamplingMethod):
    """methods/bias.py:17-73"""

    snapshot_flags = {"positions", "indices"}

    def __init__(self, cvs, center, **kwargs):
        super().__init__(cvs, **kwargs)
        self.cv_dimension = len(self.cvs)
        self.center = center

    @property
    def center(self):
        return self._center

    @center.setter
    def center(self, center):
        center = np.asarray(center, dtype=np.float64)
        if center.shape == ():
            center = center.reshape(1)
        if len(center.shape) != 1 or center.shape[0] != self.cv_dimension:
            raise RuntimeError(f"Invalid center shape expected {self.cv_dimension} got {center.shape}.")
        self._center = center


class HarmonicBias(Bias):
    """methods/harmonic_bias.py:42-132: F = K (xi - center), bias = -J^T F."""

    _kind = N.METHOD_HARMONIC
    _state_type = HarmonicBiasState

    def __init__(self, cvs, kspring, center, **kwargs):
        super().__init__(cvs, center, **kwargs)
        self.kspring = kspring

    @property
    def kspring(self):
        return self._kspring

    @kspring.setter
    def kspring(self, kspring):
        # harmonic_bias.py:78-107
        kspring = np.asarray(kspring, dtype=np.float64)
        shape = kspring.shape
        n = self.cv_dimension
        if len(shape) > 2:
            raise RuntimeError(f"Wrong kspring shape {shape} (expected scalar, 1D or 2D)")
        if len(shape) == 2:
            if shape != (n, n):
                raise RuntimeError(f"2D kspring with wrong shape, expected ({n}, {n}), got {shape}")
            if not np.allclose(kspring, kspring.T):
                raise RuntimeError("Spring matrix is not symmetric")
            self._kspring = kspring
        else:
            if kspring.size not in (n, 1):
                raise RuntimeError(f"Wrong kspring size, expected 1 or {n}, got {kspring.size}.")
            self._kspring = np.identity(n) * kspring

    def describe(self, helpers):
        d = self._base_desc(helpers)
        n = self.cv_dimension
        if self.ncomponents != n:
            raise ValueError("HarmonicBias needs scalar collective variables")
        for a in range(n):
            d.center[a] = float(self._center[a])
            for b in range(n):
                d.kspring[a * N.MAX_K + b] = float(self._kspring[a, b])
        return d


class Metadynamics(SamplingMethod):
    """methods/metad.py:72-205."""

    snapshot_flags = {"positions", "indices"}
    _kind = N.METHOD_METAD
    _state_type = MetadynamicsState

    def __init__(self, cvs, height, sigma, stride, ngaussians, deltaT=None, **kwargs):
        if deltaT is not None and "kB" not in kwargs:  # metad.py:133-138
            raise KeyError(
                "For well-tempered Metadynamics a keyword argument `kB` for "
                "the value of the Boltzmann constant (that matches the "
                "internal units of the backend) must be provided."
            )
        kwargs["grid"] = kwargs.get("grid", None)
        kwargs["restraints"] = kwargs.get("restraints", None)
        check_dims(cvs, kwargs["grid"])
        super().__init__(cvs, **kwargs)
        self.grid = kwargs["grid"]
        self.restraints = canonicalize(kwargs["restraints"], self.cvs)
        self.height = height
        self.sigma = sigma
        self.stride = stride
        self.ngaussians = ngaussians
        self.deltaT = deltaT
        self.kB = kwargs.get("kB", None)
        self.shared_hills = bool(kwargs.get("shared_hills", False))

    def describe(self, helpers):
        d = self._base_desc(helpers)
        k = self.ncomponents
        sigma = np.asarray(self.sigma, dtype=np.float64).flatten()
        if sigma.size == 1:
            sigma = np.repeat(sigma, k)
        if sigma.size != k:
            raise ValueError("sigma must have one entry per collective variable")
        d.height = float(self.height)
        periods = _colvars.get_periods(self.cvs)
        for i in range(k):
            d.sigma[i] = float(sigma[i])
            d.periods[i] = periods[i]
        d.stride = int(self.stride)
        d.ngaussians = int(self.ngaussians)
        d.well_tempered = int(self.deltaT is not None)
        d.shared_hills = int(self.shared_hills)
        d.deltaT = float(self.deltaT) if self.deltaT is not None else 0.0
        d.kB = float(self.kB) if self.kB is not None else 0.0
        self._set_grid(d)
        return d

    def build(self, snapshot, helpers, *args, **kwargs):
        """With `shared_hills=True` (multiple walkers, one process per walker / GPU under torchrun) every
        deposit step becomes psb_walker_begin -> all-gather of the hill records -> psb_walker_finish;
        with a single process the exchange is the identity and the result equals plain metadynamics."""
        snapshot, initialize, update = super().build(snapshot, helpers, *args, **kwargs)
        if not self.shared_hills:
            return snapshot, initialize, update
        from . import parallel

        native = update.native
        if native.host_buffers:
            raise NotImplementedError("shared-hill walkers need device-resident engine arrays")
        engine = parallel.NativeWalkerEngine(native)
        exchange = kwargs.get("exchange") or parallel.RecordExchange()

        def walker_update(snapshot, state, timestep=0):
            parallel.walker_step(engine, snapshot, exchange, timestep)
            return self._make_state(native)

        walker_update.native = native
        walker_update.exchange = exchange
        return snapshot, initialize, walker_update

    def _make_state(self, native):
        k = native.k
        bias = native.view("bias", (native.natoms, 3)) if native.dense_outputs else None
        gshape = tuple(int(s) for s in self.grid.shape) if self.grid is not None else None
        gp = native.view("grid_potential", gshape) if gshape else None
        gg = native.view("grid_gradient", (*gshape, k)) if gshape else None
        return MetadynamicsState(
            native.view("xi", (1, k)), bias, native.view("heights"), native.view("centers", (-1, k)),
            native.view("sigmas", (1, -1))[:, :k], gp, gg, native.view("idx", ()), native.ncalls,
        )


class ABF(GriddedSamplingMethod):
    """methods/abf.py:75-227."""

    snapshot_flags = {"positions", "indices", "momenta"}
    _kind = N.METHOD_ABF
    _state_type = ABFState

    def __init__(self, cvs, grid, **kwargs):
        super().__init__(cvs, grid, **kwargs)
        self.N = np.asarray(self.kwargs.get("N", 500))
        self.use_pinv = self.kwargs.get("use_pinv", False)

    def describe(self, helpers):
        d = self._base_desc(helpers)
        d.abf_N = int(self.N)
        d.use_pinv = int(bool(self.use_pinv))
        self._set_grid(d)
        return d

    def _make_state(self, native):
        k = native.k
        gshape = tuple(int(s) for s in self.grid.shape)
        bias = native.view("bias", (native.natoms, 3)) if native.dense_outputs else None
        return ABFState(
            native.view("xi", (1, k)), bias, native.view("hist", gshape), native.view("Fsum", (*gshape, k)),
            native.view("force")[:k], native.view("Wp")[:k], native.view("Wp_")[:k], native.ncalls,
        )


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


class NNSamplingMethod(GriddedSamplingMethod):
    """methods/core.py:144-153"""

    def __init__(self, cvs, grid, topology, **kwargs):
        super().__init__(cvs, grid, **kwargs)
        self.topology = topology


class FUNN(NNSamplingMethod):
    """
    methods/funn.py:97-296.  Every step is the fused ABF step (hist / Fsum accumulation, W p solve,
    force scatter); once `ncalls > 2 * train_freq` the kernel takes the generalized force from the
    network handed over with psb_set_force_net (funn.py:190,279-284).  The refit (funn.py:213-249:
    binned mean force -> normalize -> Blackman smoothing -> Levenberg-Marquardt) runs every
    `train_freq` calls on the method's device, before the step that first uses its result.
    """

    snapshot_flags = {"positions", "indices", "momenta"}
    _kind = N.METHOD_ABF

    def __init__(self, cvs, grid, topology, **kwargs):
        from . import ml

        super().__init__(cvs, grid, topology, **kwargs)
        self.N = np.asarray(self.kwargs.get("N", 500))
        self.train_freq = self.kwargs.get("train_freq", 5000)
        dims = grid.shape.size
        self.model = ml.MLP(dims, dims, topology, grid.lower, grid.upper, seed=self.kwargs.get("seed", 0),
                            parameters=self.kwargs.get("initial_params", None))
        default_optimizer = ml.LevenbergMarquardt(reg=ml.L2Regularization(1e-6))
        self.optimizer = self.kwargs.get("optimizer", default_optimizer)
        self.use_pinv = self.kwargs.get("use_pinv", False)

    describe = ABF.describe

    def build(self, snapshot, helpers, *args, **kwargs):
        from . import ml
        from .grids import compute_mesh

        native = NativeStep(self, snapshot, helpers, dense_outputs=self.dense_bias)
        self.native = native
        grid, model, train_freq = self.grid, self.model, int(self.train_freq)
        k = native.k
        gshape = tuple(int(s) for s in grid.shape)
        dev = native.device
        inputs = torch.as_tensor(compute_mesh(grid), dtype=torch.float64, device=dev)
        kernel = ml.blackman_kernel(k, 7)
        padding = "wrap" if grid.is_periodic else "edge"
        fit = ml.build_fitting_function(model, self.optimizer)
        zeros = torch.zeros(k, dtype=torch.float64, device=dev)
        holder = {"nn": ml.NNData(torch.as_tensor(model.parameters, device=dev), zeros, zeros.clone())}

        def push(nn):
            native.set_force_net(model.widths, nn.params, nn.mean, nn.std, predict_after=2 * train_freq)

        def learn():  # funn.py:228-239 on the state as it is before this step's sample is added
            hist = native.view("hist", gshape).to(torch.float64).unsqueeze(-1)
            y = native.view("Fsum", (*gshape, k)) / torch.clamp(hist, min=1.0)
            axes = tuple(range(y.ndim - 1))
            y, mean, std = ml.normalize(y, axes=axes)
            reference = ml.smooth(y, kernel, padding)
            params = fit(holder["nn"].params, inputs, reference)
            return ml.NNData(params, mean, std / reference.std(dim=axes, unbiased=False))

        def make_state():
            bias = native.view("bias", (native.natoms, 3)) if native.dense_outputs else None
            return FUNNState(
                native.view("xi", (1, k)), bias, native.view("hist", gshape), native.view("Fsum", (*gshape, k)),
                native.view("force")[:k], native.view("Wp")[:k], native.view("Wp_")[:k], holder["nn"], native.ncalls,
            )

        def initialize():
            native.evaluate(snapshot)
            push(holder["nn"])  # untrained network: never used before the first fit (funn.py:190-191)
            return make_state()

        def update(snapshot, state, timestep=0):
            ncalls = native.ncalls + 1
            if ncalls > 2 * train_freq and ncalls % train_freq == 1:  # funn.py:190-193
                holder["nn"] = learn()
                push(holder["nn"])
            native.step(snapshot, timestep)
            return make_state()

        def set_network(nn):
            """Restart / tests: install a network (NNData) as if it had just been fitted."""
            holder["nn"] = ml.NNData(*(torch.as_tensor(np.asarray(t) if not isinstance(t, torch.Tensor) else t,
                                                       dtype=torch.float64, device=dev) for t in nn))
            push(holder["nn"])

        def restore(prev):
            """Restart (methods/core.py:327-413): push a saved FUNNState into this handle."""
            for name, fld in (("xi", "xi"), ("hist", "hist"), ("Fsum", "Fsum"), ("force", "F"), ("Wp", "Wp"),
                              ("Wp_", "Wp_"), ("ncalls", "ncalls")):
                value = getattr(prev, fld)
                native.set_state(name, value.cpu().numpy() if hasattr(value, "cpu") else value)
            set_network(prev.nn)
            return make_state()

        update.native = native
        update.set_network = set_network
        update.restore = restore
        return snapshot, initialize, update


class ANNState(NamedTuple):  # ann.py:38-62
    xi: Any
    bias: Any
    hist: Any
    phi: Any
    prob: Any
    nn: Any
    ncalls: int = 0


class ANN(NNSamplingMethod):
    """
    methods/ann.py:65-258.  Per step the fused kernel (the histogram-only variant of the ABF instantiation:
    no momenta, no force sums) evaluates the CVs, counts the visit and applies `std * d net / d xi`
    (back-propagation through the network inside the kernel) once `ncalls > train_freq` and xi is inside the
    grid, zero force otherwise.  Every `train_freq` calls the network is refitted on the device to the
    smoothed, normalized free-energy estimate `kT log(prob)` (ann.py:170-217) and the histogram is reset.
    """

    snapshot_flags = {"positions", "indices"}
    _kind = N.METHOD_ABF

    def __init__(self, cvs, grid, topology, kT, **kwargs):
        from numbers import Real

        from . import ml

        assert isinstance(kT, Real)  # ann.py:110
        super().__init__(cvs, grid, topology, **kwargs)
        self.kT = kT
        self.train_freq = self.kwargs.get("train_freq", 5000)
        dims = grid.shape.size
        self.model = ml.MLP(dims, 1, topology, grid.lower, grid.upper, seed=self.kwargs.get("seed", 0),
                            parameters=self.kwargs.get("initial_params", None))
        self.optimizer = self.kwargs.get("optimizer", ml.LevenbergMarquardt(reg=ml.L2Regularization(1e-6)))

    def describe(self, helpers):
        d = self._base_desc(helpers)
        d.variant = 1  # histogram only
        d.abf_N = 1
        grid = self.grid
        d.grid = grid.describe()
        return d

    def build(self, snapshot, helpers, *args, **kwargs):
        from . import ml
        from .grids import compute_mesh

        native = NativeStep(self, snapshot, helpers, dense_outputs=self.dense_bias)
        self.native = native
        grid, model, train_freq, kT = self.grid, self.model, int(self.train_freq), float(self.kT)
        k = native.k
        gshape = tuple(int(s) for s in grid.shape)
        shape = gshape if k > 1 else (*gshape, 1)
        dev = native.device
        inputs = torch.as_tensor(compute_mesh(grid), dtype=torch.float64, device=dev)
        kernel = ml.blackman_kernel(k, 7)
        padding = "wrap" if grid.is_periodic else "edge"
        fit = ml.build_fitting_function(model, self.optimizer)
        one = lambda v: torch.tensor(float(v), dtype=torch.float64, device=dev)
        holder = dict(nn=ml.NNData(torch.as_tensor(model.parameters, device=dev), one(0.0), one(1.0)),
                      phi=torch.zeros(shape, dtype=torch.float64, device=dev),
                      prob=torch.ones(shape, dtype=torch.float64, device=dev))

        def push(nn):
            native.set_force_net(model.widths, nn.params, [0.0], nn.std.reshape(1), predict_after=train_freq,
                                 gradient=True)

        def smooth(y):  # ann.py:189: conv if dims > 1 else vmap(conv)(y.T).T
            return ml.convolve(y, kernel, padding) if k > 1 else ml.convolve(y[:, 0], kernel, padding)[:, None]

        def learn():  # ann.py:194-211 on the histogram as it is before this step's visit is counted
            hist = native.view("hist", gshape)
            prob = holder["prob"] + hist.to(torch.float64).reshape(shape) * torch.exp(holder["phi"] / kT)
            phi = kT * torch.log(prob)
            y, mean, std = ml.normalize(phi)
            reference = smooth(y)
            params = fit(holder["nn"].params, inputs, reference)
            nn = ml.NNData(params, mean, std / reference.std(unbiased=False))
            phi = nn.std * model.apply(params, inputs).reshape(phi.shape)
            holder.update(nn=nn, prob=prob, phi=phi - phi.min())
            hist.zero_()

        def make_state():
            bias = native.view("bias", (native.natoms, 3)) if native.dense_outputs else None
            return ANNState(native.view("xi", (1, k)), bias, native.view("hist", gshape), holder["phi"], holder["prob"],
                            holder["nn"], native.ncalls)

        def initialize():
            native.evaluate(snapshot)
            push(holder["nn"])  # untrained network: not used before the first fit (ann.py:155-158)
            return make_state()

        def update(snapshot, state, timestep=0):
            ncalls = native.ncalls + 1
            if ncalls > train_freq and ncalls % train_freq == 1:
                learn()
                push(holder["nn"])
            native.step(snapshot, timestep)
            return make_state()

        def restore(prev):
            to_t = lambda v: torch.as_tensor(np.asarray(v.cpu() if hasattr(v, "cpu") else v), dtype=torch.float64,
                                             device=dev)
            for name in ("xi", "hist", "ncalls"):
                value = getattr(prev, name)
                native.set_state(name, value.cpu().numpy() if hasattr(value, "cpu") else value)
            holder.update(phi=to_t(prev.phi).reshape(shape), prob=to_t(prev.prob).reshape(shape),
                          nn=ml.NNData(to_t(prev.nn.params), to_t(prev.nn.mean), to_t(prev.nn.std)))
            push(holder["nn"])
            return make_state()

        update.native = native
        update.restore = restore
        return snapshot, initialize, update


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


class CFF(NNSamplingMethod):
    """
    methods/cff.py:109-350.  Every step is the fused ABF step plus a second histogram (`variant = 2`); past
    `train_freq` calls the kernel takes the force from the force network (psb_set_force_net, cff.py:328-332).
    Every `train_freq` calls, on the method's device (cff.py:249-313): frequency-based free energy
    `kT log(max(1, prob))`, binned mean force, normalisation, Blackman smoothing, an energy network fitted with
    the Sobolev loss (values and input-gradients) and a force network fitted to the mean force.
    """

    snapshot_flags = {"positions", "indices", "momenta"}
    _kind = N.METHOD_ABF

    def __init__(self, cvs, grid, topology, kT, **kwargs):
        from numbers import Real

        from . import ml

        assert isinstance(kT, Real)  # cff.py:151
        super().__init__(cvs, grid, topology, **kwargs)
        self.kT = kT
        self.N = np.asarray(self.kwargs.get("N", 500))
        self.train_freq = self.kwargs.get("train_freq", 5000)
        dims = grid.shape.size
        seed = self.kwargs.get("seed", 0)
        reg = ml.L2Regularization(1e-4)
        self.model = ml.MLP(dims, 1, topology, grid.lower, grid.upper, seed=seed)
        self.fmodel = ml.MLP(dims, dims, topology, grid.lower, grid.upper, seed=seed)
        self.optimizer = self.kwargs.get("optimizer", ml.LevenbergMarquardt(loss=ml.Sobolev1SSE(), reg=reg, max_iters=1000))
        self.foptimizer = self.kwargs.get("foptimizer", ml.LevenbergMarquardt(reg=reg))
        self.use_pinv = self.kwargs.get("use_pinv", False)

    def describe(self, helpers):
        d = ABF.describe(self, helpers)
        d.variant = 2  # ABF + histp
        return d

    def build(self, snapshot, helpers, *args, **kwargs):
        from . import ml
        from .grids import compute_mesh

        native = NativeStep(self, snapshot, helpers, dense_outputs=self.dense_bias)
        self.native = native
        grid, model, fmodel = self.grid, self.model, self.fmodel
        train_freq, kT = int(self.train_freq), float(self.kT)
        k = native.k
        gshape = tuple(int(s) for s in grid.shape)
        shape = gshape if k > 1 else (*gshape, 1)
        dev = native.device
        f32 = lambda t: t.to(torch.float32).to(torch.float64)
        inputs = f32(torch.as_tensor(compute_mesh(grid), dtype=torch.float64, device=dev))  # cff.py:263
        kernel = ml.blackman_kernel(k, 7).astype(np.float32).astype(np.float64)
        padding = "wrap" if grid.is_periodic else "edge"
        fit = ml.build_fitting_function(model, self.optimizer)
        ffit = ml.build_fitting_function(fmodel, self.foptimizer)
        one = lambda v: torch.tensor(float(v), dtype=torch.float64, device=dev)
        holder = dict(
            nn=ml.NNData(torch.as_tensor(model.parameters, device=dev), one(0.0), one(1.0)),
            fnn=ml.NNData(torch.as_tensor(fmodel.parameters, device=dev), torch.zeros(k, dtype=torch.float64, device=dev), one(1.0)),
            prob=torch.zeros(shape, dtype=torch.float64, device=dev),
            fe=torch.zeros(shape, dtype=torch.float64, device=dev),
        )

        def push(fnn):
            native.set_force_net(fmodel.widths, fnn.params, fnn.mean, fnn.std.reshape(1), predict_after=train_freq)

        def vsmooth(y):
            return ml.smooth(y, kernel, padding)

        def smooth(y):
            return ml.convolve(y, kernel, padding) if k > 1 else vsmooth(y)

        def learn():  # cff.py:280-307 on the state as it is before this step's sample is added
            histp = native.view("histp", gshape)
            prob = holder["prob"] + histp.to(torch.float64).reshape(shape) * torch.exp(holder["fe"] / kT)
            fe = kT * torch.log(torch.clamp(prob, min=1.0))
            hist = native.view("hist", gshape).to(torch.float64).unsqueeze(-1)
            force = native.view("Fsum", (*gshape, k)) / torch.clamp(hist, min=1.0)
            axes = tuple(range(force.ndim - 1))
            dy, dy_mean, dy_std = ml.normalize(force, axes=axes)
            s = torch.maximum(fe.std(unbiased=False), dy_std.max())
            y = smooth(f32(fe / s))
            dy = vsmooth(f32(dy * dy_std / s))
            params = fit(holder["nn"].params, inputs, (y, dy))
            fparams = ffit(holder["fnn"].params, inputs, dy)
            nn, fnn = ml.NNData(params, holder["nn"].mean, s), ml.NNData(fparams, dy_mean, s)
            fe = nn.std * model.apply(params, inputs).reshape(fe.shape)
            holder.update(nn=nn, fnn=fnn, prob=prob, fe=fe - fe.min())
            histp.zero_()

        def make_state():
            bias = native.view("bias", (native.natoms, 3)) if native.dense_outputs else None
            return CFFState(
                native.view("xi", (1, k)), bias, native.view("hist", gshape), native.view("histp", gshape), holder["prob"],
                holder["fe"], native.view("Fsum", (*gshape, k)), native.view("force")[:k], native.view("Wp")[:k],
                native.view("Wp_")[:k], holder["nn"], holder["fnn"], native.ncalls,
            )

        def initialize():
            native.evaluate(snapshot)
            push(holder["fnn"])  # untrained network: not used before the first fit (cff.py:226-228)
            return make_state()

        def update(snapshot, state, timestep=0):
            ncalls = native.ncalls + 1
            if ncalls > train_freq and ncalls % train_freq == 1:
                learn()
                push(holder["fnn"])
            native.step(snapshot, timestep)
            return make_state()

        def restore(prev):
            to_t = lambda v: torch.as_tensor(np.asarray(v.cpu() if hasattr(v, "cpu") else v), dtype=torch.float64,
                                             device=dev)
            for name in ("xi", "hist", "histp", "Fsum", "force", "Wp", "Wp_", "ncalls"):
                value = getattr(prev, name)
                native.set_state(name, value.cpu().numpy() if hasattr(value, "cpu") else value)
            holder.update(prob=to_t(prev.prob).reshape(shape), fe=to_t(prev.fe).reshape(shape),
                          nn=ml.NNData(*(to_t(t) for t in prev.nn)), fnn=ml.NNData(*(to_t(t) for t in prev.fnn)))
            push(holder["fnn"])
            return make_state()

        update.native = native
        update.restore = restore
        return snapshot, initialize, update


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


class Sirens(NNSamplingMethod):
    """
    methods/sirens.py:95-420.  Every step is the fused ABF step (mode "cff": plus `histp`); past `train_freq`
    calls the kernel back-propagates through the staged Siren network and applies `nn.std * d net / d xi`
    (psb_set_force_net with sine layers and `gradient = 1`; restraints outside the grid first).  Every
    `train_freq` calls the free-energy network is refitted on the device: gradients loss on the binned mean
    force (mode "abf") or Sobolev loss on force + frequency-based free energy (mode "cff").
    """

    snapshot_flags = {"positions", "indices", "momenta"}
    _kind = N.METHOD_ABF

    def __init__(self, cvs, grid, topology, **kwargs):
        from numbers import Real

        from . import ml

        mode = kwargs.get("mode", "abf")
        kT = kwargs.get("kT", None)
        loss = ml.GradientsSSE() if mode == "abf" else ml.Sobolev1SSE()
        optimizer = kwargs.get("optimizer", ml.LevenbergMarquardt(loss=loss, reg=ml.L2Regularization(1e-4), max_iters=250))
        # sirens.py:181-196
        if mode not in ("abf", "cff"):
            raise ValueError(f"Invalid mode {mode}. Possible values are 'abf' and 'cff'")
        if mode == "cff":
            if kT is None:
                raise KeyError("When running in 'cff' mode, a keyword argument `kT` in the same "
                               "units as the backend internal energy units must be provided.")
            assert isinstance(kT, Real)
            assert isinstance(optimizer.loss, ml.Sobolev1SSE)
        else:
            assert isinstance(optimizer.loss, ml.GradientsSSE)
        super().__init__(cvs, grid, topology, **kwargs)
        self.mode, self.kT = mode, kT
        self.N = np.asarray(self.kwargs.get("N", 500))
        self.train_freq = self.kwargs.get("train_freq", 5000)
        self.model = ml.Siren(grid.shape.size, 1, topology, grid.lower, grid.upper, omega_0=self.kwargs.get("omega_0", 16),
                              omega=self.kwargs.get("omega", 1), seed=self.kwargs.get("seed", 0))
        self.optimizer = optimizer
        self.use_pinv = self.kwargs.get("use_pinv", False)

    def describe(self, helpers):
        d = ABF.describe(self, helpers)
        d.variant = 2 if self.mode == "cff" else 0
        return d

    def build(self, snapshot, helpers, *args, **kwargs):
        from . import ml
        from .grids import compute_mesh

        native = NativeStep(self, snapshot, helpers, dense_outputs=self.dense_bias)
        self.native = native
        grid, model, mode = self.grid, self.model, self.mode
        train_freq = int(self.train_freq)
        kT = None if self.kT is None else float(self.kT)
        k = native.k
        gshape = tuple(int(s) for s in grid.shape)
        shape = gshape if k > 1 else (*gshape, 1)
        dev = native.device
        f32 = lambda t: t.to(torch.float32).to(torch.float64)
        inputs = f32(torch.as_tensor(compute_mesh(grid), dtype=torch.float64, device=dev))  # sirens.py:290
        kernel = ml.blackman_kernel(k, 5).astype(np.float32).astype(np.float64)  # sirens.py:291
        padding = "wrap" if grid.is_periodic else "edge"
        fit = ml.build_fitting_function(model, self.optimizer)
        one = lambda v: torch.tensor(float(v), dtype=torch.float64, device=dev)
        holder = dict(nn=ml.NNData(torch.as_tensor(model.parameters, device=dev), one(0.0), one(1.0)), prob=None, fe=None)
        if mode == "cff":
            holder.update(prob=torch.zeros(shape, dtype=torch.float64, device=dev),
                          fe=torch.zeros(shape, dtype=torch.float64, device=dev))

        def push(nn):
            native.set_force_net(model.widths, nn.params, [0.0], nn.std.reshape(1), predict_after=train_freq,
                                 activation=N.ACT_SIN, omega0=model.omega_0, omega=model.omega, gradient=True)

        def smooth(y):
            return ml.convolve(y, kernel, padding) if k > 1 else ml.smooth(y, kernel, padding)

        def normalize_gradients(data):  # sirens.py:399-417
            axes = tuple(range(data.ndim - 1))
            std = data.std(dim=axes, unbiased=False).max()
            return (data - data.mean(dim=axes), std) if grid.is_periodic else (data, std)

        def learn():  # sirens.py:340-347 on the state as it is before this step's sample is added
            hist = native.view("hist", gshape).to(torch.float64).unsqueeze(-1)
            force = native.view("Fsum", (*gshape, k)) / torch.clamp(hist, min=1.0)
            if mode == "cff":
                histp = native.view("histp", gshape)
                prob = holder["prob"] + histp.to(torch.float64).reshape(shape) * torch.exp(holder["fe"] / kT)
                fe = kT * torch.log(torch.clamp(prob, min=1.0))
                dy, dy_std = normalize_gradients(force)
                s = torch.maximum(fe.std(unbiased=False), dy_std)
                params = fit(holder["nn"].params, inputs, (smooth(f32(fe / s)), dy / s))
                nn = ml.NNData(params, holder["nn"].mean, s)
                fe = nn.std * model.apply(params, inputs).reshape(fe.shape)
                holder.update(nn=nn, prob=prob, fe=fe - fe.min())
                histp.zero_()
            else:
                dy, s = normalize_gradients(force)
                holder["nn"] = ml.NNData(fit(holder["nn"].params, inputs, dy / s), holder["nn"].mean, s)

        def make_state():
            bias = native.view("bias", (native.natoms, 3)) if native.dense_outputs else None
            histp = native.view("histp", gshape) if mode == "cff" else None
            return SirensState(
                native.view("xi", (1, k)), bias, native.view("hist", gshape), histp, holder["prob"], holder["fe"],
                native.view("Fsum", (*gshape, k)), native.view("force")[:k], native.view("Wp")[:k],
                native.view("Wp_")[:k], holder["nn"], native.ncalls,
            )

        def initialize():
            native.evaluate(snapshot)
            push(holder["nn"])  # untrained network: not used before the first fit (sirens.py:255-258)
            return make_state()

        def update(snapshot, state, timestep=0):
            ncalls = native.ncalls + 1
            if ncalls > train_freq and ncalls % train_freq == 1:
                learn()
                push(holder["nn"])
            native.step(snapshot, timestep)
            return make_state()

        def restore(prev):
            to_t = lambda v: torch.as_tensor(np.asarray(v.cpu() if hasattr(v, "cpu") else v), dtype=torch.float64,
                                             device=dev)
            names = ("xi", "hist", "Fsum", "force", "Wp", "Wp_", "ncalls") + (("histp",) if mode == "cff" else ())
            for name in names:
                value = getattr(prev, name)
                native.set_state(name, value.cpu().numpy() if hasattr(value, "cpu") else value)
            if mode == "cff":
                holder.update(prob=to_t(prev.prob).reshape(shape), fe=to_t(prev.fe).reshape(shape))
            holder["nn"] = ml.NNData(*(to_t(t) for t in prev.nn))
            push(holder["nn"])
            return make_state()

        update.native = native
        update.restore = restore
        return snapshot, initialize, update


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


class Fun(NamedTuple):  # approxfun/core.py:18-27
    scale: Any
    coefficients: Any
    c0: Any = 0.0


class SpectralABF(GriddedSamplingMethod):
    """
    methods/spectral_abf.py:95-268.  Every step is the fused ABF step; past `fit_threshold` calls the
    kernel takes the force from the gradient of a Fourier (periodic grid) or monomial (grid converted to
    Chebyshev, spectral_abf.py:140) series (psb_set_force_series).  The refit -- binned mean force,
    normalisation, one pseudo-inverse product (approxfun/core.py:232-258) -- runs every `fit_freq` calls on
    the method's device and never touches the host.  Outside the grid: restraints, or no force at all.
    """

    snapshot_flags = {"positions", "indices", "momenta"}
    _kind = N.METHOD_ABF

    def __init__(self, cvs, grid, **kwargs):
        from . import approxfun
        from .grids import Chebyshev, Grid, convert

        super().__init__(cvs, grid, **kwargs)
        self.N = np.asarray(self.kwargs.get("N", 500))
        self.fit_freq = self.kwargs.get("fit_freq", 100)
        self.fit_threshold = self.kwargs.get("fit_threshold", 500)
        self.grid = self.grid if self.grid.is_periodic else convert(self.grid, Grid[Chebyshev])
        self.model = approxfun.SpectralGradientFit(self.grid)
        self.use_pinv = self.kwargs.get("use_pinv", False)

    describe = ABF.describe

    def build(self, snapshot, helpers, *args, **kwargs):
        from . import approxfun

        native = NativeStep(self, snapshot, helpers, dense_outputs=self.dense_bias)
        self.native = native
        grid, model = self.grid, self.model
        fit_freq, fit_threshold = int(self.fit_freq), int(self.fit_threshold)
        k = native.k
        gshape = tuple(int(s) for s in grid.shape)
        fit = approxfun.build_fitter(model, native.device)
        kind = N.SERIES_FOURIER if grid.is_periodic else N.SERIES_MONOMIAL
        holder = {}

        def push(values):
            holder["values"] = values
            native.set_force_series(kind, model.exponents, values, predict_after=fit_threshold, oob_zero=True)

        def refit():  # spectral_abf.py:216-219 on the state as it is before this step's sample is added
            return fit.from_state(native.view("hist", gshape), native.view("Fsum", (*gshape, k)))

        def make_state():
            bias = native.view("bias", (native.natoms, 3)) if native.dense_outputs else None
            v = holder["values"]
            return SpectralABFState(
                native.view("xi", (1, k)), bias, native.view("hist", gshape), native.view("Fsum", (*gshape, k)),
                native.view("force")[:k], native.view("Wp")[:k], native.view("Wp_")[:k], Fun(v[0], v[1:], 0.0),
                native.ncalls,
            )

        def initialize():
            native.evaluate(snapshot)
            push(fit(native.view("Fsum", (*gshape, k))))  # spectral_abf.py:178: fit of the empty force sums
            return make_state()

        def update(snapshot, state, timestep=0):
            ncalls = native.ncalls + 1
            if ncalls > fit_threshold and ncalls % fit_freq == 1:  # spectral_abf.py:184-185
                push(refit())
            native.step(snapshot, timestep)
            return make_state()

        def restore(prev):
            for name in ("xi", "hist", "Fsum", "force", "Wp", "Wp_", "ncalls"):
                value = getattr(prev, name)
                native.set_state(name, value.cpu().numpy() if hasattr(value, "cpu") else value)
            to_t = lambda v: torch.as_tensor(np.asarray(v.cpu() if hasattr(v, "cpu") else v), dtype=torch.float64,
                                             device=native.device)
            push(torch.cat([to_t(prev.fun.scale).reshape(1), to_t(prev.fun.coefficients).reshape(-1)]))
            return make_state()

        update.native = native
        update.restore = restore
        return snapshot, initialize, update


class UmbrellaIntegration(SamplingMethod):
    """methods/umbrella_integration.py:37-125: one HarmonicBias window per centre + histograms."""

    def __init__(self, cvs_or_biasers, *args, **kwargs):
        from .utils import HistogramLogger, listify

        first = list(cvs_or_biasers)
        if first and isinstance(first[0], Bias):  # umbrella_integration.py:88-114
            biasers = first
            hist_periods = args[0] if args else kwargs.pop("hist_periods")
            hist_offsets = args[1] if len(args) > 1 else kwargs.pop("hist_offsets", 0)
            cvs = biasers[0].cvs
            for b in biasers[1:]:
                if b.cvs != cvs:
                    raise RuntimeError("Attempted run of UmbrellaSampling with different CVs for the individual biaser.")
            super().__init__(cvs, **kwargs)
            self.submethods = biasers
        else:  # umbrella_integration.py:51-86
            names = ("ksprings", "centers", "hist_periods", "hist_offsets")
            vals = dict(zip(names, args))
            vals.update({n: kwargs.pop(n) for n in names if n in kwargs})
            ksprings, centers = vals["ksprings"], vals["centers"]
            hist_periods, hist_offsets = vals["hist_periods"], vals.get("hist_offsets", 0)
            super().__init__(first, **kwargs)
            replicas = len(centers)
            ksprings = listify(ksprings, replicas, "ksprings", float)
            self.submethods = [HarmonicBias(self.cvs, k, c) for (k, c) in zip(ksprings, centers)]
        replicas = len(self.submethods)
        periods = listify(hist_periods, replicas, "hist_periods", int)
        offsets = listify(hist_offsets, replicas, "hist_offsets", int)
        self.histograms = [HistogramLogger(p, o) for (p, o) in zip(periods, offsets)]

    def build(self):  # pylint: disable=arguments-differ
        """The sampling work is delegated to the HarmonicBias windows (umbrella_integration.py:123-125)."""


# names the reference exports from pysages.methods (methods/__init__.py)
from .utils import HistogramLogger, MetaDLogger, ReplicasConfiguration, SerialExecutor  # noqa: E402,F401
