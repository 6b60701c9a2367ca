"""
TEST INFRASTRUCTURE -- the REAL reference (PySAGES on JAX) as the oracle, wherever it can be imported.

`available()` probes for `jax`, `plum` and the reference package itself (installed without its
dependencies into `baseline/_ref/` by `python -m pip install --no-index --no-build-isolation --no-deps
--target baseline/_ref <copy of /root/reference>`, or importable from `/root/reference` in the build
container).  In this image jax and plum-dispatch are absent from the interpreter and from /opt/wheelhouse,
so every caller (tests/test_reference_live.py, bench.py's reference arm) falls back to the restatement in
`oracle/*.py` and says so; nothing in `pysages_b200/` imports this module.

The driver follows SURVEY.md 8(c): no MD engine is needed to run the reference's hot path --
    Snapshot(jnp arrays)                              pysages/backends/snapshot.py:34-46
    SnapshotMethods(positions, indices, momenta, ...) pysages/backends/hoomd.py:176-202 (HOOMD layout)
    HelperMethods(build_data_querier(...), lambda: 3) pysages/backends/snapshot.py:59-62, 98-107
    _, initialize, update = method.build(snapshot, helpers)   pysages/methods/core.py:108-116, 463-474
    state = update(snapshot, state)                   one jitted step, as Sampler._update_callback runs it
and applies the bias like pysages/backends/hoomd.py:219-230 (`forces[:, :3] += bias`) on a numpy copy.
Precedents inside the reference: tests/test_cvs.py:15-32, backends/ase.py:153-165.
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEARCH = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")

_state = {}


def probe():
    """-> (modules dict or None, reason).  Never raises."""
    if "result" in _state:
        return _state["result"]
    missing = []
    for name in ("jax", "plum"):
        try:
            importlib.import_module(name)
        except Exception as e:  # ImportError, or a broken install
            missing.append(f"{name} ({type(e).__name__})")
    if missing:
        _state["result"] = (None, "not importable: " + ", ".join(missing))
        return _state["result"]
    err = None
    for path in SEARCH:
        if not os.path.isdir(os.path.join(path, "pysages")):
            continue
        sys.path.insert(0, path)
        if not os.path.exists(os.path.join(path, "pysages", "_version.py")):
            # written by setuptools_scm in a git checkout; neither the plain source tree nor the --no-deps install has it
            import types

            stub = types.ModuleType("pysages._version")
            stub.version, stub.version_tuple = "0+unknown", (0, 0, 0)
            sys.modules["pysages._version"] = stub
        try:
            pysages = importlib.import_module("pysages")
            _state["result"] = (dict(pysages=pysages, path=path), f"reference imported from {path}")
            return _state["result"]
        except Exception as e:
            err = f"pysages from {path}: {type(e).__name__}: {e}"
            sys.path.remove(path)
    _state["result"] = (None, err or "reference package not found under " + " or ".join(SEARCH))
    return _state["result"]


def available():
    return probe()[0] is not None


def reason():
    return probe()[1]


# ---- building the reference's objects from the same specs the parity harness uses ---------------------
def _grid(spec):
    import pysages

    lower, upper, shape, kind = spec
    if kind == "periodic":
        return pysages.Grid(lower=lower, upper=upper, shape=shape, periodic=True)
    if kind == "chebyshev":
        from pysages.grids import Chebyshev

        return pysages.Grid[Chebyshev](lower=lower, upper=upper, shape=shape)
    return pysages.Grid(lower=lower, upper=upper, shape=shape)


def make_method(cv_specs, method_spec):
    """(name, args, kwargs) CV specs + (name, kwargs) method spec -> reference method object."""
    import pysages
    from pysages import colvars, methods

    cvs = [getattr(colvars, name)(*args, **kw) for (name, args, kw) in cv_specs]
    name, kw = method_spec
    kw = dict(kw)
    if kw.get("grid") is not None:
        kw["grid"] = _grid(kw["grid"])
    if kw.get("restraints") is not None:
        kw["restraints"] = pysages.CVRestraints(*kw["restraints"])
    if name == "HarmonicBias":
        return methods.HarmonicBias(cvs, kw.pop("kspring"), kw.pop("center"), **kw)
    if name == "ABF":
        return methods.ABF(cvs, kw.pop("grid"), **kw)
    if name == "Metadynamics":
        args = [kw.pop(k) for k in ("height", "sigma", "stride", "ngaussians")]
        return methods.Metadynamics(cvs, *args, kw.pop("deltaT", None), **kw)
    if name == "Unbiased":
        return methods.Unbiased(cvs, **kw)
    raise ValueError(name)


class LiveRun:
    """The reference's `update` stepped over HOOMD-layout snapshots given as numpy arrays
    (pysages_b200.synthetic.make_snapshot format)."""

    def __init__(self, snap, cv_specs, method_spec):
        import jax.numpy as jnp
        from pysages.backends.snapshot import (
            Box,
            HelperMethods,
            Snapshot,
            SnapshotMethods,
            build_data_querier,
        )

        if snap["layout"] != "hoomd":
            raise ValueError("the live driver restates backends/hoomd.py:176-202 only")
        self.jnp = jnp
        self.method = make_method(cv_specs, method_spec)
        a = snap["arrays"]
        self.box = Box(snap["box_H"], snap["origin"])
        self.dt = snap["dt"]
        self._static = {k: jnp.asarray(a[k]) for k in ("vel_mass", "rtags", "images")}
        self.forces = a["forces"].copy()
        unwrap = self.method.requires_box_unwrapping

        # backends/hoomd.py:176-202
        def positions(s):
            if unwrap:
                return s.positions[:, :3] + jnp.diag(s.box.H) * s.extras["images"]
            return s.positions

        def momenta(s):
            return (s.vel_mass[:, 3:] * s.vel_mass[:, :3]).flatten()

        sm = SnapshotMethods(positions, lambda s: s.ids, momenta, lambda s: s.vel_mass[:, 3:])
        self.helpers = HelperMethods(build_data_querier(sm, self.method.snapshot_flags), lambda: 3)
        self._Snapshot = Snapshot
        self.snapshot = self._snapshot(a["positions"])
        _, initialize, self.update = self.method.build(self.snapshot, self.helpers)
        self.state = initialize()

    def _snapshot(self, positions):
        s = self._static
        return self._Snapshot(self.jnp.asarray(positions), s["vel_mass"], self.jnp.asarray(self.forces), s["rtags"],
                              self.box, self.dt, dict(images=s["images"]))

    def step(self, positions=None):
        import numpy as np

        if positions is not None:
            self.snapshot = self._snapshot(positions)
        self.state = self.update(self.snapshot, self.state)
        bias = getattr(self.state, "bias", None)
        if bias is not None:  # backends/hoomd.py:219-230
            self.forces[:, :3] += np.asarray(bias.block_until_ready()).astype(self.forces.dtype)
        return self.state

    def jacobian(self):
        """(k, 3N) Jacobian the reference's autodiff produces (colvars/core.py:209-225)."""
        import numpy as np

        _, J = self.method.cv(self.helpers.query(self.snapshot))  # set by SamplingMethod.__init__ (core.py:97)
        return np.asarray(J)

    def steps_per_second(self, nsteps, warmup=2):
        """Timing leg for bench.py --impl reference (JAX-CPU, after JIT warm-up)."""
        import time

        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            self.step()
        return nsteps / (time.perf_counter() - t0)
