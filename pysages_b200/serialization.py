"""
save / load of simulation results (pysages/serialization.py:44-93), the step either side of the hot
path that checkpoint / restart goes through.

`save` pickles a `Result` whose states and snapshots have been copied to host numpy arrays (the live
objects are zero-copy views of device memory owned by a native handle, which cannot be pickled);
`load` returns the same `Result`, ready for `pysages_b200.run(result, context_generator, timesteps)`,
which pushes every state field back into a fresh handle (`psb_set_state`).
"""

import copy as _copy
import pickle

import numpy as np
import torch

from .backends.snapshot import Snapshot
from .backends.snapshot import copy as _copy_snapshot
from .core import Result


def _to_host(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().copy()
    if isinstance(x, tuple) and hasattr(x, "_fields"):  # State named tuples
        return type(x)(*(_to_host(v) for v in x))
    if isinstance(x, (list, tuple)):
        return type(x)(_to_host(v) for v in x)
    if isinstance(x, dict):
        return {k: _to_host(v) for k, v in x.items()}
    return x


def _portable_method(method):
    """The method without its native handle (and the same for UmbrellaIntegration's windows)."""
    m = _copy.copy(method)
    m.__dict__.pop("native", None)
    if hasattr(m, "submethods"):
        m.submethods = [_portable_method(s) for s in m.submethods]
    return m


def save(result, filename):
    """serialization.py:76-93"""
    if not isinstance(result, Result):
        raise TypeError("Only saving of `Result` objects is supported.")
    snapshots = [(_copy_snapshot(s, to_cpu=True) if isinstance(s, Snapshot) else _to_host(s)) for s in result.snapshots]
    portable = Result(_portable_method(result.method), [_to_host(s) for s in result.states], result.callbacks, snapshots)
    with open(filename, "wb") as io:
        pickle.dump(portable, io)


def load(filename):
    """serialization.py:44-73"""
    with open(filename, "rb") as f:
        result = pickle.load(f)
    if not isinstance(result, Result):
        raise TypeError("Only loading of `Result` objects is supported.")
    return result


def numpyfy(state):
    """Host copy of a State (for analysis / plotting)."""
    return _to_host(state)


__all__ = ["save", "load", "numpyfy", "np"]
