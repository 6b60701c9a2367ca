"""
Engine-independent half of the adapters (hoomd.py, openmm.py, lammps.py, mock.py).

`SamplerCore` carries what PySAGES calls a sampler -- `state`, `snapshot`, `callback`,
`take_snapshot()`, `restore()` (pysages/backends/hoomd.py:93-173, openmm.py:32-89,
lammps.py:37-127) -- around ONE native handle.  Its per-step entry point, `advance()`, is what the
engine hooks call between the force computation and the integrator:

    read the raw device pointers out of the plugin's DLPack capsules (no tensor objects),
    refresh a pre-filled `psb_snapshot` only where a pointer changed,
    one `psb_step` on the engine's stream,
    order the engine after the bias (policy per engine, see `SamplerCore.after_step`),
    user callback (the Snapshot of torch views is only built when somebody asks for it).

The reference does `method_update` (XLA) -> `sync_backend` -> `block_until_ready` -> cupy add ->
`sync_forces` here; everything but the ordering is inside the one native call.
"""

import ctypes

import torch

from .. import _native as N
from .snapshot import HelperMethods, Layout, Snapshot, copy, describe_snapshot
from .snapshot import restore as restore_arrays

LAMMPS_CONVERSION = {"real": 2390.0573615334906, "metal": 1.0364269e-4, "electron": 1.06657236}  # lammps.py:33

LAYOUTS = {
    "hoomd": Layout(N.LAYOUT_HOOMD, "hoomd"),
    "openmm-cuda": Layout(N.LAYOUT_OPENMM_CUDA, "openmm-cuda"),
    "openmm-cpu": Layout(N.LAYOUT_PLAIN3, "openmm-cpu"),
    "lammps": Layout(N.LAYOUT_LAMMPS, "lammps"),
}


def fused_helpers(layout):
    """HelperMethods for the fused native step: the snapshot gather and the in-place force update
    are part of psb_step, so `query` has nothing to return and `bias` nothing left to do."""

    def query(snapshot):
        raise NotImplementedError(
            "the snapshot gather (positions / indices / momenta) is fused into the CUDA step; "
            "there is no separate query stage"
        )

    factor = layout.force_unit_factor
    return HelperMethods(query, lambda: 3, (lambda x: factor * x), layout)


def tensor_from_capsule(capsule):
    """Zero-copy torch view of a DLPack capsule (the plugins hand out one-shot capsules)."""
    from torch.utils.dlpack import from_dlpack

    return from_dlpack(capsule)


# ---- DLPack capsules without tensor objects ---------------------------------------------------------
# dlpack.h (v0.x, what the *-dlext plugins export): DLManagedTensor { DLTensor dl_tensor; void* manager_ctx;
# void (*deleter)(DLManagedTensor*); }.  Reading the struct leaves the capsule unconsumed: its own destructor
# still runs the deleter when the capsule is collected, as the protocol prescribes.
class _DLDevice(ctypes.Structure):
    _fields_ = [("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32)]


class _DLDataType(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16)]


class _DLTensor(ctypes.Structure):
    _fields_ = [
        ("data", ctypes.c_void_p),
        ("device", _DLDevice),
        ("ndim", ctypes.c_int32),
        ("dtype", _DLDataType),
        ("shape", ctypes.POINTER(ctypes.c_int64)),
        ("strides", ctypes.POINTER(ctypes.c_int64)),
        ("byte_offset", ctypes.c_uint64),
    ]


_get_pointer = ctypes.pythonapi.PyCapsule_GetPointer
_get_pointer.restype = ctypes.c_void_p
_get_pointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
KDL_CPU, KDL_CUDA = 1, 2


def capsule_address(capsule):
    """(address of the first element, leading dimension, device type) of an unconsumed DLPack capsule."""
    t = _DLTensor.from_address(_get_pointer(capsule, b"dltensor"))
    return (t.data or 0) + t.byte_offset, (t.shape[0] if t.ndim else 1), t.device.device_type


_u64_at = ctypes.c_uint64.from_address
_OFF_BYTE_OFFSET = _DLTensor.byte_offset.offset


def capsule_data(capsule):
    """Address of the first element only: the per-step path (two 8-byte reads, no struct object)."""
    a = _get_pointer(capsule, b"dltensor")
    return _u64_at(a).value + _u64_at(a + _OFF_BYTE_OFFSET).value


class SamplerCore:
    """The engine-independent sampler (see the module docstring).  Engine adapters derive from it (or own
    one) and provide `_views()` -> Snapshot of zero-copy torch views of the engine's current buffers."""

    #: what to do after the bias step has been enqueued on torch's current stream
    #:   "stream" -- nothing: the engine issues its next kernels on the same stream (HOOMD: legacy default stream)
    #:   "host"   -- block the host until the bias has landed, like the reference's sync_forces()
    #:               (backends/utils.py:24-26 after `forces += bias`, openmm.py:160-170, lammps.py:163-170):
    #:               engines that integrate on a stream of their own (OpenMM's context stream, Kokkos instances)
    #:               cannot be ordered after ours in any cheaper way through the plugin API.
    ordering = "stream"

    def _bind_method(self, sampling_method, layout, callback):
        self.callback = callback
        self.layout = layout
        self.snapshot = self.take_snapshot()
        _, initialize, self._method_update = sampling_method.build(self.snapshot, fused_helpers(layout))
        self._method = sampling_method
        self._native = self._method_update.native
        self._fused = bool(getattr(self._method_update, "fused", False)) and not self._native.host_buffers
        self._desc = N.SnapshotDesc()
        self._desc_key = None
        self._state = initialize()
        self._state_stale = False

    # -- reference-facing attributes ---------------------------------------------------------------
    @property
    def state(self):
        if self._state_stale:  # built on demand: the fused path does not touch Python objects per step
            self._state = self._method._make_state(self._native)
            self._state_stale = False
        return self._state

    @state.setter
    def state(self, value):
        self._state, self._state_stale = value, False

    def take_snapshot(self):
        return copy(self._views())

    def restore(self, prev_snapshot):
        restore_arrays(self._views(), prev_snapshot)
        self.snapshot = prev_snapshot

    # -- per-step driver ---------------------------------------------------------------------------
    def after_step(self):
        if self.ordering == "host" and not self._native.host_buffers:
            torch.cuda.current_stream(self._native.device).synchronize()

    def advance_pointers(self, timestep, key, fill):
        """Fused fast path: `key` identifies the engine's current buffers (tuple of addresses); `fill(desc)`
        writes them into the psb_snapshot when the key changed."""
        if key != self._desc_key:
            fill(self._desc)
            self._desc_key = key
        self.step_current(timestep)

    def step_current(self, timestep):
        """psb_step on the psb_snapshot as it stands, ordering, user callback."""
        self._native.step_desc(self._desc, timestep)
        self._state_stale = True
        if self.ordering == "host":
            self.after_step()
        if self.callback:
            self.callback(self._views(), self.state, timestep)

    def advance_snapshot(self, snapshot, timestep):
        """General path (methods with cadenced host-side work around the step, host-resident engines)."""
        self.state = self._method_update(snapshot, self.state, timestep)
        self.after_step()
        if self.callback:
            self.callback(snapshot, self.state, timestep)

    def describe_into(self, snapshot):
        return describe_snapshot(snapshot, self.layout, self._desc)


__all__ = ["LAMMPS_CONVERSION", "LAYOUTS", "Snapshot", "SamplerCore", "capsule_address", "capsule_data", "fused_helpers",
           "tensor_from_capsule", "KDL_CPU", "KDL_CUDA"]
