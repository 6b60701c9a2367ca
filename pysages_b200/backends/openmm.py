"""
OpenMM adapter on top of the `openmm_dlext` plugin (the role of pysages/backends/openmm.py).

CUDA platform: the plugin's views persist for the life of the context, so the `psb_snapshot` is filled
once -- positions / velocities (Np, 4), fixed-point forces (3, Np) int64, atom ids (slot -> atom; the
per-step `ids.argsort()` of openmm.py:106-107 happens inside the native step) -- and every step is one
native call.  Ordering: OpenMM computes forces and integrates on its context's own stream, which the
plugin does not expose.  Before the step `view.synchronize()` makes OpenMM's force kernels visible; after
it the host waits for the bias kernel, exactly where the reference calls `sync_forces()`
(openmm.py:160-170, backends/utils.py:24-26) -- otherwise the integrator, launched on the other stream,
could read the force buffer before the bias has been added.
CPU / Reference platforms (the reference's own alanine-dipeptide example): the host arrays are staged
through the GPU every step (psb_step_host, synchronous) -- the bias is still computed on the device.
"""

import numpy as np
import openmm_dlext as dlext

from . import _common as C
from .snapshot import Box, Snapshot

try:
    import openmm
    import openmm.unit as unit
except ImportError:  # releases that still live under simtk
    from simtk import openmm, unit  # type: ignore


class Sampler(C.SamplerCore):
    ordering = "host"

    def __init__(self, context, sampling_method, callback=None):
        self._force = dlext.Force()
        self._force.add_to(context)  # OpenMM owns the force object from here on
        self.context = context
        self.view = self._force.view(context)
        self.on_gpu = self.view.device_type() == dlext.DeviceType.GPU
        self._arrays = None
        self._bind_method(sampling_method, C.LAYOUTS["openmm-cuda" if self.on_gpu else "openmm-cpu"], callback)
        self._force.set_callback_in(context, self.on_forces_ready)

    def _views(self):
        """The plugin's buffers as one Snapshot of zero-copy views (taken once: they persist, openmm.py:64-89)."""
        if self._arrays is None:
            t, v = C.tensor_from_capsule, self.view
            velocities = t(dlext.velocities(v))
            vel_mass = velocities if self.on_gpu else (velocities, t(dlext.inverse_masses(v)).reshape(-1, 1))
            cols = [c.value_in_unit(unit.nanometer) for c in self.context.getSystem().getDefaultPeriodicBoxVectors()]
            H = np.asarray(cols, dtype=np.float64).T  # box vectors are the columns of H
            dt = self.context.getIntegrator().getStepSize() / unit.picosecond
            self._arrays = Snapshot(t(dlext.positions(v)), vel_mass, t(dlext.forces(v)), t(dlext.atom_ids(v)),
                                    Box(H, np.zeros(3)), dt)
        return self._arrays

    def on_forces_ready(self, timestep=0):
        views = self._views()
        if self.on_gpu:
            self.view.synchronize()  # OpenMM's force kernels (its own stream) before our reads
        if self._fused:
            self.advance_pointers(timestep, "persistent", lambda d: self.describe_into(views))
        else:
            self.advance_snapshot(views, timestep)


def check_integrator(context):
    """Variable-step integrators change dt behind the method's back (openmm.py:183-188)."""
    if isinstance(context.getIntegrator(), (openmm.VariableLangevinIntegrator, openmm.VariableVerletIntegrator)):
        raise ValueError("Variable step size integrators are not supported")


def bind(sampling_context, callback, **kwargs):
    """SamplingContext entry point; the wrapped object is an openmm.app.Simulation (openmm.py:191-199)."""
    simulation = sampling_context.context
    check_integrator(simulation.context)
    sampler = Sampler(simulation.context, sampling_context.method, callback)
    sampling_context.run = simulation.step
    return sampler
