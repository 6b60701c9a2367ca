"""
OpenMM adapter (drop-in for pysages/backends/openmm.py) on top of the `openmm_dlext` plugin.

CUDA platform: the plugin's persistent views are used in place -- positions / velocities (Np, 4),
fixed-point forces (3, Np) int64, atom ids (slot -> atom); the per-step `ids.argsort()` of
openmm.py:106-107 runs on the device.  CPU / Reference platforms (the reference's own
alanine-dipeptide example): the host arrays are staged through the GPU every step
(psb_step_host) -- the bias is still computed on the device, there is no CPU path.
"""

import numpy as np
import openmm_dlext as dlext
from openmm_dlext import DeviceType, Force

from ._common import LAYOUTS, fused_helpers, tensor_from_capsule
from .snapshot import Box, Snapshot, copy
from .snapshot import restore as _restore

try:
    import openmm
    import openmm.unit as unit
except ImportError:  # older releases
    from simtk import openmm, unit  # type: ignore


def is_on_gpu(view):
    return view.device_type() == DeviceType.GPU


class Sampler:
    def __init__(self, context, sampling_method, callback=None):
        force = Force()
        force.add_to(context)  # OpenMM owns the force from here on
        self.context = context
        self.view = force.view(context)
        self.callback = callback
        self.snapshot = self._take_snapshot()
        layout = LAYOUTS["openmm-cuda" if is_on_gpu(self.view) else "openmm-cpu"]
        _, initialize, method_update = sampling_method.build(self.snapshot, fused_helpers(layout))
        self.state = initialize()
        self._method_update = method_update

        def update(timestep=0):
            if is_on_gpu(self.view):
                self.view.synchronize()  # the plugin's stream is not exposed: order after OpenMM's force kernels
            self.state = method_update(self.snapshot, self.state, timestep)
            if self.callback:
                self.callback(self.snapshot, self.state, timestep)

        force.set_callback_in(context, update)

    def restore(self, prev_snapshot):
        _restore(self.snapshot, prev_snapshot)

    def take_snapshot(self):
        return copy(self.snapshot)

    def _take_snapshot(self):
        view, t = self.view, tensor_from_capsule
        positions = t(dlext.positions(view))
        forces = t(dlext.forces(view))
        ids = t(dlext.atom_ids(view))
        velocities = t(dlext.velocities(view))
        if is_on_gpu(view):
            vel_mass = velocities
        else:
            vel_mass = (velocities, t(dlext.inverse_masses(view)).reshape(-1, 1))
        vectors = self.context.getSystem().getDefaultPeriodicBoxVectors()
        a, b, c = (v.value_in_unit(unit.nanometer) for v in vectors)
        H = ((a[0], b[0], c[0]), (a[1], b[1], c[1]), (a[2], b[2], c[2]))
        dt = self.context.getIntegrator().getStepSize() / unit.picosecond
        return Snapshot(positions, vel_mass, forces, ids, Box(np.asarray(H), np.zeros(3)), dt)


def check_integrator(context):
    """openmm.py:183-188"""
    integrator = context.getIntegrator()
    if isinstance(integrator, (openmm.VariableLangevinIntegrator, openmm.VariableVerletIntegrator)):
        raise ValueError("Variable step size integrators are not supported")


def bind(sampling_context, callback, **kwargs):
    """openmm.py:191-199: the context object is an openmm.app.Simulation"""
    simulation = sampling_context.context
    check_integrator(simulation.context)
    sampler = Sampler(simulation.context, sampling_context.method, callback)
    sampling_context.run = simulation.step
    return sampler
