"""
HOOMD-blue adapter (drop-in for pysages/backends/hoomd.py): binds a sampling method to a HOOMD
simulation through the `hoomd.dlext` plugin.  Same public surface as the reference -- `Sampler`
with `state`, `snapshot`, `callback`, `take_snapshot()`, `restore()`, and `bind()` / `detach()` --
but the per-step callback makes ONE native call on the arrays HOOMD exposes (zero copy) instead of
`method_update` (XLA) + `sync_backend` + `block_until_ready` + cupy add (hoomd.py:168-173,219-230).

Requires `hoomd` and `hoomd.dlext` (third-party, not shipped here); importing this module without
them raises ImportError, and a CPU-only HOOMD build is rejected: there is no CPU bias path.
HOOMD launches its kernels on the legacy default stream, which is also torch's current stream in
the stepping thread, so the bias step is ordered after the force computation and before the second
half of the integrator without any host synchronisation.
"""

from warnings import warn

import hoomd
from hoomd import md
from hoomd.dlext import AccessLocation, AccessMode, DLExtSampler, SystemView

from ._common import LAYOUTS, fused_helpers, tensor_from_capsule
from .snapshot import Box, Snapshot, copy
from .snapshot import restore as _restore

SamplerBase = DLExtSampler
CONTEXTS_SAMPLERS = {}

# ---- HOOMD 2.x / 3.x+ API differences (hoomd.py:33-78) -------------------------------------------
if getattr(hoomd, "__version__", "").startswith("2."):

    def is_on_gpu(context):
        return context.on_gpu()

    def get_integrator(context):
        return context.integrator

    def get_run_method(context):
        return hoomd.run

    def get_system(context):
        return context.system

    def set_half_step_hook(context, hook):
        context.integrator.cpp_integrator.setHalfStepHook(hook)

    def remove_half_step_hook(context):
        context.integrator.cpp_integrator.removeHalfStepHook()

else:
    if not hasattr(hoomd.dlext, "__version__"):

        class SamplerBase(DLExtSampler, md.HalfStepHook):  # noqa: F811
            def __init__(self, sysview, update, location, mode):
                md.HalfStepHook.__init__(self)
                DLExtSampler.__init__(self, sysview, update, location, mode)

    def is_on_gpu(context):
        return not isinstance(context.device, hoomd.device.CPU)

    def get_integrator(context):
        return context.operations.integrator

    def get_run_method(context):
        context.run(0)  # make sure the simulation is fully initialised
        return context.run

    def get_system(context):
        return context._cpp_sys

    def set_half_step_hook(context, hook):
        context.operations.integrator.half_step_hook = hook

    def remove_half_step_hook(context):
        context.operations.integrator.half_step_hook = None


def default_location():
    if not hasattr(AccessLocation, "OnDevice"):
        raise RuntimeError("pysages_b200 needs a CUDA-enabled HOOMD-blue / hoomd-dlext build (no CPU bias path)")
    return AccessLocation.OnDevice


class Sampler(SamplerBase):
    """HalfStepHook connecting a pysages_b200 sampling method to a HOOMD-blue simulation."""

    def __init__(self, context, sampling_method, callback=None, location=None):
        location = default_location() if location is None else location
        if not is_on_gpu(context) or location != AccessLocation.OnDevice:
            raise RuntimeError("pysages_b200 runs the bias step on the GPU only: use a hoomd GPU device")
        self.context = context
        self.callback = callback
        self.location = location
        self.view = SystemView(get_system(context))
        self.box = Box(*get_global_box(self.view))
        self.dt = get_timestep(context)
        self.update_box = lambda: self.box  # NOTE: extend for NPT simulations (hoomd.py:123)
        super().__init__(self.view, self._update_callback, location, AccessMode.Read)

        snapshot = self.take_snapshot()
        _, initialize, self._method_update = sampling_method.build(snapshot, fused_helpers(LAYOUTS["hoomd"]))
        self.state = initialize()

    def restore(self, prev_snapshot):
        self.snapshot = prev_snapshot
        self.forward_data(self._restore_callback, self.location, AccessMode.Overwrite, 0)

    def take_snapshot(self):
        self.forward_data(self._snapshot_callback, self.location, AccessMode.Read, 0)
        return self.snapshot

    def _pack_snapshot(self, positions, vel_mass, forces, rtags, images):
        t = tensor_from_capsule
        return Snapshot(t(positions), t(vel_mass), t(forces), t(rtags), self.update_box(), self.dt,
                        dict(images=t(images)))

    # the plugin's argument order differs from Snapshot's field order (hoomd.py:159)
    def _restore_callback(self, positions, vel_mass, rtags, images, forces, _):
        _restore(self._pack_snapshot(positions, vel_mass, forces, rtags, images), self.snapshot)

    def _snapshot_callback(self, positions, vel_mass, rtags, images, forces, _):
        self.snapshot = copy(self._pack_snapshot(positions, vel_mass, forces, rtags, images))

    def _update_callback(self, positions, vel_mass, rtags, images, forces, timestep):
        snapshot = self._pack_snapshot(positions, vel_mass, forces, rtags, images)
        self.state = self._method_update(snapshot, self.state, timestep)  # CVs + bias + in-place force update
        if self.callback:
            self.callback(snapshot, self.state, timestep)


def get_global_box(system_view):
    """hoomd.py:243-253"""
    box = system_view.particle_data.getGlobalBox()
    L = box.getL()
    xy, xz, yz = box.getTiltFactorXY(), box.getTiltFactorXZ(), box.getTiltFactorYZ()
    lo = box.getLo()
    H = ((L.x, xy * L.y, xz * L.z), (0.0, L.y, yz * L.z), (0.0, 0.0, L.z))
    return H, (lo.x, lo.y, lo.z)


def get_timestep(context):
    return get_integrator(context).dt


def bind(sampling_context, callback, **kwargs):
    """hoomd.py:261-271"""
    context = sampling_context.context
    sampler = Sampler(context, sampling_context.method, callback)
    set_half_step_hook(context, sampler)
    CONTEXTS_SAMPLERS[context] = sampler
    sampling_context.run = get_run_method(context)
    return sampler


def detach(context):
    """hoomd.py:274-284"""
    if context in CONTEXTS_SAMPLERS:
        remove_half_step_hook(context)
        del CONTEXTS_SAMPLERS[context]
    else:
        warn("This context has no sampler bound to it.")
