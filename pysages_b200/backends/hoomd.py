"""
HOOMD-blue adapter: binds a sampling method to a HOOMD simulation through the `hoomd.dlext` plugin
(the role of pysages/backends/hoomd.py; same `bind()` / `detach()` / `Sampler` surface).

Per step the plugin calls `Sampler.on_half_step(positions, vel_mass, rtags, images, forces, timestep)`
with five DLPack capsules (argument order of hoomd.py:159-168).  The adapter does not turn them into
arrays: it reads the five device addresses out of the capsules, refreshes the pre-filled `psb_snapshot`
when HOOMD has moved a buffer (it double-buffers while sorting particles), and makes one native call.
When most of the system is selected the step streams HOOMD's arrays in memory order; it then also asks the
plugin for the slot -> tag array (`hoomd.dlext.tags`, ParticleData's `tag`: the inverse of `rtags`), which
saves the kernel its tag -> slot inversion pass.
HOOMD issues its kernels on the legacy default stream, which is also torch's current stream in the
stepping thread, so the bias lands between the force computation and the second half step by stream
order alone: no `sync_backend()` / `block_until_ready()` / `sync_forces()` (hoomd.py:219-230).

Needs `hoomd` and `hoomd.dlext` (third party, not shipped here); a CPU-only HOOMD build is rejected,
there is no CPU bias path.
"""

from warnings import warn

import hoomd
from hoomd import dlext as _dlext

from . import _common as C
from .snapshot import Box, Snapshot

AccessLocation, AccessMode = _dlext.AccessLocation, _dlext.AccessMode
_tags_of = getattr(_dlext, "tags", None)  # tags(sysview, location, mode) -> DLPack capsule of slot -> tag


class _ApiV2:
    """hoomd 2.x: global `hoomd.run`, the SimulationContext owns system and integrator (hoomd.py:33-56)."""

    @staticmethod
    def on_gpu(sim):
        return sim.on_gpu()

    @staticmethod
    def integrator(sim):
        return sim.integrator

    @staticmethod
    def runner(sim):
        return hoomd.run  # acts on the currently entered context: SamplingContext forwards __enter__/__exit__

    @staticmethod
    def cpp_system(sim):
        return sim.system

    @staticmethod
    def install(sim, hook):
        sim.integrator.cpp_integrator.setHalfStepHook(hook)

    @staticmethod
    def remove(sim):
        sim.integrator.cpp_integrator.removeHalfStepHook()


class _ApiV3:
    """hoomd >= 3: everything hangs off `hoomd.Simulation` (hoomd.py:58-90)."""

    @staticmethod
    def on_gpu(sim):
        return not isinstance(sim.device, hoomd.device.CPU)

    @staticmethod
    def integrator(sim):
        return sim.operations.integrator

    @staticmethod
    def runner(sim):
        sim.run(0)  # attaches operations, so the first real step finds the hook in place
        return sim.run

    @staticmethod
    def cpp_system(sim):
        return sim._cpp_sys

    @staticmethod
    def install(sim, hook):
        sim.operations.integrator.half_step_hook = hook

    @staticmethod
    def remove(sim):
        sim.operations.integrator.half_step_hook = None


API = _ApiV2 if str(getattr(hoomd, "__version__", "3")).startswith("2.") else _ApiV3

# hoomd-dlext releases without a version attribute predate its own HalfStepHook base: the sampler then has to
# derive from hoomd.md.HalfStepHook as well to be accepted by the integrator (hoomd.py:60-67)
_BASES = (_dlext.DLExtSampler,)
if API is _ApiV3 and not hasattr(_dlext, "__version__"):
    _BASES = (_dlext.DLExtSampler, hoomd.md.HalfStepHook)

SAMPLERS = {}            # simulation -> Sampler
CONTEXTS_SAMPLERS = SAMPLERS  # the reference's name for the same table


class Sampler(C.SamplerCore, *_BASES):
    """Half-step hook that runs the fused bias step on HOOMD's own buffers."""

    ordering = "stream"

    def __init__(self, simulation, sampling_method, callback=None, location=None):
        if not hasattr(AccessLocation, "OnDevice"):
            raise RuntimeError("pysages_b200 needs a CUDA-enabled HOOMD-blue / hoomd-dlext build (no CPU bias path)")
        location = AccessLocation.OnDevice if location is None else location
        if location != AccessLocation.OnDevice or not API.on_gpu(simulation):
            raise RuntimeError("pysages_b200 runs the bias step on the GPU only: use a hoomd GPU device")
        self.context = simulation
        self.location = location
        self.view = _dlext.SystemView(API.cpp_system(simulation))
        self.box = Box(*global_box(self.view))
        self.dt = API.integrator(simulation).dt
        self.update_box = lambda: self.box  # constant-volume runs; replace for NPT (hoomd.py:123)
        for base in _BASES[1:]:
            base.__init__(self)
        _dlext.DLExtSampler.__init__(self, self.view, self.on_half_step, location, AccessMode.Read)
        self._last = None
        self._want_tags = _tags_of is not None  # the snapshot the method is built on says so (psb_layout_desc.has_tags)
        self._bind_method(sampling_method, C.LAYOUTS["hoomd"], callback)
        self._want_tags = _tags_of is not None and self._native.wants_tags  # per step: only if the handle uses them

    # ---- plugin callbacks: (positions, vel_mass, rtags, images, forces, timestep) -----------------
    def on_half_step(self, positions, vel_mass, rtags, images, forces, timestep):
        if not self._fused:
            return self.advance_snapshot(self._wrap(positions, vel_mass, rtags, images, forces), timestep)
        d = C.capsule_data
        # the capsule keeps HOOMD's array handle until it is dropped at the end of this call
        tags = _tags_of(self.view, self.location, AccessMode.Read) if self._want_tags else None
        key = (d(positions), d(vel_mass), d(rtags), d(images), d(forces), d(tags) if tags is not None else None)
        self._last = (positions, vel_mass, rtags, images, forces)
        try:
            if key != self._desc_key:  # HOOMD moved a buffer (it double-buffers while sorting): check, then refill
                _, n, dev = C.capsule_address(positions)
                if n != self._native.natoms or dev != C.KDL_CUDA:
                    raise RuntimeError(f"HOOMD now exposes {n} particles on device type {dev}; the method was built "
                                       f"for {self._native.natoms} on the GPU")
                ds = self._desc
                ds.positions, ds.vel_mass, ds.ids, ds.images, ds.forces, ds.tags = key
                H = self.update_box().H
                ds.box_diag[0], ds.box_diag[1], ds.box_diag[2] = float(H[0][0]), float(H[1][1]), float(H[2][2])
                ds.dt = float(self.dt)
                self._desc_key = key
            self.step_current(timestep)
        finally:
            self._last = None  # HOOMD's capsules are only valid inside the half step (hoomd.py:164-166)

    def _wrap(self, positions, vel_mass, rtags, images, forces):
        t = C.tensor_from_capsule
        extras = dict(images=t(images))
        if self._want_tags:
            extras["tags"] = t(_tags_of(self.view, self.location, AccessMode.Read))
        return Snapshot(t(positions), t(vel_mass), t(forces), t(rtags), self.update_box(), self.dt, extras)

    def _views(self):
        """Zero-copy views of HOOMD's current buffers: inside a half step the capsules of that step, otherwise a
        fresh `forward_data` round trip (hoomd.py:141-146)."""
        if self._last is not None:
            caps, self._last = self._last, None
            return self._wrap(*caps)
        out = []
        self.forward_data(lambda p, v, r, i, f, _: out.append(self._wrap(p, v, r, i, f)), self.location,
                          AccessMode.Read, 0)
        return out[0]

    def restore(self, prev_snapshot):
        def overwrite(p, v, r, i, f, _):
            C.restore_arrays(self._wrap(p, v, r, i, f), prev_snapshot)

        self.forward_data(overwrite, self.location, AccessMode.Overwrite, 0)
        self.snapshot = prev_snapshot


def global_box(system_view):
    """Upper-triangular box matrix and origin from HOOMD's BoxDim (hoomd.py:243-253)."""
    dim = system_view.particle_data.getGlobalBox()
    L, lo = dim.getL(), dim.getLo()
    xy, xz, yz = dim.getTiltFactorXY(), dim.getTiltFactorXZ(), dim.getTiltFactorYZ()
    return ((L.x, xy * L.y, xz * L.z), (0.0, L.y, yz * L.z), (0.0, 0.0, L.z)), (lo.x, lo.y, lo.z)


get_global_box = global_box


def bind(sampling_context, callback, **kwargs):
    """Entry point used by SamplingContext (hoomd.py:261-271)."""
    sim = sampling_context.context
    sampler = Sampler(sim, sampling_context.method, callback)
    API.install(sim, sampler)
    SAMPLERS[sim] = sampler
    sampling_context.run = API.runner(sim)
    return sampler


def detach(simulation):
    """hoomd.py:274-284"""
    if SAMPLERS.pop(simulation, None) is None:
        warn("This context has no sampler bound to it.")
    else:
        API.remove(simulation)
