"""
LAMMPS adapter (drop-in for pysages/backends/lammps.py) on top of the `lammps.dlext` fix.

With Kokkos-CUDA the fix exposes device arrays that are used in place; otherwise the host arrays
are staged through the GPU every step (psb_step_host).  Layout: x, v, f (N, 3) doubles, `tags_map`
(tag t at index t + 1), packed image flags, per-type masses + per-atom types; force units follow
lammps.py:33,137-139.
"""

import weakref

from lammps import dlext
from lammps.dlext import ExecutionSpace, FixDLExt, LAMMPSView, has_kokkos_cuda_enabled

from ._common import LAMMPS_CONVERSION, LAYOUTS, fused_helpers, tensor_from_capsule
from .snapshot import Box, Snapshot, copy
from .snapshot import restore as _restore

kDefaultLocation = dlext.kOnHost if not hasattr(ExecutionSpace, "kOnDevice") else dlext.kOnDevice


class Sampler(FixDLExt):
    """LAMMPS fix connecting a pysages_b200 sampling method to a LAMMPS instance."""

    def __init__(self, context, sampling_method, callback=None, location=kDefaultLocation):
        super().__init__(context)
        on_gpu = (location != dlext.kOnHost) and bool(has_kokkos_cuda_enabled(context))
        self.context = context
        self.location = location if on_gpu else dlext.kOnHost
        self.view = LAMMPSView(context)
        self.callback = callback
        self.snapshot = self.take_snapshot()
        layout = LAYOUTS["lammps"]
        units = context.extract_global("units")
        if units in LAMMPS_CONVERSION:
            layout = layout._replace(force_unit_factor=LAMMPS_CONVERSION[units])
        _, initialize, method_update = sampling_method.build(self.snapshot, fused_helpers(layout))
        self.state = initialize()
        self._update_box = lambda: self.snapshot.box

        def update(timestep):
            self.view.synchronize()
            self.snapshot = self._update_snapshot()
            self.state = method_update(self.snapshot, self.state, timestep)
            if self.callback:
                self.callback(self.snapshot, self.state, timestep)

        self.set_callback(update)

    def _partial_snapshot(self, include_masses=False):
        v, loc, t = self.view, self.location, tensor_from_capsule
        positions, types = t(dlext.positions(v, loc)), t(dlext.types(v, loc))
        velocities, forces = t(dlext.velocities(v, loc)), t(dlext.forces(v, loc))
        tags_map, images = t(dlext.tags_map(v, loc)), t(dlext.images(v, loc))
        masses = t(dlext.masses(v, loc)) if include_masses else None
        return Snapshot(positions, (velocities, (masses, types)), forces, tags_map, None, None, dict(images=images))

    def _update_snapshot(self):
        s = self._partial_snapshot()
        velocities, (_, types) = s.vel_mass
        _, (masses, _) = self.snapshot.vel_mass
        return Snapshot(s.positions, (velocities, (masses, types)), s.forces, s.ids[1:], self._update_box(),
                        self.snapshot.dt, s.extras)

    def restore(self, prev_snapshot):
        _restore(self.snapshot, prev_snapshot)

    def take_snapshot(self):
        s = self._partial_snapshot(include_masses=True)
        box = Box(*get_global_box(self.context))
        # the arrays stay views of the engine's buffers; use `copy()` for a persistent snapshot
        return Snapshot(s.positions, s.vel_mass, s.forces, s.ids[1:], box, get_timestep(self.context), s.extras)

    def copy_snapshot(self):
        return copy(self.snapshot)


def get_global_box(context):
    """lammps.py:225-233"""
    boxlo, boxhi, xy, yz, xz, *_ = context.extract_box()
    Lx, Ly, Lz = (boxhi[i] - boxlo[i] for i in range(3))
    H = ((Lx, xy * Ly, xz * Lz), (0.0, Ly, yz * Lz), (0.0, 0.0, Lz))
    return H, boxlo


def get_timestep(context):
    return context.extract_global("dt")


def bind(sampling_context, callback, **kwargs):
    """lammps.py:241-268"""
    context = sampling_context.context
    sampler = Sampler(context, sampling_context.method, callback)
    sampling_context.run = lambda n, **kw: context.command(f"run {n}")
    context.__exit__ = lambda *args: None  # keep the instance alive after `with`
    weakref.finalize(context, context.finalize)
    return sampler
