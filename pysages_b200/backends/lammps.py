"""
LAMMPS adapter on top of the `lammps.dlext` fix (the role of pysages/backends/lammps.py).

With Kokkos-CUDA the fix exposes device arrays that are used in place; otherwise the host arrays are
staged through the GPU every step (psb_step_host).  Layout: x, v, f (N, 3) doubles, `tags_map` (tag t at
index t + 1), packed image flags, per-type masses + per-atom types; force units follow lammps.py:33,137-139.
Ordering on the device: `view.synchronize()` before the step, and the host waits for the bias kernel after
it (the reference's `sync_forces()`, lammps.py:163-170): Kokkos may integrate on an execution-space
instance of its own.
"""

import weakref

import torch
from lammps import dlext

from . import _common as C
from .snapshot import Box, Snapshot

_tags_of = getattr(dlext, "tags", None)  # tags(view, location) -> atom->tag (slot -> 1-based id); optional

# The native step unpacks LAMMPS' SMALLBIG image word: 10 + 10 + 12 bits, offset 512 (lammps.py:188-202 reads
# the same constants from the plugin).  A BIGBIG build (64-bit words, 21-bit fields) would be unwrapped wrongly.
_IMAGE_PACKING = dict(kImgBitSize=32, kImgBits=10, kImg2Bits=20, kImgMask=1023, kImgMax=512)


def check_image_packing():
    seen = {name: getattr(dlext, name) for name in _IMAGE_PACKING if hasattr(dlext, name)}
    bad = {name: v for name, v in seen.items() if int(v) != _IMAGE_PACKING[name]}
    if bad:
        raise RuntimeError(f"this LAMMPS build packs image flags as {bad}; pysages_b200 supports the 32-bit "
                           f"SMALLBIG packing {_IMAGE_PACKING} only")


class Sampler(C.SamplerCore, dlext.FixDLExt):
    ordering = "host"

    def __init__(self, lmp, sampling_method, callback=None, location=None):
        dlext.FixDLExt.__init__(self, lmp)
        check_image_packing()
        device_ok = hasattr(dlext.ExecutionSpace, "kOnDevice") and bool(dlext.has_kokkos_cuda_enabled(lmp))
        wanted = dlext.kOnDevice if location is None and device_ok else (dlext.kOnHost if location is None else location)
        self.location = wanted if device_ok else dlext.kOnHost
        self.context = lmp
        self.view = dlext.LAMMPSView(lmp)
        self.box = Box(*global_box(lmp))
        self.dt = lmp.extract_global("dt")
        self._masses = None
        self._use_tags = self.location != dlext.kOnHost  # device arrays only; decided for good once the handle exists
        layout = C.LAYOUTS["lammps"]
        factor = C.LAMMPS_CONVERSION.get(lmp.extract_global("units"))
        if factor is not None:
            layout = layout._replace(force_unit_factor=factor)
        self._bind_method(sampling_method, layout, callback)
        self._use_tags = self._use_tags and _tags_of is not None and self._native.wants_tags
        self.set_callback(self.on_post_force)

    def _views(self):
        """LAMMPS may reallocate per-atom arrays between steps: the capsules are asked for again every time
        (lammps.py:87-113); per-type masses do not change."""
        t, v, loc = C.tensor_from_capsule, self.view, self.location
        if self._masses is None:
            self._masses = t(dlext.masses(v, loc))
        tags_map = t(dlext.tags_map(v, loc))
        extras = dict(images=t(dlext.images(v, loc)))
        if _tags_of is not None and self._use_tags:  # atom->tag (slot -> id): the slot-ordered class streams with it
            tags = t(_tags_of(v, loc))
            if tags.dtype == torch.int32:  # (BIGBIG builds carry 64-bit ids: the tags_map inversion pass stays)
                extras["tags"] = tags
        return Snapshot(t(dlext.positions(v, loc)), (t(dlext.velocities(v, loc)), (self._masses, t(dlext.types(v, loc)))),
                        t(dlext.forces(v, loc)), tags_map[1:], self.box, self.dt, extras)

    def on_post_force(self, timestep):
        self.view.synchronize()
        views = self._views()
        if self._fused:
            tags = views.extras.get("tags")
            key = (views.positions.data_ptr(), views.forces.data_ptr(), views.ids.data_ptr(),
                   views.vel_mass[0].data_ptr(), views.vel_mass[1][1].data_ptr(), views.extras["images"].data_ptr(),
                   tags.data_ptr() if tags is not None else None)
            self._current = views
            self.advance_pointers(timestep, key, lambda d: self.describe_into(views))
        else:
            self.advance_snapshot(views, timestep)


def global_box(lmp):
    """Box matrix and lower corner from `extract_box` (lammps.py:225-233)."""
    lo, hi, xy, yz, xz, *_ = lmp.extract_box()
    Lx, Ly, Lz = (hi[d] - lo[d] for d in range(3))
    return ((Lx, xy * Ly, xz * Lz), (0.0, Ly, yz * Lz), (0.0, 0.0, Lz)), lo


get_global_box = global_box


def bind(sampling_context, callback, **kwargs):
    """SamplingContext entry point (lammps.py:241-268)."""
    lmp = sampling_context.context
    sampler = Sampler(lmp, sampling_context.method, callback)
    sampling_context.run = lambda n, **kw: lmp.command(f"run {int(n)}")
    # `with SamplingContext(...)` must not close the instance (results are read afterwards): it is finalized
    # when the last reference goes away instead
    lmp.__exit__ = lambda *exc: None
    weakref.finalize(lmp, lmp.finalize)
    return sampler
