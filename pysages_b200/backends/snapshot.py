"""
Snapshot glue (pysages/backends/snapshot.py:23-107): zero-copy views of the engine's buffers.

A `Snapshot` holds torch CUDA tensors created with `torch.from_dlpack` from the capsules the
`*-dlext` plugins hand to the per-step callback.  Nothing is copied or re-laid-out: the native
step reads the engine's own arrays (tag -> slot map, positions, images, velocities) only at the
selected atoms and adds the bias into the engine's force array in place.
"""

from typing import Any, Callable, Dict, NamedTuple, Optional, Tuple, Union

import numpy as np
import torch

from .. import _native as N


class Box(NamedTuple):
    """Simulation box (snapshot.py:23-31): transform matrix H (3,3) and origin (3,), host floats."""

    H: Any
    origin: Any

    def __repr__(self):
        return "PySAGES " + type(self).__name__


class Snapshot(NamedTuple):
    """snapshot.py:34-46"""

    positions: Any
    vel_mass: Union[Any, Tuple[Any, Any]]
    forces: Any
    ids: Any
    box: Box
    dt: Union[float, Any]
    extras: Optional[Dict[str, Any]] = None

    def __repr__(self):
        return "PySAGES " + type(self).__name__


class Layout(NamedTuple):
    """Engine memory layout understood by the native step (include/pysages_b200.h psb_layout_kind)."""

    kind: int                # N.LAYOUT_*
    name: str
    force_unit_factor: float = 1.0   # helpers.to_force_units (lammps.py:33,137-139)


class HelperMethods(NamedTuple):
    """snapshot.py:59-62, extended with the engine layout the fused kernel is specialised for."""

    query: Callable
    dimensionality: Callable[[], int]
    to_force_units: Callable = lambda x: x
    layout: Layout = Layout(N.LAYOUT_HOOMD, "hoomd")


def _clone(x, to_cpu=False):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().copy() if to_cpu else x.detach().clone()
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _clone(v, to_cpu) for k, v in x.items()}
    if isinstance(x, Box):
        return Box(np.array(x.H, dtype=np.float64), np.array(x.origin, dtype=np.float64))
    if isinstance(x, tuple):
        return tuple(_clone(v, to_cpu) for v in x)
    return x


def copy(snapshot, to_cpu=False):
    """pysages.utils.copy for Snapshots (snapshot.py:65-73; tests/test_snapshots.py): deep copy on
    the device, or to numpy with `to_cpu=True`."""
    return Snapshot(*(_clone(x, to_cpu) for x in snapshot))


def restore(snapshot, prev_snapshot):
    """snapshot.py:81-95: overwrite the engine's buffers with a previous snapshot (restart path)."""

    def put(dst, src):
        if isinstance(dst, (tuple, list)):
            for d, s in zip(dst, src):
                put(d, s)
        elif dst is not None and src is not None:
            dst.copy_(torch.as_tensor(src, device=dst.device).reshape(dst.shape))

    put(snapshot.positions, prev_snapshot.positions)
    put(snapshot.forces, prev_snapshot.forces)
    put(snapshot.ids, prev_snapshot.ids)
    put(snapshot.vel_mass, prev_snapshot.vel_mass)
    if snapshot.extras and prev_snapshot.extras and "images" in snapshot.extras:
        put(snapshot.extras["images"], prev_snapshot.extras["images"])
    # HOOMD's slot -> tag array goes back together with rtags (`ids`): the two are each other's inverse
    if snapshot.extras and prev_snapshot.extras and "tags" in snapshot.extras and "tags" in prev_snapshot.extras:
        put(snapshot.extras["tags"], prev_snapshot.extras["tags"])


def _ptr(t):
    return None if t is None else int(t.data_ptr())


def _real_bytes(t):
    if t.dtype == torch.float32:
        return 4
    if t.dtype == torch.float64:
        return 8
    raise ValueError(f"positions must be float32 or float64 (got {t.dtype})")


def natoms(snapshot):
    return int(snapshot.positions.shape[0])


def describe_layout(snapshot, layout, unwrap, need_momenta, dense_outputs=False):
    """-> psb_layout_desc for this engine layout and snapshot."""
    d = N.LayoutDesc()
    d.layout = layout.kind
    d.real_bytes = _real_bytes(snapshot.positions)
    d.natoms = natoms(snapshot)
    d.unwrap = int(bool(unwrap))
    d.need_momenta = int(bool(need_momenta))
    d.dense_outputs = int(bool(dense_outputs))
    if layout.kind == N.LAYOUT_LAMMPS and isinstance(snapshot.vel_mass, tuple):
        masses = snapshot.vel_mass[1][0]
        d.ntypes = 0 if masses is None else int(masses.numel())
    # HOOMD snapshots that carry the slot -> tag array: the handle is planned for it (psb_layout_desc.has_tags)
    d.has_tags = int(layout.kind in (N.LAYOUT_HOOMD, N.LAYOUT_LAMMPS) and bool(snapshot.extras) and snapshot.extras.get("tags") is not None)
    return d


def describe_snapshot(snapshot, layout, out=None):
    """Fill a psb_snapshot with the raw device pointers of `snapshot` (no copies, no syncs)."""
    s = out if out is not None else N.SnapshotDesc()
    kind = layout.kind
    s.positions = _ptr(snapshot.positions)
    s.forces = _ptr(snapshot.forces)
    images = snapshot.extras.get("images") if snapshot.extras else None
    s.images = _ptr(images)
    # HOOMD / LAMMPS: the engine's slot -> tag array, when the adapter got one (psb_snapshot.tags)
    tags = snapshot.extras.get("tags") if (snapshot.extras and kind in (N.LAYOUT_HOOMD, N.LAYOUT_LAMMPS)) else None
    s.tags = _ptr(tags)
    if kind == N.LAYOUT_LAMMPS:
        velocities, (masses, types) = snapshot.vel_mass
        s.vel_mass = _ptr(velocities)
        s.masses = _ptr(masses)
        s.types = _ptr(types)
        # snapshot.ids is tags_map[1:] (lammps.py:113); the kernel indexes the full map at tag + 1
        s.ids = int(snapshot.ids.data_ptr()) - snapshot.ids.element_size()
    elif kind == N.LAYOUT_PLAIN3:
        velocities, inv_masses = snapshot.vel_mass
        s.vel_mass = _ptr(velocities)
        s.masses = _ptr(inv_masses)
        s.types = None
        s.ids = _ptr(snapshot.ids)
    else:
        s.vel_mass = _ptr(snapshot.vel_mass)
        s.masses = None
        s.types = None
        s.ids = _ptr(snapshot.ids)
    H = np.asarray(snapshot.box.H, dtype=np.float64)
    for i in range(3):
        s.box_diag[i] = float(H[i, i])
    s.dt = float(snapshot.dt) if snapshot.dt is not None else 0.0
    return s
