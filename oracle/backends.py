"""
Oracle (test infrastructure): snapshot glue and force apply, per engine layout.

Restates pysages/backends/snapshot.py:34-107, hoomd.py:176-240,
openmm.py:96-180 and lammps.py:130-223 on numpy / torch CPU arrays.  No engine
and no DLPack are involved: snapshots are plain arrays laid out exactly as the
dlext plugins export them.
"""

from typing import Any, Callable, NamedTuple

import numpy as np
import torch


class Box(NamedTuple):  # snapshot.py:23-31
    H: Any
    origin: Any


class Snapshot(NamedTuple):  # snapshot.py:34-46
    positions: Any
    vel_mass: Any
    forces: Any
    ids: Any
    box: Any
    dt: Any
    extras: Any = None


class HelperMethods(NamedTuple):  # snapshot.py:59-62
    query: Callable
    dimensionality: Callable
    to_force_units: Callable = lambda x: x


def _t(x):
    return torch.as_tensor(np.asarray(x))


def build_data_querier(snapshot_methods, flags):
    """snapshot.py:98-107 -- fields sorted alphabetically into a namedtuple"""
    flags = sorted(flags)
    ParticleData = NamedTuple("ParticleData", [(f, Any) for f in flags])
    getters = [snapshot_methods[f] for f in flags]

    def query_data(snapshot):
        return ParticleData(*(g(snapshot) for g in getters))

    return query_data


# --------------------------------------------------------------------------- #
# HOOMD-blue  (hoomd.py:176-240)
# --------------------------------------------------------------------------- #


def hoomd_snapshot_methods(requires_box_unwrapping=True):
    if requires_box_unwrapping:

        def positions(s):  # hoomd.py:179-181: f32/f64 xyz + f64 diag(H) * i32 images -> f64
            L = torch.diag(_t(np.asarray(s.box.H, dtype=np.float64)))
            return _t(s.positions)[:, :3] + L * _t(s.extras["images"])

    else:

        def positions(s):  # hoomd.py:185-186
            return _t(s.positions)

    def indices(s):  # hoomd.py:189-190
        return _t(s.ids)

    def momenta(s):  # hoomd.py:192-196 -- product in the array's own dtype
        vm = _t(s.vel_mass)
        return (vm[:, 3:] * vm[:, :3]).flatten()

    def masses(s):
        return _t(s.vel_mass)[:, 3:]

    return dict(positions=positions, indices=indices, momenta=momenta, masses=masses)


def hoomd_bias(snapshot, bias):
    """hoomd.py:219-230: forces[:, :3] += biases (in place, cast to the force dtype)"""
    if bias is None:
        return
    f = snapshot.forces
    f[:, :3] = (f[:, :3].astype(np.float64) + np.asarray(bias, dtype=np.float64)).astype(f.dtype)


# --------------------------------------------------------------------------- #
# OpenMM  (openmm.py:96-180)
# --------------------------------------------------------------------------- #


def openmm_snapshot_methods(on_gpu):
    def safe_divide(v, invm):  # openmm.py:96-97 (vmapped per atom)
        return torch.where(invm == 0, v, v / torch.where(invm == 0, torch.ones_like(invm), invm))

    if on_gpu:

        def unpack(vel_mass):  # openmm.py:103-104
            vm = _t(vel_mass)
            return vm[:, :3], vm[:, 3:]

        def indices(s):  # openmm.py:106-107 -- argsort every step
            return torch.argsort(_t(np.asarray(s.ids).astype(np.int64)), stable=True)

    else:

        def unpack(vel_mass):
            return _t(vel_mass[0]), _t(vel_mass[1])

        def indices(s):  # openmm.py:115-116
            return _t(s.ids)

    def positions(s):  # openmm.py:122-123 -- raw, never unwrapped
        return _t(s.positions)

    def momenta(s):  # openmm.py:125-127
        V, invM = unpack(s.vel_mass)
        return safe_divide(V, invM).flatten()

    def masses(s):
        _, invM = unpack(s.vel_mass)
        return 1 / invM

    return dict(positions=positions, indices=indices, momenta=momenta, masses=masses)


def openmm_bias(snapshot, bias, on_gpu):
    """openmm.py:141-143,160-170: CUDA forces are int64 fixed point, shape (3, Np)"""
    if bias is None:
        return
    if on_gpu:
        adapted = (2.0**32 * np.asarray(bias, dtype=np.float64).T).astype(np.int64)  # truncates
        snapshot.forces[:, : adapted.shape[1]] += adapted
    else:
        snapshot.forces[...] = (
            snapshot.forces.astype(np.float64) + np.asarray(bias, dtype=np.float64)
        ).astype(snapshot.forces.dtype)


# --------------------------------------------------------------------------- #
# LAMMPS  (lammps.py:130-223)
# --------------------------------------------------------------------------- #

# image packing constants of a default (SMALLBIG) LAMMPS build, lmptype.h
LAMMPS_IMG = dict(kImgBits=10, kImg2Bits=20, kImgMask=1023, kImgMax=512, kImgBitSize=32)
LAMMPS_CONVERSION = {"real": 2390.0573615334906, "metal": 1.0364269e-4, "electron": 1.06657236}


def lammps_snapshot_methods(requires_box_unwrapping=True, img=LAMMPS_IMG):
    if requires_box_unwrapping:
        dtype = np.int64 if img["kImgBitSize"] == 64 else np.int32
        bits = np.asarray((0, img["kImgBits"], img["kImg2Bits"]), dtype=dtype)
        mask = np.asarray((img["kImgMask"], img["kImgMask"], -1), dtype=dtype)
        offset = img["kImgMax"]

        def positions(s):  # lammps.py:197-202
            L = torch.diag(_t(np.asarray(s.box.H, dtype=np.float64)))
            image = np.asarray(s.extras["images"]).reshape(-1, 1)
            unpacked = ((image >> bits) & mask) - offset
            return _t(s.positions)[:, :3] + L * _t(unpacked)

    else:

        def positions(s):
            return _t(s.positions)

    def indices(s):  # lammps.py:209-210 (ids == tags_map[1:], lammps.py:113)
        return _t(s.ids)

    def momenta(s):  # lammps.py:213-217
        V, (masses, types) = s.vel_mass
        M = _t(masses)[_t(np.asarray(types).astype(np.int64))].reshape(-1, 1)
        return (M * _t(V)).flatten()

    return dict(positions=positions, indices=indices, momenta=momenta, masses=None)


def lammps_bias(snapshot, bias):
    """lammps.py:163-170"""
    if bias is None:
        return
    f = snapshot.forces
    f[:, :3] = (f[:, :3].astype(np.float64) + np.asarray(bias, dtype=np.float64)).astype(f.dtype)


def build_helpers(engine, flags, requires_box_unwrapping=True, units=None):
    """-> (HelperMethods, bias(snapshot, bias_array)) for engine in
    {"hoomd", "openmm-cuda", "openmm-cpu", "lammps"}"""
    if engine == "hoomd":
        sm, bias = hoomd_snapshot_methods(requires_box_unwrapping), hoomd_bias
        to_force_units = lambda x: x  # noqa: E731
    elif engine == "openmm-cuda":
        sm = openmm_snapshot_methods(True)
        bias = lambda s, b: openmm_bias(s, b, True)  # noqa: E731
        to_force_units = lambda x: x  # noqa: E731
    elif engine == "openmm-cpu":
        sm = openmm_snapshot_methods(False)
        bias = lambda s, b: openmm_bias(s, b, False)  # noqa: E731
        to_force_units = lambda x: x  # noqa: E731
    elif engine == "lammps":
        sm, bias = lammps_snapshot_methods(requires_box_unwrapping), lammps_bias
        factor = LAMMPS_CONVERSION.get(units)  # lammps.py:33,137-139
        to_force_units = (lambda x: x) if factor is None else (lambda x: factor * x)
    else:
        raise ValueError(f"Invalid backend {engine}")  # backends/core.py:42-44
    helpers = HelperMethods(build_data_querier(sm, flags), lambda: 3, to_force_units)
    return helpers, bias
