"""
Mock MD engines + their samplers.

No MD engine (HOOMD-blue, OpenMM, LAMMPS) nor its `*-dlext` plugin is installable in the build
or benchmark environment, so these classes stand in for them at the plugin boundary: a
`MockEngine` owns CUDA tensors laid out exactly like the engine's buffers and, once per step,
calls the installed hook with DLPack capsules the way the plugins do
(hoomd.dlext.DLExtSampler.forward_data -> cb(positions, vel_mass, rtags, images, forces, timestep),
pysages/backends/hoomd.py:159-173; openmm_dlext.Force callback, openmm.py:50-56; lammps FixDLExt,
lammps.py:79-95).  The `Sampler` below drives the same
`SamplerCore` as the real adapters: same attributes (`state`, `snapshot`, `callback`, `take_snapshot`,
`restore`) and the same callback order (fused method update + bias -> user callback).
"""

import numpy as np
import torch
from torch.utils.dlpack import from_dlpack, to_dlpack

from . import _common as C
from ._common import LAMMPS_CONVERSION, LAYOUTS  # noqa: F401
from .snapshot import Box, Snapshot


class MockEngine:
    """
    Device-resident stand-in for an MD engine.

    arrays: dict of numpy arrays in the engine's layout (see `pysages_b200.synthetic`):
      hoomd       positions (N,4), vel_mass (N,4), forces (N,4), rtags (N,) i32, images (N,3) i32
      openmm-cuda positions (Np,4), vel_mass (Np,4) [w = 1/m], forces (3,Np) i64, ids (Np,) i32
      openmm-cpu  positions (N,3), velocities (N,3), inv_masses (N,1), forces (N,3), ids (N,) i32
      lammps      positions (N,3), velocities (N,3), forces (N,3), tags_map (N+1,) i32,
                  images (N,) i32 packed, types (N,) i32, masses (ntypes+1,) f64
    trajectory: optional (T, *positions.shape) array; step t exposes frame t % T (zero copy).
    """

    def __init__(self, layout, arrays, box_H, origin=(0.0, 0.0, 0.0), dt=0.005, device="cuda",
                 trajectory=None, integrate=False, units=None):
        if layout not in LAYOUTS:
            raise ValueError(f"Invalid backend {layout}")
        self.layout_name = layout
        self.device = torch.device(device)
        self.t = {k: torch.as_tensor(np.ascontiguousarray(v)).to(self.device) for k, v in arrays.items()}
        self.box_H = np.asarray(box_H, dtype=np.float64)
        self.origin = np.asarray(origin, dtype=np.float64)
        self.dt = float(dt)
        self.units = units
        self.trajectory = None if trajectory is None else torch.as_tensor(np.ascontiguousarray(trajectory)).to(self.device)
        self.integrate = integrate
        self.timestep = 0
        self.hook = None
        self._base_positions = self.t["positions"]

    # ---- what the dlext plugins export --------------------------------------------------------
    def capsules(self):
        t = self.t
        if self.layout_name == "hoomd":
            return tuple(to_dlpack(t[k]) for k in ("positions", "vel_mass", "rtags", "images", "forces"))
        if self.layout_name == "openmm-cuda":
            return tuple(to_dlpack(t[k]) for k in ("positions", "vel_mass", "forces", "ids"))
        if self.layout_name == "openmm-cpu":
            return tuple(to_dlpack(t[k]) for k in ("positions", "velocities", "inv_masses", "forces", "ids"))
        return tuple(to_dlpack(t[k]) for k in ("positions", "velocities", "forces", "tags_map", "images", "types", "masses"))

    def get_global_box(self):
        return self.box_H, self.origin

    def synchronize(self):
        torch.cuda.current_stream(self.device).synchronize()

    # ---- stepping -----------------------------------------------------------------------------
    def _advance(self):
        if self.trajectory is not None:
            self.t["positions"] = self.trajectory[self.timestep % self.trajectory.shape[0]]
        elif self.integrate:
            self._integrate()

    def _integrate(self):
        """Free particles + bias force: symplectic Euler in the engine's own arrays (tests only)."""
        t, dt = self.t, self.dt
        if self.layout_name == "hoomd":
            f, vm, x = t["forces"], t["vel_mass"], t["positions"]
            vm[:, :3] += f[:, :3] / vm[:, 3:] * dt
            x[:, :3] += vm[:, :3] * dt
            f.zero_()
        elif self.layout_name == "openmm-cuda":
            f, vm, x = t["forces"], t["vel_mass"], t["positions"]
            vm[:, :3] += (f.T.double() / 2.0**32).to(vm.dtype) * vm[:, 3:] * dt
            x[:, :3] += vm[:, :3] * dt
            f.zero_()
        elif self.layout_name == "openmm-cpu":
            t["velocities"] += t["forces"] * t["inv_masses"] * dt
            t["positions"] += t["velocities"] * dt
            t["forces"].zero_()
        else:
            m = t["masses"][t["types"].long()].reshape(-1, 1)
            t["velocities"] += t["forces"] / m * dt
            t["positions"] += t["velocities"] * dt
            t["forces"].zero_()

    def run(self, timesteps, **kwargs):
        for _ in range(int(timesteps)):
            self._advance()
            if self.hook is not None:
                self.hook(self.timestep)
            self.timestep += 1


class Sampler(C.SamplerCore):
    """Per-step driver for the stand-in engines: the same `SamplerCore` the real adapters use (hoomd.py, openmm.py,
    lammps.py), fed with the capsules `MockEngine.capsules()` hands out the way the plugins do."""

    ordering = "stream"  # MockEngine works on torch's current stream

    def __init__(self, context, sampling_method, callback=None):
        self.context = context
        layout = LAYOUTS[context.layout_name]
        if context.layout_name == "lammps" and context.units in LAMMPS_CONVERSION:
            layout = layout._replace(force_unit_factor=LAMMPS_CONVERSION[context.units])
        self.box = Box(*context.get_global_box())
        self.dt = context.dt
        self.update_box = lambda: self.box  # constant-volume runs
        self._bind_method(sampling_method, layout, callback)

    def _pack_snapshot(self, caps=None):
        """Wrap the plugin's capsules (zero copy), like hoomd.py:148-157 / lammps.py:87-101."""
        name = self.context.layout_name
        caps = [from_dlpack(c) for c in (self.context.capsules() if caps is None else caps)]
        if name == "hoomd":
            positions, vel_mass, rtags, images, forces = caps
            return Snapshot(positions, vel_mass, forces, rtags, self.update_box(), self.dt, dict(images=images))
        if name == "openmm-cuda":
            positions, vel_mass, forces, ids = caps
            return Snapshot(positions, vel_mass, forces, ids, self.update_box(), self.dt)
        if name == "openmm-cpu":
            positions, velocities, inv_masses, forces, ids = caps
            return Snapshot(positions, (velocities, inv_masses), forces, ids, self.update_box(), self.dt)
        positions, velocities, forces, tags_map, images, types, masses = caps
        return Snapshot(positions, (velocities, (masses, types)), forces, tags_map[1:], self.update_box(),
                        self.dt, dict(images=images))

    _views = _pack_snapshot

    def on_hoomd_half_step(self, positions, vel_mass, rtags, images, forces, timestep):
        """The hoomd.dlext callback signature (hoomd.py:159-168) on the pointer-only fast path."""
        d = C.capsule_data
        key = (d(positions), d(vel_mass), d(rtags), d(images), d(forces))
        if key != self._desc_key:
            ds = self._desc
            ds.positions, ds.vel_mass, ds.ids, ds.images, ds.forces = key
            H = self.box.H
            ds.box_diag[0], ds.box_diag[1], ds.box_diag[2] = float(H[0][0]), float(H[1][1]), float(H[2][2])
            ds.dt = float(self.dt)
            self._desc_key = key
        self.step_current(timestep)

    def _update_callback(self, timestep):
        caps = self.context.capsules()
        if self._fused and self.context.layout_name == "hoomd":
            return self.on_hoomd_half_step(*caps, timestep)
        self.advance_snapshot(self._pack_snapshot(caps), timestep)


def bind(sampling_context, callback, **kwargs):
    """SamplingContext entry point"""
    context = sampling_context.context
    sampler = Sampler(context, sampling_context.method, callback)
    context.hook = sampler._update_callback
    sampling_context.run = context.run
    return sampler
