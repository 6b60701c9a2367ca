"""
Stand-ins for the third-party engine plugins (hoomd + hoomd.dlext, openmm + openmm_dlext,
lammps + lammps.dlext), installed into sys.modules so the REAL adapters
(pysages_b200/backends/{hoomd,openmm,lammps}.py) can be exercised without the engines: they expose
the plugin APIs the adapters call (names and call order as used by the reference backends) on top of
`MockEngine` tensors.  Test infrastructure only.
"""
import sys
import types

import numpy as np
import torch
from torch.utils.dlpack import to_dlpack

from pysages_b200.backends.mock import MockEngine


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _V3:
    def __init__(self, x, y, z):
        self.x, self.y, self.z = float(x), float(y), float(z)


# ---------------------------------------------------------------------------------------------------
# HOOMD-blue (v3+/v4 API) + hoomd.dlext
# ---------------------------------------------------------------------------------------------------
def install_hoomd(with_tags=True):
    class AccessLocation:
        OnHost, OnDevice = 0, 1

    class AccessMode:
        Read, ReadWrite, Overwrite = 0, 1, 2

    class _GlobalBox:
        def __init__(self, H, origin):
            self.H, self.origin = np.asarray(H, float), np.asarray(origin, float)

        def getL(self):
            return _V3(self.H[0, 0], self.H[1, 1], self.H[2, 2])

        def getTiltFactorXY(self):
            return self.H[0, 1] / self.H[1, 1]

        def getTiltFactorXZ(self):
            return self.H[0, 2] / self.H[2, 2]

        def getTiltFactorYZ(self):
            return self.H[1, 2] / self.H[2, 2]

        def getLo(self):
            return _V3(*self.origin)

    class SystemView:
        def __init__(self, cpp_sys):
            self.engine = cpp_sys
            self.particle_data = types.SimpleNamespace(getGlobalBox=lambda: _GlobalBox(cpp_sys.box_H, cpp_sys.origin))

        def synchronize(self):
            self.engine.synchronize()

    class DLExtSampler:
        def __init__(self, sysview, update, location, mode):
            self._sysview, self._update, self._location, self._mode = sysview, update, location, mode

        def forward_data(self, cb, location, mode, timestep):
            t = self._sysview.engine.t
            cb(*(to_dlpack(t[k]) for k in ("positions", "vel_mass", "rtags", "images", "forces")), timestep)

        def update(self, timestep):  # what the C++ integrator calls between the two half steps
            self.forward_data(self._update, self._location, self._mode, timestep)

    def tags(sysview, location, mode):
        # hoomd-dlext: ParticleData's slot -> tag array (the inverse of rtags; HOOMD keeps both current)
        import torch

        eng = sysview.engine
        rt = eng.t["rtags"]
        inv = torch.empty_like(rt)
        inv[rt.long()] = torch.arange(rt.numel(), device=rt.device, dtype=rt.dtype)
        eng.tags_calls = getattr(eng, "tags_calls", 0) + 1
        return to_dlpack(inv)

    class CPU:
        pass

    class GPU:
        pass

    class Simulation:
        def __init__(self, engine):
            self.device = GPU()
            self._cpp_sys = engine
            self.operations = types.SimpleNamespace(integrator=types.SimpleNamespace(dt=engine.dt, half_step_hook=None))
            self.timestep = 0

        def run(self, n):
            for _ in range(int(n)):
                self._cpp_sys._advance()
                hook = self.operations.integrator.half_step_hook
                if hook is not None:
                    hook.update(self.timestep)
                self._cpp_sys.timestep += 1
                self.timestep += 1

    Simulation.__module__ = "hoomd.simulation"
    hoomd = _module("hoomd", __version__="4.0.0", Simulation=Simulation)
    hoomd.md = _module("hoomd.md", HalfStepHook=type("HalfStepHook", (), {}))
    hoomd.device = _module("hoomd.device", CPU=CPU, GPU=GPU)
    hoomd.dlext = _module("hoomd.dlext", __version__="0.4.0", AccessLocation=AccessLocation, AccessMode=AccessMode,
                          DLExtSampler=DLExtSampler, SystemView=SystemView)
    if with_tags:  # older hoomd-dlext builds: no accessor for ParticleData's tag array
        hoomd.dlext.tags = tags
    sys.modules.pop("pysages_b200.backends.hoomd", None)
    return Simulation


# ---------------------------------------------------------------------------------------------------
# OpenMM + openmm_dlext
# ---------------------------------------------------------------------------------------------------
def install_openmm(own_stream=False):
    class DeviceType:
        CPU, GPU = 0, 1

    class ContextView:
        def __init__(self, context):
            self.context = context

        def device_type(self):
            return DeviceType.GPU if self.context.engine.layout_name == "openmm-cuda" else DeviceType.CPU

        def synchronize(self):  # openmm_dlext: waits for OpenMM's OWN context stream, not for anybody else's
            if self.context.engine.device.type == "cuda":
                stream = self.context.engine_stream
                stream.synchronize() if stream is not None else self.context.engine.synchronize()

    class Force:
        def add_to(self, context):
            context.forces.append(self)

        def view(self, context):
            return ContextView(context)

        def set_callback_in(self, context, fn):
            context.callback = fn

    def _cap(view, key):
        return to_dlpack(view.context.engine.t[key])

    class _Vec(tuple):
        def value_in_unit(self, unit):
            return tuple(self)

    class Context:
        def __init__(self, engine):
            self.engine, self.forces, self.callback = engine, [], None
            # own_stream: like the real CUDA platform, force computation and integration run on a non-blocking
            # stream of the context; `seen_forces` is what the "integrator" read right after the callback
            self.engine_stream = torch.cuda.Stream() if (own_stream and engine.device.type == "cuda") else None
            self.seen_forces = None

        def getSystem(self):
            H = self.engine.box_H
            return types.SimpleNamespace(getDefaultPeriodicBoxVectors=lambda: [_Vec(H[:, 0]), _Vec(H[:, 1]), _Vec(H[:, 2])])

        def getIntegrator(self):
            return types.SimpleNamespace(getStepSize=lambda: self.engine.dt)

    class Simulation:
        def __init__(self, engine):
            self.context = Context(engine)
            self.currentStep = 0

        def step(self, n):
            eng = self.context.engine
            for _ in range(int(n)):
                # OpenMM's buffers persist (the plugin's views are taken once): update them in place
                if eng.trajectory is not None:
                    eng.t["positions"].copy_(eng.trajectory[eng.timestep % eng.trajectory.shape[0]])
                es = self.context.engine_stream
                if es is not None:
                    es.wait_stream(torch.cuda.current_stream())  # the position update above
                    # work already queued on the stream the bias step will use (another library's kernels):
                    # the bias kernel starts late, the engine's integrator below does not wait for it by itself
                    torch.cuda._sleep(40_000_000)
                if self.context.callback is not None:
                    self.context.callback(self.currentStep)
                if es is not None:
                    with torch.cuda.stream(es):  # "integrator": reads the force buffer on the engine's stream
                        self.context.seen_forces = eng.t["forces"].clone()
                eng.timestep += 1
                self.currentStep += 1

    Simulation.__module__ = "openmm.app.simulation"
    openmm = _module("openmm", VariableLangevinIntegrator=type("VariableLangevinIntegrator", (), {}),
                     VariableVerletIntegrator=type("VariableVerletIntegrator", (), {}))
    openmm.unit = _module("openmm.unit", nanometer=1.0, picosecond=1.0)
    _module("openmm_dlext", DeviceType=DeviceType, ContextView=ContextView, Force=Force,
            positions=lambda v: _cap(v, "positions"), forces=lambda v: _cap(v, "forces"),
            atom_ids=lambda v: _cap(v, "ids"),
            velocities=lambda v: _cap(v, "vel_mass" if v.context.engine.layout_name == "openmm-cuda" else "velocities"),
            inverse_masses=lambda v: _cap(v, "inv_masses"))
    sys.modules.pop("pysages_b200.backends.openmm", None)
    return Simulation


# ---------------------------------------------------------------------------------------------------
# LAMMPS + lammps.dlext
# ---------------------------------------------------------------------------------------------------
def install_lammps(kokkos_cuda=True, own_stream=False, with_tags=True):
    class ExecutionSpace:
        kOnHost, kOnDevice = 0, 1

    class FixDLExt:
        def __init__(self, context):
            context.fixes.append(self)
            self._callback = None

        def set_callback(self, fn):
            self._callback = fn

    class LAMMPSView:
        def __init__(self, context):
            self.context = context

        def synchronize(self):  # Kokkos fence of LAMMPS' own execution-space instance
            if self.context.engine.device.type == "cuda":
                stream = self.context.engine_stream
                stream.synchronize() if stream is not None else self.context.engine.synchronize()

    def _cap(view, key):
        return to_dlpack(view.context.engine.t[key])

    def tags(view, location):
        # atom->tag: the 1-based id of the atom in every slot (tags_map is its inverse: tags_map[id] = slot)
        eng = view.context.engine
        tm = eng.t["tags_map"]
        ids = torch.arange(1, tm.numel(), device=tm.device, dtype=tm.dtype)
        out = torch.zeros(tm.numel() - 1, device=tm.device, dtype=tm.dtype)
        out[tm[1:].long()] = ids
        eng.tags_calls = getattr(eng, "tags_calls", 0) + 1
        return to_dlpack(out)

    class lammps:  # noqa: N801  (the real class is lower-case)
        def __init__(self, engine):
            self.engine, self.fixes, self.closed = engine, [], False
            self.engine_stream = torch.cuda.Stream() if (own_stream and engine.device.type == "cuda") else None
            self.seen_forces = None

        def extract_global(self, name):
            return {"units": self.engine.units or "lj", "dt": self.engine.dt}[name]

        def extract_box(self):
            H, lo = self.engine.box_H, self.engine.origin
            hi = lo + np.array([H[0, 0], H[1, 1], H[2, 2]])
            return list(lo), list(hi), 0.0, 0.0, 0.0, [1, 1, 1], 0

        def command(self, cmd):
            word, n = cmd.split()
            assert word == "run"
            for _ in range(int(n)):
                self.engine._advance()
                if self.engine_stream is not None:
                    self.engine_stream.wait_stream(torch.cuda.current_stream())
                    torch.cuda._sleep(40_000_000)  # see install_openmm
                for fix in self.fixes:
                    if fix._callback is not None:
                        fix._callback(self.engine.timestep)
                if self.engine_stream is not None:
                    with torch.cuda.stream(self.engine_stream):
                        self.seen_forces = self.engine.t["forces"].clone()
                self.engine.timestep += 1

        def finalize(self):
            self.closed = True

    lammps.__module__ = "lammps.core"
    lmp = _module("lammps", lammps=lammps)
    lmp.dlext = _module("lammps.dlext", ExecutionSpace=ExecutionSpace, FixDLExt=FixDLExt, LAMMPSView=LAMMPSView,
                        kOnHost=0, kOnDevice=1, has_kokkos_cuda_enabled=lambda ctx: kokkos_cuda,
                        kImgBitSize=32, kImgBits=10, kImg2Bits=20, kImgMask=1023, kImgMax=512,
                        positions=lambda v, loc: _cap(v, "positions"), velocities=lambda v, loc: _cap(v, "velocities"),
                        forces=lambda v, loc: _cap(v, "forces"), types=lambda v, loc: _cap(v, "types"),
                        tags_map=lambda v, loc: _cap(v, "tags_map"), images=lambda v, loc: _cap(v, "images"),
                        masses=lambda v, loc: _cap(v, "masses"))
    if with_tags:
        lmp.dlext.tags = tags
    sys.modules.pop("pysages_b200.backends.lammps", None)
    return lammps


def engine(snap, device, trajectory=None, units=None):
    return MockEngine(snap["layout"], snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device=device,
                      trajectory=trajectory, units=units)
