"""
Shared harness of the GPU parity tests: builds the SAME collective variables / method on the
oracle (CPU, torch-fp64 autodiff) and on pysages_b200 (CUDA, through the C ABI) from one spec,
steps both over the same snapshots and compares every observable of the hot path.
"""
import os

import numpy as np
import torch

import oracle
import pysages_b200 as ps
from pysages_b200.backends import mock
from pysages_b200.backends.snapshot import Box, HelperMethods, Snapshot

X64_RTOL = 1e-10  # BASELINE.json north_star: <= 1e-10 relative in x64 mode
F32_RTOL = 1e-5   #                          <= 1e-5  in fp32

GRID_KINDS = {
    "regular": (oracle.grids.Regular, ps.grids.Regular),
    "periodic": (oracle.grids.Periodic, ps.grids.Periodic),
    "chebyshev": (oracle.grids.Chebyshev, ps.grids.Chebyshev),
}


def make_cvs(mod, specs):
    return [getattr(mod, name)(*args, **kwargs) for (name, args, kwargs) in specs]


def make_grid(which, spec):
    if spec is None:
        return None
    lower, upper, shape, kind = spec
    ok, pk = GRID_KINDS[kind]
    if which == "oracle":
        return oracle.grids.Grid(lower, upper, shape, kind=ok)
    return ps.Grid[pk](lower, upper, shape)


def make_method(which, spec, cvs):
    name, kwargs = spec
    kwargs = dict(kwargs)
    mod = oracle.methods if which == "oracle" else ps.methods
    if "grid" in kwargs:
        kwargs["grid"] = make_grid(which, kwargs["grid"])
    if "restraints" in kwargs and kwargs["restraints"] is not None:
        kwargs["restraints"] = mod.CVRestraints(*kwargs["restraints"])
    if which != "oracle":
        kwargs["dense_bias"] = True
    if name == "HarmonicBias":
        return mod.HarmonicBias(cvs, kwargs.pop("kspring"), kwargs.pop("center"), **kwargs)
    if name == "Metadynamics":
        args = [kwargs.pop(k) for k in ("height", "sigma", "stride", "ngaussians")]
        return mod.Metadynamics(cvs, *args, kwargs.pop("deltaT", None), **kwargs)
    if name == "ABF":
        return mod.ABF(cvs, kwargs.pop("grid"), **kwargs)
    if name == "Sirens":
        if which != "oracle" and "max_iters" in kwargs:
            from pysages_b200 import ml
            loss = ml.Sobolev1SSE() if kwargs.get("mode", "abf") == "cff" else ml.GradientsSSE()
            kwargs["optimizer"] = ml.LevenbergMarquardt(loss=loss, reg=ml.L2Regularization(1e-4),
                                                        max_iters=kwargs.pop("max_iters"))
        return mod.Sirens(cvs, kwargs.pop("grid"), kwargs.pop("topology"), **kwargs)
    if name == "CFF":
        if which != "oracle":
            from pysages_b200 import ml
            reg = ml.L2Regularization(1e-4)
            kwargs["optimizer"] = ml.LevenbergMarquardt(loss=ml.Sobolev1SSE(), reg=reg, max_iters=kwargs.pop("max_iters", 1000))
            kwargs["foptimizer"] = ml.LevenbergMarquardt(reg=reg, max_iters=kwargs.pop("fmax_iters", 500))
        return mod.CFF(cvs, kwargs.pop("grid"), kwargs.pop("topology"), kwargs.pop("kT"), **kwargs)
    if name == "ANN":
        if which != "oracle" and "max_iters" in kwargs:
            from pysages_b200 import ml
            kwargs["optimizer"] = ml.LevenbergMarquardt(reg=ml.L2Regularization(kwargs.pop("reg", 1e-6)),
                                                        max_iters=kwargs.pop("max_iters"))
        return mod.ANN(cvs, kwargs.pop("grid"), kwargs.pop("topology"), kwargs.pop("kT"), **kwargs)
    if name == "SpectralABF":
        return mod.SpectralABF(cvs, kwargs.pop("grid"), **kwargs)
    if name == "FUNN":
        if which != "oracle" and "max_iters" in kwargs:  # the oracle takes the LM iteration cap directly
            from pysages_b200 import ml
            kwargs["optimizer"] = ml.LevenbergMarquardt(reg=ml.L2Regularization(kwargs.pop("reg", 1e-6)),
                                                        max_iters=kwargs.pop("max_iters"))
        return mod.FUNN(cvs, kwargs.pop("grid"), kwargs.pop("topology"), **kwargs)
    raise ValueError(name)


# ---- snapshots ---------------------------------------------------------------------------------
def oracle_snapshot(snap, positions=None):
    a = {k: np.array(v, copy=True) for k, v in snap["arrays"].items()}
    if positions is not None:
        a["positions"] = np.array(positions, copy=True)
    box = oracle.backends.Box(snap["box_H"], snap["origin"])
    L = snap["layout"]
    if L == "hoomd":
        return oracle.backends.Snapshot(a["positions"], a["vel_mass"], a["forces"], a["rtags"], box, snap["dt"],
                                        dict(images=a["images"]))
    if L == "openmm-cuda":
        return oracle.backends.Snapshot(a["positions"], a["vel_mass"], a["forces"], a["ids"], box, snap["dt"])
    if L == "openmm-cpu":
        return oracle.backends.Snapshot(a["positions"], (a["velocities"], a["inv_masses"]), a["forces"], a["ids"],
                                        box, snap["dt"])
    if L == "lammps":
        return oracle.backends.Snapshot(a["positions"], (a["velocities"], (a["masses"], a["types"])), a["forces"],
                                        a["tags_map"][1:], box, snap["dt"], dict(images=a["images"]))
    raise ValueError(L)


def device_snapshot(snap, device="cuda", positions=None):
    a = {k: torch.as_tensor(np.ascontiguousarray(v)).to(device) for k, v in snap["arrays"].items()}
    if positions is not None:
        a["positions"] = torch.as_tensor(np.ascontiguousarray(positions)).to(device)
    box = Box(snap["box_H"], snap["origin"])
    L = snap["layout"]
    if L == "hoomd":
        extras = dict(images=a["images"])
        if os.environ.get("PSB_TEST_TAGS") == "1":  # HOOMD's slot -> tag array (hoomd-dlext `tags`): inverse of rtags
            rt = a["rtags"]
            extras["tags"] = torch.empty_like(rt)
            extras["tags"][rt.long()] = torch.arange(rt.numel(), device=rt.device, dtype=rt.dtype)
        return Snapshot(a["positions"], a["vel_mass"], a["forces"], a["rtags"], box, snap["dt"], extras)
    if L == "openmm-cuda":
        return Snapshot(a["positions"], a["vel_mass"], a["forces"], a["ids"], box, snap["dt"])
    if L == "openmm-cpu":
        return Snapshot(a["positions"], (a["velocities"], a["inv_masses"]), a["forces"], a["ids"], box, snap["dt"])
    if L == "lammps":
        extras = dict(images=a["images"])
        if os.environ.get("PSB_TEST_TAGS") == "1" and a["positions"].is_cuda:  # atom->tag: 1-based id of the atom in each slot
            tm = a["tags_map"]
            extras["tags"] = torch.zeros(tm.numel() - 1, device=tm.device, dtype=tm.dtype)
            extras["tags"][tm[1:].long()] = torch.arange(1, tm.numel(), device=tm.device, dtype=tm.dtype)
        return Snapshot(a["positions"], (a["velocities"], (a["masses"], a["types"])), a["forces"], a["tags_map"][1:],
                        box, snap["dt"], extras)
    raise ValueError(L)


def device_helpers(layout_name, units=None):
    layout = mock.LAYOUTS[layout_name]
    if units in mock.LAMMPS_CONVERSION:
        layout = layout._replace(force_unit_factor=mock.LAMMPS_CONVERSION[units])
    return HelperMethods(None, lambda: 3, lambda x: x, layout)


def fp32_compute(snap):
    """The reference computes CVs in the array dtype when no fp64 constant enters (SURVEY.md trap 13):
    fp32 OpenMM arrays are never unwrapped, so there parity is the fp32 tolerance."""
    return snap["layout"] == "openmm-cuda" and snap["arrays"]["positions"].dtype == np.float32


def close(a, b, rtol, scale=None, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ref = np.max(np.abs(b)) if scale is None else scale
    err = np.max(np.abs(a - b)) if a.size else 0.0
    assert err <= rtol * max(ref, 1e-300) + 1e-300, f"{what}: max abs err {err:.3e} vs scale {ref:.3e} (rtol {rtol:g})"


def forces_xyz(snap_layout, forces):
    f = np.asarray(forces)
    if snap_layout == "openmm-cuda":
        return f
    return f[:, :3]


class ParityRun:
    """Steps the oracle and the CUDA implementation side by side."""

    def __init__(self, snap, cv_specs, method_spec, device="cuda", units=None):
        self.snap = snap
        # CVs whose Jacobian entries are products, not differences of large terms: their bias is gated elementwise
        self.elementwise = all(name in ("Component", "Distance", "Displacement", "Angle", "DihedralAngle",
                                        "RadiusOfGyration") for (name, _, _) in cv_specs)
        self.layout = snap["layout"]
        self.device = device
        self.rtol = F32_RTOL if fp32_compute(snap) else X64_RTOL
        # oracle
        ocvs = make_cvs(oracle.colvars, cv_specs)
        self.omethod = make_method("oracle", method_spec, ocvs)
        self.osnap = oracle_snapshot(snap)
        ohelpers, self.obias = oracle.backends.build_helpers(
            self.layout, self.omethod.snapshot_flags, self.omethod.requires_box_unwrapping, units)
        _, oinit, self.oupdate = self.omethod.build(self.osnap, ohelpers)
        self.ostate = oinit()
        self.ohelpers = ohelpers
        # ours
        cvs = make_cvs(ps.colvars, cv_specs)
        self.method = make_method("ours", method_spec, cvs)
        self.dsnap = device_snapshot(snap, device)
        _, init, self.update = self.method.build(self.dsnap, device_helpers(self.layout, units))
        self.native = self.update.native
        self.state = init()
        torch.cuda.synchronize()
        close(self.state.xi.cpu().numpy(), self.ostate.xi.numpy(), self.rtol, what="initial xi")

    def step(self, positions=None, check=True):
        """One bias step on both sides (optionally with new positions); returns (ours, oracle) states."""
        if positions is not None:
            self.osnap = self.osnap._replace(positions=np.array(positions, copy=True))
            self.dsnap = self.dsnap._replace(
                positions=torch.as_tensor(np.ascontiguousarray(positions)).to(self.device))
        self.ostate = self.oupdate(self.osnap, self.ostate)
        bias = None if self.ostate.bias is None else self.ostate.bias.numpy()
        self.obias(self.osnap, bias)
        self.state = self.update(self.dsnap, self.state)
        torch.cuda.synchronize()
        if check:
            self.check()
        return self.state, self.ostate

    def jacobian(self):
        """Dense (k, 3N) Jacobian of the oracle on the current snapshot."""
        _, J = self.omethod.cv(self.ohelpers.query(self.osnap))
        return J.numpy()

    def check(self, jacobian=True):
        """Gates.  rtol = 1e-10 (x64) / 1e-5 (fp32 compute), always relative to the magnitude
        of the TERMS a quantity is summed from -- the forward-error scale of a floating-point sum:
          xi, J, heights, centers, grids ... max |reference| of the field;
          bias (N,3)                      (a) max |reference| overall, and (b) ELEMENTWISE, after taking out the
                                          generalized-force difference dF (gated by (a)): |d bias + J^T dF| <=
                                          rtol * (|J|^T |F|) for every atom and axis (CVs whose Jacobian entries
                                          are products: point CVs and RadiusOfGyration; the RMSD / gyration /
                                          puckering gradients are differences of O(|r|) terms and keep (a));
          Wp, Wp_ = solve(J J^T, J p)     |(J J^T)^-1| (|J| |p|): J p is a signed sum over 3N products that
                                          cancels (thermal momenta), two summation orders differ by eps * sum |terms|;
          Fsum, force, F                  max |field| or the finite difference (1.5 Wp - 2 Wp' + 0.5 Wp'') / dt
                                          of abf.py:213 evaluated on that Wp scale, per accumulated call."""
        s, o, rtol = self.state, self.ostate, self.rtol
        close(s.xi.cpu().numpy(), o.xi.numpy(), rtol, what="xi")
        assert int(s.ncalls) == int(o.ncalls)
        obias = o.bias.numpy()
        scale = max(np.max(np.abs(obias)), 1e-300)
        sbias = s.bias.cpu().numpy()
        close(sbias, obias, rtol, scale=scale, what="bias")
        Jo = None
        if jacobian:
            J = self.native.view("jacobian", (self.native.k, -1)).cpu().numpy()
            Jo = self.jacobian()
            close(J, Jo, rtol, what="jacobian")
        if jacobian and self.elementwise:
            # (b) elementwise: bias = -J^T F (harmonic_bias.py:127, metad.py:200, abf.py:223)
            F = self.native.view("cv_force").cpu().numpy().ravel()[: Jo.shape[0]]
            Fo = np.linalg.lstsq(-Jo.T, obias.ravel(), rcond=None)[0] if np.any(Jo) else np.zeros_like(F)
            resid = (sbias - obias).ravel() + Jo.T @ (F - Fo)
            terms = np.abs(Jo).T @ np.maximum(np.abs(F), np.abs(Fo))
            worst = np.max(np.abs(resid) - rtol * terms)
            assert worst <= 1e-13 * scale + 1e-300, f"bias, elementwise: {worst:.3e} beyond rtol * |J|^T |F| (scale {scale:.3e})"
        # forces after the in-place apply
        f = forces_xyz(self.layout, self.dsnap.forces.cpu().numpy())
        fo = forces_xyz(self.layout, self.osnap.forces)
        if self.layout == "openmm-cuda":
            # fixed point 2^-32: the truncating cast may differ by one unit when the fp64 bias differs in
            # its last bits (SURVEY.md trap 12)
            assert np.max(np.abs(f.astype(np.int64) - fo.astype(np.int64))) <= max(2, int(2**32 * scale * rtol) + 2)
        else:
            ftol = 2e-7 if f.dtype == np.float32 else 1e-13
            close(f, fo, max(ftol, rtol if f.dtype == np.float64 else ftol), what="forces")
        w_scale = None
        if hasattr(o, "Wp") and getattr(o, "Wp") is not None:
            data = self.ohelpers.query(self.osnap)
            Jd = Jo if Jo is not None else self.jacobian()
            p = np.abs(np.asarray(data.momenta, dtype=np.float64).ravel())
            absJp = np.abs(Jd) @ p
            w_scale = float(np.max(np.abs(np.linalg.pinv(Jd @ Jd.T)) @ absJp))
            self._w_scale = max(getattr(self, "_w_scale", 0.0), w_scale)
        for name in ("heights", "centers", "Fsum", "force", "F", "Wp", "Wp_", "grid_potential", "grid_gradient"):
            if hasattr(o, name) and getattr(o, name) is not None:
                ov = getattr(o, name).numpy()
                sv = getattr(s, name).cpu().numpy().reshape(ov.shape)
                ref = None
                if w_scale is not None and name in ("Wp", "Wp_"):
                    ref = max(np.max(np.abs(ov)), self._w_scale)
                elif w_scale is not None and name in ("Fsum", "force", "F"):
                    unit = abs(float(self.ohelpers.to_force_units(1.0)))
                    calls = int(o.ncalls) if name == "Fsum" else 1
                    ref = max(np.max(np.abs(ov)), calls * 4.0 * self._w_scale * unit / float(self.osnap.dt))
                close(sv, ov, rtol, scale=ref, what=name)
        if hasattr(o, "hist"):
            assert np.array_equal(s.hist.cpu().numpy().astype(np.int64), o.hist.numpy()), "hist must be bit-exact"
        if hasattr(o, "idx"):
            assert int(s.idx) == int(o.idx)
