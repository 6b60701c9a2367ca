"""
TEST INFRASTRUCTURE (never imported by the product path): restatement of multiple-walker
metadynamics with shared hills.  The reference has no such method (replicas never communicate,
methods/core.py:169-236); the semantics are defined in SURVEY.md 8(e) as the natural extension of
metad.py:187-323: on a deposit step every walker computes its own hill exactly like the
single-walker update would (height from ITS OWN potential before any of this step's hills,
metad.py:219-229; in grid mode only when its xi is inside the grid, metad.py:258-271), the hills of
all W walkers are appended to every walker's arrays in rank order (and added to its grids,
metad.py:251-256), and the generalized force is evaluated afterwards (metad.py:282-314).
**Parity unpinned**: there is no reference output for this.
"""
import numpy as np
import torch

from . import colvars, grids
from .methods import F64, MetadynamicsState, _gather, _natoms, apply_restraints, gaussian, sum_of_gaussians


class SharedHillsWalker:
    """One walker; same two-phase interface as pysages_b200.parallel.NativeWalkerEngine."""

    def __init__(self, method, snapshot, helpers, bias_fn=None):
        self.m = method
        self.helpers = helpers
        self.bias_fn = bias_fn
        _, init, self._update = method.build(snapshot, helpers)
        self.state = init()
        self.periods = colvars.get_periods(method.cvs)
        self.k = int(self.state.xi.numel())
        if method.grid is not None:
            self.mesh = torch.as_tensor(grids.compute_mesh(method.grid))
            self.index = grids.build_indexer(method.grid)
            self.gshape = tuple(int(s) for s in method.grid.shape)
        self._ctx = None

    def _oob(self, I):
        return any(int(i) == s for i, s in zip(I, self.gshape))

    def next_step_deposits(self):
        n = self.state.ncalls + 1
        return n > 1 and n % self.m.stride == 1

    def step(self, snapshot, timestep=0):
        assert not self.next_step_deposits()
        self.state = self._update(snapshot, self.state)
        if self.bias_fn is not None:
            self.bias_fn(snapshot, self.state.bias.numpy())

    def walker_begin(self, snapshot, timestep=0):
        m, s = self.m, self.state
        xi, Jxi = m.cv(self.helpers.query(snapshot))
        I = self.index(xi.detach().numpy()) if m.grid is not None else None
        valid = m.grid is None or not self._oob(I)
        h = 0.0
        if valid:
            if m.deltaT is None:
                h = float(m.height)
            else:
                if m.grid is None:
                    V = sum_of_gaussians(xi.to(F64), s.heights, s.centers, s.sigmas, self.periods)
                else:
                    V = _gather(s.grid_potential, I)
                h = float(m.height * torch.exp(-V / (m.deltaT * m.kB)))
        self._ctx = (xi, Jxi, I)
        rec = np.zeros(self.k + 2)
        rec[: self.k] = xi.detach().numpy().ravel()
        rec[self.k] = h
        rec[self.k + 1] = 1.0 if valid else 0.0
        return torch.as_tensor(rec)

    def walker_finish(self, snapshot, records):
        m, s = self.m, self.state
        xi, Jxi, I = self._ctx
        k = self.k
        heights, centers = s.heights.clone(), s.centers.clone()
        gp = None if s.grid_potential is None else s.grid_potential.clone()
        gg = None if s.grid_gradient is None else s.grid_gradient.clone()
        idx = s.idx
        for rec in torch.as_tensor(records, dtype=F64).reshape(-1, k + 2):
            if float(rec[k + 1]) == 0.0:
                continue
            if idx < heights.shape[0]:  # out-of-range scatters are dropped (JAX default)
                heights[idx] = rec[k]
                centers[idx] = rec[:k]
            idx += 1
            if m.grid is not None:
                mesh = self.mesh.clone().requires_grad_(True)
                delta = colvars.wrap(mesh - rec[:k].reshape(1, -1), self.periods)
                vals = gaussian(rec[k], s.sigmas, delta)
                (grads,) = torch.autograd.grad(vals.sum(), mesh)
                if gp is not None:
                    gp = gp + vals.detach().reshape(gp.shape)
                gg = gg + grads.reshape(gg.shape)
        if m.grid is None:
            x = xi.detach().clone().to(F64).requires_grad_(True)
            V = sum_of_gaussians(x, heights, centers, s.sigmas, self.periods)
            (F,) = torch.autograd.grad(V, x)
        elif self._oob(I):
            if m.restraints:
                lo, hi, kl, kh = m.restraints
                F = apply_restraints(lo, hi, kl, kh, xi.reshape(len(self.gshape)).to(F64))
            else:
                F = torch.zeros(len(self.gshape), dtype=F64)
        else:
            F = _gather(gg, I)
        bias = (-Jxi.T.to(F64) @ F.flatten()).reshape(_natoms(snapshot), -1)
        self.state = MetadynamicsState(xi, bias, heights, centers, s.sigmas, gp, gg, idx, s.ncalls + 1)
        if self.bias_fn is not None:
            self.bias_fn(snapshot, bias.numpy())
