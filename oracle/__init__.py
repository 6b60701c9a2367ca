"""
oracle/ -- TEST INFRASTRUCTURE ONLY.  Not shipped, not on any product path.

CPU restatement (torch-fp64 autograd + numpy) of the PySAGES per-timestep bias
hot path, written function-for-function after the reference sources (each
function cites the reference file:line it follows).  The reference computes
Jacobians with `jax.grad`/`jax.jacrev` over a dense (N, 3) position array; this
restatement does the same with `torch.autograd`, so it is independent of the
closed-form gradients used by the CUDA kernels in `pysages_b200/csrc`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this package, and there only as the checker
or the timed CPU baseline.  `pysages_b200` never imports it.

Pinning status (see DESIGN.md "Oracle"):
  * JAX / plum are not installable in the build container, so the reference
    itself cannot be executed here.  The restatement is pinned against every
    known-answer value the reference's own tests hold for this path
    (tests/test_cvs.py, tests/test_rmsd.py, tests/test_grids.py) -- see
    tests/test_oracle_golden.py -- and its autodiff Jacobians are cross-checked
    against central finite differences.
  * PARITY UNPINNED (no reference implementation or test exists):
    RadiusOfGyration Jacobian (reference cannot differentiate it),
    CoordinationNumber (absent from the reference), shared-hill multiple
    walkers (absent from the reference).  For these the oracle is the defining
    restatement of the semantics documented in DESIGN.md.
"""

from . import backends, colvars, grids, methods, walkers  # noqa: F401
