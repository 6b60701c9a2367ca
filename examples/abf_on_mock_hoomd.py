"""
ABF on two Distance CVs, driven exactly like a PySAGES script -- `run(method, generate_context, steps)` --
with a stand-in HOOMD engine that owns device arrays in HOOMD's layout (no MD engine is installable in
the build image; with hoomd-dlext present, return a real `hoomd.Simulation` from `generate_context`
and nothing else changes).  Needs a CUDA device.

    python examples/abf_on_mock_hoomd.py
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import pysages_b200 as pysages  # noqa: E402
from pysages_b200 import synthetic  # noqa: E402
from pysages_b200.backends import mock  # noqa: E402
from pysages_b200.colvars import Distance  # noqa: E402
from pysages_b200.methods import ABF, CVRestraints, HistogramLogger  # noqa: E402

N_ATOMS, STEPS = 20_000, 2_000


def generate_context(**kwargs):
    snap = synthetic.make_snapshot(N_ATOMS, layout="hoomd", dtype=np.float32, seed=1)
    traj = synthetic.make_trajectory(snap, frames=64, step=0.02)  # the "MD": a random walk, replayed cyclically
    return mock.MockEngine("hoomd", snap["arrays"], snap["box_H"], snap["origin"], snap["dt"], device="cuda",
                           trajectory=traj)


def main():
    quarter = N_ATOMS // 4
    groups = [list(range(i * quarter, (i + 1) * quarter)) for i in range(4)]
    cvs = [Distance([groups[0], groups[1]]), Distance([groups[2], groups[3]])]
    grid = pysages.Grid(lower=(0.0, 0.0), upper=(2.0, 2.0), shape=(32, 32))
    restraints = CVRestraints(lower=(0.0, 0.0), upper=(2.0, 2.0), kl=(10.0, 10.0), ku=(10.0, 10.0))
    method = ABF(cvs, grid, N=100, restraints=restraints)
    logger = HistogramLogger(period=10)
    result = pysages.run(method, generate_context, STEPS, callback=logger)
    out = pysages.analyze(result)
    state = result.states[0]
    print("steps:", int(state.ncalls), " visited bins:", int((out["histogram"] > 0).sum()),
          " logged CV samples:", len(logger.data))
    print("last xi:", state.xi.cpu().numpy().ravel(), " generalized force:", state.force.cpu().numpy())
    pysages.save(result, "abf_result.pkl")  # restart with pysages.run(pysages.load("abf_result.pkl"), ...)


if __name__ == "__main__":
    main()
