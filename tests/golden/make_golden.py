#!/usr/bin/env python
"""
Builds tests/golden/reference_fixtures.npz from the reference checkout's own test
fixtures and known-answer values (run in the build container, where
/root/reference exists; the GPU box only sees the committed .npz).

Sources (relative to /root/reference):
  tests/ci2_1.txt, ci2_2.txt, ci2_1t.txt, ci2_12.txt   columns 6:9 (x, y, z), 1064 atoms
      -> RMSD known answers of tests/test_rmsd.py:12-14
  tests/test_cvs.py:7-12,28,32                          2-atom Distance known answer
  tests/test_grids.py:70-79,91-92                       bin-index known answers
  examples/inputs/alanine-dipeptide/adp-vacuum.pdb      22-atom ADP coordinates (Angstrom)
Only numeric test DATA is extracted; no reference source code is copied.
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.npz")


def ci2(name):
    return np.loadtxt(os.path.join(REF, "tests", name), dtype=str)[:, 6:9].astype(np.float64)


def pdb_xyz(path):
    xyz = []
    with open(path) as f:
        for line in f:
            if line.startswith(("ATOM", "HETATM")):
                xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.asarray(xyz, dtype=np.float64)


pi = np.pi
data = dict(
    ci2_1=ci2("ci2_1.txt"),
    ci2_2=ci2("ci2_2.txt"),
    ci2_1t=ci2("ci2_1t.txt"),
    ci2_12=ci2("ci2_12.txt"),
    # tests/test_rmsd.py:12-14  (f1, f2, expected)
    rmsd_expected=np.array([11.7768, 0.000493294127, 15.422431]),
    # tests/test_cvs.py:7-12, 28, 32
    distance_positions=np.array([[0.0929926, 0.966452, 0.424666], [0.415969, 0.485482, 0.408579]]),
    distance_expected=np.array(0.57957285),
    # tests/test_grids.py:70-79 : 1-D grid on [-pi, pi], 64 bins
    grid1d_x_regular=np.array([-pi - 1, -pi, 0.0]),
    grid1d_i_regular=np.array([64, 0, 32], dtype=np.uint32),
    grid1d_x_cheb=np.array([-pi - 1, -pi, -pi * 0.996, 0.0]),
    grid1d_i_cheb=np.array([64, 0, 1, 32], dtype=np.uint32),
    grid1d_x_periodic=np.array([-pi * (1 + 0.99 / 32), -pi, pi, 0.0]),
    grid1d_i_periodic=np.array([63, 0, 0, 32], dtype=np.uint32),
    # tests/test_grids.py:81-92 : 2-D grid on [-pi, pi] x [-1, 1], 64 x 32 bins
    grid2d_x=np.array([[-pi, 1.0], [pi, -2.0]]),
    grid2d_i=np.array([[0, 32], [64, 32]], dtype=np.uint32),
    adp_xyz_angstrom=pdb_xyz(os.path.join(REF, "examples/inputs/alanine-dipeptide/adp-vacuum.pdb")),
)
np.savez_compressed(OUT, **data)
print("wrote", OUT, {k: v.shape for k, v in data.items()})
