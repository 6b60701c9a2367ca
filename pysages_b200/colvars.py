"""
Collective variables: host-side descriptors for the fused CUDA step.

Mirrors the reference constructors (pysages/colvars/core.py:24-147, coordinates.py, angles.py,
shape.py:14-35, orientation.py:98-126): same names, same index / group conventions, same errors.
No numerical work happens here; `describe()` lowers a CV to the `psb_cv_desc` consumed by
psb_create (include/pysages_b200.h) and the kernels in csrc/psb_step.cuh evaluate the CV and its
closed-form Jacobian on the device.
"""

import math
from numbers import Integral

import numpy as np

from . import _native as N


def _is_group(obj):
    # colvars/core.py:302-314
    if isinstance(obj, (Integral, np.integer, range)):
        return False
    if isinstance(obj, (list, tuple, np.ndarray)):
        return True
    raise ValueError(f"Invalid indices or group: {obj}")


def _flatten(obj, out):
    if isinstance(obj, (Integral, np.integer)):
        out.append(int(obj))
    elif isinstance(obj, range):
        out.extend(obj)
    else:
        for o in obj:
            _flatten(o, out)


def _process_groups(indices):
    """colvars/core.py:277-295: flat uint32 indices + one arange slice per group."""
    total = 0
    collected, groups = [], []
    for obj in indices:
        found = _is_group(obj)
        before = len(collected)
        _flatten(obj, collected)
        n = len(collected) - before
        if found:
            groups.append(np.arange(total, total + n, dtype=np.uint32))
        total += n
    if any(i < 0 for i in collected):
        raise ValueError("particle indices must be non-negative")
    return np.asarray(collected, dtype=np.uint32), groups


def _check_groups_size(indices, groups, group_length):
    # colvars/core.py:171-176
    input_length = np.size(indices, 0) - sum(np.size(g, 0) - 1 for g in groups)
    if input_length != group_length:
        raise ValueError(
            f"Exactly {group_length} indices or groups must be provided (got {input_length})"
        )


class CollectiveVariable:
    """colvars/core.py:24-63"""

    _multicomponent = False
    _kind = None
    _nargs = 1  # arity of the reference's `function` (colvars/core.py:190-207)

    def __init__(self, indices, group_length=None):
        indices, groups = _process_groups(indices)
        if group_length is not None:
            _check_groups_size(indices, groups, group_length)
        self.indices = indices
        self.groups = groups
        self.requires_box_unwrapping = True

    @property
    def multicomponent(self):
        return self._multicomponent

    @property
    def ncomponents(self):
        return 3 if self._multicomponent else 1

    def points(self):
        """
        The atom groups whose barycentres are the CV's arguments, following the three
        evaluation modes of colvars/core.py:190-207:
          1-argument functions get the whole gathered block;
          without groups every index is one point;
          with groups only the groups are used (bare indices are silently dropped -- quirk kept).
        """
        if self._nargs == 1:
            return [self.indices]
        if len(self.groups) == 0:
            return [self.indices[i : i + 1] for i in range(len(self.indices))]
        return [self.indices[g] for g in self.groups]

    def describe(self, keep):
        """-> psb_cv_desc; `keep` collects the numpy buffers the descriptor points into."""
        pts = self.points()
        if len(pts) != self._nargs:
            raise ValueError(
                f"Exactly {self._nargs} indices or groups must be provided (got {len(pts)})"
            )
        tags = np.ascontiguousarray(np.concatenate(pts), dtype=np.uint32)
        offsets = np.zeros(len(pts) + 1, dtype=np.uint32)
        offsets[1:] = np.cumsum([len(p) for p in pts])
        keep += [tags, offsets]
        d = N.CvDesc()
        d.kind = self._kind
        d.axis = getattr(self, "axis", 0)
        d.n_groups = len(pts)
        d.tags = tags.ctypes.data_as(N.POINTER(N.c_uint32))
        d.group_offsets = offsets.ctypes.data_as(N.POINTER(N.c_uint32))
        return d


class AxisCV(CollectiveVariable):
    """colvars/core.py:67-86"""

    def __init__(self, indices, axis, group_length=None):
        if axis not in (0, 1, 2):
            raise RuntimeError(f"Invalid Cartesian axis {axis} index choose 0 (X), 1 (Y), 2 (Z)")
        super().__init__(indices, group_length)
        self.axis = axis


class TwoPointCV(CollectiveVariable):
    _nargs = 2

    def __init__(self, indices):
        super().__init__(indices, 2)


class ThreePointCV(CollectiveVariable):
    _nargs = 3

    def __init__(self, indices):
        super().__init__(indices, 3)


class FourPointCV(CollectiveVariable):
    _nargs = 4

    def __init__(self, indices):
        super().__init__(indices, 4)


class Component(AxisCV):
    """colvars/coordinates.py:56-71: Cartesian component of the barycentre of the selection."""

    _kind = N.CV_COMPONENT


class Distance(TwoPointCV):
    """colvars/coordinates.py:74-109: distance between two points / group barycentres."""

    _kind = N.CV_DISTANCE


class Displacement(TwoPointCV):
    """colvars/coordinates.py:112-148 (multicomponent, k = 3): r2 - r1."""

    _kind = N.CV_DISPLACEMENT
    _multicomponent = True


class Angle(ThreePointCV):
    """colvars/angles.py:24-74.  Groups are taken as barycentres (extension: the reference only
    works with three bare indices)."""

    _kind = N.CV_ANGLE


class DihedralAngle(FourPointCV):
    """colvars/angles.py:77-128."""

    _kind = N.CV_DIHEDRAL


class RadiusOfGyration(CollectiveVariable):
    """
    colvars/shape.py:14-57.  Value parity with the reference quirk: sum_i r_i.r_i / n (uncentred,
    squared, no sqrt); the reference returns it replicated on a 3-vector and cannot differentiate
    it, here it is one scalar component with Jacobian 2 r_i / n (parity unpinned).
    """

    _kind = N.CV_ROG


class _GyrationCV(CollectiveVariable):
    """Functions of the eigenvalues of the reference's (uncentred) gyration tensor sum_i r_i r_i^T / n
    (colvars/shape.py:118-176): one scalar component, Jacobian (2/n) W r_i."""

    _kind = N.CV_GYRATION
    _mode = 0
    _pair = (0, 0)

    def describe(self, keep):
        d = super().describe(keep)
        d.axis = self._mode
        d.nn, d.mm = self._pair
        return d


class PrincipalMoment(_GyrationCV):
    """colvars/shape.py:85-176: eigenvalue `axis` (ascending) of the gyration tensor."""

    def __init__(self, indices, axis, group_length=None):
        if axis not in (0, 1, 2):
            raise RuntimeError(f"Invalid Cartesian axis {axis} index choose 0 (X), 1 (Y), 2 (Z)")
        super().__init__(indices, group_length)
        self.axis = axis
        self._mode = axis


class Asphericity(_GyrationCV):
    """colvars/shape.py:179-221: l3 - (l1 + l2) / 2."""

    _mode = N.GYR_ASPHERICITY


class Acylindricity(_GyrationCV):
    """colvars/shape.py:224-287: l_k - l_j.  The reference indexes the (0-based) eigenvalue array with
    the pairs xy -> (1, 2), yz -> (2, 3), xz -> (1, 3); JAX clamps the out-of-range 3 to 2, so "yz" is
    identically 0 and "xz" equals "xy" -- reproduced here."""

    _mode = N.GYR_ACYLINDRICITY
    symmetry_axes = {"xy": (1, 2), "yz": (2, 3), "xz": (1, 3)}

    def __init__(self, indices, axes="xy", group_length=None):
        axes = "".join(sorted(axes.lower()))
        if axes not in self.symmetry_axes:
            raise RuntimeError(f"Invalid acylindrity axes specification {axes}."
                               f" Valid options are: {tuple(self.symmetry_axes.keys())}.")
        super().__init__(indices, group_length)
        self.axes = axes
        j, k = self.symmetry_axes[axes]
        self._pair = (min(j, 2), min(k, 2))


class ShapeAnisotropy(_GyrationCV):
    """colvars/shape.py:290-344: (3 sum l^2 / (sum l)^2 - 1) / 2."""

    _mode = N.GYR_ANISOTROPY


class _RingCV(CollectiveVariable):
    """Cremer-Pople puckering (m = 2) of a monocyclic ring (colvars/angles.py:131-281): the indices
    are the ring atoms in ring order.  Gradient in closed form (csrc/psb_math.cuh: ring_cv)."""

    _kind = N.CV_RING
    _mode = N.RING_AMPLITUDE

    def describe(self, keep):
        if len(self.indices) < 3:
            raise ValueError("a ring needs at least 3 atoms")
        d = super().describe(keep)
        d.axis = self._mode
        return d


class RingAmplitude(_RingCV):
    """colvars/angles.py:245-281: puckering amplitude q (same unit as the coordinates)."""

    _mode = N.RING_AMPLITUDE


class RingPhaseAngle(_RingCV):
    """colvars/angles.py:208-242: puckering phase phi in (-pi, pi] (not treated as periodic by
    get_periods, like the reference: colvars/utils.py:13-19 only knows Angle / DihedralAngle)."""

    _mode = N.RING_PHASE


class RingPuckeringCoordinates(_RingCV):
    """colvars/angles.py:131-155 (multicomponent, k = 2): [q, phi].  Lowered to two kernel-level CVs
    over the same atoms (amplitude, phase)."""

    _multicomponent = True

    @property
    def ncomponents(self):
        return 2

    def describe(self, keep):
        out = []
        for mode in (N.RING_AMPLITUDE, N.RING_PHASE):
            self._mode = mode
            out.append(super().describe(keep))
        return out


class RMSD(CollectiveVariable):
    """colvars/orientation.py:98-126 (Kabsch-fitted RMSD to `references`)."""

    _kind = N.CV_RMSD

    def __init__(self, indices, references, weights=None):
        if weights is not None and len(indices) != len(weights):
            raise RuntimeError("Indices and weights must be of the same length")
        super().__init__(indices)
        self.references = np.ascontiguousarray(references, dtype=np.float64)
        if self.references.shape != (len(self.indices), 3):
            raise ValueError("references must have shape (len(indices), 3)")
        self.weights = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)

    def describe(self, keep):
        d = super().describe(keep)
        keep.append(self.references)
        d.reference = self.references.ctypes.data_as(N.POINTER(N.c_double))
        if self.weights is not None:
            keep.append(self.weights)
            d.weights = self.weights.ctypes.data_as(N.POINTER(N.c_double))
        return d


class ERMSD(CollectiveVariable):
    """colvars/orientation.py:129-175: eRMSD of n nucleotides to `references` (Bottaro et al., NAR 2014).
    `indices`: 3 ring atoms per base -- C2, C6, C4 for A / G and C2, C4, C6 for U / C -- as one flat list;
    `references` (3n, 3) in the same order.  Evaluated by one CTA with hand-written adjoints
    (csrc/psb_ermsd.cuh); up to 256 bases."""

    _kind = N.CV_ERMSD

    def __init__(self, indices, references, cutoff=2.4, a=0.5, b=0.3):
        super().__init__(indices)
        self.references = np.ascontiguousarray(references, dtype=np.float64)
        if len(self.indices) % 3 != 0:
            raise ValueError(f"Please make sure the indices are in dimension 3 * n_nucleotides! Now it's {len(self.indices)}.")
        if self.references.shape != (len(self.indices), 3):
            raise ValueError("references must have shape (len(indices), 3)")
        self.cutoff, self.a, self.b = float(cutoff), float(a), float(b)

    def describe(self, keep):
        d = super().describe(keep)
        ab = np.array([self.a, self.b], dtype=np.float64)
        keep += [self.references, ab]
        d.reference = self.references.ctypes.data_as(N.POINTER(N.c_double))
        d.weights = ab.ctypes.data_as(N.POINTER(N.c_double))
        d.r0 = self.cutoff
        return d


class ERMSDCG(ERMSD):
    """colvars/orientation.py:380-497: eRMSD of a coarse-grained RNA model.  `indices`: the three base sites
    B1 / B2 / B3 per nucleotide (A: C8, N6, C2; U: C6, O4, O2; G: C8, O6, N2; C: C6, N4, O2); `sequence`: 0 (A),
    1 (U), 2 (G), 3 (C) per nucleotide; `local_coordinates` (4, 3, 3): the ideal C2 / C6 / C4 (A, G) or C2 / C4 / C6
    (U, C) atoms in the frame of the sites (default: the reference's RACER -> SPQR table, nm).  The kernel
    reconstructs those atoms, runs the eRMSD on them and back-propagates through the reconstruction."""

    DEFAULT_LOCAL_COORDINATES = [
        [[-0.1333021, -0.1762476, -0.0000001], [-0.0857174, 0.0499809, -0.0000033], [0.0795015, -0.1210861, -0.0001387]],
        [[-0.0253117, -0.1224880, 0.0003341], [-0.0236469, 0.1239022, -0.0006548], [0.1811280, 0.0000000, -0.0000000]],
        [[-0.1167570, -0.1372227, -0.0002264], [-0.0409667, 0.0959927, -0.0012938], [0.1013132, -0.0983306, 0.0005419]],
        [[-0.0267224, -0.1188717, 0.0004939], [-0.0254680, 0.1138277, 0.0008671], [0.1815304, -0.0000000, -0.0000000]],
    ]

    def __init__(self, indices, reference, sequence, local_coordinates=None, cutoff=2.4, a=0.5, b=0.3):
        super().__init__(indices, reference, cutoff, a, b)
        self.reference = self.references
        self.sequence = np.asarray(sequence, dtype=np.int64)
        if self.sequence.shape != (len(self.indices) // 3,) or self.sequence.min() < 0 or self.sequence.max() > 3:
            raise ValueError("sequence needs one entry in {0 (A), 1 (U), 2 (G), 3 (C)} per nucleotide")
        lc = self.DEFAULT_LOCAL_COORDINATES if local_coordinates is None else local_coordinates
        self.local_coordinates = np.ascontiguousarray(lc, dtype=np.float64)
        if self.local_coordinates.shape != (4, 3, 3):
            raise ValueError("local_coordinates must have shape (4, 3, 3)")

    def describe(self, keep):
        d = super().describe(keep)
        packed = np.concatenate([[self.a, self.b], self.local_coordinates.ravel(), self.sequence.astype(np.float64)])
        keep.append(packed)
        d.weights = packed.ctypes.data_as(N.POINTER(N.c_double))
        d.axis = 1
        return d


class CoordinationNumber(TwoPointCV):
    """
    Pairwise coordination number between two groups (extension; absent from the reference):
    xi = sum_{i in A, j in B} (1 - x^n) / (1 - x^m), x = r_ij / r0, defaults n = 6, m = 12.
    """

    _kind = N.CV_COORDINATION

    def __init__(self, indices, r0=1.0, n=6, m=12):
        super().__init__(indices)
        if len(self.groups) != 2:
            raise ValueError("CoordinationNumber needs two groups of atoms")
        self.r0, self.n, self.m = float(r0), int(n), int(m)

    def describe(self, keep):
        d = super().describe(keep)
        d.r0, d.nn, d.mm = self.r0, self.n, self.m
        return d


def get_periods(cvs):
    """colvars/utils.py:13-19: 2*pi for EXACT Angle / DihedralAngle types, inf otherwise;
    one entry per CV component."""
    out = []
    for cv in cvs:
        p = 2 * math.pi if type(cv) in (Angle, DihedralAngle) else math.inf
        out += [p] * cv.ncomponents
    return out


def wrap(x, P):
    """colvars/utils.py:22-26: x - [|x| > P/2] * sign(x) * P, identity where P is inf (host helper for
    analysis code; on the device the same rule is `wrap_period` in csrc/psb_math.cuh)."""
    x = np.asarray(x, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    finite = np.isfinite(P)
    Pf = np.where(finite, P, 0.0)
    return np.where(finite & (np.abs(x) > Pf / 2), x - np.sign(x) * Pf, x)
