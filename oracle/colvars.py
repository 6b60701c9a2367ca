"""
Oracle (test infrastructure): collective variables and their autodiff Jacobians.

Restates pysages/colvars/{core,coordinates,angles,shape,orientation,utils}.py
with torch (CPU) in place of jax.numpy and torch.autograd in place of
jax.grad / jax.jacrev.  dtype promotion follows JAX's x64 rules: Python floats
are weakly typed, f32 (+) f64 arrays -> f64.
"""

import math
from inspect import signature

import numpy as np
import torch

# --------------------------------------------------------------------------- #
# index / group processing  (pysages/colvars/core.py:277-295, 178-181)
# --------------------------------------------------------------------------- #


def _is_group(obj):
    # core.py:302-314 -- an index is an int or a range; a group is a sequence of them
    if isinstance(obj, (int, np.integer, range)):
        return False
    if isinstance(obj, (list, tuple, np.ndarray)):
        return True
    raise ValueError(f"Invalid indices or group: {obj}")


def _group_size(obj):
    # core.py:317-329
    if isinstance(obj, (int, np.integer)):
        return 1
    if isinstance(obj, range):
        return len(obj)
    return sum(_group_size(o) for o in obj)


def _flatten(obj, out):
    if isinstance(obj, (int, np.integer)):
        out.append(int(obj))
    elif isinstance(obj, range):
        out.extend(obj)
    else:
        for o in obj:
            _flatten(o, out)


def process_groups(indices):
    """core.py:277-295 -> (flat uint32 index array, list of arange slices, one per group)"""
    total = 0
    collected = []
    groups = []
    for obj in indices:
        found = _is_group(obj)
        _flatten(obj, collected)
        n = _group_size(obj)
        if found:
            groups.append(np.arange(total, total + n, dtype=np.int64))
        total += n
    return np.asarray(collected, dtype=np.uint32), groups


def _check_groups_size(indices, groups, group_length):
    # core.py:171-176
    input_length = np.size(indices, 0) - sum(np.size(g, 0) - 1 for g in groups)
    if input_length != group_length:
        raise ValueError(
            f"Exactly {group_length} indices or groups must be provided (got {input_length})"
        )


class CollectiveVariable:
    """core.py:24-63"""

    multicomponent = False

    def __init__(self, indices, group_length=None):
        indices, groups = process_groups(indices)
        if group_length is not None:
            _check_groups_size(indices, groups, group_length)
        self.indices = indices
        self.groups = groups
        self.requires_box_unwrapping = True


# --------------------------------------------------------------------------- #
# coordinates.py
# --------------------------------------------------------------------------- #


def barycenter(positions):
    # coordinates.py:14-28
    return torch.sum(positions, dim=0) / positions.shape[0]


def weighted_barycenter(positions, weights):
    # coordinates.py:31-53 (the reference loops over atoms; same sum, vectorised)
    return torch.sum(weights.reshape(-1, 1) * positions, dim=0)


def distance(r1, r2):
    # coordinates.py:91-109
    return torch.linalg.norm(r1 - r2)


class Component(CollectiveVariable):
    # core.py:67-86 (AxisCV) + coordinates.py:56-71
    def __init__(self, indices, axis, group_length=None):
        if axis not in (0, 1, 2):
            raise RuntimeError(f"Invalid Cartesian axis {axis} index choose 0 (X), 1 (Y), 2 (Z)")
        super().__init__(indices, group_length)
        self.axis = axis

    @property
    def function(self):
        return lambda rs: barycenter(rs)[self.axis]


class Distance(CollectiveVariable):
    # coordinates.py:74-88
    def __init__(self, indices):
        super().__init__(indices, 2)

    @property
    def function(self):
        if len(self.groups) == 0:
            return distance
        return lambda r1, r2: distance(barycenter(r1), barycenter(r2))


class Displacement(CollectiveVariable):
    # coordinates.py:112-148
    multicomponent = True

    def __init__(self, indices):
        super().__init__(indices, 2)

    @property
    def function(self):
        if len(self.groups) == 0:
            return lambda r1, r2: r2 - r1
        return lambda r1, r2: barycenter(r2) - barycenter(r1)


# --------------------------------------------------------------------------- #
# angles.py
# --------------------------------------------------------------------------- #


def angle(p1, p2, p3):
    # angles.py:46-74
    q = p1 - p2
    r = p3 - p2
    return torch.atan2(torch.linalg.norm(torch.linalg.cross(q, r)), torch.dot(q, r))


def dihedral_angle(p1, p2, p3, p4):
    # angles.py:95-128
    q = p3 - p2
    r = torch.linalg.cross(p2 - p1, q)
    s = torch.linalg.cross(q, p4 - p3)
    return torch.atan2(
        torch.dot(torch.linalg.cross(r, s), q), torch.dot(r, s) * torch.linalg.norm(q)
    )


class Angle(CollectiveVariable):
    # angles.py:24-43
    def __init__(self, indices):
        super().__init__(indices, 3)

    @property
    def function(self):
        return angle


class DihedralAngle(CollectiveVariable):
    # angles.py:77-92
    def __init__(self, indices):
        super().__init__(indices, 4)

    @property
    def function(self):
        return dihedral_angle


# --------------------------------------------------------------------------- #
# shape.py
# --------------------------------------------------------------------------- #


def radius_of_gyration(positions):
    # shape.py:38-57: sum_i r_i.r_i / n, broadcast onto a 3-vector (quirk kept)
    n = positions.shape[0]
    rog = torch.zeros(3, dtype=positions.dtype) + torch.sum(positions * positions)
    return rog / n


class RadiusOfGyration(CollectiveVariable):
    """
    shape.py:14-35.  The reference returns a (3,) vector whose components are all
    equal and cannot differentiate it (not @multicomponent, so jax.grad rejects
    it).  PARITY UNPINNED for the Jacobian: the oracle defines the scalar CV as
    component 0 of the reference vector and differentiates that.
    """

    @property
    def function(self):
        return lambda rs: radius_of_gyration(rs)[0]

    @property
    def reference_vector_function(self):
        return radius_of_gyration


def _local_reference_systems(rs):
    # orientation.py:177-216: rs (n, 3, 3) -> unit axes (n, 3, 3) [x, y, z] and origins (n, 3)
    origins = rs.mean(dim=1)
    x = rs[:, 0] - origins
    x_hat = x / torch.linalg.norm(x, dim=-1, keepdim=True)
    z = torch.linalg.cross(x_hat, rs[:, 1] - origins, dim=1)
    z_hat = z / torch.linalg.norm(z, dim=-1, keepdim=True)
    y = torch.linalg.cross(z_hat, x_hat, dim=1)
    y_hat = y / torch.linalg.norm(y, dim=1, keepdim=True)
    return torch.stack((x_hat, y_hat, z_hat), dim=1), origins


def _g_vector(systems, origins, cutoff, a, b):
    # orientation.py:219-288
    pairs = origins[:, None] - origins[None, :]
    r = torch.einsum("ijk,jlk->ijl", pairs, systems)
    n = r.shape[0]
    eye = torch.eye(n, dtype=torch.bool)
    r = torch.where(eye[:, :, None], torch.full_like(r, 10.0), r)  # masked self entries: outside the cutoff
    r_tilde = r * torch.tensor([1 / a, 1 / a, 1 / b], dtype=r.dtype)
    norm = torch.linalg.norm(r_tilde, dim=-1)
    gamma = math.pi / cutoff
    inv = torch.where(norm != 0, 1 / norm, torch.zeros_like(norm))
    G123 = r_tilde * (torch.sin(gamma * norm) * inv)[:, :, None]
    G4 = 1 + torch.cos(gamma * norm)
    G = torch.cat((G123, G4[:, :, None]), dim=-1)
    return G * ((cutoff - norm > 0).to(r.dtype) / gamma)[:, :, None]


def ermsd(rs, reference, cutoff, a, b):
    # orientation.py:291-378
    n = rs.shape[0] // 3
    sys_, org = _local_reference_systems(rs.reshape(n, 3, 3))
    ref = torch.as_tensor(np.asarray(reference, dtype=np.float64)).reshape(n, 3, 3).to(rs.dtype)
    rsys, rorg = _local_reference_systems(ref)
    Gs, Gr = _g_vector(sys_, org, cutoff, a, b), _g_vector(rsys, rorg, cutoff, a, b)
    return torch.sqrt(torch.sum((Gs - Gr) ** 2) / n)


class ERMSD(CollectiveVariable):
    # orientation.py:129-175
    def __init__(self, indices, references, cutoff=2.4, a=0.5, b=0.3):
        super().__init__(indices)
        self.references = np.asarray(references)
        self.cutoff, self.a, self.b = cutoff, a, b

    @property
    def function(self):
        return lambda r: ermsd(r, self.references, self.cutoff, self.a, self.b)


ERMSDCG_LOCAL_COORDINATES = [  # orientation.py:440-461 (RACER-like B1/B2/B3 -> C2/C6/C4 or C2/C4/C6, nm)
    [[-0.1333021, -0.1762476, -0.0000001], [-0.0857174, 0.0499809, -0.0000033], [0.0795015, -0.1210861, -0.0001387]],
    [[-0.0253117, -0.1224880, 0.0003341], [-0.0236469, 0.1239022, -0.0006548], [0.1811280, 0.0000000, -0.0000000]],
    [[-0.1167570, -0.1372227, -0.0002264], [-0.0409667, 0.0959927, -0.0012938], [0.1013132, -0.0983306, 0.0005419]],
    [[-0.0267224, -0.1188717, 0.0004939], [-0.0254680, 0.1138277, 0.0008671], [0.1815304, -0.0000000, -0.0000000]],
]


def _infer_base_coordinates(systems, origins, sequence, local_coordinates):
    # orientation.py:499-540: einsum("jlk,jml->jmk", systems, local_coordinates[sequence]) + origins
    lc = torch.as_tensor(np.asarray(local_coordinates, dtype=np.float64))[torch.as_tensor(np.asarray(sequence, dtype=np.int64))]
    return torch.einsum("jlk,jml->jmk", systems, lc.to(systems.dtype)) + origins[:, None, :]


def ermsd_cg(rs, reference, sequence, local_coordinates, cutoff, a, b):
    # orientation.py:543-590
    n = rs.shape[0] // 3
    ref = torch.as_tensor(np.asarray(reference, dtype=np.float64)).reshape(n, 3, 3).to(rs.dtype)
    sys_, org = _local_reference_systems(rs.reshape(n, 3, 3))
    rsys, rorg = _local_reference_systems(ref)
    atoms = _infer_base_coordinates(sys_, org, sequence, local_coordinates)
    ratoms = _infer_base_coordinates(rsys, rorg, sequence, local_coordinates)
    s2, o2 = _local_reference_systems(atoms)
    rs2, ro2 = _local_reference_systems(ratoms)
    Gs, Gr = _g_vector(s2, o2, cutoff, a, b), _g_vector(rs2, ro2, cutoff, a, b)
    return torch.sqrt(torch.sum((Gs - Gr) ** 2) / n)


class ERMSDCG(CollectiveVariable):
    # orientation.py:380-497 (decorated @multicomponent in the reference although it returns one number)
    def __init__(self, indices, reference, sequence, local_coordinates=None, cutoff=2.4, a=0.5, b=0.3):
        super().__init__(indices)
        self.reference = np.asarray(reference)
        self.sequence = np.asarray(sequence)
        self.local_coordinates = np.asarray(ERMSDCG_LOCAL_COORDINATES if local_coordinates is None else local_coordinates)
        self.cutoff, self.a, self.b = cutoff, a, b

    @property
    def function(self):
        return lambda r: ermsd_cg(r, self.reference, self.sequence, self.local_coordinates, self.cutoff, self.a, self.b)


def ring_puckering_coordinates(rs):
    """angles.py:158-205: Cremer-Pople amplitude q and phase phi (m = 2) of a monocyclic ring."""
    N = rs.shape[0]
    rc = rs - rs.mean(dim=0)  # barycenter, coordinates.py:14-28
    theta = 2 * math.pi * torch.arange(N, dtype=torch.float64) / N
    R1 = rc.T @ torch.sin(theta).to(rs.dtype)  # imag(exp(i theta))
    R2 = rc.T @ torch.cos(theta).to(rs.dtype)
    n = torch.linalg.cross(R1, R2)
    n = n / torch.linalg.norm(n)
    z = rc @ n
    a = math.sqrt(2 / N) * torch.sum(z * torch.cos(2 * theta).to(rs.dtype))
    b = -math.sqrt(2 / N) * torch.sum(z * torch.sin(2 * theta).to(rs.dtype))
    q = torch.sqrt(a**2 + b**2)
    phi = torch.atan2(b, a)
    return torch.stack([q, phi])


class RingPuckeringCoordinates(CollectiveVariable):
    # angles.py:131-155 (@multicomponent: [q, phi])
    multicomponent = True

    @property
    def function(self):
        return ring_puckering_coordinates


class RingPhaseAngle(CollectiveVariable):
    # angles.py:208-242
    @property
    def function(self):
        return lambda rs: ring_puckering_coordinates(rs)[1]


class RingAmplitude(CollectiveVariable):
    # angles.py:245-281
    @property
    def function(self):
        return lambda rs: ring_puckering_coordinates(rs)[0]


def gyration_tensor(positions):
    # shape.py:118-135: sum_i outer(r_i, r_i) / n  (uncentred)
    return positions.T @ positions / positions.shape[0]


def principal_moments(positions):
    # shape.py:161-176: eigvalsh (ascending)
    return torch.linalg.eigvalsh(gyration_tensor(positions))


def asphericity(positions):
    # shape.py:203-221
    l1, l2, l3 = principal_moments(positions)
    return l3 - (l1 + l2) / 2


def acylindricity(positions, axes):
    # shape.py:267-287; jnp indexing clamps out-of-range indices (the reference's pairs go up to 3)
    lambdas = principal_moments(positions)
    l1, l2 = [lambdas[min(i, 2)] for i in axes]
    return l2 - l1


def shape_anisotropy(positions):
    # shape.py:318-344
    l1, l2, l3 = principal_moments(positions)
    return (3 * (l1**2 + l2**2 + l3**2) / (l1 + l2 + l3) ** 2 - 1) / 2


class PrincipalMoment(Component):
    # shape.py:85-115 (AxisCV)
    @property
    def function(self):
        return lambda rs: principal_moments(rs)[self.axis]


class Asphericity(CollectiveVariable):
    # shape.py:179-200
    @property
    def function(self):
        return asphericity


class Acylindricity(CollectiveVariable):
    # shape.py:224-264
    symmetry_axes = {"xy": (1, 2), "yz": (2, 3), "xz": (1, 3)}

    def __init__(self, indices, axes="xy", group_length=None):
        axes = "".join(sorted(axes.lower()))
        if axes not in self.symmetry_axes:
            raise RuntimeError(f"Invalid acylindrity axes specification {axes}.")
        super().__init__(indices, group_length)
        self.axes = axes

    @property
    def function(self):
        return lambda rs: acylindricity(rs, self.symmetry_axes[self.axes])


class ShapeAnisotropy(CollectiveVariable):
    # shape.py:290-315
    @property
    def function(self):
        return shape_anisotropy


# --------------------------------------------------------------------------- #
# orientation.py (RMSD)
# --------------------------------------------------------------------------- #


def fitted_positions(positions, weights):
    # orientation.py:27-30
    return positions - weighted_barycenter(positions, weights)


def kabsch(P, Q):
    # orientation.py:33-73
    C = P.T @ Q
    V, S, W = torch.linalg.svd(C, full_matrices=False)
    dh = torch.linalg.det(V) * torch.linalg.det(W)
    sgn = torch.sign(dh)
    scale = torch.ones(3, dtype=P.dtype)
    scale[-1] = sgn  # V[:, -1] *= sign(dh)
    V = V * scale.reshape(1, 3)
    return V @ W


def rmsd(r1, Q, w_0, optimal_rotation):
    # orientation.py:76-95
    P = fitted_positions(r1, w_0)
    U = optimal_rotation(P, Q)
    optimal_P = P @ U
    return torch.sqrt(torch.sum(torch.square(optimal_P - Q)) / P.shape[0])


class RMSD(CollectiveVariable):
    # orientation.py:98-126
    def __init__(self, indices, references, weights=None):
        if weights is not None and len(indices) != len(weights):
            raise RuntimeError("Indices and weights must be of the same length")
        super().__init__(indices)
        self.references = torch.as_tensor(np.asarray(references, dtype=np.float64))
        n = len(indices)
        if weights is None:
            self.weights = torch.ones(n, dtype=torch.float64) / n
        else:
            self.weights = torch.as_tensor(np.asarray(weights, dtype=np.float64))
        self.Q = fitted_positions(self.references, self.weights)

    @property
    def function(self):
        return lambda r: rmsd(r, self.Q, self.weights, kabsch)


# --------------------------------------------------------------------------- #
# CoordinationNumber -- ABSENT from the reference (PARITY UNPINNED).
# Defined here (and in DESIGN.md) as the PLUMED-style rational switching
# function summed over all A x B pairs, pushed through the reference's
# CV-extension hook (a CollectiveVariable whose `function` takes two groups).
# --------------------------------------------------------------------------- #


class CoordinationNumber(CollectiveVariable):
    def __init__(self, indices, r0=1.0, n=6, m=12):
        super().__init__(indices, 2)
        if len(self.groups) != 2:
            raise ValueError("CoordinationNumber needs two groups of atoms")
        self.r0, self.n, self.m = float(r0), int(n), int(m)

    @property
    def function(self):
        r0, n, m = self.r0, self.n, self.m

        def coordination(ra, rb, chunk=512):
            total = torch.zeros((), dtype=ra.dtype)
            for s in range(0, ra.shape[0], chunk):
                d = ra[s : s + chunk, None, :] - rb[None, :, :]
                x2 = torch.sum(d * d, dim=-1) / (r0 * r0)
                if m == 2 * n and n % 2 == 0:
                    sw = 1.0 / (1.0 + x2 ** (n // 2))
                else:
                    x = torch.sqrt(x2)
                    sw = (1.0 - x**n) / (1.0 - x**m)
                total = total + torch.sum(sw)
            return total

        return coordination


# --------------------------------------------------------------------------- #
# build  (core.py:184-274)
# --------------------------------------------------------------------------- #


def _get_nargs(function):
    return len(signature(function).parameters)


def _evaluator(cv):
    xi = cv.function
    idx = torch.as_tensor(cv.indices.astype(np.int64))
    gps = [torch.as_tensor(g) for g in cv.groups]

    if _get_nargs(xi) == 1:  # core.py:190-194

        def evaluate(positions, ids):
            pos = positions[ids[idx]]
            return xi(pos)

    elif len(gps) == 0:  # core.py:196-200

        def evaluate(positions, ids):
            pos = positions[ids[idx]]
            return xi(*pos)

    else:  # core.py:202-207 (bare indices mixed with groups are silently dropped)

        def evaluate(positions, ids):
            pos = positions[ids[idx]]
            pos = [pos[g] for g in gps]
            return xi(*pos)

    return evaluate


def _build(cv, differentiate=True):
    evaluate = _evaluator(cv)

    def apply(pos, ids):
        if not differentiate:
            with torch.no_grad():
                return evaluate(pos, ids).reshape(1, -1)
        p = pos.detach().clone().requires_grad_(True)
        val = evaluate(p, ids)
        xi = val.detach().reshape(1, -1)
        comps = val.reshape(-1)
        rows = []
        for c in range(comps.shape[0]):  # grad (scalar) / jacrev (multicomponent), core.py:298-299
            (g,) = torch.autograd.grad(comps[c], p, retain_graph=c + 1 < comps.shape[0])
            rows.append(g.reshape(1, -1))
        return xi, torch.cat(rows, dim=0)  # core.py:216-217

    return apply


def build(*cvs, differentiate=True):
    """core.py:228-274 -> apply(data) with data.positions (N, >=3), data.indices (N,)"""
    fns = [_build(cv, differentiate) for cv in cvs]

    def apply(data):
        pos = torch.as_tensor(data.positions)[:, :3]
        ids = torch.as_tensor(np.asarray(data.indices).astype(np.int64))
        if differentiate:
            xis, grads = [], []
            for f in fns:
                xi, J = f(pos, ids)
                xis.append(xi)
                grads.append(J)
            return torch.hstack(xis), torch.vstack(grads)
        return torch.hstack([f(pos, ids) for f in fns])

    return apply


# --------------------------------------------------------------------------- #
# utils.py
# --------------------------------------------------------------------------- #


def get_periods(cvs):
    # colvars/utils.py:13-19 -- exact type match, subclasses are NOT periodic
    return torch.tensor(
        [2 * math.pi if type(cv) in (Angle, DihedralAngle) else math.inf for cv in cvs],
        dtype=torch.float64,
    )


def wrap(x, P):
    # colvars/utils.py:22-26 (the inf branch is masked out before the arithmetic so that
    # autograd never sees 0 * inf; the selected values are identical)
    Pf = torch.where(torch.isinf(P), torch.ones_like(P), P)
    return torch.where(torch.isinf(P), x, x - (torch.abs(x) > Pf / 2) * torch.sign(x) * Pf)
