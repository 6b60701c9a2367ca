"""
Synthetic snapshots in the engines' native layouts: the shared input generator of
the parity tests, the smoke test and bench.py.  Pure numpy; nothing here is on the hot path.

Seeds follow SURVEY.md: numpy.random.Generator(PCG64(seed)), seed = 20240 + 1000*config + replica.
"""

import numpy as np

LAMMPS_IMG_MAX = 512  # SMALLBIG image packing: 10 / 10 / 10 bits, offset 512


def box_length(natoms, density=0.8):
    return float((natoms / density) ** (1.0 / 3.0))


def make_snapshot(natoms, layout="hoomd", dtype=np.float32, seed=20240, density=0.8, globule=False,
                  ntypes=4, dt=0.005):
    """
    -> dict(arrays=..., box_H=(3,3), origin=(3,), dt=float, L=float, layout=str)

    Cubic box L = (N / density)^(1/3), origin -L/2; positions U[-L/2, L/2) (or a compact globule
    N(0, (0.3 N^(1/3))^2) for the RMSD / RoG configuration); velocities N(0,1); masses U[1,16];
    forces N(0,1); tag->slot map a random permutation; images U{-1,0,1}.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n = int(natoms)
    L = box_length(n, density)
    if globule:
        xyz = rng.normal(0.0, 0.3 * n ** (1.0 / 3.0), size=(n, 3))
    else:
        xyz = rng.uniform(-L / 2, L / 2, size=(n, 3))
    vel = rng.normal(size=(n, 3))
    mass = rng.uniform(1.0, 16.0, size=n)
    frc = rng.normal(size=(n, 3))
    perm = rng.permutation(n).astype(np.int32)  # slot of tag t
    images = rng.integers(-1, 2, size=(n, 3)).astype(np.int32)
    H = np.diag([L, L, L]).astype(np.float64)
    origin = np.full(3, -L / 2)
    out = dict(box_H=H, origin=origin, dt=float(dt), L=L, layout=layout, natoms=n)

    if layout == "hoomd":
        pos4 = np.zeros((n, 4), dtype=dtype); pos4[:, :3] = xyz
        vm4 = np.zeros((n, 4), dtype=dtype); vm4[:, :3] = vel; vm4[:, 3] = mass
        f4 = np.zeros((n, 4), dtype=dtype); f4[:, :3] = frc
        # arrays are stored by SLOT: slot s holds the particle whose tag t satisfies perm[t] == s
        inv = np.argsort(perm)
        out["arrays"] = dict(positions=pos4[inv], vel_mass=vm4[inv], forces=f4[inv], rtags=perm,
                             images=images[inv])
    elif layout == "openmm-cuda":
        npad = (n + 31) // 32 * 32
        ids = rng.permutation(npad).astype(np.int32)  # atom index held by each slot
        pos4 = np.zeros((npad, 4), dtype=dtype); vm4 = np.zeros((npad, 4), dtype=dtype)
        f = np.zeros((3, npad), dtype=np.int64)
        xyz_p = np.zeros((npad, 3)); xyz_p[:n] = xyz
        vel_p = np.zeros((npad, 3)); vel_p[:n] = vel
        invm = np.zeros(npad); invm[:n] = 1.0 / mass
        frc_p = np.zeros((npad, 3)); frc_p[:n] = frc
        pos4[:, :3] = xyz_p[ids]; vm4[:, :3] = vel_p[ids]; vm4[:, 3] = invm[ids]
        f[:] = (frc_p[ids].T * 2.0**32).astype(np.int64)
        out["arrays"] = dict(positions=pos4, vel_mass=vm4, forces=f, ids=ids)
        out["natoms_padded"] = npad
    elif layout == "openmm-cpu":
        out["arrays"] = dict(positions=xyz.astype(dtype), velocities=vel.astype(dtype),
                             inv_masses=(1.0 / mass).reshape(n, 1).astype(dtype),
                             forces=frc.astype(dtype), ids=np.arange(n, dtype=np.int32))
    elif layout == "lammps":
        inv = np.argsort(perm)
        types = rng.integers(1, ntypes + 1, size=n).astype(np.int32)
        masses = np.zeros(ntypes + 1); masses[1:] = rng.uniform(1.0, 16.0, size=ntypes)
        tags_map = np.empty(n + 1, dtype=np.int32); tags_map[0] = -1; tags_map[1:] = perm
        im = images[inv]
        packed = ((im[:, 0] + LAMMPS_IMG_MAX) | ((im[:, 1] + LAMMPS_IMG_MAX) << 10)
                  | ((im[:, 2] + LAMMPS_IMG_MAX) << 20)).astype(np.int32)
        out["arrays"] = dict(positions=xyz[inv].astype(np.float64), velocities=vel[inv].astype(np.float64),
                             forces=frc[inv].astype(np.float64), tags_map=tags_map, images=packed,
                             types=types, masses=masses)
    else:
        raise ValueError(f"Invalid backend {layout}")
    return out


def make_trajectory(snap, frames=8, step=0.02, seed=1):
    """(frames, *positions.shape): x_t = x_0 + step * cumulative N(0,1) displacements."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = snap["arrays"]["positions"]
    traj = np.repeat(pos[None], frames, axis=0).copy()
    disp = np.cumsum(step * rng.normal(size=(frames, pos.shape[0], 3)), axis=0)
    traj[:, :, :3] += disp.astype(pos.dtype)
    return traj


def adp_snapshot(xyz_angstrom, layout="openmm-cpu", dtype=np.float64, seed=20240 + 1000, dt=0.002):
    """Alanine dipeptide in vacuum (22 atoms, nm), the reference's CPU-runnable configuration
    (examples/openmm/metad/alanine-dipeptide.py:97-114)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xyz = np.asarray(xyz_angstrom, dtype=np.float64) / 10.0
    n = xyz.shape[0]
    mass = rng.uniform(1.0, 16.0, size=n)
    vel = rng.normal(scale=0.5, size=(n, 3))
    frc = rng.normal(scale=10.0, size=(n, 3))
    H = np.diag([3.0, 3.0, 3.0])
    out = dict(box_H=H, origin=np.zeros(3), dt=float(dt), L=3.0, layout=layout, natoms=n)
    if layout == "openmm-cpu":
        out["arrays"] = dict(positions=xyz.astype(dtype), velocities=vel.astype(dtype),
                             inv_masses=(1.0 / mass).reshape(n, 1).astype(dtype), forces=frc.astype(dtype),
                             ids=np.arange(n, dtype=np.int32))
    elif layout == "hoomd":
        pos4 = np.zeros((n, 4), dtype=dtype); pos4[:, :3] = xyz
        vm4 = np.zeros((n, 4), dtype=dtype); vm4[:, :3] = vel; vm4[:, 3] = mass
        f4 = np.zeros((n, 4), dtype=dtype); f4[:, :3] = frc
        out["arrays"] = dict(positions=pos4, vel_mass=vm4, forces=f4, rtags=np.arange(n, dtype=np.int32),
                             images=np.zeros((n, 3), dtype=np.int32))
    else:
        raise ValueError(layout)
    return out
