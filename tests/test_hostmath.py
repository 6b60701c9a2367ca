"""
CPU unit tests of the scalar fp64 building blocks the CUDA kernels are made of
(pysages_b200/csrc/psb_math.cuh compiled for the host by tests/hostmath), checked against the
oracle's autodiff and the reference's known answers.  No GPU needed.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import colvars, grids

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(HERE, "hostmath", "hostmath.cpp")
    out = os.path.join(HERE, "hostmath", "_hostmath.so")
    hdr = os.path.join(HERE, "..", "pysages_b200", "csrc", "psb_math.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-ffp-contract=off", src, "-o", out])
    lib = ctypes.CDLL(out)
    dp = ctypes.POINTER(ctypes.c_double)
    for name in ("hm_distance", "hm_angle", "hm_dihedral"):
        getattr(lib, name).restype = ctypes.c_double
        getattr(lib, name).argtypes = [dp, dp]
    lib.hm_kabsch.argtypes = [dp, dp]
    lib.hm_gyration.restype = ctypes.c_double
    lib.hm_gyration.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp]
    lib.hm_kabsch_eig.argtypes = [dp, dp]
    lib.hm_ring.restype = ctypes.c_double
    lib.hm_ring.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp]
    lib.hm_grid_index.restype = ctypes.c_uint
    lib.hm_grid_index.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int]
    lib.hm_mesh_coord.restype = ctypes.c_double
    lib.hm_mesh_coord.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int]
    lib.hm_chol_solve.argtypes = [ctypes.c_int, dp, dp, dp]
    lib.hm_spd_solve.argtypes = [ctypes.c_int, dp, dp, dp]
    lib.hm_grid_index_slow.restype = ctypes.c_uint
    lib.hm_grid_index_slow.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int]
    lib.hm_pinv_solve.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double]
    lib.hm_wrap.restype = ctypes.c_double
    lib.hm_wrap.argtypes = [ctypes.c_double, ctypes.c_double]
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@pytest.mark.parametrize("name, fn, npts", [("hm_distance", colvars.distance, 2), ("hm_angle", colvars.angle, 3),
                                            ("hm_dihedral", colvars.dihedral_angle, 4)])
def test_point_cv_gradients_match_autodiff(hm, name, fn, npts):
    rng = np.random.default_rng(7)
    for _ in range(200):
        p = rng.normal(size=(npts, 3)) * rng.uniform(0.5, 5.0)
        g = np.zeros((npts, 3))
        val = getattr(hm, name)(_ptr(p), _ptr(g))
        t = torch.tensor(p, requires_grad=True)
        ref = fn(*t)
        (gref,) = torch.autograd.grad(ref, t)
        assert abs(val - ref.item()) <= 1e-13 * max(1.0, abs(ref.item()))
        assert np.allclose(g, gref.numpy(), rtol=1e-10, atol=1e-13)


def test_gyration_tensor_cvs_match_autodiff(hm):
    """shape.py:85-344: value and per-atom gradient (2/n) W r_i against torch autograd through eigvalsh."""
    rng = np.random.default_rng(21)
    fns = [(0, 0, 0, lambda r: colvars.principal_moments(r)[0]), (1, 0, 0, lambda r: colvars.principal_moments(r)[1]),
           (2, 0, 0, lambda r: colvars.principal_moments(r)[2]), (3, 0, 0, colvars.asphericity),
           (4, 1, 2, lambda r: colvars.acylindricity(r, (1, 2))), (4, 2, 2, lambda r: colvars.acylindricity(r, (2, 3))),
           (4, 1, 2, lambda r: colvars.acylindricity(r, (1, 3))), (5, 0, 0, colvars.shape_anisotropy)]
    for n in (5, 40, 700):
        r = torch.tensor(rng.normal(size=(n, 3)) * np.array([1.0, 2.0, 0.5]) + rng.normal(size=3), requires_grad=True)
        G = (r.detach().numpy().T @ r.detach().numpy()) / n
        g6 = np.array([G[0, 0], G[0, 1], G[0, 2], G[1, 1], G[1, 2], G[2, 2]])
        for mode, ja, jb, fn in fns:
            W = np.zeros(9)
            val = hm.hm_gyration(mode, ja, jb, _ptr(g6), _ptr(W))
            ref = fn(r)
            if ref.requires_grad:
                (gref,) = torch.autograd.grad(ref, r)
            else:
                gref = torch.zeros_like(r)
            assert abs(val - ref.item()) <= 1e-12 * max(1.0, abs(ref.item()))
            grad = (2.0 / n) * r.detach().numpy() @ W.reshape(3, 3).T
            assert np.allclose(grad, gref.numpy(), rtol=1e-9, atol=1e-12 * np.abs(gref.numpy()).max() + 1e-15)


def test_kabsch_matches_svd_reference(hm, golden):
    rng = np.random.default_rng(3)
    pairs = [(golden["ci2_1"], golden["ci2_2"]), (golden["ci2_1"], golden["ci2_1t"]), (golden["ci2_2"], golden["ci2_12"])]
    for _ in range(20):
        P = rng.normal(size=(30, 3)); Q = rng.normal(size=(30, 3))
        pairs.append((P, Q))
    for P, Q in pairs:
        P = P - P.mean(0); Q = Q - Q.mean(0)
        C = np.ascontiguousarray(P.T @ Q)
        U = np.zeros((3, 3))
        hm.hm_kabsch(_ptr(C), _ptr(U))
        Uref = colvars.kabsch(torch.tensor(P), torch.tensor(Q)).numpy()
        assert np.allclose(U, Uref, atol=1e-12)
        assert abs(np.linalg.det(U) - 1) < 1e-12


def test_kabsch_polar_fast_path_equals_eigen_path(hm):
    """Newton polar iteration (fast path) vs Horn's quaternion eigenproblem on well- and badly-
    conditioned covariances, near-identity rotations, near-planar point sets and large scales."""
    rng = np.random.default_rng(17)
    cases = []
    for n, noise, scale in ((1000, 0.1, 1.0), (1000, 1e-6, 1.0), (50, 1.0, 1e3), (50, 0.3, 1e-4), (2000, 0.05, 30.0)):
        for _ in range(6):
            P = scale * rng.normal(size=(n, 3)); P -= P.mean(0)
            R, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            if np.linalg.det(R) < 0:
                R[:, 0] *= -1
            Q = P @ R + scale * noise * rng.normal(size=(n, 3)); Q -= Q.mean(0)
            cases.append(np.ascontiguousarray(P.T @ Q))
    for n, noise in ((500, 0.05), (40, 0.5), (3000, 1e-3)):  # improper covariances: reflection fix on the fast path
        for _ in range(8):
            P = rng.normal(size=(n, 3)) * rng.uniform(0.5, 2.0, size=3); P -= P.mean(0)
            R, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            if np.linalg.det(R) > 0:
                R[:, 0] *= -1
            Q = P @ R + noise * rng.normal(size=(n, 3)); Q -= Q.mean(0)
            C = np.ascontiguousarray(P.T @ Q)
            assert np.linalg.det(C) < 0
            cases.append(C)
    for _ in range(20):  # arbitrary full-rank matrices of either handedness
        cases.append(np.ascontiguousarray(rng.normal(size=(3, 3)) * 10 ** rng.uniform(-3, 3)))
    flat = rng.normal(size=(200, 3)) * np.array([1.0, 1.0, 1e-5]); flat -= flat.mean(0)  # nearly planar: eigen path
    cases.append(np.ascontiguousarray(flat.T @ (flat + 1e-7 * rng.normal(size=flat.shape))))
    cases.append(np.eye(3) * 5.0)
    for C in cases:
        U, V = np.zeros((3, 3)), np.zeros((3, 3))
        hm.hm_kabsch(_ptr(C), _ptr(U))
        hm.hm_kabsch_eig(_ptr(C), _ptr(V))
        sv = np.linalg.svd(C, compute_uv=False)
        if np.linalg.det(C) < 0 and (sv[1] - sv[2]) < 1e-3 * sv[0]:
            continue  # (near-)degenerate reflection direction: the optimum itself is ill-defined
        assert np.allclose(U, V, atol=5e-12 * max(1.0, sv[0] / max(sv[1] - sv[2], 1e-300) if np.linalg.det(C) < 0 else 1.0)), np.abs(U - V).max()
        assert np.allclose(U @ U.T, np.eye(3), atol=1e-13) and abs(np.linalg.det(U) - 1) < 1e-13
        W, _, Zt = np.linalg.svd(C)
        if np.linalg.det(W @ Zt) > 0 and np.linalg.cond(C) < 1e5:
            assert np.allclose(U, W @ Zt, atol=1e-11)


def test_kabsch_reflection_case(hm):
    # covariance with negative determinant: the proper rotation, not the reflection
    rng = np.random.default_rng(5)
    P = rng.normal(size=(8, 3)); P -= P.mean(0)
    Q = P * np.array([1.0, 1.0, -1.0]) + 0.01 * rng.normal(size=(8, 3)); Q -= Q.mean(0)
    C = np.ascontiguousarray(P.T @ Q)
    assert np.linalg.det(C) < 0
    U = np.zeros((3, 3))
    hm.hm_kabsch(_ptr(C), _ptr(U))
    Uref = colvars.kabsch(torch.tensor(P), torch.tensor(Q)).numpy()
    assert np.allclose(U, Uref, atol=1e-10)


def test_grid_index_bit_exact(hm, golden):
    pi = np.pi
    for tag, kind in (("regular", 1), ("periodic", 2), ("cheb", 3)):
        for x, i in zip(golden[f"grid1d_x_{tag}"], golden[f"grid1d_i_{tag}"]):
            assert hm.hm_grid_index(kind, float(x), -pi, pi, 64) == int(i)
    rng = np.random.default_rng(11)
    for kind, gk in ((1, grids.Regular), (2, grids.Periodic), (3, grids.Chebyshev)):
        for lower, upper, shape in ((-pi, pi, 50), (0.0, 1.0, 10), (-2.5, 7.25, 64), (0.0, 23.2, 64)):
            g = grids.Grid((lower,), (upper,), (shape,), kind=gk)
            get_index = grids.build_indexer(g)
            h = (upper - lower) / shape
            xs = np.concatenate([lower + h * np.arange(-2, shape + 3), rng.uniform(lower - 1, upper + 1, 2000),
                                 np.nextafter(lower + h * np.arange(shape + 1), np.inf),
                                 np.nextafter(lower + h * np.arange(shape + 1), -np.inf)])
            for x in xs:
                assert hm.hm_grid_index(kind, float(x), lower, upper, shape) == int(get_index(x)[0]), (kind, x)


def test_grid_index_fast_path_equals_literal_restatement(hm):
    """The division-only bin index (device fast path) against the fmod-based restatement of
    jnp.floor_divide on edges, near-edges, far out-of-range values and non-finite inputs."""
    rng = np.random.default_rng(5)
    for kind in (1, 2):
        for lower, upper, shape in ((-np.pi, np.pi, 50), (0.0, 1.0, 10), (0.0, 23.2, 64), (-1e-3, 1e3, 4096),
                                    (0.1, 0.7, 3), (-5.0, -1.0, 1)):
            h = (upper - lower) / shape
            edges = lower + h * np.arange(-3 * shape, 4 * shape + 1)
            xs = np.concatenate([edges, np.nextafter(edges, np.inf), np.nextafter(edges, -np.inf),
                                 lower + np.arange(-shape, 2 * shape) * h + 0.5 * h,
                                 rng.uniform(lower - 3 * (upper - lower), upper + 3 * (upper - lower), 4000),
                                 [1e300, -1e300, 1e17, -1e17, 3e9 * h, -3e9 * h, np.inf, -np.inf, np.nan, 0.0, -0.0]])
            for x in xs:
                a = hm.hm_grid_index(kind, float(x), lower, upper, shape)
                b = hm.hm_grid_index_slow(kind, float(x), lower, upper, shape)
                assert a == b, (kind, x, lower, upper, shape, a, b)


def test_mesh_and_wrap(hm):
    g = grids.Grid((-np.pi,), (np.pi,), (50,))
    mesh = grids.compute_mesh(g)[:, 0]
    got = np.array([hm.hm_mesh_coord(1, k, -np.pi, np.pi, 50) for k in range(50)])
    assert np.array_equal(mesh, got)
    gc = grids.Grid((-1.0,), (3.0,), (16,), kind=grids.Chebyshev)
    got = np.array([hm.hm_mesh_coord(3, k, -1.0, 3.0, 16) for k in range(16)])
    assert np.allclose(grids.compute_mesh(gc)[:, 0], got, rtol=1e-15, atol=1e-15)
    P = torch.tensor([2 * np.pi, np.inf], dtype=torch.float64)
    for x in (-7.0, -3.2, -0.1, 0.0, 2.9, 3.3, 6.0):
        w = colvars.wrap(torch.tensor([x, x], dtype=torch.float64), P).numpy()
        assert hm.hm_wrap(x, 2 * np.pi) == w[0] and hm.hm_wrap(x, np.inf) == w[1]


def test_small_solvers(hm):
    rng = np.random.default_rng(2)
    for k in (1, 2, 3, 4):
        J = rng.normal(size=(k, 12))
        A = J @ J.T
        b = rng.normal(size=k)
        A4 = np.zeros((4, 4)); A4[:k, :k] = A
        x = np.zeros(4)
        assert hm.hm_chol_solve(k, _ptr(A4), _ptr(b), _ptr(x)) == 1
        assert np.allclose(x[:k], np.linalg.solve(A, b), rtol=1e-11)
        x3 = np.zeros(4)
        assert hm.hm_spd_solve(k, _ptr(A4), _ptr(b), _ptr(x3)) == 1
        assert np.allclose(x3[:k], np.linalg.solve(A, b), rtol=1e-11)
        x2 = np.zeros(4)
        hm.hm_pinv_solve(k, _ptr(A4), _ptr(b), _ptr(x2), 1e-14)
        assert np.allclose(x2[:k], np.linalg.pinv(A) @ b, rtol=1e-9)
    A4 = np.zeros((4, 4)); A4[0, 0] = -1.0
    assert hm.hm_chol_solve(1, _ptr(A4), _ptr(np.ones(1)), _ptr(np.zeros(4))) == 0
    assert hm.hm_spd_solve(1, _ptr(A4), _ptr(np.ones(1)), _ptr(np.zeros(4))) == 0
    A4 = np.zeros((4, 4)); A4[:2, :2] = [[1.0, 2.0], [2.0, 1.0]]  # indefinite
    assert hm.hm_spd_solve(2, _ptr(A4), _ptr(np.ones(2)), _ptr(np.zeros(4))) == 0


@pytest.mark.parametrize("n", [5, 6, 7, 9])
def test_ring_puckering_matches_autodiff(hm, n):
    """Cremer-Pople q / phi (angles.py:158-205) and their closed-form gradients against the oracle's
    autodiff of the reference formula, for rings far from the origin too (the centring drops out)."""
    from oracle import colvars as oc

    rng = np.random.default_rng(n)
    for shift in (0.0, 150.0):
        ang = 2 * np.pi * np.arange(n) / n
        ring = np.stack([np.cos(ang), np.sin(ang), 0.35 * rng.normal(size=n)], axis=1) * 1.4
        ring += 0.1 * rng.normal(size=(n, 3)) + shift
        Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        pos = np.ascontiguousarray(ring @ Q.T)
        r = torch.tensor(pos, dtype=torch.float64, requires_grad=True)
        ref = oc.ring_puckering_coordinates(r)
        for mode in (0, 1):
            grad = np.zeros((n, 3))
            val = hm.hm_ring(mode, n, _ptr(pos), _ptr(grad))
            (gref,) = torch.autograd.grad(ref[mode], r, retain_graph=True)
            assert abs(val - ref[mode].item()) <= 1e-11 * max(1.0, abs(ref[mode].item())), (n, mode, shift)
            scale = np.abs(gref.numpy()).max()
            assert np.allclose(grad, gref.numpy(), rtol=1e-8, atol=1e-10 * scale), (n, mode, shift)
