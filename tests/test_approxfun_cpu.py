"""
SpectralABF's series machinery on the CPU: the product's set-up / fit code (pysages_b200/approxfun.py)
against the numpy restatement of pysages/approxfun/core.py (oracle/approxfun.py), and the restatement
against functions whose gradients are known in closed form.
"""

import numpy as np
import pytest
import torch

import pysages_b200 as ps
from oracle import approxfun as oaf
from oracle import grids as og
from pysages_b200 import approxfun as paf

CASES = [
    ((-np.pi, -np.pi), (np.pi, np.pi), (16, 12), True),
    ((-1.0,), (1.0,), (20,), True),
    ((0.0,), (4.0,), (32,), False),
    ((0.0, -1.0), (4.0, 2.0), (10, 8), False),
]


def grids_of(lo, hi, shape, periodic):
    g = ps.Grid[ps.Periodic if periodic else ps.Chebyshev](lo, hi, shape)
    o = og.Grid(lo, hi, shape, kind=og.Periodic if periodic else og.Chebyshev)
    return g, o


def test_exponents_follow_the_reference_rule():
    # approxfun/core.py:73-96: periodic r = sum(shape) / 4, otherwise sum(shape^2) / 16; n = floor(sqrt(r)) + 1
    g, _ = grids_of((-np.pi, -np.pi), (np.pi, np.pi), (64, 64), True)
    ns = paf.collect_exponents(g)
    assert ns.shape == (50, 2) and ns[:5].tolist() == [[0, i] for i in range(1, 6)]
    assert all(a * a + b * b <= 32 for a, b in ns[10:]) and all(a >= 1 and b != 0 for a, b in ns[10:])
    g, _ = grids_of((0.0,), (1.0,), (32,), False)
    assert paf.collect_exponents(g).ravel().tolist() == list(range(1, 9))  # r = 64 -> n = 9
    with pytest.raises(ValueError, match="Only 1D and 2D grids"):
        paf.collect_exponents(ps.Grid((0.0,) * 3, (1.0,) * 3, (4, 4, 4)))


@pytest.mark.parametrize("lo, hi, shape, periodic", CASES)
def test_setup_and_fit_match_oracle(lo, hi, shape, periodic):
    g, o = grids_of(lo, hi, shape, periodic)
    model = paf.SpectralGradientFit(g)
    ns = oaf.collect_exponents(o)
    assert np.array_equal(model.exponents, ns)
    np.testing.assert_allclose(model.mesh, oaf.unit_mesh(o), rtol=0, atol=0)
    P = oaf.gradient_pinv(o, ns)
    np.testing.assert_allclose(model.pinv, P, rtol=1e-10, atol=1e-12 * np.abs(P).max())
    dy = np.random.default_rng(0).normal(size=(*shape, len(shape))) + 0.3
    v = paf.build_fitter(model, "cpu")(torch.as_tensor(dy)).numpy()
    f = oaf.fit_gradient(o, P, dy)
    np.testing.assert_allclose(v[0], f.scale, rtol=1e-13)
    np.testing.assert_allclose(v[1:], f.coefficients, rtol=1e-9, atol=1e-12 * np.abs(f.coefficients).max())
    zero = paf.build_fitter(model, "cpu")(torch.zeros((*shape, len(shape)), dtype=torch.float64)).numpy()
    assert zero[0] == 1.0 and not np.any(zero[1:])  # std == 0 -> scale 1 (core.py:245-246), the fit of initialize()


def test_fit_recovers_known_gradients():
    # periodic: A = cos x + 0.5 sin 2y -> (-sin x, cos 2y); the Fourier basis contains both exactly
    _, o = grids_of((-np.pi, -np.pi), (np.pi, np.pi), (16, 16), True)
    ns, P = oaf.collect_exponents(o), None
    P = oaf.gradient_pinv(o, ns)
    mesh = (oaf.unit_mesh(o) + 1) / 2 * o.size + o.lower
    dy = np.stack([-np.sin(mesh[:, 0]), np.cos(2 * mesh[:, 1])], -1).reshape(16, 16, 2)
    fun = oaf.fit_gradient(o, P, dy)
    x = np.random.default_rng(1).uniform(-np.pi, np.pi, size=(9, 2))
    np.testing.assert_allclose(oaf.get_gradient(o, ns, fun, x), np.stack([-np.sin(x[:, 0]), np.cos(2 * x[:, 1])], -1),
                               atol=1e-12)
    # non-periodic: dA/dx = 3 x^2 - 2 is inside the monomial basis
    _, o = grids_of((0.0,), (4.0,), (32,), False)
    ns = oaf.collect_exponents(o)
    P = oaf.gradient_pinv(o, ns)
    m = (oaf.unit_mesh(o) + 1) / 2 * o.size + o.lower
    fun = oaf.fit_gradient(o, P, (3 * m[:, 0] ** 2 - 2).reshape(32, 1))
    xx = np.array([[0.2], [1.0], [3.3]])
    np.testing.assert_allclose(oaf.get_gradient(o, ns, fun, xx).ravel(), 3 * xx.ravel() ** 2 - 2, rtol=1e-9)


@pytest.mark.parametrize("kind", ["periodic", "chebyshev"])
def test_series_value_is_the_antiderivative_of_the_kernel_series(kind):
    """spectral_abf.py:304-346 `analyze`: the fitted series evaluated as a VALUE (approxfun/core.py:149-166, 282-308)
    must be the function whose gradient the step kernel evaluates: fit the sampled gradient of a known f, then
    compare values up to the constant the fit cannot see."""
    import pysages_b200 as ps
    from pysages_b200 import analysis, approxfun
    import torch

    if kind == "periodic":
        grid = ps.Grid[ps.Periodic]((-np.pi, -np.pi), (np.pi, np.pi), (16, 16))
        f = lambda x: np.cos(x[:, 0]) * np.sin(x[:, 1]) + 0.3 * np.cos(2 * x[:, 0])
        df = lambda x: np.stack([-np.sin(x[:, 0]) * np.sin(x[:, 1]) - 0.6 * np.sin(2 * x[:, 0]), np.cos(x[:, 0]) * np.cos(x[:, 1])], -1)
    else:
        grid = ps.Grid[ps.Chebyshev]((0.0, -1.0), (2.0, 3.0), (16, 16))
        f = lambda x: 0.5 * x[:, 0] ** 2 * x[:, 1] - x[:, 1] ** 3 / 7
        df = lambda x: np.stack([x[:, 0] * x[:, 1], 0.5 * x[:, 0] ** 2 - 3 * x[:, 1] ** 2 / 7], -1)
    model = approxfun.SpectralGradientFit(grid)
    pts = np.asarray(grid.lower) + (model.mesh + 1) / 2 * np.asarray(grid.size)  # the fit's own sample points
    fit = approxfun.build_fitter(model, "cpu")
    v = fit(torch.as_tensor(df(pts).reshape(16, 16, 2))).numpy()

    class Fun:
        scale, coefficients, c0 = v[0], v[1:], 0.0

    got = analysis.series_value_evaluator(model)(Fun, pts).reshape(-1)
    want = f(pts)
    resid = (got - got.mean()) - (want - want.mean())
    assert np.abs(resid).max() < 1e-6 * np.abs(want - want.mean()).max()
