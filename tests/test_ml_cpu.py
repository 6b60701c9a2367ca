"""
Host-side fit of the ABF family's force network (pysages_b200/ml.py, torch) against the numpy
restatement of the reference's ml/ package (oracle/ml.py).  Runs on CPU: the fit is cadenced work
outside the per-step kernel, and torch executes it on whatever device the method state lives on.
"""

import numpy as np
import pytest
import torch

from oracle import ml as oml
from pysages_b200 import ml as pml

LOWER, UPPER = np.array([-1.0, 0.5]), np.array([2.0, 3.0])


def _model(topology=(6, 5), dims=2, seed=3):
    return pml.MLP(dims, dims, topology, LOWER[:dims], UPPER[:dims], seed=seed)


def test_initial_parameters_match_oracle_layout():
    m = _model()
    assert m.widths == oml.layer_sizes(2, (6, 5), 2)
    assert m.parameters.size == oml.number_of_params(m.widths)
    np.testing.assert_array_equal(m.parameters, oml.init_params(m.widths, 3))


def test_apply_and_jacobian_match_oracle():
    m = _model()
    rng = np.random.default_rng(0)
    x = rng.uniform(LOWER, UPPER, size=(17, 2))
    ps = torch.as_tensor(m.parameters)
    out, J = m.apply(ps, torch.as_tensor(x), with_jacobian=True)
    ref = oml.mlp_apply(m.parameters, m.widths, x, LOWER, UPPER)
    np.testing.assert_allclose(out.numpy(), ref, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(J.numpy(), oml.mlp_jacobian(m.parameters, m.widths, x, LOWER, UPPER), rtol=1e-12, atol=1e-15)
    # the closed-form Jacobian against central differences of the oracle's forward pass
    h = 1e-6
    for j in rng.choice(m.parameters.size, size=8, replace=False):
        dp = np.zeros_like(m.parameters)
        dp[j] = h
        fd = (oml.mlp_apply(m.parameters + dp, m.widths, x, LOWER, UPPER)
              - oml.mlp_apply(m.parameters - dp, m.widths, x, LOWER, UPPER)).ravel() / (2 * h)
        np.testing.assert_allclose(J.numpy()[:, j], fd, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("dims", [1, 2, 3])
def test_blackman_kernel(dims):
    k = pml.blackman_kernel(dims, 7)
    assert k.shape == (5,) * dims
    np.testing.assert_allclose(k.sum(), 1.0, rtol=1e-14)
    np.testing.assert_array_equal(k, oml.blackman_kernel(dims, 7))
    np.testing.assert_allclose(k, np.flip(k), rtol=1e-15)
    centre = (2,) * dims
    assert k[centre] == k.max()


@pytest.mark.parametrize("boundary", ["edge", "wrap"])
@pytest.mark.parametrize("shape", [(11,), (7, 9), (4, 5, 6)])
def test_convolve_matches_scipy(shape, boundary):
    rng = np.random.default_rng(len(shape))
    data = rng.normal(size=shape)
    kernel = rng.uniform(size=(5,) * len(shape))  # asymmetric on purpose: convolution flips it
    ours = pml.convolve(torch.as_tensor(data), kernel, boundary).numpy()
    np.testing.assert_allclose(ours, oml.convolve(data, kernel, boundary), rtol=1e-12, atol=1e-14)
    assert ours.shape == shape


def test_normalize_and_smooth():
    rng = np.random.default_rng(5)
    y = rng.normal(size=(8, 6, 2)) * [3.0, 0.2] + [1.0, -4.0]
    yn, mean, std = pml.normalize(torch.as_tensor(y), axes=(0, 1))
    rn, rmean, rstd = oml.normalize(y, axes=(0, 1))
    np.testing.assert_allclose(yn.numpy(), rn, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(mean.numpy(), rmean, rtol=1e-13)
    np.testing.assert_allclose(std.numpy(), rstd, rtol=1e-13)
    k = pml.blackman_kernel(2, 7)
    np.testing.assert_allclose(pml.smooth(yn, k, "edge").numpy(), oml.smooth(rn, k, "edge"), rtol=1e-12, atol=1e-14)


def _target(x):
    u = oml.scale(x, LOWER, UPPER)
    return np.stack([np.sin(2.0 * u[:, 0]) * u[:, 1], 0.5 * u[:, 0] ** 2 - u[:, 1]], axis=-1)


def test_levenberg_marquardt_follows_oracle_iterates():
    m = _model(topology=(5,), seed=11)
    g = np.stack(np.meshgrid(np.linspace(-0.9, 1.9, 9), np.linspace(0.6, 2.9, 8), indexing="ij"), -1).reshape(-1, 2)
    y = _target(g).reshape(9, 8, 2)
    opt = pml.LevenbergMarquardt(reg=pml.L2Regularization(1e-6), max_iters=40)
    fit = pml.build_fitting_function(m, opt)
    hist_p, hist_o = [], []
    p = fit(torch.as_tensor(m.parameters), torch.as_tensor(g), torch.as_tensor(y), history=hist_p).numpy()
    po = oml.lm_fit(m.parameters, m.widths, g, y, LOWER, UPPER, reg=1e-6, max_iters=40, history=hist_o)
    assert len(hist_p) == len(hist_o) > 10
    # the first iterations agree to rounding; later ones drift apart only through the float32 residuals
    for (pp, cp, mp), (qo, co, mo) in list(zip(hist_p, hist_o))[:5]:
        np.testing.assert_allclose(pp.numpy(), qo, rtol=1e-7, atol=1e-9)
        assert mp == mo
        np.testing.assert_allclose(cp, co, rtol=1e-6)
    np.testing.assert_allclose(hist_p[-1][1], hist_o[-1][1], rtol=1e-3)
    # and the fit is a fit: cost far below the initial one, prediction close to the target
    e0 = oml.mlp_apply(m.parameters, m.widths, g, LOWER, UPPER) - y.reshape(-1, 2)
    e1 = oml.mlp_apply(p, m.widths, g, LOWER, UPPER) - y.reshape(-1, 2)
    assert np.sum(e1 ** 2) < 0.05 * np.sum(e0 ** 2)
    assert np.all(np.isfinite(po))


def test_levenberg_marquardt_rejects_bad_steps():
    """A step that raises the cost is undone and mu grows by mu_c (optimizers.py:223-227)."""
    m = _model(topology=(3,), seed=2)
    g = np.random.default_rng(1).uniform(LOWER, UPPER, size=(30, 2))
    y = _target(g)
    opt = pml.LevenbergMarquardt(reg=pml.L2Regularization(1e-6), max_iters=60)
    hist = []
    pml.build_fitting_function(m, opt)(torch.as_tensor(m.parameters), torch.as_tensor(g), torch.as_tensor(y), history=hist)
    costs = [c for _, c, _ in hist]
    assert all(b <= a * (1 + 1e-12) for a, b in zip(costs, costs[1:]))  # monotone: rejected steps keep the cost
    mus = [mu for _, _, mu in hist]
    assert min(mus) >= 1e-8 * (1 - 1e-6) and max(mus) <= 1e8


def test_sobolev_levenberg_marquardt_follows_oracle():
    """cff.py:164: LevenbergMarquardt(loss=Sobolev1SSE()) -- values AND input-gradients are fitted.  The product
    differentiates the input-gradients forward-over-reverse, the oracle reverse-over-reverse."""
    m = pml.MLP(2, 1, (4,), LOWER, UPPER, seed=5)
    g = np.stack(np.meshgrid(np.linspace(-0.9, 1.9, 6), np.linspace(0.6, 2.9, 5), indexing="ij"), -1).reshape(-1, 2)
    u = oml.scale(g, LOWER, UPPER)
    y = (np.sin(2 * u[:, 0]) * u[:, 1]).reshape(6, 5, 1)
    dy = np.stack([2 * np.cos(2 * u[:, 0]) * u[:, 1] * 2 / 3.0, np.sin(2 * u[:, 0]) * 2 / 2.5], -1).reshape(6, 5, 2)
    opt = pml.LevenbergMarquardt(loss=pml.Sobolev1SSE(), reg=pml.L2Regularization(1e-4), max_iters=12)
    hp, ho = [], []
    pml.build_fitting_function(m, opt)(torch.as_tensor(m.parameters), torch.as_tensor(g),
                                       (torch.as_tensor(y), torch.as_tensor(dy)), history=hp)
    oml.lm_fit_sobolev(m.parameters, m.widths, g, y, dy, LOWER, UPPER, reg=1e-4, max_iters=12, history=ho)
    assert len(hp) == len(ho) >= 8
    for (pp, cp, mp), (qo, co, mo) in list(zip(hp, ho))[:4]:
        np.testing.assert_allclose(pp.numpy(), qo, rtol=1e-6, atol=1e-8)
        assert mp == mo
    assert hp[-1][1] < 0.1 * hp[0][1]  # and it fits: the cost drops by more than 10x


def test_gradient_learning_recovers_a_known_free_energy():
    """analysis.py:44-175 (ABF / FUNN `analyze`): a network whose input-gradient is fitted to binned mean forces.
    Forces = gradient of a known A(x) on a periodic 1-D grid -> fes_fn reproduces max A - A."""
    import pysages_b200 as ps
    from pysages_b200 import analysis
    from pysages_b200.grids import compute_mesh

    grid = ps.Grid[ps.Periodic]((-np.pi,), (np.pi,), (32,))
    x = compute_mesh(grid).reshape(-1)
    A = 2.0 * np.cos(x) + 0.5 * np.sin(2 * x)
    dA = -2.0 * np.sin(x) + 1.0 * np.cos(2 * x)
    hist = np.full(32, 7, dtype=np.int64)
    Fsum = (dA * 7).reshape(32, 1)
    fes_fn, nn = analysis.gradient_learning_fes(grid, hist, Fsum, (8, 8), device="cpu", pre_iters=60, iters=300)
    got = fes_fn(x.reshape(-1, 1)).reshape(-1)
    want = A.max() - A
    assert np.abs(got - want).max() < 0.05 * (want.max() - want.min())
    t = analysis.grid_transposer(ps.Grid[ps.Regular]((0.0, 0.0), (1.0, 1.0), (3, 5)))
    a = np.arange(30).reshape(3, 5, 2)
    np.testing.assert_array_equal(t(a), a.transpose(1, 0, 2))
    np.testing.assert_array_equal(t(np.arange(15)), np.arange(15).reshape(3, 5).T)
