"""
psb_lm_fit (csrc/psb_refit.cu): the Levenberg-Marquardt refit of FUNN / ANN / CFF's force network in hand-written
kernels (ml/optimizers.py:188-232, ml/objectives.py: float32 residuals, L2 regularisation) against the torch
implementation of the same loop (pysages_b200/ml.py, itself gated against oracle/ml.py on the CPU).
"""
import numpy as np
import pytest
import torch

from pysages_b200 import ml

pytestmark = pytest.mark.gpu


def problem(dims, outdim, topology, n, seed):
    rng = np.random.default_rng(seed)
    lower, upper = -np.ones(dims) * 2.0, np.ones(dims) * 3.0
    x = rng.uniform(lower, upper, size=(n, dims))
    if outdim == dims:
        y = np.stack([np.sin(x[:, d]) * (1 + 0.3 * np.cos(x[:, (d + 1) % dims])) for d in range(dims)], -1)
    else:
        y = np.stack([np.cos(x).prod(-1) + 0.1 * j for j in range(outdim)], -1)
    model = ml.MLP(dims, outdim, topology, lower, upper, seed=seed)
    dev = torch.device("cuda")
    return model, torch.as_tensor(x, device=dev), torch.as_tensor(y, device=dev), torch.as_tensor(model.parameters, device=dev)


CASES = [(2, 2, (14, 14), 4096), (1, 1, (8,), 64), (2, 1, (8, 8), 1024), (3, 3, (15, 23), 1000), (1, 1, (33, 5, 40), 257)]


def conditioning(model, p, x, shift):
    """cond(J^T J + shift I) at p: the amplification of rounding differences by one LM solve."""
    _, J = model.apply(p, x, True)
    ev = torch.linalg.eigvalsh(J.T @ J)
    return float((ev[-1] + shift) / (ev[0].clamp(min=0) + shift))


@pytest.mark.parametrize("dims, outdim, topology, n", CASES)
def test_one_native_lm_step_from_points_along_a_fit(dims, outdim, topology, n):
    """Single iterations started from the host loop's own iterates (parameters and damping of iterations 0, 2, 5, 9):
    the step itself is compared, at a tolerance scaled by the conditioning of the damped normal equations."""
    model, x, y, p0 = problem(dims, outdim, topology, n, seed=3)
    reg = ml.L2Regularization(1e-6)
    hist = []
    ml.build_fitting_function(model, ml.LevenbergMarquardt(reg=reg, max_iters=10), force_host=True)(p0, x, y, history=hist)
    starts = [(p0, ml.LevenbergMarquardt().params.mu_0)] + [(hist[i][0], hist[i][2]) for i in (1, 4, 8) if i < len(hist)]
    for ps, mu in starts:
        opt = ml.LevenbergMarquardt(params=ml.LevenbergMarquardtParams(mu_0=float(mu)), reg=reg, max_iters=1)
        host = ml.build_fitting_function(model, opt, force_host=True)
        native = ml.build_fitting_function(model, opt)
        assert native.native and not host.native
        h1 = []
        ph, pn = host(ps, x, y, history=h1), native(ps, x, y)
        it, cost, mu1, loops = native.workspace.result()
        cond = conditioning(model, ps, x, float(np.float32(1e-6) + np.float32(mu)))
        step = float((ph - ps).abs().max())
        err = float((pn - ph).abs().max())
        print(f"one LM step: |native - host| = {err:.2e}, |step| = {step:.2e}, cond = {cond:.2e}")
        assert err <= 64 * 2.3e-16 * cond * max(step, float(ps.abs().max()) * 1e-3) + 1e-14
        accepted = not torch.equal(h1[-1][0], ps)
        assert it == int(accepted) and loops == 1 and mu1 == np.float32(h1[-1][2])
        if accepted:
            assert abs(cost - h1[-1][1]) <= max(1e-10, 1e-15 * cond) * abs(h1[-1][1])


@pytest.mark.parametrize("dims, outdim, topology, n", CASES)
@pytest.mark.parametrize("iters", [1, 3, 8])
def test_native_lm_follows_the_host_loop(dims, outdim, topology, n, iters):
    """Whole trajectories (accepted and rejected steps, damping schedule).  Rounding differences of the two solvers
    are amplified by every ill-conditioned solve and by the float32 rounding of the residuals, so beyond the first
    iteration the gate is on the cost; the parameter difference is printed."""
    model, x, y, p0 = problem(dims, outdim, topology, n, seed=3)
    opt = ml.LevenbergMarquardt(reg=ml.L2Regularization(1e-6), max_iters=iters)
    host = ml.build_fitting_function(model, opt, force_host=True)
    native = ml.build_fitting_function(model, opt)
    hist = []
    ph = host(p0, x, y, history=hist)
    pn = native(p0, x, y)
    it, cost, mu, loops = native.workspace.result()
    scale = float(ph.abs().max())
    err = float((pn - ph).abs().max())
    print(f"native vs host LM: max |dp| = {err:.3e} (scale {scale:.3g}), iters {it}, loops {loops} (host {len(hist)})")
    assert abs(cost - hist[-1][1]) <= 2e-3 * abs(hist[-1][1])
    assert 1 <= it <= iters and loops >= it
    if iters == 1:
        assert err <= 1e-9 * scale and loops == len(hist) and mu == np.float32(hist[-1][2])
    assert np.array_equal(p0.cpu().numpy(), model.parameters)  # the input parameters are not modified


def test_native_lm_converges_and_stops_without_host_synchronisation():
    model, x, y, p0 = problem(2, 2, (14, 14), 64 * 64, seed=5)
    opt = ml.LevenbergMarquardt(reg=ml.L2Regularization(1e-6), max_iters=60)
    fit = ml.build_fitting_function(model, opt)
    p = fit(p0, x, y)
    it, cost, mu, loops = fit.workspace.result()
    e0 = float(((model.apply(p0, x) - y) ** 2).sum() / 2)
    assert 1 <= it <= 60 and cost < 0.05 * e0 and loops >= it
    pred = model.apply(p, x)
    assert abs(float(((pred - y).float().double() ** 2).sum() / 2 + 1e-6 * (p @ p) / 2) - cost) <= 1e-6 * cost
    p2 = fit(p0, x, y)  # same problem again: the captured iteration graph is reused, identical result
    assert torch.equal(p, p2)


def test_gradient_and_sobolev_losses_and_sirens_stay_on_the_host_path():
    model, x, y, p0 = problem(1, 1, (8,), 32, seed=1)
    assert not ml.build_fitting_function(model, ml.LevenbergMarquardt(loss=ml.GradientsSSE())).native
    assert not ml.build_fitting_function(model, ml.LevenbergMarquardt(loss=ml.Sobolev1SSE())).native
    siren = ml.Siren(1, 1, (8,), [-1.0], [1.0])
    assert not ml.build_fitting_function(siren, ml.LevenbergMarquardt()).native


@pytest.mark.parametrize("shape, periodic", [((32,), True), ((24, 20), True), ((16,), False), ((12, 10), False)])
def test_series_fit_kernels_against_the_torch_form(shape, periodic):
    """psb_series_fit (SpectralABF refit: mean force, normalisation, projection on the pseudo-inverse) against
    approxfun.build_fitter's torch implementation of approxfun/core.py:232-258 on the same state."""
    import pysages_b200 as ps
    from pysages_b200 import approxfun
    from pysages_b200.grids import Chebyshev, Grid, convert

    d = len(shape)
    if periodic:
        grid = ps.Grid(lower=(-np.pi,) * d, upper=(np.pi,) * d, shape=shape, periodic=True)
    else:
        grid = convert(ps.Grid(lower=(0.0,) * d, upper=(2.0,) * d, shape=shape), Grid[Chebyshev])
    model = approxfun.SpectralGradientFit(grid)
    rng = np.random.default_rng(7)
    dev = torch.device("cuda")
    hist = torch.as_tensor(rng.integers(0, 40, size=shape).astype(np.int32), device=dev)
    Fsum = torch.as_tensor(rng.normal(size=(*shape, d)) * 50.0 + 3.0, device=dev)
    native = approxfun.build_fitter(model, dev)
    host = approxfun.build_fitter(model, dev, force_host=True)
    assert native.native and not host.native
    a, b = native.from_state(hist, Fsum), host.from_state(hist, Fsum)
    torch.cuda.synchronize()
    assert a.shape == b.shape == (1 + model.pinv.shape[0],)
    assert abs(float(a[0] - b[0])) <= 1e-13 * abs(float(b[0]))
    assert float((a[1:] - b[1:]).abs().max()) <= 1e-12 * float(b[1:].abs().max())
    zero = native.from_state(torch.zeros_like(hist), torch.zeros_like(Fsum))  # spectral_abf.py:178: scale 1, zero series
    assert float(zero[0]) == 1.0 and float(zero[1:].abs().max()) == 0.0


@pytest.mark.parametrize("topology, iters", [((5,), 1), ((5,), 4), ((8, 8), 2)])
def test_native_lm_against_the_oracle_restatement(topology, iters):
    """psb_lm_fit against oracle/ml.py:lm_fit (numpy / scipy restatement of ml/optimizers.py:188-232 with the float32
    residual and damping conventions of the reference): parameters, cost and damping after the first iterations."""
    from oracle import ml as oml

    lower, upper = np.array([-1.0, 0.5]), np.array([2.0, 3.0])
    g = np.stack(np.meshgrid(np.linspace(-0.9, 1.9, 9), np.linspace(0.6, 2.9, 8), indexing="ij"), -1).reshape(-1, 2)
    y = np.stack([np.sin(g[:, 0]) * g[:, 1], np.cos(g[:, 1]) + 0.3 * g[:, 0]], -1)
    model = ml.MLP(2, 2, topology, lower, upper, seed=11)
    hist = []
    po = oml.lm_fit(model.parameters, model.widths, g, y, lower, upper, reg=1e-6, max_iters=iters, history=hist)
    fit = ml.build_fitting_function(model, ml.LevenbergMarquardt(reg=ml.L2Regularization(1e-6), max_iters=iters))
    dev = torch.device("cuda")
    pn = fit(torch.as_tensor(model.parameters, device=dev), torch.as_tensor(g, device=dev), torch.as_tensor(y, device=dev))
    it, cost, mu, loops = fit.workspace.result()
    assert fit.native and loops == len(hist) and mu == np.float32(hist[-1][2])
    np.testing.assert_allclose(pn.cpu().numpy(), po, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(cost, hist[-1][1], rtol=1e-6)
