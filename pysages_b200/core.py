"""
`run` / `Result` orchestration (pysages/methods/core.py:34-59, 160-413; umbrella_integration.py:128-203).

One simulation = one process = one GPU; replicas are independent tasks on an Executor exactly as in
the reference (`ReplicasConfiguration`).  For multi-GPU runs launch one process per GPU
(`torchrun`) and use `pysages_b200.parallel` to partition windows / share hills.
"""

from copy import deepcopy

from .backends.core import SamplingContext
from .methods import SamplingMethod, UmbrellaIntegration
from .utils import ReplicasConfiguration, SerialExecutor


class Result:
    """methods/core.py:34-59"""

    def __init__(self, method, states, callbacks, snapshots):
        self.method = method
        self.states = states
        self.callbacks = callbacks
        self.snapshots = snapshots


class ReplicaResult(Result):
    pass


def _run(method_or_result, context_generator, timesteps, context_args=None, callback=None,
         post_run_action=None, **kwargs):
    """methods/core.py:327-413 (fresh run, or restart from a ReplicaResult)."""
    timesteps = int(timesteps)
    context_args = {} if context_args is None else context_args
    restart = isinstance(method_or_result, ReplicaResult)
    method = method_or_result.method if restart else method_or_result
    if restart:
        callback = method_or_result.callbacks
    sampling_context = SamplingContext(method, context_generator, callback, context_args)
    context_args["context"] = sampling_context.context
    sampler = sampling_context.sampler
    if restart:
        sampler.restore(method_or_result.snapshots)
        native = sampler._method_update.native
        prev = method_or_result.states
        restore = getattr(sampler._method_update, "restore", None)
        if restore is not None:  # methods whose state holds more than flat device arrays (FUNN's network)
            sampler.state = restore(prev)
        else:
            for field in prev._fields:
                value = getattr(prev, field)
                if value is not None and field != "bias":
                    native.set_state(field, value.cpu().numpy() if hasattr(value, "cpu") else value)
            sampler.state = method._make_state(native)
    with sampling_context:
        sampling_context.run(timesteps, **kwargs)
        if post_run_action:
            post_run_action(**context_args)
    return ReplicaResult(method, sampler.state, callback, sampler.take_snapshot())


def run(method_or_result, context_generator, timesteps, callback=None, context_args=None,
        post_run_action=None, config=None, executor=None, executor_shutdown=True, **kwargs):
    """pysages.run: dispatches on the method type like the reference's plum table."""
    context_args = {} if context_args is None else context_args
    method = method_or_result.method if isinstance(method_or_result, Result) else method_or_result
    if isinstance(method, UmbrellaIntegration):
        return _run_umbrella(method_or_result, context_generator, timesteps, context_args,
                             post_run_action, executor or SerialExecutor(), executor_shutdown, **kwargs)
    if not isinstance(method, SamplingMethod):
        raise TypeError(f"{method!r} is not a SamplingMethod")
    config = config or ReplicasConfiguration()
    if isinstance(method_or_result, Result):  # methods/core.py:239-318
        result = method_or_result
        futures = [
            config.executor.submit(
                _run, ReplicaResult(method, result.states[i], _nth(result.callbacks, i), result.snapshots[i]),
                context_generator, timesteps, deepcopy(context_args), None, post_run_action, **kwargs)
            for i in range(len(result.states))
        ]
    else:
        futures = [
            config.executor.submit(_run, method, context_generator, timesteps, deepcopy(context_args),
                                   callback, post_run_action, **kwargs)
            for _ in range(config.copies)
        ]
    results = [f.result() for f in futures]
    callbacks = None if (callback is None and not isinstance(method_or_result, Result)) else [r.callbacks for r in results]
    return Result(method, [r.states for r in results], callbacks, [r.snapshots for r in results])


def _nth(seq, i):
    return None if seq is None else seq[i]


def _run_umbrella_lockstep(method, context_generator, timesteps, context_args, post_run_action, todo):
    """All windows of this process advance together, one launch per MD step for all their bias steps
    (parallel.WindowBatch / psb_batch_step) instead of one replica after the other."""
    from . import parallel

    contexts = []
    for n in todo:
        replica_args = deepcopy(context_args)
        replica_args["replica_num"] = n
        contexts.append(SamplingContext(method.submethods[n], context_generator, method.histograms[n], replica_args))
    parallel.run_lockstep(contexts, timesteps)
    if post_run_action:
        for c in contexts:
            post_run_action(context=c.context)
    return Result(method, [c.sampler.state for c in contexts], [c.sampler.callback for c in contexts],
                  [c.sampler.take_snapshot() for c in contexts])


def _run_umbrella(method_or_result, context_generator, timesteps, context_args, post_run_action,
                  executor, executor_shutdown, windows=None, lockstep=False, **kwargs):
    """umbrella_integration.py:128-203; `windows` restricts the run to a subset (multi-GPU partition);
    `lockstep=True` steps the windows together with one launch per MD step (engines steppable from outside)."""
    restart = isinstance(method_or_result, Result)
    method = method_or_result.method if restart else method_or_result
    todo = list(range(len(method.submethods))) if windows is None else list(windows)
    if lockstep and not restart:
        return _run_umbrella_lockstep(method, context_generator, int(timesteps), context_args, post_run_action, todo)
    futures = []
    for n in todo:
        replica_args = deepcopy(context_args)
        replica_args["replica_num"] = n
        if restart:
            r = method_or_result
            arg = ReplicaResult(method.submethods[n], r.states[n], method.histograms[n], r.snapshots[n])
            futures.append(executor.submit(_run, arg, context_generator, timesteps, replica_args, None,
                                           post_run_action, **kwargs))
        else:
            futures.append(executor.submit(_run, method.submethods[n], context_generator, timesteps,
                                           replica_args, method.histograms[n], post_run_action, **kwargs))
    results = [f.result() for f in futures]
    if executor_shutdown:
        executor.shutdown()
    return Result(method, [r.states for r in results], [r.callbacks for r in results],
                  [r.snapshots for r in results])


def _metapotential(state, periods):
    """metad.py:318-323 sum_of_gaussians, batched over the query points, on the device the hills live on."""
    import numpy as np
    import torch

    as_t = lambda v: v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
    heights, centers, sigmas = as_t(state.heights), as_t(state.centers), as_t(state.sigmas).reshape(-1)
    nh = min(int(state.idx), heights.shape[0])  # slots beyond idx hold zero heights anyway
    heights, centers = heights[:nh].to(torch.float64), centers[:nh].to(torch.float64)
    P = torch.as_tensor(np.asarray(periods, dtype=np.float64), device=heights.device)

    def metapotential(x):
        x = torch.as_tensor(np.asarray(x, dtype=np.float64), device=heights.device).reshape(-1, centers.shape[1])
        out = torch.zeros(x.shape[0], dtype=torch.float64, device=x.device)
        for lo in range(0, x.shape[0], 4096):  # (points, hills, k) blocks
            d = x[lo:lo + 4096, None, :] - centers[None, :, :]
            finite = torch.isfinite(P)
            Pf = torch.where(finite, P, torch.zeros_like(P))
            d = torch.where(finite & (d.abs() > Pf / 2), d - torch.sign(d) * Pf, d)  # colvars/utils.py:22-26
            out[lo:lo + 4096] = (heights[None, :] * torch.exp(-((d / sigmas) ** 2).sum(-1) / 2)).sum(-1)
        return out.cpu().numpy()

    return metapotential


def analyze(result, **kwargs):
    """Post-run analysis.  UmbrellaIntegration: umbrella_integration.py:220-261.  Metadynamics: metad.py:326-383
    (heights + the deposited bias potential as a callable).  ABF / FUNN: gradient learning (analysis.py:44-175: a
    network whose input-gradient is fitted to the binned mean force; `topology=` as in abf.py:305, funn.py:342).
    CFF / Sirens / ANN: the free-energy network the run trained.  SpectralABF: the fitted series as a value.
    Grid-mapped arrays come back transposed along the grid axes, like the reference's."""
    import numpy as np

    from . import analysis as _an
    from . import methods as _m
    from .colvars import get_periods
    from .grids import compute_mesh
    from .serialization import numpyfy

    method = result.method
    if isinstance(method, _m.Metadynamics):
        P = get_periods(method.cvs)
        states = list(result.states)
        heights = [numpyfy(s).heights for s in states]
        pots = [_metapotential(s, P) for s in states]
        if len(states) == 1:
            return dict(heights=heights[0], metapotential=pots[0])
        return dict(heights=heights, metapotential=pots)
    if isinstance(method, _m.ANN):  # ann.py:261-322: histogram, free energy on the grid, the network and fes_fn
        import torch

        out = dict(histogram=[], free_energy=[], nn=[], fes_fn=[], mesh=compute_mesh(method.grid))
        model = method.model

        def build_fes_fn(nn):
            def fes_fn(x):
                ps = torch.as_tensor(np.asarray(nn.params), dtype=torch.float64)
                A = float(np.asarray(nn.std)) * model.apply(ps, torch.as_tensor(np.asarray(x, dtype=np.float64))) + float(
                    np.asarray(nn.mean))
                return (A.max() - A).numpy()

            return fes_fn

        for s in result.states:
            h = numpyfy(s)
            out["histogram"].append(np.asarray(h.hist))
            out["free_energy"].append(np.asarray(h.phi).max() - np.asarray(h.phi))
            out["nn"].append(h.nn)
            out["fes_fn"].append(build_fes_fn(h.nn))
        if len(result.states) == 1:
            for key in ("histogram", "free_energy", "nn", "fes_fn"):
                out[key] = out[key][0]
        return out
    if isinstance(method, (_m.ABF, _m.FUNN, _m.SpectralABF, _m.CFF, _m.Sirens)):
        grid = method.grid
        mesh = compute_mesh(grid)
        transpose = _an.grid_transposer(grid)
        d = mesh.shape[-1]
        hists, forces, energies, fns, extra = [], [], [], [], {}
        spectral_eval = _an.series_value_evaluator(method.model) if isinstance(method, _m.SpectralABF) else None
        for s in result.states:
            h = numpyfy(s)
            if isinstance(method, _m.SpectralABF):  # spectral_abf.py:304-346
                fes_fn = (lambda fun: (lambda x: (lambda A: A.max() - A)(spectral_eval(fun, x))))(h.fun)
                extra.setdefault("fun", []).append(h.fun)
            elif isinstance(method, (_m.CFF, _m.Sirens)):  # cff.py:392-424, sirens.py:467-498: the run's own network
                fes_fn = _an.network_fes_fn(method.model, h.nn)
                extra.setdefault("nn", []).append(h.nn)
                if isinstance(method, _m.CFF):
                    extra.setdefault("fnn", []).append(h.fnn)
            else:  # ABF, FUNN: gradient learning on the binned mean force
                topology = kwargs.get("topology", getattr(method, "topology", (8, 8)))
                dev = s.hist.device if hasattr(s.hist, "device") else None
                fes_fn, _ = _an.gradient_learning_fes(grid, s.hist, s.Fsum, topology, device=dev,
                                                      pre_iters=kwargs.get("pre_iters", 250), iters=kwargs.get("iters", 1000))
                if isinstance(method, _m.FUNN):
                    extra.setdefault("nn", []).append(h.nn)
            hists.append(transpose(h.hist))
            forces.append(transpose(_an.average_forces(h.hist, h.Fsum)))
            if isinstance(method, _m.CFF):
                energies.append(transpose(np.asarray(h.fe).max() - np.asarray(h.fe)))
            else:
                energies.append(transpose(fes_fn(mesh)))
            fns.append(fes_fn)
        out = dict(histogram=_an.first_or_all(hists), mean_force=_an.first_or_all(forces),
                   free_energy=_an.first_or_all(energies), fes_fn=_an.first_or_all(fns),
                   mesh=transpose(mesh).reshape(-1, d).squeeze())
        for key, vals in extra.items():
            out[key] = _an.first_or_all(vals)
        return out
    if not isinstance(method, UmbrellaIntegration):
        raise NotImplementedError(f"analyze() is not implemented for {type(method).__name__} (DESIGN.md scope)")
    ksprings = [s.kspring for s in method.submethods]
    centers = [s.center for s in method.submethods]
    hist_means = [cb.get_means() for cb in result.callbacks]
    mean_forces, free_energy = [], [0.0]
    for i, (k, c, mean) in enumerate(zip(ksprings, centers, hist_means)):
        mean_forces.append(-(k @ (mean - c)))  # eq. 13 of doi:10.1063/1.3175798
        if i > 0:
            free_energy.append(free_energy[i - 1] + mean_forces[i - 1].T @ (centers[i] - centers[i - 1]))
    return dict(ksprings=np.asarray(ksprings), centers=np.asarray(centers), histograms=result.callbacks,
                histogram_means=np.asarray(hist_means), mean_forces=np.asarray(mean_forces),
                free_energy=np.asarray(free_energy, dtype=np.float64))
