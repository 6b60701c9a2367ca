"""
SamplingContext: finds out which engine a user-built simulation object belongs to and binds the
sampler to it (the role of pysages/backends/core.py:10-73).
"""

import importlib

# module-name prefix of the simulation object -> adapter module of this package
ENGINES = (
    (("hoomd",), "hoomd"),
    (("openmm", "simtk.openmm"), "openmm"),
    (("lammps",), "lammps"),
    (("pysages_b200.backends.mock",), "mock"),
)


def supported_backends():
    return tuple(name for _, name in ENGINES)


supported = supported_backends()


def backend_of(simulation):
    module = type(simulation).__module__
    for prefixes, name in ENGINES:
        if module.startswith(prefixes):
            return name
    raise ValueError(f"Invalid backend {module}: supported options are ({', '.join(supported_backends())})")


class SamplingContext:
    """Owns the engine's simulation object, the method and the bound sampler; `run` is set by the adapter."""

    def __init__(self, method, context_generator, callback=None, context_args=None, **kwargs):
        self.context = context_generator(**({} if context_args is None else context_args))
        self._backend_name = backend_of(self.context)
        self.method = method
        self.run = None
        adapter = importlib.import_module(f"{__package__}.{self._backend_name}")
        self.sampler = adapter.bind(self, callback, **kwargs)
        if self.run is None:
            raise RuntimeError(f"the {self._backend_name} adapter did not provide a run function")

    @property
    def backend_name(self):
        return self._backend_name

    def get_backend_name(self):
        return self._backend_name

    # `with sampling_context:` enters the engine's own context where it has one (hoomd 2.x SimulationContext:
    # the global `hoomd.run` steps whichever context is entered, so replicas / umbrella windows would otherwise
    # step the wrong simulation; core.py:59-73)
    def __enter__(self):
        enter = getattr(self.context, "__enter__", None)
        return enter() if enter is not None else self.context

    def __exit__(self, exc_type, exc_value, exc_traceback):
        leave = getattr(self.context, "__exit__", None)
        if leave is not None:
            leave(exc_type, exc_value, exc_traceback)
