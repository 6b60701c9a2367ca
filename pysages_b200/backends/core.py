"""
SamplingContext (pysages/backends/core.py:10-73): detects the engine a user-built context belongs
to and binds the sampler to it.
"""

import importlib

supported = ("hoomd", "openmm", "lammps", "mock")


def supported_backends():
    return supported


class SamplingContext:
    """backends/core.py:10-73"""

    def __init__(self, method, context_generator, callback=None, context_args=None, **kwargs):
        context_args = {} if context_args is None else context_args
        self._backend_name = None
        context = context_generator(**context_args)
        module_name = type(context).__module__
        if module_name.startswith("hoomd"):
            self._backend_name = "hoomd"
        elif module_name.startswith(("openmm", "simtk.openmm")):
            self._backend_name = "openmm"
        elif module_name.startswith("lammps"):
            self._backend_name = "lammps"
        elif module_name.startswith("pysages_b200.backends.mock"):
            self._backend_name = "mock"
        if self._backend_name is None:
            backends = ", ".join(supported)
            raise ValueError(f"Invalid backend {module_name}: supported options are ({backends})")
        self.context = context
        self.method = method
        self.run = None
        backend = importlib.import_module("." + self._backend_name, package="pysages_b200.backends")
        self.sampler = backend.bind(self, callback, **kwargs)
        assert self.run is not None

    def get_backend_name(self):
        return self._backend_name

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        pass
