"""
pysages_b200 -- B200-native implementation of the PySAGES per-timestep bias hot path.

Keeps the `pysages.run` / method / CV API (see DESIGN.md for the exact scope); all numerical work
runs in hand-written sm_100a CUDA kernels behind the C ABI of include/pysages_b200.h.
"""

from . import backends, colvars, grids, methods  # noqa: F401
from .backends.core import supported_backends  # noqa: F401
from .core import Result, analyze, run  # noqa: F401
from .grids import Chebyshev, Grid, Periodic, Regular  # noqa: F401
from .methods import CVRestraints, freeze  # noqa: F401
from .serialization import load, save  # noqa: F401
from .utils import HistogramLogger, MetaDLogger, ReplicasConfiguration, SerialExecutor  # noqa: F401

__version__ = "0.1.0"
