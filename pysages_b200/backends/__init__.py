"""Engine bindings: snapshot glue, per-engine samplers and the mock engines used without an MD code."""

from .core import SamplingContext, supported_backends  # noqa: F401
from .snapshot import Box, HelperMethods, Layout, Snapshot, copy, restore  # noqa: F401
