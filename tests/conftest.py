import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) where no CUDA device exists, whatever -m says."""
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter):
    """Which oracle pinned this session: the live reference (JAX) or the restatement under oracle/."""
    try:
        from oracle import ref_jax

        mods, why = ref_jax.probe()
        branch = "LIVE reference (PySAGES on JAX)" if mods else "restatement in oracle/ (live reference unavailable)"
        terminalreporter.write_line(f"reference oracle: {branch} -- {why}")
    except Exception as e:  # pragma: no cover
        terminalreporter.write_line(f"reference oracle: probe failed ({e})")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
