"""Pieces shared by the engine adapters (hoomd.py, openmm.py, lammps.py, mock.py)."""

from .. import _native as N
from .snapshot import HelperMethods, Layout

LAMMPS_CONVERSION = {"real": 2390.0573615334906, "metal": 1.0364269e-4, "electron": 1.06657236}  # lammps.py:33

LAYOUTS = {
    "hoomd": Layout(N.LAYOUT_HOOMD, "hoomd"),
    "openmm-cuda": Layout(N.LAYOUT_OPENMM_CUDA, "openmm-cuda"),
    "openmm-cpu": Layout(N.LAYOUT_PLAIN3, "openmm-cpu"),
    "lammps": Layout(N.LAYOUT_LAMMPS, "lammps"),
}


def fused_helpers(layout):
    """HelperMethods for the fused native step: the snapshot gather and the in-place force update
    are part of psb_step, so `query` has nothing to return and `bias` nothing left to do."""

    def query(snapshot):
        raise NotImplementedError(
            "the snapshot gather (positions / indices / momenta) is fused into the CUDA step; "
            "there is no separate query stage"
        )

    factor = layout.force_unit_factor
    return HelperMethods(query, lambda: 3, (lambda x: factor * x), layout)


def tensor_from_capsule(capsule):
    """Zero-copy torch view of a DLPack capsule (the plugins hand out one-shot capsules)."""
    from torch.utils.dlpack import from_dlpack

    return from_dlpack(capsule)
