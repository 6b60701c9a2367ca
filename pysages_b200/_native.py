"""
ctypes binding of libpysages_b200.so (include/pysages_b200.h).

The library is the product: there is NO Python / CPU fallback.  If the shared library is
missing or no sm_100 device is present, every compute entry point raises.
"""

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int32, c_int64, c_uint32, c_void_p

MAX_CVS, MAX_K, MAX_GROUPS, MAX_GRID_DIMS = 8, 8, 32, 4
if os.environ.get("PSB_ABI_LIMITS"):  # experiment builds with -DPSB_MAX_CVS=... (tools/ab_limits.sh): same numbers here
    MAX_CVS, MAX_K, MAX_GROUPS = (int(x) for x in os.environ["PSB_ABI_LIMITS"].split(","))

OK, ERR_VALUE, ERR_RUNTIME, ERR_KEY, ERR_CUDA, ERR_UNSUPPORTED = range(6)

CV_COMPONENT, CV_DISTANCE, CV_DISPLACEMENT, CV_ANGLE, CV_DIHEDRAL, CV_ROG, CV_RMSD, CV_COORDINATION, CV_GYRATION, CV_RING = range(1, 11)
RING_AMPLITUDE, RING_PHASE = 0, 1
CV_ERMSD = 11
GYR_ASPHERICITY, GYR_ACYLINDRICITY, GYR_ANISOTROPY = 3, 4, 5
GRID_NONE, GRID_REGULAR, GRID_PERIODIC, GRID_CHEBYSHEV = range(4)
METHOD_UNBIASED, METHOD_HARMONIC, METHOD_METAD, METHOD_ABF = range(4)
LAYOUT_HOOMD, LAYOUT_OPENMM_CUDA, LAYOUT_LAMMPS, LAYOUT_PLAIN3 = range(1, 5)


class CvDesc(Structure):
    _fields_ = [
        ("kind", c_int32),
        ("axis", c_int32),
        ("n_groups", c_int32),
        ("reserved", c_int32),
        ("tags", POINTER(c_uint32)),
        ("group_offsets", POINTER(c_uint32)),
        ("reference", POINTER(c_double)),
        ("weights", POINTER(c_double)),
        ("r0", c_double),
        ("nn", c_int32),
        ("mm", c_int32),
    ]


class GridDesc(Structure):
    _fields_ = [
        ("kind", c_int32),
        ("ndim", c_int32),
        ("lower", c_double * MAX_GRID_DIMS),
        ("upper", c_double * MAX_GRID_DIMS),
        ("shape", c_int32 * MAX_GRID_DIMS),
    ]


class MethodDesc(Structure):
    _fields_ = [
        ("kind", c_int32),
        ("variant", c_int32),
        ("kspring", c_double * (MAX_K * MAX_K)),
        ("center", c_double * MAX_K),
        ("height", c_double),
        ("sigma", c_double * MAX_K),
        ("periods", c_double * MAX_K),
        ("stride", c_int64),
        ("ngaussians", c_int64),
        ("well_tempered", c_int32),
        ("shared_hills", c_int32),
        ("deltaT", c_double),
        ("kB", c_double),
        ("abf_N", c_int64),
        ("use_pinv", c_int32),
        ("has_restraints", c_int32),
        ("r_lower", c_double * MAX_K),
        ("r_upper", c_double * MAX_K),
        ("r_kl", c_double * MAX_K),
        ("r_ku", c_double * MAX_K),
        ("force_unit_factor", c_double),
        ("grid", GridDesc),
    ]


MAX_DENSE = 8
MAX_WIDTH = 128
ACT_TANH, ACT_SIN = 0, 1


class ForceNetDesc(Structure):  # psb_force_net
    _fields_ = [
        ("n_dense", c_int32),
        ("widths", c_int32 * (MAX_DENSE + 1)),
        ("activation", c_int32),
        ("params_on_device", c_int32),
        ("gradient", c_int32),
        ("reserved", c_int32),
        ("omega0", c_double),
        ("omega", c_double),
        ("predict_after", c_int64),
        ("params", c_void_p),
        ("mean", c_double * MAX_K),
        ("std", c_double * MAX_K),
    ]


SERIES_FOURIER, SERIES_MONOMIAL = 1, 2


class ForceSeriesDesc(Structure):  # psb_force_series
    _fields_ = [
        ("kind", c_int32),
        ("n_terms", c_int32),
        ("oob_zero", c_int32),
        ("values_on_device", c_int32),
        ("predict_after", c_int64),
        ("exponents", c_void_p),
        ("values", c_void_p),
    ]


class LayoutDesc(Structure):
    _fields_ = [
        ("layout", c_int32),
        ("real_bytes", c_int32),
        ("natoms", c_int64),
        ("unwrap", c_int32),
        ("need_momenta", c_int32),
        ("dense_outputs", c_int32),
        ("ntypes", c_int32),
        ("has_tags", c_int32),
        ("reserved", c_int32),
    ]


class SnapshotDesc(Structure):
    _fields_ = [
        ("positions", c_void_p),
        ("vel_mass", c_void_p),
        ("forces", c_void_p),
        ("ids", c_void_p),
        ("images", c_void_p),
        ("masses", c_void_p),
        ("types", c_void_p),
        ("box_diag", c_double * 3),
        ("dt", c_double),
        ("tags", c_void_p),
    ]


class SeriesFitDesc(Structure):
    _fields_ = [
        ("npoints", c_int64),
        ("k", c_int32),
        ("nterms", c_int32),
        ("periodic", c_int32),
        ("reserved", c_int32),
        ("hist", c_void_p),
        ("Fsum", c_void_p),
        ("pinv", c_void_p),
        ("out", c_void_p),
        ("work", c_void_p),
    ]


class LmProblem(Structure):
    _fields_ = [
        ("n_dense", c_int32),
        ("widths", c_int32 * (8 + 1)),
        ("activation", c_int32),
        ("max_iters", c_int32),
        ("max_loops", c_int32),
        ("reserved", c_int32),
        ("n", c_int64),
        ("x", c_void_p),
        ("y", c_void_p),
        ("params", c_void_p),
        ("lower", c_double * MAX_K),
        ("upper", c_double * MAX_K),
        ("reg", c_double),
        ("mu_0", ctypes.c_float),
        ("mu_c", ctypes.c_float),
        ("mu_min", ctypes.c_float),
        ("mu_max", ctypes.c_float),
        ("rho_c", ctypes.c_float),
        ("rho_min", ctypes.c_float),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "psb_create": (c_int32, [POINTER(CvDesc), c_int32, POINTER(MethodDesc), POINTER(LayoutDesc), POINTER(c_void_p)]),
    "psb_destroy": (None, [c_void_p]),
    "psb_last_error": (c_char_p, []),
    "psb_step": (c_int32, [c_void_p, POINTER(SnapshotDesc), c_int64, c_void_p]),
    "psb_evaluate_cvs": (c_int32, [c_void_p, POINTER(SnapshotDesc), c_void_p]),
    "psb_run_cycle": (c_int32, [c_void_p, POINTER(SnapshotDesc), c_int32, c_int64, c_void_p]),
    "psb_step_host": (c_int32, [c_void_p, POINTER(SnapshotDesc), c_int64, c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    "psb_host_release": (c_int32, [c_void_p]),
    "psb_evaluate_cvs_host": (c_int32, [c_void_p, POINTER(SnapshotDesc), c_void_p]),
    "psb_batch_create": (c_int32, [POINTER(c_void_p), c_int32, POINTER(c_void_p)]),
    "psb_batch_destroy": (None, [c_void_p]),
    "psb_batch_step": (c_int32, [c_void_p, POINTER(SnapshotDesc), c_int64, c_void_p]),
    "psb_batch_run_cycle": (c_int32, [c_void_p, POINTER(SnapshotDesc), c_int32, c_int64, c_void_p]),
    "psb_batch_launch_count": (c_int64, [c_void_p]),
    "psb_state_info": (c_int32, [c_void_p, c_char_p, POINTER(c_int64), POINTER(c_int32), POINTER(c_void_p)]),
    "psb_get_state": (c_int32, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "psb_set_state": (c_int32, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "psb_walker_begin": (c_int32, [c_void_p, POINTER(SnapshotDesc), c_int64, c_void_p, c_void_p]),
    "psb_walker_finish": (c_int32, [c_void_p, POINTER(SnapshotDesc), c_void_p, c_int32, c_void_p]),
    "psb_walker_record_doubles": (c_int32, [c_void_p]),
    "psb_next_step_deposits": (c_int32, [c_void_p]),
    "psb_grid_index": (c_int32, [POINTER(GridDesc), POINTER(c_double), c_int64, POINTER(c_uint32)]),
    "psb_set_option": (c_int32, [c_void_p, c_char_p, c_int64]),
    "psb_set_force_net": (c_int32, [c_void_p, POINTER(ForceNetDesc), c_void_p]),
    "psb_set_force_series": (c_int32, [c_void_p, POINTER(ForceSeriesDesc), c_void_p]),
    "psb_launch_count": (c_int64, [c_void_p]),
    "psb_cycle_uses_graph": (c_int32, [c_void_p]),
    "psb_launch_info": (c_int32, [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int64)]),
    "psb_launch_floor": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_int64, c_void_p]),
    "psb_latency_probe": (c_int32, [POINTER(c_double)]),
    "psb_fma_peak": (c_int32, [POINTER(c_double)]),
    "psb_lm_fit": (c_int32, [POINTER(LmProblem), POINTER(c_void_p), c_void_p]),
    "psb_lm_result": (c_int32, [c_void_p, POINTER(c_int32), POINTER(c_double), POINTER(ctypes.c_float), POINTER(c_int32), c_void_p]),
    "psb_lm_destroy": (None, [c_void_p]),
    "psb_abi_version": (c_int32, []),
    "psb_wants_tags": (c_int32, [c_void_p]),
    "psb_series_fit": (c_int32, [POINTER(SeriesFitDesc), c_void_p]),
}

LIB_PATH = os.environ.get("PSB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libpysages_b200.so")  # PSB_LIB: A/B runs of two builds on one box
_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load(path=None):
    """Load libpysages_b200.so and declare every entry point of include/pysages_b200.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise NativeLibraryError(
            f"{p} is missing: build it with `python -m pysages_b200.build` "
            "(pysages_b200 has no CPU fallback; the CUDA extension is required)"
        )
    lib = ctypes.CDLL(p)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if path is None:
        _lib = lib
    return lib


_EXC = {ERR_VALUE: ValueError, ERR_RUNTIME: RuntimeError, ERR_KEY: KeyError, ERR_CUDA: RuntimeError,
        ERR_UNSUPPORTED: NotImplementedError}


def check(status):
    """Map a psb_status to the reference's exception types."""
    if status != OK:
        msg = load().psb_last_error().decode("utf8", "replace")
        raise _EXC.get(status, RuntimeError)(msg)
