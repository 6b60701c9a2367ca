"""
CPU-only checks of the drop-in boundary: the shared library loads, exports EVERY function that
include/pysages_b200.h declares, the ctypes table covers exactly that set, the struct layouts agree
with the header (sizes computed by the C compiler), and nothing computes without a GPU.
"""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pysages_b200.h")


@pytest.fixture(scope="module")
def native():
    from pysages_b200 import _native, build

    build.build()  # nvcc cross-compiles without a GPU; a no-op when the library is current
    return _native


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native):
    lib = native.load()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/pysages_b200.h but not exported"
    assert sorted(native.EXPORTS) == names, "pysages_b200/_native.py binds a different set than the header declares"
    assert lib.psb_abi_version() == 2


def test_struct_layouts_match_the_header(native, tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text(
        '#include <stdio.h>\n#include "pysages_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(psb_cv_desc), sizeof(psb_grid_desc), '
        "sizeof(psb_method_desc), sizeof(psb_layout_desc), sizeof(psb_snapshot)); return 0;}\n"
    )
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    got = [ctypes.sizeof(t) for t in (native.CvDesc, native.GridDesc, native.MethodDesc, native.LayoutDesc, native.SnapshotDesc)]
    assert got == sizes


def test_no_cpu_fallback(native):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = native.load()
    cv = native.CvDesc()
    handle = ctypes.c_void_p()
    st = lib.psb_create(ctypes.byref(cv), 1, ctypes.byref(native.MethodDesc()), ctypes.byref(native.LayoutDesc()),
                        ctypes.byref(handle))
    assert st == native.ERR_CUDA and b"no CPU fallback" in lib.psb_last_error()
    import pysages_b200 as ps
    from pysages_b200.backends.snapshot import Box, HelperMethods, Snapshot
    import numpy as np

    snap = Snapshot(torch.zeros(4, 4), torch.zeros(4, 4), torch.zeros(4, 4), torch.arange(4, dtype=torch.int32),
                    Box(np.eye(3), np.zeros(3)), 0.005, dict(images=torch.zeros(4, 3, dtype=torch.int32)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ps.methods.HarmonicBias([ps.colvars.Component([0], 0)], 1.0, [0.0]).build(snap, HelperMethods(None, lambda: 3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pysages_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), os.path.join(d, f)
