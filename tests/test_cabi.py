"""CPU-side checks of the boundary: the library builds, loads, and exports every symbol
include/ionob200.h declares; the host logic fails loudly without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ionob200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iono_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from ionotomo_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
        assert s in _lib.SIGNATURES, "ctypes prototype missing for " + s
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.iono_version() == 1
    assert lib.iono_misfit_scratch_elems() > 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import ionotomo_b200 as ib
    from ionotomo_b200._lib import IonoError
    tci = ib.TriCubic(np.linspace(0, 1, 4), np.linspace(0, 1, 4), np.linspace(0, 1, 4), np.zeros((4, 4, 4)))
    with pytest.raises(IonoError):
        tci.interp(np.array([0.5]), np.array([0.5]), np.array([0.5]))
    with pytest.raises(IonoError):
        ib.forward_equation(np.zeros((1, 1, 1, 4, 8)), 1e11, tci, 0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ionotomo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(from|import)\s+oracle|oracle[./]|ionotomo_oracle", txt), (dirpath, f)


def test_host_bisection_matches_reference(golden):
    import numpy as np
    from ionotomo_b200.geometry.tri_cubic import bisection
    g = golden("tricubic")
    np.testing.assert_array_equal([bisection(g["xvec"], v) for v in g["bvals"]], g["bidx"])
