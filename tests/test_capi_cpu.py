"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/egg_cuda.h declares, carries the reference's constants as defaults and refuses to run
without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "egg_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(egg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import eggshell_b200 as E
    L = E.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in egg_cuda.h but not exported"
    assert set(E.EXPORTS) == set(names)


def test_desc_defaults_are_the_reference_constants():
    import eggshell_b200 as E
    from eggshell_b200.batch import EggDesc
    d = EggDesc()
    assert E.lib().egg_desc_default(C.byref(d), 7, 10, 3) == 0
    assert (d.n_worlds, d.n_bodies, d.n_joints) == (7, 10, 3)
    assert d.k_max == 500 and d.tol == 1e-9            # sparse_iterations.cc:19, constants.h:5
    assert d.cfm == 0.01 and d.erp == 0.2              # ensembles.cc:14, ensembles.h:166
    assert list(d.gravity) == [0.0, 0.0, -9.8]         # constants.h:8
    assert d.min_constraint_dist == 1e-6               # ensembles.cc:15
    assert d.solver == E.SOLVER_DENSE_MURTY            # ensembles.cc:21 kSparseImplementation = false
    assert d.quirks == E.QUIRKS_REFERENCE and d.precision == 64


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import eggshell_b200 as E
    with pytest.raises(E.EggError) as ei:
        E.Batch(2, 4, 0, solver=E.SOLVER_PGS)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "eggshell_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|pyoracle|liboracle|#include\s+\"[^\"]*orc_|CDLL\([^)]*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f


def test_scene_shapes_and_algorithmic_bytes():
    import eggshell_b200.scenes as S
    assert S.algorithmic_bytes_per_world_step(10, 0) == 4168
    assert S.algorithmic_bytes_per_world_step(64, 0) == 26632     # C3
    assert S.algorithmic_bytes_per_world_step(32, 32) == 15112    # C4
    assert S.algorithmic_bytes_per_world_step(20, 19) == 9392     # C5
    for sc, n, nj in ((S.stack10(3), 10, 0), (S.pile64(2), 64, 0), (S.chain32(2), 32, 32), (S.legged20(2), 20, 19), (S.chain(1), 10, 10)):
        assert sc["p"].shape == (sc["W"], n, 3) and sc["R"].shape == (sc["W"], n, 3, 3) and sc["nj"] == nj
        RtR = np.einsum("wnij,wnik->wnjk", sc["R"], sc["R"])
        assert np.allclose(RtR, np.eye(3), atol=1e-12)
