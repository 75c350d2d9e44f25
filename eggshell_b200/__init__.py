"""eggshell_b200 — B200-native batched rigid-body step, drop-in for eggshell's Ensemble::Step path.

The compute lives in ``libeggshell_b200.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/egg_cuda.h``).  This package is the Python host mirror of that ABI plus the synthetic
scene builders; it never imports anything from ``oracle/`` and has no CPU fallback.
"""
from .batch import (Batch, EggError, lib, lib_path, pinned_empty, EXPORTS,  # noqa: F401
                    SOLVER_DENSE_MURTY, SOLVER_PGS, SOLVER_JACOBI, SOLVER_SOR,
                    CFM_AUTO, CFM_ALWAYS, CFM_NEVER, QUIRKS_REFERENCE, QUIRK_GS_BOUNDS_SHIFT,
                    QUIRK_DENSE_IGNORES_BOUNDS, OPEN_DYNAMICS_ENGINE)
from . import scenes  # noqa: F401

__version__ = "0.1"
