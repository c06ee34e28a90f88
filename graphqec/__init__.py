"""``graphqec`` -- the reference's import name, served by the B200-native implementation.

"The Python API stays unchanged" (BASELINE.json north_star): a user of adelshb/graphqec writes
``from graphqec import RepetitionCode, ThresholdLAB`` (``/root/reference/README.md:62,75``) or imports the
sub-modules (``graphqec.codes.base_code``, ``graphqec.lab.threshold.threshold_lab``, ``graphqec.math`` ...;
layout of ``/root/reference/graphqec/``).  This package re-exports ``graphqec_b200`` under those names: every
``graphqec.<path>`` module object IS the ``graphqec_b200.<path>`` one, nothing is duplicated.
"""
import importlib as _importlib
import sys as _sys

import graphqec_b200 as _impl
from graphqec_b200 import *  # noqa: F401,F403
from graphqec_b200 import (  # noqa: F401
    BaseCode, BivariateBicycleCode, Circuit, CssCode, DetectorErrorModel, RepetitionCode, RotatedSurfaceCode,
    Task, TaskStats, ThresholdLAB, UnrotatedSurfaceCode, target_rec, wilson_interval,
)

__version__ = _impl.__version__

for _name in (
    "codes", "codes.base_code", "codes.repetition_code", "codes.rotated_surface_code", "codes.unrotated_surface_code",
    "codes.css_code", "codes.bivariate_bicycle_code", "lab", "lab.threshold", "lab.threshold.threshold_lab",
    "math", "measurement", "stab", "circuit", "dem", "stats", "sinter_compat", "backend",
):
    _mod = _importlib.import_module("graphqec_b200." + _name)
    _sys.modules["graphqec." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod
del _name, _mod
